"""CPU oracle for program semantics: a direct, one-witness-at-a-time interpreter of CVM text.

TEST INFRASTRUCTURE ONLY (only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
leg may import this).  The reference has no CVM interpreter; the meaning of
every CVM instruction is therefore taken from the C++ the *same bucket* emits through
`impl WriteC` and from the runtime it links against:

  get_signal/set_signal       signalValues[mySignalStart + i]      load_bucket.rs:249-457, store_bucket.rs:444-648
  get_cmp_signal/set_cmp_*    componentMemory[mySubcomponents[c]]  store_bucket.rs:662-800 (inputCounter, run on 0)
  ff.load/ff.store            lvar[i]                              c_code_generator.rs:76-81
  ff.<op>                     Fr_<op>                              oracle/fr_model.py <- generic/fr.cpp
  ff.wrap_i64                 Fr_toInt                             generic/fr.cpp:1102-1170
  loop/if/else/end/break      while(Fr_isTrue) / if(Fr_isTrue)     loop_bucket.rs:78-94, branch_bucket.rs:101-124
  error 0                     failed assert                        assert_bucket.rs:71-86
  ff.call / return            lvarcall arena + destination copy    call_bucket.rs:465-847, return_bucket.rs:98-122
  component creation          T_create / run-if-no-inputs          template.rs:240-331, create_component_bucket.rs:206-354
  witness extraction          signalValues[witness2Signal[i]]      common/main.cpp:324-330

Program-level parity with the reference runtime is pinned separately (tests/test_ref_runtime.py:
the same programs emitted as C++ in the WriteC shapes and linked against the reference's
calcwit.cpp/main.cpp/fr.cpp must produce byte-identical .wtns).
"""
from __future__ import annotations

import re

from . import fr_model as M

SPR_BASE = 1 << 30

ST_OK, ST_ASSERT, ST_TOINT, ST_DIVZERO, ST_INPUT = 0, 1, 2, 3, 4


class WitnessError(Exception):
    def __init__(self, status, msg=""):
        super().__init__(msg)
        self.status = status


FF_BIN = {"ff.add": M.add, "ff.sub": M.sub, "ff.mul": M.mul, "ff.idiv": M.idiv, "ff.rem": M.mod, "ff.pow": M.pow_,
          "ff.shl": M.shl, "ff.shr": M.shr, "ff.band": M.band, "ff.bor": M.bor, "ff.bxor": M.bxor,
          "ff.lt": M.lt, "ff.le": M.leq, "ff.gt": M.gt, "ff.ge": M.geq, "ff.eq": M.eq, "ff.neq": M.neq,
          "ff.and": M.land, "ff.or": M.lor}
I64_BIN = {"i64.add": lambda a, b: a + b, "i64.sub": lambda a, b: a - b, "i64.mul": lambda a, b: a * b,
           "i64.lt": lambda a, b: int(a < b), "i64.le": lambda a, b: int(a <= b), "i64.eq": lambda a, b: int(a == b),
           "i64.gt": lambda a, b: int(a > b), "i64.ge": lambda a, b: int(a >= b), "i64.neq": lambda a, b: int(a != b)}


class Code:
    def __init__(self, header):
        self.header = header
        self.ins = []          # (op, dst, args)
        self.match = {}        # pc of if -> (else_pc|None, end_pc) ; loop -> end_pc ; end -> ('if'|'loop', start)
        self.loop_of = {}      # pc of break/continue -> loop start pc
        self.n_inputs = 0
        self.n_outputs = 0
        self.n_signals = 0
        self.n_subcmps = 0
        self.local_memory = 0
        self.is_function = False


_ARG_RE = re.compile(r"^(i64\.memory|signal|subcmpsignal)\((.*)\)$")


def _dims_size(decls):
    """'[ ff 0  ff 1 3 ]' payload -> total number of field elements."""
    toks = decls.split()
    total, k = 0, 0
    while k < len(toks):
        nd = int(toks[k + 1])
        size = 1
        for d in toks[k + 2:k + 2 + nd]:
            size *= int(d)
        total += size
        k += 2 + nd
    return total


class Program:
    def __init__(self, text):
        self.prime = None
        self.n_signals = 0
        self.start = None
        self.witness = []
        self.codes = {}
        self.parse(text)

    def parse(self, text):
        cur = None
        for raw in text.split("\n"):
            line = raw.strip()
            if not line:
                continue
            if line.startswith(";;%%create_cmp"):
                t = line.split()
                cur.ins.append(("create_cmp", None, [int(t[1]), t[2].lstrip("$")] + [int(x) for x in t[3:8]]))
                continue
            if line.startswith(";;") or line.startswith("//"):
                continue
            if line.startswith("%%"):
                t = line.split()
                d = t[0]
                if d == "%%prime":
                    self.prime = int(t[1])
                    if self.prime != M.Q:
                        raise ValueError("only BN254 is supported")
                elif d == "%%signals":
                    self.n_signals = int(t[1])
                elif d == "%%start":
                    self.start = t[1]
                elif d == "%%witness":
                    self.witness = [int(x) for x in t[1:]]
                elif d in ("%%template", "%%function"):
                    cur = Code(t[1])
                    self.codes[t[1]] = cur
                    br = re.findall(r"\[([^\]]*)\]", line)
                    if d == "%%template":
                        cur.n_inputs = _dims_size(br[0])          # bracket 1 = Input wires (build.rs:87-104)
                        cur.n_outputs = _dims_size(br[1])         # bracket 2 = Output wires
                        cur.n_signals = int(br[2])
                        cur.n_subcmps = int(br[3])
                    else:
                        cur.is_function = True
                continue
            t = line.split()
            if t[0] == "local.memory":
                cur.local_memory = int(t[1])
                continue
            if len(t) >= 3 and t[1] == "=":
                if t[0].startswith("i64.") or t[0].startswith("ff."):
                    # emitter defect: a literal used as a mutable register has no
                    # consistent meaning; refuse instead of guessing.
                    raise ValueError("assignment to literal operand %r" % t[0])
                if len(t) == 3:
                    cur.ins.append(("mov", t[0], [t[2]]))
                else:
                    cur.ins.append((t[2], t[0], t[3:]))
            else:
                cur.ins.append((t[0], None, t[1:]))
        for c in self.codes.values():
            self.link(c)

    @staticmethod
    def link(c):
        stack = []
        for pc, (op, _d, _a) in enumerate(c.ins):
            if op == "if":
                stack.append(["if", pc, None])
            elif op == "loop":
                stack.append(["loop", pc, None])
            elif op == "else":
                stack[-1][2] = pc
            elif op == "end":
                kind, start, els = stack.pop()
                if kind == "if":
                    c.match[start] = (els, pc)
                    if els is not None:
                        c.match[els] = pc
                else:
                    c.match[start] = pc
                c.match[pc] = (kind, start)
            elif op in ("break", "continue"):
                for fr in reversed(stack):
                    if fr[0] == "loop":
                        c.loop_of[pc] = fr[1]
                        break
        assert not stack, "unbalanced control flow in " + c.header


class Component:
    __slots__ = ("code", "start", "counter", "subs")

    def __init__(self, code, start):
        self.code, self.start, self.counter = code, start, code.n_inputs
        self.subs = {}


class Machine:
    """Runs one witness.  `counters` accumulates dynamic op counts (N_mul for the roofline)."""

    def __init__(self, prog):
        self.p = prog
        self.counters = {"mul": 0, "div": 0, "ops": 0}
        self.max_ops = 500_000_000

    def witness(self, inputs):
        """inputs: canonical ints for main's input signals in signal order -> witness values."""
        p = self.p
        self.sig = [0] * p.n_signals
        self.sig[0] = 1                                            # calcwit.cpp:34
        main = Component(p.codes[p.start], 1)                      # circuit.rs:539
        if len(inputs) != main.code.n_inputs:
            raise WitnessError(ST_INPUT, "expected %d inputs" % main.code.n_inputs)
        n_out = main.code.n_outputs          # main inputs sit right after main's outputs (A.5)
        for k, v in enumerate(inputs):
            self.sig[1 + n_out + k] = v % M.Q
        self.run(main)
        return [self.sig[s] for s in p.witness]

    # ---- operands
    @staticmethod
    def lit(tok):
        if tok.startswith("i64."):
            return int(tok[4:])
        if tok.startswith("ff."):
            return int(tok[3:]) % M.Q
        if tok.startswith("i64") and tok[3:].lstrip("-").isdigit():     # emitter defect: "i64<n>" (A.4 defect 1)
            return int(tok[3:])
        return int(tok)

    def val(self, regs, tok):
        v = regs.get(tok)
        if v is not None:
            return v
        if tok == "spr":
            return SPR_BASE
        return self.lit(tok)

    # ---- execution
    def run(self, comp):
        self.exec(comp.code, comp, {}, {})

    def exec(self, code, comp, regs, lvar, dest=None):
        ins, match = code.ins, code.match
        sig = self.sig
        pc, n = 0, len(ins)
        cnt = self.counters
        while pc < n:
            op, dst, a = ins[pc]
            pc += 1
            cnt["ops"] += 1
            if cnt["ops"] > self.max_ops:
                raise RuntimeError("CVM oracle: op budget exceeded (runaway loop?)")
            f = FF_BIN.get(op)
            if f is not None:
                x, y = self.val(regs, a[0]), self.val(regs, a[1])
                if op == "ff.mul":
                    cnt["mul"] += 1
                try:
                    regs[dst] = f(x % M.Q, y % M.Q)
                except M.FrError as e:
                    raise WitnessError(ST_DIVZERO, str(e))
                continue
            f = I64_BIN.get(op)
            if f is not None:
                regs[dst] = f(self.val(regs, a[0]), self.val(regs, a[1]))
                continue
            if op == "mov":
                regs[dst] = self.val(regs, a[0])
            elif op == "ff.div":
                cnt["div"] += 1
                cnt["mul"] += 1
                y = self.val(regs, a[1]) % M.Q
                regs[dst] = M.div(self.val(regs, a[0]) % M.Q, y)      # a / 0 = 0, as the reference's Fr_div (fr_model.inv)
            elif op == "ff.eqz":
                regs[dst] = int(self.val(regs, a[0]) % M.Q == 0)
            elif op == "ff.bnot":
                regs[dst] = M.bnot(self.val(regs, a[0]) % M.Q)
            elif op == "ff.wrap_i64":
                try:
                    regs[dst] = M.to_int(self.val(regs, a[0]) % M.Q)
                except M.FrError as e:
                    raise WitnessError(ST_TOINT, str(e))
            elif op == "ff.load":
                regs[dst] = lvar.get(self.val(regs, a[0]), 0)
            elif op == "ff.store":
                lvar[self.val(regs, a[0])] = self.val(regs, a[1]) % M.Q
            elif op == "get_signal":
                regs[dst] = sig[comp.start + self.val(regs, a[0])]
            elif op == "set_signal":
                sig[comp.start + self.val(regs, a[0])] = self.val(regs, a[1]) % M.Q
            elif op == "get_cmp_signal":
                sub = comp.subs[self.val(regs, a[0])]
                regs[dst] = sig[sub.start + self.val(regs, a[1])]
            elif op.startswith("set_cmp_input"):
                sub = comp.subs[self.val(regs, a[0])]
                sig[sub.start + self.val(regs, a[1])] = self.val(regs, a[2]) % M.Q
                if op == "set_cmp_input_cnt":
                    sub.counter -= 1
                elif op == "set_cmp_input_run":
                    self.run(sub)
                elif op == "set_cmp_input_cnt_check":
                    sub.counter -= 1
                    if sub.counter == 0:
                        self.run(sub)
            elif op == "create_cmp":
                slot, hdr, so, sj, _co, _cj, num = a
                tcode = self.p.codes[hdr]
                for k in range(num):
                    sub = Component(tcode, comp.start + so + k * sj)
                    comp.subs[slot + k] = sub
                    if tcode.n_inputs == 0:                    # template.rs:326-331
                        self.run(sub)
            elif op == "if":
                if self.val(regs, a[0]) % M.Q == 0:
                    els, end = match[pc - 1]
                    pc = (els + 1) if els is not None else (end + 1)
            elif op == "else":
                pc = match[pc - 1] + 1
            elif op == "end":
                pass                                           # falling out of an if, or out of a loop
            elif op == "loop":
                pass
            elif op == "continue":
                pc = code.loop_of[pc - 1] + 1
            elif op == "break":
                pc = match[code.loop_of[pc - 1]] + 1
            elif op == "error":
                raise WitnessError(ST_ASSERT, "error %s in %s" % (a[0], code.header))
            elif op == "ff.call":
                self.call(comp, regs, lvar, a)
            elif op == "return":
                dlv, daddr, dsize = dest
                if a[1] == "1":
                    dlv[daddr] = self.val(regs, a[0]) % M.Q
                else:
                    src = self.val(regs, a[0])
                    for k in range(min(self.val(regs, a[1]), dsize)):
                        dlv[daddr + k] = lvar.get(src + k, 0)
                return
            elif op in ("get_template_id", "get_template_signal_position", "get_template_signal_size",
                        "get_template_signal_dimension", "get_template_signal_type"):
                raise NotImplementedError("mapped (mixed component array) accesses need the io-map, which the "
                                          "fork's .cvm does not carry (SURVEY.md F3)")
            else:
                raise ValueError("unknown CVM instruction %r" % op)

    def call(self, comp, regs, lvar, a):
        fn = self.p.codes[a[0].lstrip("$")]
        daddr, dsize = self.val(regs, a[1]), self.val(regs, a[2])
        flv = {}
        pos = 0
        for tok in a[3:]:
            m = _ARG_RE.match(tok)
            if m is None:
                flv[pos] = self.val(regs, tok) % M.Q
                pos += 1
                continue
            parts = m.group(2).split(",")
            n = int(parts[-1])
            if m.group(1) == "i64.memory":
                base = self.val(regs, parts[0])
                for k in range(n):
                    flv[pos + k] = lvar.get(base + k, 0)
            elif m.group(1) == "signal":
                base = comp.start + self.val(regs, parts[0])
                for k in range(n):
                    flv[pos + k] = self.sig[base + k]
            else:
                sub = comp.subs[self.val(regs, parts[0])]
                base = sub.start + self.val(regs, parts[1])
                for k in range(n):
                    flv[pos + k] = self.sig[base + k]
            pos += n
        fregs = {"destination": daddr, "destination_size": dsize}
        self.exec(fn, comp, fregs, flv, dest=(lvar, daddr, dsize))


def load(path_or_text):
    text = path_or_text
    if "\n" not in path_or_text:
        with open(path_or_text) as f:
            text = f.read()
    return Program(text)


def compute_witness(prog, inputs):
    return Machine(prog).witness(inputs)
