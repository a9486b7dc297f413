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
import struct

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
        self.io_map = {}              # template-instance id -> [(offset, lengths[1..], size, busId)] by signal code
        self.parse(text)

    def parse(self, text):
        cur = None
        for raw in text.split("\n"):
            line = raw.strip()
            if not line:
                continue
            if line.startswith(";;%%create_cmp"):
                t = line.split()
                pos = [int(x) for x in t[9:]] if len(t) > 8 else None            # "| p0 p1 ...": arrays with undefined positions
                cur.ins.append(("create_cmp", None, [int(t[1]), t[2].lstrip("$")] + [int(x) for x in t[3:8]] + [pos]))
                continue
            if line.startswith(";;%%io_map"):
                # ;;%%io_map <template id> <n> { offset len <len lengths[1..]> size busId }*  (the .dat record as text)
                t = [int(x) for x in line.split()[1:]]
                self.io_map[t[0]] = _io_defs(t, 2, t[1])[0]
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
                if len(t) == 3:
                    cur.ins.append(("mov", t[0], [t[2]]))
                else:
                    cur.ins.append((t[2], t[0], t[3:]))
            else:
                cur.ins.append((t[0], None, t[1:]))
        for c in self.codes.values():
            self.fix_array_eq(c)
            self.fix_literal_registers(c)
            self.link_returns(c)
            self.link(c)

    # ---- emitter defects that are honoured (same rules as csrc/cvm_parse.hpp, written independently)
    @staticmethod
    def _is_literal(tok):
        return (tok.startswith("i64.") and tok[4:].lstrip("-").isdigit()) or (tok.startswith("ff.") and tok[3:].isdigit())

    @staticmethod
    def _loops(c):
        """-> {loop pc: end pc}, and for every pc the innermost loop whose body holds it (or None)"""
        ends, inner, stack = {}, [], []
        for pc, (op, _d, _a) in enumerate(c.ins):
            inner.append(next((s for k, s in reversed(stack) if k == "loop"), None))
            if op in ("if", "loop"):
                stack.append((op, pc))
            elif op == "end":
                k, s = stack.pop()
                if k == "loop":
                    ends[s] = pc
        return ends, inner

    def fix_literal_registers(self, c):
        """copy loops increment their address operands textually (store_bucket.rs:1016-1035, call_bucket.rs:975-994):
        `i64.5 = i64.add i64.5 i64.1`.  Shape:
            loop / if cnt / GET src / SET dest / cnt = i64.sub cnt i64.1 / src = i64.add src i64.1 / dest = i64.add dest i64.1 /
            continue / end / break / end  [ GET src / set_cmp_input_{run,cnt_check} c dest v ]   <- peeled last element
        The first increment means GET's address, the second SET's address (by position: both may be the same literal).  A
        literal address becomes a fresh register, set to the literal right before `loop` and used only by GET / SET / the
        peeled pair."""
        if not any(d is not None and self._is_literal(d) for (_o, d, _a) in c.ins):
            return
        GET = {"ff.load": 0, "get_signal": 0, "get_cmp_signal": 1}
        SET = {"ff.store": 0, "set_signal": 0, "set_cmp_input": 1, "set_cmp_input_cnt": 1, "set_cmp_input_run": 1,
               "set_cmp_input_cnt_check": 1}
        I = c.ins
        inits = []
        for L in range(len(I) - 10):
            if I[L][0] != "loop" or I[L + 1][0] != "if" or I[L + 2][0] not in GET or I[L + 3][0] not in SET:
                continue
            if [x[0] for x in I[L + 4:L + 11]] != ["i64.sub", "i64.add", "i64.add", "continue", "end", "break", "end"]:
                continue
            end = L + 10
            peeled = (end + 2 < len(I) and I[end + 1][0] == I[L + 2][0]
                      and I[end + 2][0] in ("set_cmp_input_run", "set_cmp_input_cnt_check"))
            for k in range(2):
                op, d, a = I[L + 5 + k]
                if not self._is_literal(d):
                    continue
                uop, ud, ua = I[L + 2 + k]
                ai = (SET if k else GET)[uop]
                if ua[ai] != d or a[0] != d:
                    raise ValueError("copy loop increments a literal that is not the address of its load/store")
                name = "%s@%s%d" % (d, "dst" if k else "src", L)
                ua = list(ua)
                ua[ai] = name
                I[L + 2 + k] = (uop, ud, ua)
                I[L + 5 + k] = (op, name, [name] + list(a[1:]))
                if peeled:
                    pop, pd, pa = I[end + 1 + k]
                    pi = 1 if k else GET[pop]
                    if pa[pi] == d:
                        pa = list(pa)
                        pa[pi] = name
                        I[end + 1 + k] = (pop, pd, pa)
                inits.append((L, name, d))
        for (_o, d, _a) in I:
            if d is not None and self._is_literal(d):
                raise ValueError("assignment to literal operand %r outside a loop of the copy-loop shape" % d)
        for (L, name, lit) in sorted(inits, key=lambda x: -x[0]):
            I.insert(L, ("mov", name, [lit]))

    @staticmethod
    def link_returns(c):
        """a multi-element return passes the VALUE of the first element (return_bucket.rs:131): use that load's address"""
        for pc, (op, d, a) in enumerate(c.ins):
            if op != "return" or a[1] == "1":
                continue
            for k in range(pc - 1, -1, -1):
                if c.ins[k][1] == a[0]:
                    if c.ins[k][0] == "ff.load":
                        c.ins[pc] = (op, d, [a[0], a[1], c.ins[k][2][0]])
                    break

    @staticmethod
    def fix_array_eq(c):
        """the emitter's array-equality shape (compute_bucket.rs:538-586) compares the first elements, increments those
        values as addresses and leaves the result in a register nobody reads.  It is replaced by one `array_eq`
        instruction that computes what the C++ twin does (compute_bucket.rs:375-407) into the register the consumer
        reads (the one allocated two before the loop's own result register)."""
        i = 0
        while i + 12 < len(c.ins):
            I = c.ins
            ok = (I[i][0] == "mov" and I[i][2][0].startswith("i64.") and I[i + 1][0] == "loop" and I[i + 2][0] == "if"
                  and I[i + 2][2][0] == I[i][1] and I[i + 3][0] == "ff.eq" and I[i + 4][0] == "if"
                  and I[i + 4][2][0] == I[i + 3][1] and I[i + 5][0] == "i64.sub" and I[i + 5][1] == I[i][1]
                  and I[i + 6][0] == "i64.add" and I[i + 6][1] == I[i + 3][2][0] and I[i + 6][2] == [I[i + 3][2][0], "i64.1"]
                  and I[i + 7][0] == "i64.add" and I[i + 7][1] == I[i + 3][2][1] and I[i + 7][2] == [I[i + 3][2][1], "i64.1"]
                  and [x[0] for x in I[i + 8:i + 13]] == ["continue", "end", "end", "break", "end"])
            if not ok:
                i += 1
                continue
            srcs = []
            for r in I[i + 3][2]:
                d = next((k for k in range(i - 1, -1, -1) if I[k][1] == r), None)
                if d is None or I[d][0] not in ("ff.load", "get_signal", "get_cmp_signal"):
                    raise ValueError("array-equality loop whose operands are not loads")
                srcs.append((I[d][0], list(I[d][2])))
            r2 = I[i + 3][1]
            assert r2.startswith("x_"), r2
            c.ins[i:i + 13] = [("array_eq", "x_%d" % (int(r2[2:]) - 2), [srcs[0], srcs[1], int(I[i][2][0][4:])])]
            i += 1

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


def _io_defs(words, pos, n):
    """n IODef records from a list of u32 (c_code_generator.rs:617-674; reader main.cpp:70-86) -> (defs, next position)"""
    defs = []
    for _ in range(n):
        offset, ln = words[pos], words[pos + 1]
        tail = list(words[pos + 2:pos + 2 + ln])
        pos += 2 + ln
        defs.append((offset, tail, words[pos], words[pos + 1]))
        pos += 2
    return defs, pos


def read_dat_io_map(prog, cpp_text, dat):
    """The io-map section of <circuit>.dat (present only with mixed component arrays).  The section sizes are not in the
    file but in the generated C++ (`uint get_size_of_*() {return N;}`, circuit.rs:481-497; main.cpp:22-92 uses them)."""
    def size_of(name):
        m = re.search(r"uint get_size_of_%s\(\) \{return (\d+);\}" % name, cpp_text)
        return int(m.group(1)) if m else 0
    n_io = size_of("io_map")
    if n_io == 0:
        return
    start = size_of("input_hashmap") * 24 + size_of("witness") * 8 + size_of("constants") * 40
    words = struct.unpack("<%dI" % ((len(dat) - start) // 4), dat[start:start + (len(dat) - start) // 4 * 4])
    ids = words[:n_io]
    pos = n_io
    for tid in ids:
        n = words[pos]
        prog.io_map[tid], pos = _io_defs(words, pos + 1, n)


def _template_id(header):
    """template headers are <name>_<instance id> (executed_template.rs; the id _create stores in componentMemory)"""
    return int(header.rsplit("_", 1)[1])


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
                slot, hdr, so, sj, _co, _cj, num = a[:7]
                positions = a[7] if len(a) > 7 and a[7] is not None else list(range(num))
                tcode = self.p.codes[hdr]
                for k, at in enumerate(positions):       # offsets advance per created component (create_component_bucket.rs:339-349)
                    sub = Component(tcode, comp.start + so + k * sj)
                    comp.subs[slot + at] = sub
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
            elif op == "array_eq":
                (opa, aa), (opb, ab), n_elems = a
                cnt["ops"] += 12
                res = 1
                for k in range(n_elems):
                    v = []
                    for (lop, la) in ((opa, aa), (opb, ab)):
                        if lop == "ff.load":
                            v.append(lvar.get(self.val(regs, la[0]) + k, 0))
                        elif lop == "get_signal":
                            v.append(sig[comp.start + self.val(regs, la[0]) + k])
                        else:
                            v.append(sig[comp.subs[self.val(regs, la[0])].start + self.val(regs, la[1]) + k])
                    res &= int(v[0] == v[1])
                regs[dst] = res
            elif op == "ff.call":
                self.call(comp, regs, lvar, a)
            elif op == "return":
                dlv, daddr, dsize = dest
                if a[1] == "1":
                    dlv[daddr] = self.val(regs, a[0]) % M.Q
                else:
                    src = self.val(regs, a[2] if len(a) > 2 else a[0])      # (link_returns: the address behind a loaded value)
                    for k in range(min(self.val(regs, a[1]), dsize)):
                        dlv[daddr + k] = lvar.get(src + k, 0)
                return
            elif op == "get_template_id":                      # load_bucket.rs:262-266: componentMemory[sub].templateId
                regs[dst] = _template_id(comp.subs[self.val(regs, a[0])].code.header)
            elif op in ("get_template_signal_position", "get_template_signal_size", "get_template_signal_dimension"):
                # location_rule.rs:99-146 / load_bucket.rs:262-318: templateInsId2IOSignalInfo[id].defs[code]
                defs = self.p.io_map.get(self.val(regs, a[0]))
                if defs is None:
                    raise NotImplementedError("mapped (mixed component array) access without an io-map entry: the fork's "
                                              ".cvm does not carry the io-map (SURVEY.md F3); load with the .cpp and .dat")
                offset, tail, size, _bus = defs[self.val(regs, a[1])]
                if op == "get_template_signal_position":
                    regs[dst] = offset
                elif op == "get_template_signal_size":
                    regs[dst] = size
                else:
                    regs[dst] = tail[self.val(regs, a[2]) - 1]
            elif op in ("get_template_signal_type", "get_bus_signal_position", "get_bus_signal_size",
                        "get_bus_signal_dimension", "get_bus_signal_type"):
                raise NotImplementedError("bus accesses through the io-map are not implemented")
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


def recover_creates(prog, cpp_text):
    """The fork's --cvm emitter prints nothing for component creation (create_component_bucket.rs:356-360); the generated
    <circuit>.cpp of the same compile has it: the `<Sub>_create(...)` blocks that `impl WriteC for CreateCmpBucket` prints in
    every `_run` body (create_component_bucket.rs:206-354).  They are placed at the top of the template (creation does not
    depend on signal values)."""
    unit, found, multi = None, [], None

    def flush():
        if unit is not None and found and not any(i[0] == "create_cmp" for i in unit.ins):
            unit.ins[0:0] = found
            unit.match, unit.loop_of = {}, {}
            Program.link(unit)
    for raw in cpp_text.split("\n"):
        line = raw.strip()
        if line.startswith("void "):
            flush()
            unit, found, multi = None, [], None
            m = re.match(r"void (\w+)_run\(uint ctx_index", line)
            if m and m.group(1) in prog.codes and not prog.codes[m.group(1)].is_function:
                unit = prog.codes[m.group(1)]
            continue
        if unit is None:
            continue
        m = re.match(r"uint aux_create = (\d+);", line)
        if m:
            multi = {"slot": int(m.group(1))}
            continue
        if multi is not None and line.startswith("uint aux_positions"):
            multi["pos"] = [int(x) for x in re.findall(r"\d+", line.split("{", 1)[1])]
            continue
        if multi is not None:
            for key, pat in (("cmp", r"int aux_cmp_num = (\d+)\+ctx_index\+1;"), ("sig", r"uint csoffset = mySignalStart\+(\d+);"),
                             ("n", r"for \(uint i(?:_aux)? = 0; i(?:_aux)? < (\d+); i(?:_aux)?\+\+\) \{"), ("sj", r"csoffset \+= (\d+) ;"),
                             ("sym", r"(\w+)_create\(csoffset,aux_cmp_num,"), ("cj", r"aux_cmp_num \+= (\d+);")):
                m = re.match(pat, line)
                if m:
                    multi[key] = m.group(1)
            if "cj" in multi:
                found.append(("create_cmp", None, [multi["slot"], multi["sym"], int(multi["sig"]), int(multi["sj"]), int(multi["cmp"]),
                                                   int(multi["cj"]), int(multi["n"]), multi.get("pos")]))
                multi = None
            continue
        m = re.match(r"(\w+)_create\(mySignalStart\+(\d+),(\d+)\+ctx_index\+1,", line)
        if m:
            found.append(("create_cmp", None, [None, m.group(1), int(m.group(2)), 0, int(m.group(3)), 0, 1]))
            continue
        m = re.match(r"mySubcomponents\[(\d+)\] = ", line)
        if m and found and found[-1][2][0] is None:
            found[-1][2][0] = int(m.group(1))
    flush()


def load(path_or_text, cpp_text=None, dat=None):
    text = path_or_text
    if "\n" not in path_or_text:
        with open(path_or_text) as f:
            text = f.read()
    prog = Program(text)
    if cpp_text is not None:
        recover_creates(prog, cpp_text)
        if dat is not None:
            read_dat_io_map(prog, cpp_text, dat)
    return prog


def compute_witness(prog, inputs):
    return Machine(prog).witness(inputs)
