"""Symbolic execution of the circuit description: instantiation + R1CS derivation.

Plays the role of the reference's constraint_generation crate (execute.rs: templates are
run with concrete parameters, signals stay symbolic, `<==`/`===` leave quadratic
constraints `A*B - C = 0`, circom_algebra/src/algebra.rs:1002-1054) and of the O1
simplification pass that removes signal-to-signal equalities
(constraint_list/src/constraint_simplification.rs).  Fixture tooling only.
"""
from __future__ import annotations

from .dsl import (P, Assert, Bin, CallE, Constrain, If, Loop, NewCmp, Num, Ref, Return, Set, SigSet, Template, Un,
                  prod)

HALF = P >> 1
MASK = (1 << 254) - 1


class CircuitError(Exception):
    pass


# ------------------------------------------------------------------ constant folding
def _signed(v):
    return v - P if v > HALF else v


def _mr(t):
    t &= MASK
    return t - P if t >= P else t


def const_binop(op, a, b):
    if op == "add": return (a + b) % P
    if op == "sub": return (a - b) % P
    if op == "mul": return (a * b) % P
    if op == "div": return (a * pow(b, -1, P)) % P if b else 0     # Fr_div(a, 0) = 0 in the reference
    if op == "idiv": return a // b
    if op == "mod": return a % b
    if op == "pow": return pow(a, b, P)
    if op == "shl":
        if b < 254: return _mr(a << b)
        s = P - b
        return 0 if s >= 254 else a >> s
    if op == "shr":
        if b < 254: return a >> b
        s = P - b
        return 0 if s >= 254 else _mr(a << s)
    if op == "band": return _mr(a & b)
    if op == "bor": return _mr(a | b)
    if op == "bxor": return _mr(a ^ b)
    if op == "eq": return int(a == b)
    if op == "neq": return int(a != b)
    if op == "lt": return int(_signed(a) < _signed(b))
    if op == "gt": return int(_signed(a) > _signed(b))
    if op == "leq": return int(_signed(a) <= _signed(b))
    if op == "geq": return int(_signed(a) >= _signed(b))
    if op == "land": return int(a != 0 and b != 0)
    if op == "lor": return int(a != 0 or b != 0)
    raise CircuitError("unknown op " + op)


def const_unop(op, a):
    if op == "neg": return (-a) % P
    if op == "lnot": return int(a == 0)
    if op == "bnot": return _mr(~a & ((1 << 256) - 1))
    raise CircuitError("unknown op " + op)


# ------------------------------------------------------------------ symbolic values
class Lin:
    __slots__ = ("t", "c")

    def __init__(self, t=None, c=0):
        self.t = t or {}
        self.c = c % P

    def scaled(self, k):
        k %= P
        if k == 0:
            return 0
        return Lin({s: (v * k) % P for s, v in self.t.items()}, self.c * k)

    def plus(self, o, sign=1):
        if isinstance(o, int):
            return Lin(dict(self.t), self.c + sign * o)
        t = dict(self.t)
        for s, v in o.t.items():
            nv = (t.get(s, 0) + sign * v) % P
            if nv:
                t[s] = nv
            else:
                t.pop(s, None)
        c = (self.c + sign * o.c) % P
        if not t:
            return c
        return Lin(t, c)


class Quad:
    __slots__ = ("a", "b", "c")

    def __init__(self, a, b, c):
        self.a, self.b, self.c = a, b, c     # a*b + c ; a, b Lin ; c Lin or int


class _Unknown:
    def __repr__(self):
        return "UNK"


UNK = _Unknown()


def _aslin(x):
    return Lin({}, x) if isinstance(x, int) else x


def v_add(a, b, sign=1):
    if a is UNK or b is UNK:
        return UNK
    if isinstance(a, int) and isinstance(b, int):
        return (a + sign * b) % P
    if isinstance(a, Quad) and isinstance(b, Quad):
        return UNK
    if isinstance(a, Quad):
        c = v_add(a.c, b, sign)
        return Quad(a.a, a.b, c)
    if isinstance(b, Quad):
        nb = Quad(b.a.scaled(sign) if sign != 1 else b.a, b.b, v_mul(b.c, sign % P))
        return Quad(nb.a, nb.b, v_add(a, nb.c))
    return _aslin(a).plus(b, sign)


def v_mul(a, b):
    if a is UNK or b is UNK:
        return UNK
    if isinstance(a, int) and isinstance(b, int):
        return (a * b) % P
    if isinstance(b, int):
        a, b = b, a
    if isinstance(a, int):
        if a % P == 0:
            return 0
        if isinstance(b, Lin):
            return b.scaled(a)
        return Quad(b.a.scaled(a), b.b, v_mul(a, b.c))
    if isinstance(a, Lin) and isinstance(b, Lin):
        return Quad(a, b, 0)
    return UNK


class Instance:
    """One template instance (template name + concrete arguments)."""

    def __init__(self, name, key, tmpl):
        self.name, self.key, self.tmpl = name, key, tmpl
        self.id = None
        self.header = None
        self.n_out = self.n_in = self.n_mid = 0
        self.n_own = 0
        self.n_signals = 0           # own + all sub-components
        self.n_components = 1        # itself + all sub-components
        self.subs = []               # (cmpsym, flat, Instance, signal_offset, component_offset)
        self.constraints = []        # own constraints, symbolic keys
        self.sigsym = {}


class Program:
    """All template instances reachable from main, plus functions."""

    def __init__(self):
        self.instances = {}      # key -> Instance
        self.order = []          # post-order (children before parents)
        self.functions = {}      # name -> Function AST (registered by circuits)
        self.main = None
        self.public_inputs = []

    def instantiate(self, fn, args):
        key = (fn.__name__, _freeze(args))
        inst = self.instances.get(key)
        if inst is not None:
            return inst
        tmpl = Template(fn.__name__)
        tmpl.program = self
        fn(tmpl, *args)
        inst = Instance(fn.__name__, key, tmpl)
        inst.sigsym = {s.name: s for s in tmpl.signals}
        self.instances[key] = inst
        _Executor(self, inst).run()
        inst.id = len(self.order)
        inst.header = "%s_%d" % (inst.name, inst.id)
        self.order.append(inst)
        return inst


def _freeze(x):
    if isinstance(x, (list, tuple)):
        return tuple(_freeze(v) for v in x)
    return x


class _Executor:
    def __init__(self, prog, inst):
        self.prog, self.inst, self.t = prog, inst, inst.tmpl
        self.env = {}            # VarSym -> list of values
        self.shadow = 0          # >0 while inside a branch whose condition is unknown

    # ---- layout of own signals: outputs, inputs, intermediates (executed_template.rs:448-550)
    def layout_own(self):
        off = 0
        inst = self.inst
        for xt in ("out", "in", "mid"):
            for s in self.t.signals:
                if s.xtype == xt:
                    s.offset = off
                    off += s.size
            if xt == "out": inst.n_out = off
            elif xt == "in": inst.n_in = off - inst.n_out
        inst.n_mid = off - inst.n_out - inst.n_in
        inst.n_own = off

    def run(self):
        self.layout_own()
        for v in self.t.vars:
            self.env[v] = list(v.init) if v.is_param else [0] * v.size
        self.block(self.t.body)
        inst = self.inst
        # place sub-components after own signals, declaration order then array index
        sig_off, cmp_off = inst.n_own, 0
        slot = 0
        for c in self.t.components:
            c.offset = slot
            slot += c.size
            for flat in range(c.size):
                sub = c.instances.get(flat)
                if sub is None:
                    raise CircuitError("component %s[%d] of %s never instantiated" % (c.name, flat, inst.name))
                inst.subs.append((c, flat, sub, sig_off, cmp_off))
                sig_off += sub.n_signals
                cmp_off += sub.n_components
        inst.n_signals = sig_off
        inst.n_components = 1 + cmp_off
        inst.n_slots = slot

    # ---- expressions
    def index(self, dims, idx, dynamic_ok=False):
        """-> (flat offset, element count); with dynamic_ok an index that depends on signals gives (None, count): legal in
        `<--` code and in var computations, whose values the constraint system does not see"""
        assert len(idx) <= len(dims)
        flat = 0
        for k, d in enumerate(dims):
            flat *= d
            if k < len(idx):
                i = self.ev(idx[k])
                if not isinstance(i, int):
                    if dynamic_ok:
                        return None, prod(dims[len(idx):])
                    raise CircuitError("array index is not a compile-time constant")
                if i >= d:
                    raise CircuitError("index %d out of range %d" % (i, d))
                flat += i
        return flat, prod(dims[len(idx):])

    def sigkey(self, ref):
        """-> (list of keys, size).  key = local offset | (cmpsym, flat, child offset)."""
        sym = ref.sym
        if sym.kind == "sig":
            base, n = self.index(sym.dims, ref.idx)
            return [sym.offset + base + k for k in range(n)]
        cflat, cn = self.index(sym.dims, ref.idx)
        assert cn == 1, "component array slice used as a signal"
        sub = sym.instances.get(cflat)
        if sub is None:
            raise CircuitError("component %s[%d] used before creation" % (sym.name, cflat))
        ss = sub.sigsym[ref.sig]
        base, n = self.index(ss.dims, ref.sigidx)
        return [(sym, cflat, ss.offset + base + k) for k in range(n)]

    def ev(self, e):
        if isinstance(e, Num):
            return e.v
        if isinstance(e, Ref):
            if e.sym.kind == "var":
                flat, n = self.index(e.sym.dims, e.idx, dynamic_ok=True)
                assert n == 1, "array-valued var used as a scalar"
                return UNK if flat is None else self.env[e.sym][flat]
            if e.sym.kind == "sig" and self.index(e.sym.dims, e.idx, dynamic_ok=True)[0] is None:
                return UNK          # signal array read at a data-dependent index (only meaningful on the right of `<--`)
            keys = self.sigkey(e)
            assert len(keys) == 1, "array-valued signal used as a scalar"
            return Lin({keys[0]: 1}, 0)
        if isinstance(e, Bin):
            a, b = self.ev(e.a), self.ev(e.b)
            if isinstance(a, int) and isinstance(b, int):
                return const_binop(e.op, a, b)
            if e.op == "add": return v_add(a, b)
            if e.op == "sub": return v_add(a, b, -1)
            if e.op == "mul": return v_mul(a, b)
            if e.op == "div" and isinstance(b, int): return v_mul(a, pow(b, -1, P) if b else 0)
            return UNK
        if isinstance(e, Un):
            a = self.ev(e.a)
            if isinstance(a, int):
                return const_unop(e.op, a)
            if e.op == "neg": return v_mul(a, P - 1)
            return UNK
        if isinstance(e, CallE):
            return UNK
        raise CircuitError("bad expression %r" % (e,))

    # ---- statements
    def constraint(self, v, line):
        """record v == 0"""
        if self.shadow:
            raise CircuitError("constraint under a condition that depends on signals (line %d)" % line)
        if v is UNK:
            raise CircuitError("non-quadratic constraint in %s (line %d)" % (self.inst.name, line))
        if isinstance(v, int):
            if v % P:
                raise CircuitError("constant constraint is false (line %d)" % line)
            return
        if isinstance(v, Lin):
            a, b, c = Lin(), Lin(), v_mul(v, P - 1)
        else:
            a, b, c = v.a, v.b, v_mul(v.c, P - 1)
        self.inst.constraints.append((a, b, _aslin(c)))

    def block(self, stmts):
        for st in stmts:
            self.stmt(st)

    def stmt(self, st):
        if isinstance(st, Set):
            flat, n = self.index(st.dst.sym.dims, st.dst.idx, dynamic_ok=True)
            if flat is None:            # var[data-dependent index] = ...: any element may have changed
                self.ev(st.src)
                self.env[st.dst.sym] = [UNK] * len(self.env[st.dst.sym])
            elif n == 1:
                val = self.ev(st.src)
                self.env[st.dst.sym][flat] = UNK if self.shadow else val
            else:   # array-valued assignment (function result or array copy): contents unknown to constraints
                for k in range(n):
                    self.env[st.dst.sym][flat + k] = UNK
        elif isinstance(st, SigSet):
            keys = self.sigkey(st.dst)
            if st.constrain:
                if len(keys) == 1:
                    self.constraint(v_add(self.ev(st.src), Lin({keys[0]: 1}), -1), st.line)
                else:
                    skeys = self.sigkey(st.src)
                    for kd, ks in zip(keys, skeys):
                        self.constraint(Lin({ks: 1, kd: P - 1}), st.line)
        elif isinstance(st, Constrain):
            self.constraint(v_add(self.ev(st.l), self.ev(st.r), -1), st.line)
        elif isinstance(st, Loop):
            guard = 0
            while True:
                c = self.ev(st.cond)
                if not isinstance(c, int):
                    # trip count unknown at compile time: legal in circom as long as the body only touches vars, whose
                    # values are then unknown to the constraint system (execute.rs treats the block as "unknown")
                    self.shadow += 1
                    self.block(st.body)
                    self.shadow -= 1
                    break
                if c == 0:
                    break
                self.block(st.body)
                guard += 1
                if guard > 10_000_000:
                    raise CircuitError("runaway loop")
        elif isinstance(st, If):
            c = self.ev(st.cond)
            if isinstance(c, int):
                self.block(st.then if c else st.other)
            else:
                self.shadow += 1
                self.block(st.then)
                self.block(st.other)
                self.shadow -= 1
        elif isinstance(st, Assert):
            c = self.ev(st.e)
            if isinstance(c, int) and c == 0:
                raise CircuitError("assert failed at generation time (line %d)" % st.line)
        elif isinstance(st, NewCmp):
            flat, n = self.index(st.dst.sym.dims, st.dst.idx)
            assert n == 1
            args = []
            for a in st.args:
                if isinstance(a, (int, list, tuple)):
                    args.append(a)
                else:
                    v = self.ev(a)
                    if not isinstance(v, int):
                        raise CircuitError("template argument is not constant")
                    args.append(v)
            sub = self.prog.instantiate(st.template, args)
            st.dst.sym.instances[flat] = sub
        elif isinstance(st, Return):
            raise CircuitError("return inside a template")
        else:
            raise CircuitError("bad statement %r" % (st,))


def build_program(main_fn, args=(), public=(), functions=()):
    prog = Program()
    for f in functions:
        prog.functions[f.name] = f
    prog.main = prog.instantiate(main_fn, list(args))
    prog.public_inputs = list(public)
    _order_main_inputs(prog)
    return prog


def _order_main_inputs(prog):
    """Main's inputs: public ones first (executed_template.rs:448-550)."""
    main = prog.main
    ins = [s for s in main.tmpl.signals if s.xtype == "in"]
    pub = [s for s in ins if s.name in prog.public_inputs]
    prv = [s for s in ins if s.name not in prog.public_inputs]
    if not pub or ins == pub + prv:
        prog.n_pub_in = sum(s.size for s in pub)
        return
    raise CircuitError("declare public inputs before private ones in main")


# ------------------------------------------------------------------ flatten to global numbering
def flatten_constraints(prog):
    """-> list of (A, B, C) dicts keyed by GLOBAL signal number (0 = constant one).

    Global numbering: main starts at 1 (circuit.rs:539), sub-components follow their parent's
    own signals (create_component_bucket.rs:221-235).
    """
    out = []

    def conv(lin, start, inst, submap):
        d = {}
        for k, v in lin.t.items():
            if isinstance(k, tuple):
                g = start + submap[(k[0], k[1])] + k[2]
            else:
                g = start + k
            d[g] = v
        if lin.c:
            d[0] = lin.c
        return d

    stack = [(prog.main, 1)]
    while stack:
        inst, start = stack.pop()
        submap = {(c, flat): off for (c, flat, _s, off, _co) in inst.subs}
        for (a, b, c) in inst.constraints:
            out.append((conv(a, start, inst, submap), conv(b, start, inst, submap), conv(c, start, inst, submap)))
        for (c, flat, sub, off, _co) in reversed(inst.subs):
            stack.append((sub, start + off))
    return out


def simplify_o1(constraints, n_signals, protect):
    """Remove `x - y = 0` constraints by merging y into x (smaller id survives; ids < protect
    are never eliminated).  Returns (constraints, witness2signal list)."""
    parent = list(range(n_signals))

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x

    kept = []
    for (a, b, c) in constraints:
        if not a and not b and len(c) == 2 and 0 not in c:
            (x, vx), (y, vy) = c.items()
            if (vx + vy) % P == 0:
                rx, ry = find(x), find(y)
                if rx != ry:
                    lo, hi = min(rx, ry), max(rx, ry)
                    if hi >= protect:
                        parent[hi] = lo
                        continue
                else:
                    continue
        kept.append((a, b, c))

    def remap(d):
        r = {}
        for k, v in d.items():
            k2 = find(k)
            nv = (r.get(k2, 0) + v) % P
            if nv:
                r[k2] = nv
            else:
                r.pop(k2, None)
        return r

    res = []
    for (a, b, c) in kept:
        a2, b2, c2 = remap(a), remap(b), remap(c)
        if not a2 and not b2 and not c2:
            continue
        res.append((a2, b2, c2))
    witness = [s for s in range(n_signals) if find(s) == s]
    return res, witness
