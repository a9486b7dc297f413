"""Embedded circuit-description language used to generate fixture circuits.

Why this exists: the reference compiler is Rust and cannot be built in this image
(no cargo), and circomlib is not on disk, so no genuine .cvm/.r1cs can be produced
here.  This package is a small stand-in for the *front half* of the
reference compiler: it describes circuits with the same statement kinds circom has
(`<--`, `<==`, `===`, vars, loops, ifs, components, functions) and then

  * execute.py   derives the R1CS, like constraint_generation/src/execute.rs does,
  * translate.py lowers the same AST to the reference's bucket IR
                 (compiler/src/intermediate_representation/translate.rs),
  * emit_cvm.py / emit_cpp.py print that IR in the reference's CVM / C++ shapes.

It is fixture tooling, not part of the product path.
"""
from __future__ import annotations

P = 21888242871839275222246405745257275088548364400416034343698204186575808495617


def prod(xs):
    r = 1
    for x in xs:
        r *= x
    return r


def wrap(x):
    if isinstance(x, Expr):
        return x
    if isinstance(x, bool):
        return Num(int(x))
    if isinstance(x, int):
        return Num(x % P)
    raise TypeError("cannot use %r in a circuit expression" % (x,))


class Expr:
    def _b(self, op, o):
        return Bin(op, self, wrap(o))

    def _rb(self, op, o):
        return Bin(op, wrap(o), self)

    def __add__(self, o): return self._b("add", o)
    def __radd__(self, o): return self._rb("add", o)
    def __sub__(self, o): return self._b("sub", o)
    def __rsub__(self, o): return self._rb("sub", o)
    def __mul__(self, o): return self._b("mul", o)
    def __rmul__(self, o): return self._rb("mul", o)
    def __truediv__(self, o): return self._b("div", o)
    def __rtruediv__(self, o): return self._rb("div", o)
    def __floordiv__(self, o): return self._b("idiv", o)
    def __rfloordiv__(self, o): return self._rb("idiv", o)
    def __mod__(self, o): return self._b("mod", o)
    def __pow__(self, o): return self._b("pow", o)
    def __rpow__(self, o): return self._rb("pow", o)
    def __lshift__(self, o): return self._b("shl", o)
    def __rshift__(self, o): return self._b("shr", o)
    def __rlshift__(self, o): return self._rb("shl", o)
    def __rrshift__(self, o): return self._rb("shr", o)
    def __rmod__(self, o): return self._rb("mod", o)
    def __ror__(self, o): return self._rb("bor", o)
    def __rxor__(self, o): return self._rb("bxor", o)
    def __and__(self, o): return self._b("band", o)
    def __rand__(self, o): return self._rb("band", o)
    def __or__(self, o): return self._b("bor", o)
    def __xor__(self, o): return self._b("bxor", o)
    def __lt__(self, o): return self._b("lt", o)
    def __le__(self, o): return self._b("leq", o)
    def __gt__(self, o): return self._b("gt", o)
    def __ge__(self, o): return self._b("geq", o)
    def __neg__(self): return Un("neg", self)
    def __invert__(self): return Un("bnot", self)
    def eq(self, o): return self._b("eq", o)
    def ne(self, o): return self._b("neq", o)
    def land(self, o): return self._b("land", o)
    def lor(self, o): return self._b("lor", o)
    def lnot(self): return Un("lnot", self)
    __hash__ = object.__hash__


class Num(Expr):
    def __init__(self, v):
        self.v = v % P


class Bin(Expr):
    def __init__(self, op, a, b):
        self.op, self.a, self.b = op, a, b


class Un(Expr):
    def __init__(self, op, a):
        self.op, self.a = op, a


class CallE(Expr):
    """Function call in expression position (only legal as the whole right-hand side)."""

    def __init__(self, fn, args):
        self.fn, self.args = fn, [a if isinstance(a, Expr) else wrap(a) for a in args]


class Sym:
    def __init__(self, kind, name, dims=()):
        self.kind = kind              # 'var' | 'sig' | 'cmp'
        self.name = name
        self.dims = tuple(int(d) for d in dims)
        self.size = prod(self.dims)
        # var: is_param, init ; sig: xtype in/out/mid, offset ; cmp: slot, instances
        self.is_param = False
        self.init = None
        self.xtype = None
        self.offset = None            # lvar address / local signal offset / component slot
        self.instances = {}           # cmp: flat index -> Instance


class Ref(Expr):
    """Access path: sym[idx...] or cmp[idx...].sig[sigidx...]."""

    def __init__(self, sym, idx=(), sig=None, sigidx=()):
        self.sym, self.idx, self.sig, self.sigidx = sym, tuple(idx), sig, tuple(sigidx)

    def __getitem__(self, i):
        if self.sig is None:
            return Ref(self.sym, self.idx + (wrap(i),))
        return Ref(self.sym, self.idx, self.sig, self.sigidx + (wrap(i),))

    def pin(self, name):
        assert self.sym.kind == "cmp" and self.sig is None
        return Ref(self.sym, self.idx, name)


# ---------------------------------------------------------------- statements
class Stmt:
    line = 0


class Set(Stmt):            # var = expr   (expr may be a CallE)
    def __init__(self, dst, src): self.dst, self.src = dst, src


class SigSet(Stmt):         # sig <-- expr  /  sig <== expr  (constrain=True)
    def __init__(self, dst, src, constrain): self.dst, self.src, self.constrain = dst, src, constrain


class Constrain(Stmt):      # l === r
    def __init__(self, l, r): self.l, self.r = l, r


class Loop(Stmt):
    def __init__(self, cond): self.cond, self.body = cond, []


class If(Stmt):
    def __init__(self, cond): self.cond, self.then, self.other = cond, [], []


class Assert(Stmt):
    def __init__(self, e): self.e = e


class Return(Stmt):
    def __init__(self, e): self.e = e


class NewCmp(Stmt):         # component c[idx] = Template(args)
    def __init__(self, dst, template, args): self.dst, self.template, self.args = dst, template, args


class _Block:
    def __init__(self, ctx, target, after=None):
        self.ctx, self.target, self.after = ctx, target, after

    def __enter__(self):
        self.ctx._stack.append(self.target)
        return self

    def __exit__(self, *exc):
        if exc[0] is None and self.after is not None:
            self.after()
        self.ctx._stack.pop()
        return False


class Body:
    """Common builder for template and function bodies."""

    def __init__(self, name):
        self.name = name
        self.body = []
        self._stack = [self.body]
        self._line = 1
        self.vars = []          # VarSym in declaration order (params first)
        self.symnames = set()

    def _emit(self, st):
        st.line = self._line
        self._line += 1
        self._stack[-1].append(st)
        return st

    def _declare(self, sym):
        assert sym.name not in self.symnames, "duplicate symbol %s" % sym.name
        self.symnames.add(sym.name)
        return sym

    # ---- vars
    def param(self, name, value):
        """Template/function parameter with a known value (int or nested list)."""
        flat, dims = _flatten(value)
        s = self._declare(Sym("var", name, dims))
        s.is_param = True
        s.init = [v % P for v in flat]
        self.vars.append(s)
        return Ref(s)

    def var(self, name, dims=(), init=None):
        s = self._declare(Sym("var", name, dims))
        self.vars.append(s)
        r = Ref(s)
        if init is not None:
            if isinstance(init, (list, tuple)):
                flat, d = _flatten(init)
                assert tuple(d) == s.dims
                for k, v in enumerate(flat):
                    self._emit(Set(_index_flat(r, s.dims, k), wrap(v)))
            else:
                self._emit(Set(r, wrap(init)))
        return r

    def set(self, dst, src):
        assert dst.sym.kind == "var"
        self._emit(Set(dst, src if isinstance(src, Expr) else wrap(src)))

    # ---- control
    def loop(self, cond):
        st = self._emit(Loop(wrap(cond)))
        return _Block(self, st.body)

    def for_(self, i, start, cond, step=None):
        """for (i = start; cond; i = step or i+1)"""
        self.set(i, start)
        st = self._emit(Loop(wrap(cond)))
        nxt = (i + 1) if step is None else step
        return _Block(self, st.body, after=lambda: self.set(i, nxt))

    def if_(self, cond):
        st = self._emit(If(wrap(cond)))
        self._last_if = st
        return _Block(self, st.then)

    def else_(self):
        st = self._stack[-1][-1]        # the `if` that was just closed in THIS block (not a nested one)
        assert isinstance(st, If), "else_ must directly follow an if_ block"
        return _Block(self, st.other)

    def assert_(self, e):
        self._emit(Assert(wrap(e)))

    def call(self, fn, *args):
        return CallE(fn, list(args))


class Template(Body):
    def __init__(self, name):
        super().__init__(name)
        self.signals = []       # SigSym
        self.components = []    # CmpSym

    def _sig(self, name, dims, xtype):
        s = self._declare(Sym("sig", name, dims))
        s.xtype = xtype
        self.signals.append(s)
        return Ref(s)

    def input(self, name, dims=()): return self._sig(name, dims, "in")
    def output(self, name, dims=()): return self._sig(name, dims, "out")
    def signal(self, name, dims=()): return self._sig(name, dims, "mid")

    def component(self, name, dims=()):
        s = self._declare(Sym("cmp", name, dims))
        self.components.append(s)
        return Ref(s)

    def new(self, dst, template, *args):
        assert dst.sym.kind == "cmp"
        self._emit(NewCmp(dst, template, list(args)))

    def assign(self, dst, src):        # <--
        self._emit(SigSet(dst, src if isinstance(src, Expr) else wrap(src), False))

    def bind(self, dst, src):          # <==
        self._emit(SigSet(dst, src if isinstance(src, Expr) else wrap(src), True))

    def constrain(self, l, r):         # ===
        self._emit(Constrain(wrap(l), wrap(r)))


class Function(Body):
    def __init__(self, name):
        super().__init__(name)
        self.params = []        # (VarSym) in order; values are NOT known at build time
        self.returns = ()       # dims of the returned value

    def arg(self, name, dims=()):
        s = self._declare(Sym("var", name, dims))
        s.is_param = True
        self.vars.append(s)
        self.params.append(s)
        return Ref(s)

    def ret(self, e):
        self._emit(Return(e if isinstance(e, Expr) else wrap(e)))


def _flatten(value):
    if isinstance(value, (list, tuple)):
        if len(value) == 0:
            return [], (0,)
        subs = [_flatten(v) for v in value]
        d0 = subs[0][1]
        flat = []
        for f, d in subs:
            assert d == d0, "ragged array"
            flat.extend(f)
        return flat, (len(value),) + tuple(d0)
    return [int(value)], ()


def _index_flat(ref, dims, k):
    idx = []
    for d in reversed(dims):
        idx.append(k % d)
        k //= d
    for i in reversed(idx):
        ref = ref[i]
    return ref
