"""Small circuits: the doc examples of the reference plus circomlib-style gadgets written from
their public definitions (not in the reference tree; see SURVEY.md 'fixture gap')."""
from __future__ import annotations

from ..dsl import Function


def Multiplier2(T):
    """mkdocs/docs/getting-started/writing-circuits.md: c <== a*b."""
    a = T.input("a")
    b = T.input("b")
    c = T.output("c")
    T.bind(c, a * b)


def MultiplierN(T, n):
    """Chain of Multiplier2 components (component array, uniform): out = prod(in[i])."""
    inp = T.input("in", (n,))
    out = T.output("out")
    comp = T.component("comp", (n - 1,))
    i = T.var("i")
    with T.for_(i, 0, i < n - 1):
        T.new(comp[i], Multiplier2)
    T.bind(comp[0].pin("a"), inp[0])
    T.bind(comp[0].pin("b"), inp[1])
    with T.for_(i, 0, i < n - 2):
        T.bind(comp[i + 1].pin("a"), comp[i].pin("c"))
        T.bind(comp[i + 1].pin("b"), inp[i + 2])
    T.bind(out, comp[n - 2].pin("c"))


def Num2Bits(T, n):
    """circomlib bitify.circom Num2Bits."""
    n_ = T.param("n", n)
    inp = T.input("in")
    out = T.output("out", (n,))
    lc1 = T.var("lc1", init=0)
    e2 = T.var("e2", init=1)
    i = T.var("i")
    with T.for_(i, 0, i < n_):
        T.assign(out[i], (inp >> i) & 1)
        T.constrain(out[i] * (out[i] - 1), 0)
        T.set(lc1, lc1 + out[i] * e2)
        T.set(e2, e2 + e2)
    T.constrain(lc1, inp)


def Bits2Num(T, n):
    n_ = T.param("n", n)
    inp = T.input("in", (n,))
    out = T.output("out")
    lc1 = T.var("lc1", init=0)
    e2 = T.var("e2", init=1)
    i = T.var("i")
    with T.for_(i, 0, i < n_):
        T.set(lc1, lc1 + inp[i] * e2)
        T.set(e2, e2 + e2)
    T.bind(out, lc1)


def IsZero(T):
    """circomlib comparators.circom IsZero: inv <-- in!=0 ? 1/in : 0."""
    inp = T.input("in")
    out = T.output("out")
    inv = T.signal("inv")
    tmp = T.var("tmp")
    with T.if_(inp.ne(0)):
        T.set(tmp, 1 / inp)
    with T.else_():
        T.set(tmp, 0)
    T.assign(inv, tmp)
    T.bind(out, -inp * inv + 1)
    T.constrain(inp * out, 0)


def IsEqual(T):
    inp = T.input("in", (2,))
    out = T.output("out")
    isz = T.component("isz")
    T.new(isz, IsZero)
    T.bind(isz.pin("in"), inp[1] - inp[0])
    T.bind(out, isz.pin("out"))


def LessThan(T, n):
    """circomlib comparators.circom LessThan(n)."""
    inp = T.input("in", (2,))
    out = T.output("out")
    n2b = T.component("n2b")
    T.new(n2b, Num2Bits, n + 1)
    T.bind(n2b.pin("in"), inp[0] + (1 << n) - inp[1])
    T.bind(out, 1 - n2b.pin("out")[n])


def make_ops_function():
    """A circom function touching every field operator (the op set of compute_bucket.rs:9-36);
    used from a `<--` hint so that no constraint is involved."""
    F = Function("allops")
    a = F.arg("a")
    b = F.arg("b")
    sh = F.arg("sh")
    F.returns = (24,)
    r = F.var("r", (24,))
    exprs = [a + b, a - b, a * b, a / (b + 1), a // (b + 1), a % (b + 1), a ** sh, a << sh, a >> sh,
             a & b, a | b, a ^ b, ~a, -a, a < b, a <= b, a > b, a >= b, a.eq(b), a.ne(b),
             a.land(b), a.lor(b), a.lnot(), (a >> sh) & 1]
    for k, e in enumerate(exprs):
        F.set(r[k], e)
    F.ret(r)
    return F


ALLOPS = make_ops_function()


def OpsZoo(T):
    """Every operator through a function call with array return, array copy to signals, a
    data-dependent branch, and a static loop over a var array."""
    a = T.input("a")
    b = T.input("b")
    sh = T.input("sh")
    out = T.output("out", (24,))
    acc = T.output("acc")
    res = T.var("res", (24,))
    T.set(res, T.call(ALLOPS, a, b, sh))
    T.assign(out, res)
    s = T.var("s", init=0)
    i = T.var("i")
    with T.for_(i, 0, i < 24):
        with T.if_(res[i] > 5):
            T.set(s, s + res[i])
        with T.else_():
            T.set(s, s * 3 + 1)
    T.assign(acc, s)


def Sum3Cmp(T):
    """Parent wiring three sub-components with array ports (multi-element stores) --
    exercises Fr_copyn / the peeled copy loop of store_bucket.rs:899-1041."""
    inp = T.input("in", (4,))
    out = T.output("out", (4,))
    b2n = T.component("b2n")
    n2b = T.component("n2b")
    T.new(b2n, Bits2Num, 4)
    T.new(n2b, Num2Bits, 4)
    T.bind(b2n.pin("in"), inp)
    T.bind(n2b.pin("in"), b2n.pin("out"))
    T.bind(out, n2b.pin("out"))


def NBits(T):
    """Bit length by a `while` whose trip count depends on the input (circomlib-style `nbits`, here on a signal in
    `<--` code): exercises the bounded predicated unrolling of the trace compiler."""
    inp = T.input("in")
    out = T.output("out")
    n = T.var("n", init=0)
    r = T.var("r", init=0)
    T.set(n, inp)
    with T.loop(n.ne(0)):
        T.set(r, r + 1)
        T.set(n, n >> 1)
    T.assign(out, r)


def CountDown(T):
    """while (n != 0) n = n - 1 with a running sum: needs `in` iterations (only small inputs stay within the bound)."""
    inp = T.input("in")
    out = T.output("out")
    n = T.var("n", init=0)
    acc = T.var("acc", init=0)
    T.set(n, inp)
    with T.loop(n.ne(0)):
        T.set(acc, acc + n * n)
        T.set(n, n - 1)
    T.assign(out, acc)


def make_early_return_functions():
    """circom functions that `return` from inside data-dependent branches and from inside a loop."""
    fabs = Function("fabs")
    a = fabs.arg("a")
    fabs.returns = ()
    with fabs.if_(a < 0):
        fabs.ret(-a)
    fabs.ret(a)

    fclamp = Function("fclamp")
    x = fclamp.arg("x")
    lo = fclamp.arg("lo")
    hi = fclamp.arg("hi")
    fclamp.returns = ()
    with fclamp.if_(x < lo):
        fclamp.ret(lo)
    with fclamp.else_():
        with fclamp.if_(x > hi):
            fclamp.ret(hi)
    fclamp.ret(x * 1 + 0)

    first = Function("first_ge")
    v = first.arg("v")
    b = first.arg("b")
    first.returns = ()
    i = first.var("i")
    acc = first.var("acc", init=0)
    with first.for_(i, 0, i < 8):
        with first.if_((v >> i) <= b):
            first.ret(i * 1000 + acc)
        first.set(acc, acc + (v >> i))
    first.ret(8000 + acc)
    return fabs, fclamp, first


FABS, FCLAMP, FIRST_GE = make_early_return_functions()


def EarlyReturns(T):
    """`<--` hints through functions with early returns under data-dependent conditions (if-converted by the tracer)."""
    a = T.input("a")
    b = T.input("b")
    out = T.output("out", (4,))
    t = T.var("t")
    T.set(t, T.call(FABS, a))
    T.assign(out[0], t)
    T.set(t, T.call(FCLAMP, a, b, b + 100))
    T.assign(out[1], t)
    T.set(t, T.call(FIRST_GE, a, b))
    T.assign(out[2], t)
    T.set(t, T.call(FABS, a - b))
    T.assign(out[3], t + 1)


WIDE_SUM_TERMS = (3, 6, 11, 16, 17, 33)


def wide_sum_const(k, i):
    """Deterministic full-size coefficients, several of them with long runs of one bits."""
    q = 21888242871839275222246405745257275088548364400416034343698204186575808495617
    special = [q - 1, q - 2, (1 << 253) - 1, q - (1 << 224), (q - 1) // 2, (((q >> 224) - 1) << 224) | ((1 << 224) - 1)]
    if (k + i) % 3 == 0:
        return special[(k * 7 + i) % len(special)]
    return pow(7 + k, 1000 + 37 * i, q)


def WideSums(T):
    """Linear combinations of 3..33 inputs with full-size constant coefficients (the shape of an MDS row, widened):
    exercises fused dot products of every length, their multi-subtraction final reductions and splitting past 16 terms."""
    inp = T.input("in", (40,))
    out = T.output("out", (len(WIDE_SUM_TERMS),))
    sq = T.output("sq")
    for k, n in enumerate(WIDE_SUM_TERMS):
        e = None
        for i in range(n):
            term = inp[(7 * k + i) % 40] * wide_sum_const(k, i)
            e = term if e is None else e + term
        T.bind(out[k], e)
    T.bind(sq, out[3] * out[5])


def WeightedRows(T, n, k):
    """Sub-component of MixedArr.  `partial` comes first and has n elements, so the offsets of every later signal - and
    the second dimension of `m` - depend on n: instances with different n have different io-map entries."""
    n_ = T.param("n", n)
    k_ = T.param("k", k)
    partial = T.output("partial", (n,))
    total = T.output("total")
    m = T.input("m", (2, n))
    w = T.input("w")
    acc = T.var("acc", init=0)
    j = T.var("j")
    with T.for_(j, 0, j < n_):
        T.set(acc, acc + m[0][j] * (k_ + j) + m[1][j])
        T.bind(partial[j], acc)
    T.bind(total, acc * w)


def MixedArr(T):
    """A MIXED component array (its positions hold different template instances: n = 2, 3, 4), wired through loop
    variables: every access to c[i].<signal> is a "mapped" location (location_rule.rs:86-171) that goes through the
    io-map - scalar and 1-D outputs, a 2-D input (get_template_signal_dimension), an input behind the arrays."""
    a = T.input("a", (9,))
    b = T.input("b", (9,))
    w = T.input("w", (3,))
    out = T.output("out", (3,))
    last = T.output("last", (3,))
    c = T.component("c", (3,))
    i = T.var("i")
    j = T.var("j")
    pos = T.var("pos", init=0)
    with T.for_(i, 0, i < 3):
        T.new(c[i], WeightedRows, i + 2, 7)
    with T.for_(i, 0, i < 3):
        with T.for_(j, 0, j < i + 2):
            T.bind(c[i].pin("m")[0][j], a[pos])
            T.bind(c[i].pin("m")[1][j], b[pos] + 1)
            T.set(pos, pos + 1)
        T.bind(c[i].pin("w"), w[i])
    with T.for_(i, 0, i < 3):
        T.bind(out[i], c[i].pin("total"))
        T.bind(last[i], c[i].pin("partial")[i + 1] + c[2].pin("partial")[0])


def DynIndex(T):
    """Data-dependent array indices in `<--` code (Fr_toInt of a signal-dependent value as an address): a table lookup in a
    var array, a histogram (var[index] += ...), a signal array read at a computed index.  `sel` values outside 0..7 read /
    write nothing that exists; values that do not fit an int make the reference abort (ST_TOINT)."""
    sel = T.input("sel")
    x = T.input("x", (8,))
    out = T.output("out")
    picked = T.output("picked")
    hist_out = T.output("hist", (4,))
    table = T.var("table", (8,), init=[3, 1, 4, 1, 5, 9, 2, 6])
    hist = T.var("h", (4,), init=[0, 0, 0, 0])
    i = T.var("i")
    T.assign(out, table[sel] * 10 + table[(sel + 1) & 7])
    T.assign(picked, x[sel & 7] + x[(sel * 3 + 1) & 7])
    with T.for_(i, 0, i < 8):
        T.set(hist[x[i] & 3], hist[x[i] & 3] + 1)
    with T.for_(i, 0, i < 4):
        T.assign(hist_out[i], hist[i])
