"""EdDSAPoseidonVerifier (BASELINE config 4), restated from circomlib's public templates -- eddsaposeidon.circom,
escalarmulany.circom, escalarmulfix.circom, montgomery.circom, mux3.circom, compconstant.circom, aliascheck.circom,
bitify.circom, comparators.circom -- which are not in the reference tree (SURVEY.md 'fixture gap').

Deviations from the library's text, none of which changes a constraint or a witness value:
  * component arrays whose members take different template arguments (`segments[nsegments]` of EscalarMulAny /
    EscalarMulFix) are declared as separate components (the generator has no mixed arrays);
  * `if (i == 0)` inside construction loops is peeled; CompConstant's compile-time case analysis on the bits of
    the constant is done while building the template.
Pinned end to end by tests: signatures made by the plain-integer signer below satisfy every constraint (the final
ForceEqualIfEnabled only holds if both scalar multiplications and the hash agree with the integer model).
"""
from __future__ import annotations

from ..dsl import P
from . import babyjub as BJ
from .babyjub import BabyAdd, BabyDbl
from .basic import IsZero, Num2Bits
from .poseidon import Poseidon, poseidon_hash

MONT_A = 168698          # (2*(a+d))/(a-d)
MONT_B = 1               # 4/(a-d)


# ---------------------------------------------------------------- montgomery.circom
def Edwards2Montgomery(T):
    inp = T.input("in", (2,))
    out = T.output("out", (2,))
    T.assign(out[0], (1 + inp[1]) / (1 - inp[1]))
    T.assign(out[1], out[0] / inp[0])
    T.constrain(out[0] * (1 - inp[1]), 1 + inp[1])
    T.constrain(out[1] * inp[0], out[0])


def Montgomery2Edwards(T):
    inp = T.input("in", (2,))
    out = T.output("out", (2,))
    T.assign(out[0], inp[0] / inp[1])
    T.assign(out[1], (inp[0] - 1) / (inp[0] + 1))
    T.constrain(out[0] * inp[1], inp[0])
    T.constrain(out[1] * (inp[0] + 1), inp[0] - 1)


def MontgomeryAdd(T):
    in1 = T.input("in1", (2,))
    in2 = T.input("in2", (2,))
    out = T.output("out", (2,))
    lamda = T.signal("lamda")
    T.assign(lamda, (in2[1] - in1[1]) / (in2[0] - in1[0]))
    T.constrain(lamda * (in2[0] - in1[0]), in2[1] - in1[1])
    T.bind(out[0], MONT_B * lamda * lamda - MONT_A - in1[0] - in2[0])
    T.bind(out[1], lamda * (in1[0] - out[0]) - in1[1])


def MontgomeryDouble(T):
    inp = T.input("in", (2,))
    out = T.output("out", (2,))
    lamda = T.signal("lamda")
    x1_2 = T.signal("x1_2")
    T.bind(x1_2, inp[0] * inp[0])
    T.assign(lamda, (3 * x1_2 + 2 * MONT_A * inp[0] + 1) / (2 * MONT_B * inp[1]))
    T.constrain(lamda * (2 * MONT_B * inp[1]), 3 * x1_2 + 2 * MONT_A * inp[0] + 1)
    T.bind(out[0], MONT_B * lamda * lamda - MONT_A - 2 * inp[0])
    T.bind(out[1], lamda * (inp[0] - out[0]) - inp[1])


# ---------------------------------------------------------------- escalarmulany.circom
def Multiplexor2(T):
    sel = T.input("sel")
    inp = T.input("in", (2, 2))
    out = T.output("out", (2,))
    T.bind(out[0], (inp[1][0] - inp[0][0]) * sel + inp[0][0])
    T.bind(out[1], (inp[1][1] - inp[0][1]) * sel + inp[0][1])


def BitElementMulAny(T):
    sel = T.input("sel")
    dbl_in = T.input("dblIn", (2,))
    add_in = T.input("addIn", (2,))
    dbl_out = T.output("dblOut", (2,))
    add_out = T.output("addOut", (2,))
    doubler = T.component("doubler")
    adder = T.component("adder")
    selector = T.component("selector")
    T.new(doubler, MontgomeryDouble)
    T.new(adder, MontgomeryAdd)
    T.new(selector, Multiplexor2)
    T.bind(selector.pin("sel"), sel)
    for k in range(2):
        T.bind(doubler.pin("in")[k], dbl_in[k])
    for k in range(2):
        T.bind(adder.pin("in1")[k], doubler.pin("out")[k])
    for k in range(2):
        T.bind(adder.pin("in2")[k], add_in[k])
    for k in range(2):
        T.bind(selector.pin("in")[0][k], add_in[k])
    for k in range(2):
        T.bind(selector.pin("in")[1][k], adder.pin("out")[k])
    for k in range(2):
        T.bind(dbl_out[k], doubler.pin("out")[k])
    for k in range(2):
        T.bind(add_out[k], selector.pin("out")[k])


def SegmentMulAny(T, n):
    e = T.input("e", (n,))
    p = T.input("p", (2,))
    out = T.output("out", (2,))
    dbl = T.output("dbl", (2,))
    bits = T.component("bits", (n - 1,))
    e2m = T.component("e2m")
    m2e = T.component("m2e")
    eadder = T.component("eadder")
    last_sel = T.component("lastSel")
    i = T.var("i")
    T.new(e2m, Edwards2Montgomery)
    for k in range(2):
        T.bind(e2m.pin("in")[k], p[k])
    with T.for_(i, 0, i < n - 1):
        T.new(bits[i], BitElementMulAny)
    for k in range(2):
        T.bind(bits[0].pin("dblIn")[k], e2m.pin("out")[k])
    for k in range(2):
        T.bind(bits[0].pin("addIn")[k], e2m.pin("out")[k])
    T.bind(bits[0].pin("sel"), e[1])
    with T.for_(i, 1, i < n - 1):
        T.bind(bits[i].pin("dblIn")[0], bits[i - 1].pin("dblOut")[0])
        T.bind(bits[i].pin("dblIn")[1], bits[i - 1].pin("dblOut")[1])
        T.bind(bits[i].pin("addIn")[0], bits[i - 1].pin("addOut")[0])
        T.bind(bits[i].pin("addIn")[1], bits[i - 1].pin("addOut")[1])
        T.bind(bits[i].pin("sel"), e[i + 1])
    for k in range(2):
        T.bind(dbl[k], bits[n - 2].pin("dblOut")[k])
    T.new(m2e, Montgomery2Edwards)
    for k in range(2):
        T.bind(m2e.pin("in")[k], bits[n - 2].pin("addOut")[k])
    T.new(eadder, BabyAdd)
    T.bind(eadder.pin("x1"), m2e.pin("out")[0])
    T.bind(eadder.pin("y1"), m2e.pin("out")[1])
    T.bind(eadder.pin("x2"), -p[0])
    T.bind(eadder.pin("y2"), p[1])
    T.new(last_sel, Multiplexor2)
    T.bind(last_sel.pin("sel"), e[0])
    T.bind(last_sel.pin("in")[0][0], eadder.pin("xout"))
    T.bind(last_sel.pin("in")[0][1], eadder.pin("yout"))
    T.bind(last_sel.pin("in")[1][0], m2e.pin("out")[0])
    T.bind(last_sel.pin("in")[1][1], m2e.pin("out")[1])
    for k in range(2):
        T.bind(out[k], last_sel.pin("out")[k])


def EscalarMulAny(T, n):
    e = T.input("e", (n,))
    p = T.input("p", (2,))
    out = T.output("out", (2,))
    nsegments = (n - 1) // 148 + 1
    nlast = n - (nsegments - 1) * 148
    segments = [T.component("segments_%d" % s) for s in range(nsegments)]
    doublers = [T.component("doublers_%d" % s) for s in range(nsegments - 1)]
    m2e = [T.component("m2e_%d" % s) for s in range(nsegments - 1)]
    adders = [T.component("adders_%d" % s) for s in range(nsegments - 1)]
    zeropoint = T.component("zeropoint")
    T.new(zeropoint, IsZero)
    T.bind(zeropoint.pin("in"), p[0])
    i = T.var("i")
    for s in range(nsegments):
        nseg = 148 if s < nsegments - 1 else nlast
        T.new(segments[s], SegmentMulAny, nseg)
        with T.for_(i, 0, i < nseg):
            T.bind(segments[s].pin("e")[i], e[s * 148 + i])
        if s == 0:
            # force G8 point if input point is zero
            T.bind(segments[s].pin("p")[0], p[0] + (BJ.BASE8[0] - p[0]) * zeropoint.pin("out"))
            T.bind(segments[s].pin("p")[1], p[1] + (BJ.BASE8[1] - p[1]) * zeropoint.pin("out"))
        else:
            T.new(doublers[s - 1], MontgomeryDouble)
            T.new(m2e[s - 1], Montgomery2Edwards)
            T.new(adders[s - 1], BabyAdd)
            for k in range(2):
                T.bind(doublers[s - 1].pin("in")[k], segments[s - 1].pin("dbl")[k])
            for k in range(2):
                T.bind(m2e[s - 1].pin("in")[k], doublers[s - 1].pin("out")[k])
            for k in range(2):
                T.bind(segments[s].pin("p")[k], m2e[s - 1].pin("out")[k])
            if s == 1:
                T.bind(adders[s - 1].pin("x1"), segments[s - 1].pin("out")[0])
                T.bind(adders[s - 1].pin("y1"), segments[s - 1].pin("out")[1])
            else:
                T.bind(adders[s - 1].pin("x1"), adders[s - 2].pin("xout"))
                T.bind(adders[s - 1].pin("y1"), adders[s - 2].pin("yout"))
            T.bind(adders[s - 1].pin("x2"), segments[s].pin("out")[0])
            T.bind(adders[s - 1].pin("y2"), segments[s].pin("out")[1])
    if nsegments == 1:
        T.bind(out[0], segments[0].pin("out")[0] * (1 - zeropoint.pin("out")))
        T.bind(out[1], segments[0].pin("out")[1] + (1 - segments[0].pin("out")[1]) * zeropoint.pin("out"))
    else:
        last = adders[nsegments - 2]
        T.bind(out[0], last.pin("xout") * (1 - zeropoint.pin("out")))
        T.bind(out[1], last.pin("yout") + (1 - last.pin("yout")) * zeropoint.pin("out"))


# ---------------------------------------------------------------- mux3.circom
def MultiMux3(T, n):
    c = T.input("c", (n, 8))
    s = T.input("s", (3,))
    out = T.output("out", (n,))
    a210 = T.signal("a210", (n,))
    a21 = T.signal("a21", (n,))
    a20 = T.signal("a20", (n,))
    a2 = T.signal("a2", (n,))
    a10 = T.signal("a10", (n,))
    a1 = T.signal("a1", (n,))
    a0 = T.signal("a0", (n,))
    a = T.signal("a", (n,))
    s10 = T.signal("s10")
    T.bind(s10, s[1] * s[0])
    i = T.var("i")
    with T.for_(i, 0, i < n):
        T.bind(a210[i], (c[i][7] - c[i][6] - c[i][5] + c[i][4] - c[i][3] + c[i][2] + c[i][1] - c[i][0]) * s10)
        T.bind(a21[i], (c[i][6] - c[i][4] - c[i][2] + c[i][0]) * s[1])
        T.bind(a20[i], (c[i][5] - c[i][4] - c[i][1] + c[i][0]) * s[0])
        T.bind(a2[i], c[i][4] - c[i][0])
        T.bind(a10[i], (c[i][3] - c[i][2] - c[i][1] + c[i][0]) * s10)
        T.bind(a1[i], (c[i][2] - c[i][0]) * s[1])
        T.bind(a0[i], (c[i][1] - c[i][0]) * s[0])
        T.bind(a[i], c[i][0])
        T.bind(out[i], (a210[i] + a21[i] + a20[i] + a2[i]) * s[2] + (a10[i] + a1[i] + a0[i] + a[i]))


# ---------------------------------------------------------------- escalarmulfix.circom
def WindowMulFix(T):
    inp = T.input("in", (3,))
    base = T.input("base", (2,))
    out = T.output("out", (2,))
    out8 = T.output("out8", (2,))
    mux = T.component("mux")
    dbl2 = T.component("dbl2")
    adr = [T.component("adr%d" % k) for k in range(3, 9)]
    T.new(mux, MultiMux3, 2)
    for k in range(3):
        T.bind(mux.pin("s")[k], inp[k])
    T.new(dbl2, MontgomeryDouble)
    for a_ in adr:
        T.new(a_, MontgomeryAdd)
    # in[0] -> 1*BASE
    T.bind(mux.pin("c")[0][0], base[0])
    T.bind(mux.pin("c")[1][0], base[1])
    # in[1] -> 2*BASE
    T.bind(dbl2.pin("in")[0], base[0])
    T.bind(dbl2.pin("in")[1], base[1])
    T.bind(mux.pin("c")[0][1], dbl2.pin("out")[0])
    T.bind(mux.pin("c")[1][1], dbl2.pin("out")[1])
    prev = dbl2
    for k, a_ in enumerate(adr):          # 3*BASE ... 8*BASE
        T.bind(a_.pin("in1")[0], base[0])
        T.bind(a_.pin("in1")[1], base[1])
        T.bind(a_.pin("in2")[0], prev.pin("out")[0])
        T.bind(a_.pin("in2")[1], prev.pin("out")[1])
        if k < 5:
            T.bind(mux.pin("c")[0][k + 2], a_.pin("out")[0])
            T.bind(mux.pin("c")[1][k + 2], a_.pin("out")[1])
        prev = a_
    # adr8 is 8*BASE: only out8; mux.c[.][7] is 8*BASE in the library (adr8 feeds both)
    T.bind(mux.pin("c")[0][7], adr[5].pin("out")[0])
    T.bind(mux.pin("c")[1][7], adr[5].pin("out")[1])
    T.bind(out8[0], adr[5].pin("out")[0])
    T.bind(out8[1], adr[5].pin("out")[1])
    T.bind(out[0], mux.pin("out")[0])
    T.bind(out[1], mux.pin("out")[1])


def SegmentMulFix(T, n_windows):
    e = T.input("e", (n_windows * 3,))
    base = T.input("base", (2,))
    out = T.output("out", (2,))
    dbl = T.output("dbl", (2,))
    e2m = T.component("e2m")
    windows = T.component("windows", (n_windows,))
    adders = T.component("adders", (n_windows,))
    cadders = T.component("cadders", (n_windows,))
    dbl_last = T.component("dblLast")
    m2e = T.component("m2e")
    cm2e = T.component("cm2e")
    cadd = T.component("cAdd")
    i = T.var("i")
    j = T.var("j")
    T.new(e2m, Edwards2Montgomery)
    T.bind(e2m.pin("in")[0], base[0])
    T.bind(e2m.pin("in")[1], base[1])
    T.new(dbl_last, MontgomeryDouble)
    with T.for_(i, 0, i < n_windows):
        T.new(windows[i], WindowMulFix)
        T.new(cadders[i], MontgomeryAdd)
        T.new(adders[i], MontgomeryAdd)
    # Statement order matters for the witness program (a sub-component's outputs exist once its last input is set):
    # per window, as in the library: base, cadders.in1, in[0..2], then cadders.in2 from the window's out8.
    lw = n_windows - 1

    def feed_window(idx, first):
        if first:
            for k in range(2):
                T.bind(windows[0].pin("base")[k], e2m.pin("out")[k])
            for k in range(2):
                T.bind(cadders[0].pin("in1")[k], e2m.pin("out")[k])
        else:
            T.bind(windows[idx].pin("base")[0], windows[idx - 1].pin("out8")[0])
            T.bind(windows[idx].pin("base")[1], windows[idx - 1].pin("out8")[1])
            T.bind(cadders[idx].pin("in1")[0], cadders[idx - 1].pin("out")[0])
            T.bind(cadders[idx].pin("in1")[1], cadders[idx - 1].pin("out")[1])
        with T.for_(j, 0, j < 3):
            T.bind(windows[idx].pin("in")[j], e[3 * idx + j])

    def close_window(idx, last):
        if not last:
            T.bind(cadders[idx].pin("in2")[0], windows[idx].pin("out8")[0])
            T.bind(cadders[idx].pin("in2")[1], windows[idx].pin("out8")[1])
        else:
            # In the last step an extra doubler is added so that numbers do not match
            for k in range(2):
                T.bind(dbl_last.pin("in")[k], windows[idx].pin("out8")[k])
            for k in range(2):
                T.bind(cadders[idx].pin("in2")[k], dbl_last.pin("out")[k])

    feed_window(0, True)
    close_window(0, lw == 0)
    if lw >= 1:
        with T.for_(i, 1, i < lw):
            feed_window(i, False)
            close_window(i, False)
        feed_window(lw, False)
        close_window(lw, True)
    for k in range(2):
        T.bind(adders[0].pin("in1")[k], dbl_last.pin("out")[k])
    for k in range(2):
        T.bind(adders[0].pin("in2")[k], windows[0].pin("out")[k])
    with T.for_(i, 1, i < n_windows):
        T.bind(adders[i].pin("in1")[0], adders[i - 1].pin("out")[0])
        T.bind(adders[i].pin("in1")[1], adders[i - 1].pin("out")[1])
        T.bind(adders[i].pin("in2")[0], windows[i].pin("out")[0])
        T.bind(adders[i].pin("in2")[1], windows[i].pin("out")[1])
    T.new(m2e, Montgomery2Edwards)
    T.new(cm2e, Montgomery2Edwards)
    for k in range(2):
        T.bind(m2e.pin("in")[k], adders[lw].pin("out")[k])
    for k in range(2):
        T.bind(cm2e.pin("in")[k], cadders[lw].pin("out")[k])
    T.new(cadd, BabyAdd)
    T.bind(cadd.pin("x1"), m2e.pin("out")[0])
    T.bind(cadd.pin("y1"), m2e.pin("out")[1])
    T.bind(cadd.pin("x2"), -cm2e.pin("out")[0])
    T.bind(cadd.pin("y2"), cm2e.pin("out")[1])
    T.bind(out[0], cadd.pin("xout"))
    T.bind(out[1], cadd.pin("yout"))
    for k in range(2):
        T.bind(dbl[k], windows[lw].pin("out8")[k])


def EscalarMulFix(T, n, base):
    e = T.input("e", (n,))
    out = T.output("out", (2,))
    nsegments = (n - 1) // 246 + 1
    nlast = n - (nsegments - 1) * 249
    segments = [T.component("segments_%d" % s) for s in range(nsegments)]
    m2e = [T.component("m2e_%d" % s) for s in range(nsegments - 1)]
    adders = [T.component("adders_%d" % s) for s in range(nsegments - 1)]
    i = T.var("i")
    for s in range(nsegments):
        nseg = 249 if s < nsegments - 1 else nlast
        n_windows = (nseg - 1) // 3 + 1
        T.new(segments[s], SegmentMulFix, n_windows)
        with T.for_(i, 0, i < nseg):
            T.bind(segments[s].pin("e")[i], e[s * 249 + i])
        with T.for_(i, nseg, i < n_windows * 3):
            T.bind(segments[s].pin("e")[i], 0)
        if s == 0:
            T.bind(segments[s].pin("base")[0], base[0])
            T.bind(segments[s].pin("base")[1], base[1])
        else:
            T.new(m2e[s - 1], Montgomery2Edwards)
            T.new(adders[s - 1], BabyAdd)
            for k in range(2):
                T.bind(m2e[s - 1].pin("in")[k], segments[s - 1].pin("dbl")[k])
            for k in range(2):
                T.bind(segments[s].pin("base")[k], m2e[s - 1].pin("out")[k])
            if s == 1:
                T.bind(adders[s - 1].pin("x1"), segments[s - 1].pin("out")[0])
                T.bind(adders[s - 1].pin("y1"), segments[s - 1].pin("out")[1])
            else:
                T.bind(adders[s - 1].pin("x1"), adders[s - 2].pin("xout"))
                T.bind(adders[s - 1].pin("y1"), adders[s - 2].pin("yout"))
            T.bind(adders[s - 1].pin("x2"), segments[s].pin("out")[0])
            T.bind(adders[s - 1].pin("y2"), segments[s].pin("out")[1])
    if nsegments == 1:
        T.bind(out[0], segments[0].pin("out")[0])
        T.bind(out[1], segments[0].pin("out")[1])
    else:
        T.bind(out[0], adders[nsegments - 2].pin("xout"))
        T.bind(out[1], adders[nsegments - 2].pin("yout"))


# ---------------------------------------------------------------- compconstant.circom / aliascheck.circom / bitify.circom
def CompConstant(T, ct):
    """out = 1 iff the 254-bit little-endian number in[] is greater than ct."""
    ct %= P
    inp = T.input("in", (254,))
    out = T.output("out")
    parts = T.signal("parts", (127,))
    sout = T.signal("sout")
    num2bits = T.component("num2bits")
    b = (1 << 128) - 1
    a = 1
    e = 1
    total = 0
    for i in range(127):
        clsb = (ct >> (i * 2)) & 1
        cmsb = (ct >> (i * 2 + 1)) & 1
        slsb = inp[i * 2]
        smsb = inp[i * 2 + 1]
        if cmsb == 0 and clsb == 0:
            T.bind(parts[i], -b * smsb * slsb + b * smsb + b * slsb)
        elif cmsb == 0 and clsb == 1:
            T.bind(parts[i], a * smsb * slsb - a * slsb + b * smsb - a * smsb + a)
        elif cmsb == 1 and clsb == 0:
            T.bind(parts[i], b * smsb * slsb - a * smsb + a)
        else:
            T.bind(parts[i], -a * smsb * slsb + a)
        total = total + parts[i]
        b = b - e
        a = a + e
        e = e * 2
    T.bind(sout, total)
    T.new(num2bits, Num2Bits, 135)
    T.bind(num2bits.pin("in"), sout)
    T.bind(out, num2bits.pin("out")[127])


def AliasCheck(T):
    inp = T.input("in", (254,))
    cc = T.component("compConstant")
    T.new(cc, CompConstant, -1)
    i = T.var("i")
    with T.for_(i, 0, i < 254):
        T.bind(cc.pin("in")[i], inp[i])
    T.constrain(cc.pin("out"), 0)


def Num2Bits_strict(T):
    inp = T.input("in")
    out = T.output("out", (254,))
    alias = T.component("aliasCheck")
    n2b = T.component("n2b")
    T.new(alias, AliasCheck)
    T.new(n2b, Num2Bits, 254)
    T.bind(n2b.pin("in"), inp)
    i = T.var("i")
    with T.for_(i, 0, i < 254):
        T.bind(out[i], n2b.pin("out")[i])
        T.bind(alias.pin("in")[i], n2b.pin("out")[i])


def ForceEqualIfEnabled(T):
    enabled = T.input("enabled")
    inp = T.input("in", (2,))
    isz = T.component("isz")
    T.new(isz, IsZero)
    T.bind(isz.pin("in"), inp[1] - inp[0])
    T.constrain((1 - isz.pin("out")) * enabled, 0)


# ---------------------------------------------------------------- eddsaposeidon.circom
def EdDSAPoseidonVerifier(T):
    enabled = T.input("enabled")
    ax = T.input("Ax")
    ay = T.input("Ay")
    s_ = T.input("S")
    r8x = T.input("R8x")
    r8y = T.input("R8y")
    m = T.input("M")
    i = T.var("i")
    # Ensure S < subgroup order
    snum2bits = T.component("snum2bits")
    T.new(snum2bits, Num2Bits, 253)
    T.bind(snum2bits.pin("in"), s_)
    comp = T.component("compConstant")
    T.new(comp, CompConstant, BJ.SUBORDER - 1)
    with T.for_(i, 0, i < 253):
        T.bind(comp.pin("in")[i], snum2bits.pin("out")[i])
    T.bind(comp.pin("in")[253], 0)
    T.constrain(comp.pin("out") * enabled, 0)
    # h = H(R8, A, M)
    hash_ = T.component("hash")
    T.new(hash_, Poseidon, 5)
    T.bind(hash_.pin("inputs")[0], r8x)
    T.bind(hash_.pin("inputs")[1], r8y)
    T.bind(hash_.pin("inputs")[2], ax)
    T.bind(hash_.pin("inputs")[3], ay)
    T.bind(hash_.pin("inputs")[4], m)
    h2bits = T.component("h2bits")
    T.new(h2bits, Num2Bits_strict)
    T.bind(h2bits.pin("in"), hash_.pin("out"))
    # right2 = h * 8 * A (three doublings put the point in the subgroup)
    dbl1 = T.component("dbl1")
    dbl2 = T.component("dbl2")
    dbl3 = T.component("dbl3")
    T.new(dbl1, BabyDbl)
    T.bind(dbl1.pin("x"), ax)
    T.bind(dbl1.pin("y"), ay)
    T.new(dbl2, BabyDbl)
    T.bind(dbl2.pin("x"), dbl1.pin("xout"))
    T.bind(dbl2.pin("y"), dbl1.pin("yout"))
    T.new(dbl3, BabyDbl)
    T.bind(dbl3.pin("x"), dbl2.pin("xout"))
    T.bind(dbl3.pin("y"), dbl2.pin("yout"))
    # A is not zero
    is_zero = T.component("isZero")
    T.new(is_zero, IsZero)
    T.bind(is_zero.pin("in"), dbl3.pin("x"))
    T.constrain(is_zero.pin("out") * enabled, 0)
    mul_any = T.component("mulAny")
    T.new(mul_any, EscalarMulAny, 254)
    with T.for_(i, 0, i < 254):
        T.bind(mul_any.pin("e")[i], h2bits.pin("out")[i])
    T.bind(mul_any.pin("p")[0], dbl3.pin("xout"))
    T.bind(mul_any.pin("p")[1], dbl3.pin("yout"))
    # right = R8 + right2
    add_right = T.component("addRight")
    T.new(add_right, BabyAdd)
    T.bind(add_right.pin("x1"), r8x)
    T.bind(add_right.pin("y1"), r8y)
    T.bind(add_right.pin("x2"), mul_any.pin("out")[0])
    T.bind(add_right.pin("y2"), mul_any.pin("out")[1])
    # left = S * B8
    mul_fix = T.component("mulFix")
    T.new(mul_fix, EscalarMulFix, 253, BJ.BASE8)
    with T.for_(i, 0, i < 253):
        T.bind(mul_fix.pin("e")[i], snum2bits.pin("out")[i])
    # left == right if enabled
    eqx = T.component("eqCheckX")
    T.new(eqx, ForceEqualIfEnabled)
    T.bind(eqx.pin("enabled"), enabled)
    T.bind(eqx.pin("in")[0], mul_fix.pin("out")[0])
    T.bind(eqx.pin("in")[1], add_right.pin("xout"))
    eqy = T.component("eqCheckY")
    T.new(eqy, ForceEqualIfEnabled)
    T.bind(eqy.pin("enabled"), enabled)
    T.bind(eqy.pin("in")[0], mul_fix.pin("out")[1])
    T.bind(eqy.pin("in")[1], add_right.pin("yout"))


# ---------------------------------------------------------------- host-side signer (plain integers)
def sign(secret_scalar, nonce, msg):
    """EdDSA-Poseidon on Baby Jubjub (the scheme of circomlibjs' eddsa.signPoseidon, with the secret scalar and the
    nonce given directly): A = k*B8, R8 = r*B8, S = r + H(R8, A, M) * 8k mod l.
    -> the circuit's input vector [enabled, Ax, Ay, S, R8x, R8y, M]."""
    k = secret_scalar % BJ.SUBORDER
    r = nonce % BJ.SUBORDER
    ax, ay = BJ.mul(k, BJ.BASE8)
    r8x, r8y = BJ.mul(r, BJ.BASE8)
    hm = poseidon_hash([r8x, r8y, ax, ay, msg % P])
    s = (r + hm * 8 * k) % BJ.SUBORDER
    return [1, ax, ay, s, r8x, r8y, msg % P]


def verify(inputs):
    """The equation the circuit enforces, on integers: S*B8 == R8 + h*(8*A)."""
    _en, ax, ay, s, r8x, r8y, msg = inputs
    hm = poseidon_hash([r8x, r8y, ax, ay, msg])
    a8 = BJ.mul(8, (ax, ay))
    return BJ.mul(s, BJ.BASE8) == BJ.add((r8x, r8y), BJ.mul(hm, a8))
