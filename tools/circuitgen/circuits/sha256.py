"""SHA-256 as a circuit, following the structure of circomlib's circuits/sha256/ (sha256.circom,
sha256compression.circom, sigmaplus.circom, sigma.circom, t1.circom, t2.circom, ch.circom, maj.circom,
xor3.circom, rotate.circom, shift.circom, ../binsum.circom).  circomlib is not in the reference tree
(SURVEY.md 'fixture gap'); the templates below were rewritten from their public definitions and are pinned
by FIPS 180-4 through hashlib (tests compare out[256] with hashlib.sha256 of the input bits).

One structural difference: circomlib instantiates the round constants as components `K(t)` / `H(i)` (one
template instance per constant, i.e. a *mixed* component array).  The generator does not emit mixed arrays
yet, so the constants are bound directly with `<==` from var arrays; the constraints are the same
(`bit === constant`), the component tree is slightly flatter.
"""
from __future__ import annotations

K_CONST = [
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
    0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
    0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
    0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
    0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
    0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
    0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2]
H_CONST = [0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19]


def nbits(a):
    n, r = 1, 0
    while n - 1 < a:
        r += 1
        n *= 2
    return r


def BinSum(T, n, ops):
    """circomlib binsum.circom"""
    nout = nbits(((1 << n) - 1) * ops)
    n_ = T.param("n", n)
    ops_ = T.param("ops", ops)
    nout_ = T.param("nout", nout)
    inp = T.input("in", (ops, n))
    out = T.output("out", (nout,))
    lin = T.var("lin", init=0)
    lout = T.var("lout", init=0)
    k = T.var("k")
    j = T.var("j")
    e2 = T.var("e2", init=1)
    with T.for_(k, 0, k < n_):
        with T.for_(j, 0, j < ops_):
            T.set(lin, lin + inp[j][k] * e2)
        T.set(e2, e2 + e2)
    T.set(e2, 1)
    with T.for_(k, 0, k < nout_):
        T.assign(out[k], (lin >> k) & 1)
        T.constrain(out[k] * (out[k] - 1), 0)
        T.set(lout, lout + out[k] * e2)
        T.set(e2, e2 + e2)
    T.constrain(lin, lout)


def RotR(T, n, r):
    n_ = T.param("n", n)
    r_ = T.param("r", r)
    inp = T.input("in", (n,))
    out = T.output("out", (n,))
    i = T.var("i")
    with T.for_(i, 0, i < n_):
        T.bind(out[i], inp[(i + r_) % n_])


def ShR(T, n, r):
    n_ = T.param("n", n)
    r_ = T.param("r", r)
    inp = T.input("in", (n,))
    out = T.output("out", (n,))
    i = T.var("i")
    with T.for_(i, 0, i < n_):
        with T.if_(i + r_ >= n_):
            T.bind(out[i], 0)
        with T.else_():
            T.bind(out[i], inp[i + r_])


def Xor3(T, n):
    n_ = T.param("n", n)
    a = T.input("a", (n,))
    b = T.input("b", (n,))
    c = T.input("c", (n,))
    out = T.output("out", (n,))
    mid = T.signal("mid", (n,))
    k = T.var("k")
    with T.for_(k, 0, k < n_):
        T.bind(mid[k], b[k] * c[k])
        T.bind(out[k], a[k] * (1 - 2 * b[k] - 2 * c[k] + 4 * mid[k]) + b[k] + c[k] - 2 * mid[k])


def Ch_t(T, n):
    n_ = T.param("n", n)
    a = T.input("a", (n,))
    b = T.input("b", (n,))
    c = T.input("c", (n,))
    out = T.output("out", (n,))
    k = T.var("k")
    with T.for_(k, 0, k < n_):
        T.bind(out[k], a[k] * (b[k] - c[k]) + c[k])


def Maj_t(T, n):
    n_ = T.param("n", n)
    a = T.input("a", (n,))
    b = T.input("b", (n,))
    c = T.input("c", (n,))
    out = T.output("out", (n,))
    mid = T.signal("mid", (n,))
    k = T.var("k")
    with T.for_(k, 0, k < n_):
        T.bind(mid[k], b[k] * c[k])
        T.bind(out[k], a[k] * (b[k] + c[k] - 2 * mid[k]) + mid[k])


def SmallSigma(T, ra, rb, rc):
    inp = T.input("in", (32,))
    out = T.output("out", (32,))
    rota = T.component("rota")
    rotb = T.component("rotb")
    shrc = T.component("shrc")
    xor3 = T.component("xor3")
    T.new(rota, RotR, 32, ra)
    T.new(rotb, RotR, 32, rb)
    T.new(shrc, ShR, 32, rc)
    T.new(xor3, Xor3, 32)
    k = T.var("k")
    with T.for_(k, 0, k < 32):
        T.bind(rota.pin("in")[k], inp[k])
        T.bind(rotb.pin("in")[k], inp[k])
        T.bind(shrc.pin("in")[k], inp[k])
    with T.for_(k, 0, k < 32):
        T.bind(xor3.pin("a")[k], rota.pin("out")[k])
        T.bind(xor3.pin("b")[k], rotb.pin("out")[k])
        T.bind(xor3.pin("c")[k], shrc.pin("out")[k])
    with T.for_(k, 0, k < 32):
        T.bind(out[k], xor3.pin("out")[k])


def BigSigma(T, ra, rb, rc):
    inp = T.input("in", (32,))
    out = T.output("out", (32,))
    rota = T.component("rota")
    rotb = T.component("rotb")
    rotc = T.component("rotc")
    xor3 = T.component("xor3")
    T.new(rota, RotR, 32, ra)
    T.new(rotb, RotR, 32, rb)
    T.new(rotc, RotR, 32, rc)
    T.new(xor3, Xor3, 32)
    k = T.var("k")
    with T.for_(k, 0, k < 32):
        T.bind(rota.pin("in")[k], inp[k])
        T.bind(rotb.pin("in")[k], inp[k])
        T.bind(rotc.pin("in")[k], inp[k])
    with T.for_(k, 0, k < 32):
        T.bind(xor3.pin("a")[k], rota.pin("out")[k])
        T.bind(xor3.pin("b")[k], rotb.pin("out")[k])
        T.bind(xor3.pin("c")[k], rotc.pin("out")[k])
    with T.for_(k, 0, k < 32):
        T.bind(out[k], xor3.pin("out")[k])


def SigmaPlus(T):
    in2 = T.input("in2", (32,))
    in7 = T.input("in7", (32,))
    in15 = T.input("in15", (32,))
    in16 = T.input("in16", (32,))
    out = T.output("out", (32,))
    sigma1 = T.component("sigma1")
    sigma0 = T.component("sigma0")
    summ = T.component("sum")
    T.new(sigma1, SmallSigma, 17, 19, 10)
    T.new(sigma0, SmallSigma, 7, 18, 3)
    T.new(summ, BinSum, 32, 4)
    k = T.var("k")
    with T.for_(k, 0, k < 32):
        T.bind(sigma1.pin("in")[k], in2[k])
        T.bind(sigma0.pin("in")[k], in15[k])
    with T.for_(k, 0, k < 32):
        T.bind(summ.pin("in")[0][k], sigma1.pin("out")[k])
        T.bind(summ.pin("in")[1][k], in7[k])
        T.bind(summ.pin("in")[2][k], sigma0.pin("out")[k])
        T.bind(summ.pin("in")[3][k], in16[k])
    with T.for_(k, 0, k < 32):
        T.bind(out[k], summ.pin("out")[k])


def T1(T):
    h = T.input("h", (32,))
    e = T.input("e", (32,))
    f = T.input("f", (32,))
    g = T.input("g", (32,))
    kk = T.input("k", (32,))
    w = T.input("w", (32,))
    out = T.output("out", (32,))
    ch = T.component("ch")
    bigsigma1 = T.component("bigsigma1")
    summ = T.component("sum")
    T.new(ch, Ch_t, 32)
    T.new(bigsigma1, BigSigma, 6, 11, 25)
    T.new(summ, BinSum, 32, 5)
    ki = T.var("ki")
    with T.for_(ki, 0, ki < 32):
        T.bind(bigsigma1.pin("in")[ki], e[ki])
        T.bind(ch.pin("a")[ki], e[ki])
        T.bind(ch.pin("b")[ki], f[ki])
        T.bind(ch.pin("c")[ki], g[ki])
    with T.for_(ki, 0, ki < 32):
        T.bind(summ.pin("in")[0][ki], h[ki])
        T.bind(summ.pin("in")[1][ki], bigsigma1.pin("out")[ki])
        T.bind(summ.pin("in")[2][ki], ch.pin("out")[ki])
        T.bind(summ.pin("in")[3][ki], kk[ki])
        T.bind(summ.pin("in")[4][ki], w[ki])
    with T.for_(ki, 0, ki < 32):
        T.bind(out[ki], summ.pin("out")[ki])


def T2(T):
    a = T.input("a", (32,))
    b = T.input("b", (32,))
    c = T.input("c", (32,))
    out = T.output("out", (32,))
    bigsigma0 = T.component("bigsigma0")
    maj = T.component("maj")
    summ = T.component("sum")
    T.new(bigsigma0, BigSigma, 2, 13, 22)
    T.new(maj, Maj_t, 32)
    T.new(summ, BinSum, 32, 2)
    k = T.var("k")
    with T.for_(k, 0, k < 32):
        T.bind(bigsigma0.pin("in")[k], a[k])
        T.bind(maj.pin("a")[k], a[k])
        T.bind(maj.pin("b")[k], b[k])
        T.bind(maj.pin("c")[k], c[k])
    with T.for_(k, 0, k < 32):
        T.bind(summ.pin("in")[0][k], bigsigma0.pin("out")[k])
        T.bind(summ.pin("in")[1][k], maj.pin("out")[k])
    with T.for_(k, 0, k < 32):
        T.bind(out[k], summ.pin("out")[k])


def Sha256compression(T):
    hin = T.input("hin", (256,))
    inp = T.input("inp", (512,))
    out = T.output("out", (256,))
    regs = {n: T.signal(n, (65, 32)) for n in "abcdefgh"}
    a, b, c, d, e, f, g, h = (regs[n] for n in "abcdefgh")
    w = T.signal("w", (64, 32))
    Kc = T.param("K", K_CONST)
    sigma_plus = T.component("sigmaPlus", (48,))
    t1 = T.component("t1", (64,))
    t2 = T.component("t2", (64,))
    suma = T.component("suma", (64,))
    sume = T.component("sume", (64,))
    fsum = T.component("fsum", (8,))
    i = T.var("i")
    t = T.var("t")
    k = T.var("k")
    with T.for_(i, 0, i < 48):
        T.new(sigma_plus[i], SigmaPlus)
    with T.for_(i, 0, i < 64):
        T.new(t1[i], T1)
        T.new(t2[i], T2)
        T.new(suma[i], BinSum, 32, 2)
        T.new(sume[i], BinSum, 32, 2)
    with T.for_(i, 0, i < 8):
        T.new(fsum[i], BinSum, 32, 2)
    # message schedule
    with T.for_(t, 0, t < 64):
        with T.if_(t < 16):
            with T.for_(k, 0, k < 32):
                T.bind(w[t][k], inp[t * 32 + 31 - k])
        with T.else_():
            with T.for_(k, 0, k < 32):
                T.bind(sigma_plus[t - 16].pin("in2")[k], w[t - 2][k])
                T.bind(sigma_plus[t - 16].pin("in7")[k], w[t - 7][k])
                T.bind(sigma_plus[t - 16].pin("in15")[k], w[t - 15][k])
                T.bind(sigma_plus[t - 16].pin("in16")[k], w[t - 16][k])
            with T.for_(k, 0, k < 32):
                T.bind(w[t][k], sigma_plus[t - 16].pin("out")[k])
    # initial working variables (little-endian bit order inside a word)
    for r, reg in enumerate((a, b, c, d, e, f, g, h)):
        with T.for_(k, 0, k < 32):
            T.bind(reg[0][k], hin[32 * r + k])
    # 64 rounds
    with T.for_(t, 0, t < 64):
        with T.for_(k, 0, k < 32):
            T.bind(t1[t].pin("h")[k], h[t][k])
            T.bind(t1[t].pin("e")[k], e[t][k])
            T.bind(t1[t].pin("f")[k], f[t][k])
            T.bind(t1[t].pin("g")[k], g[t][k])
            T.bind(t1[t].pin("k")[k], (Kc[t] >> k) & 1)
            T.bind(t1[t].pin("w")[k], w[t][k])
            T.bind(t2[t].pin("a")[k], a[t][k])
            T.bind(t2[t].pin("b")[k], b[t][k])
            T.bind(t2[t].pin("c")[k], c[t][k])
        with T.for_(k, 0, k < 32):
            T.bind(sume[t].pin("in")[0][k], d[t][k])
            T.bind(sume[t].pin("in")[1][k], t1[t].pin("out")[k])
            T.bind(suma[t].pin("in")[0][k], t1[t].pin("out")[k])
            T.bind(suma[t].pin("in")[1][k], t2[t].pin("out")[k])
        with T.for_(k, 0, k < 32):
            T.bind(h[t + 1][k], g[t][k])
            T.bind(g[t + 1][k], f[t][k])
            T.bind(f[t + 1][k], e[t][k])
            T.bind(e[t + 1][k], sume[t].pin("out")[k])
            T.bind(d[t + 1][k], c[t][k])
            T.bind(c[t + 1][k], b[t][k])
            T.bind(b[t + 1][k], a[t][k])
            T.bind(a[t + 1][k], suma[t].pin("out")[k])
    # feed-forward
    for r, reg in enumerate((a, b, c, d, e, f, g, h)):
        with T.for_(k, 0, k < 32):
            T.bind(fsum[r].pin("in")[0][k], hin[32 * r + k])
            T.bind(fsum[r].pin("in")[1][k], reg[64][k])
    for r in range(8):
        with T.for_(k, 0, k < 32):
            T.bind(out[r * 32 + 31 - k], fsum[r].pin("out")[k])


def Sha256(T, n_bits):
    n_blocks = ((n_bits + 64) // 512) + 1
    inp = T.input("in", (n_bits,))
    out = T.output("out", (256,))
    padded = T.signal("paddedIn", (n_blocks * 512,))
    Hc = T.param("H", H_CONST)
    comp = T.component("sha256compression", (n_blocks,))
    i = T.var("i")
    k = T.var("k")
    with T.for_(k, 0, k < n_bits):
        T.bind(padded[k], inp[k])
    T.bind(padded[n_bits], 1)
    with T.for_(k, n_bits + 1, k < n_blocks * 512 - 64):
        T.bind(padded[k], 0)
    with T.for_(k, 0, k < 64):
        T.bind(padded[n_blocks * 512 - k - 1], (n_bits >> k) & 1)
    with T.for_(i, 0, i < n_blocks):
        T.new(comp[i], Sha256compression)
    with T.for_(i, 0, i < n_blocks):
        with T.if_(i.eq(0)):
            for r in range(8):
                with T.for_(k, 0, k < 32):
                    T.bind(comp[i].pin("hin")[32 * r + k], (Hc[r] >> k) & 1)
        with T.else_():
            with T.for_(k, 0, k < 32):
                for r in range(8):
                    T.bind(comp[i].pin("hin")[32 * r + k], comp[i - 1].pin("out")[32 * r + 31 - k])
        with T.for_(k, 0, k < 512):
            T.bind(comp[i].pin("inp")[k], padded[i * 512 + k])
    with T.for_(k, 0, k < 256):
        T.bind(out[k], comp[n_blocks - 1].pin("out")[k])


def sha256_bits(bits):
    """hashlib digest of a bit string given MSB-first, as 256 output bits MSB-first."""
    import hashlib
    assert len(bits) % 8 == 0
    data = bytes(int("".join(str(b) for b in bits[i:i + 8]), 2) for i in range(0, len(bits), 8))
    dig = hashlib.sha256(data).digest()
    return [(byte >> (7 - j)) & 1 for byte in dig for j in range(8)]
