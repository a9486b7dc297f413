"""Poseidon over BN254 Fr, as a circuit and as a plain hash.

Not in the reference tree (circomlib is an external library; SURVEY.md 'fixture gap').  The
parameters are regenerated here from the published Poseidon procedure (Grassi et al., "Poseidon",
USENIX Security 2021, reference script generate_parameters_grain: Grain LFSR in self-shrinking mode,
80-bit init = field tag 1 | s-box tag 0 | n=254 | t | R_F | R_P | 30 ones), which is what circomlib's
poseidon_constants.circom was produced with.  Pinned by the published vector
poseidon([1, 2]) = 7853200120776062878684798364095072458815029376092732009249414926327459813530.

Circuit structure: circomlib's poseidon.circom in its 0.5.x form (templates Sigma / Ark / Mix, one
component per round step, full MDS multiplication in every round).  Later circomlib versions compute
the same function with pre-multiplied sparse matrices; the hash value is identical.
"""
from __future__ import annotations

from ..dsl import P

N_ROUNDS_P = [56, 57, 56, 60, 60, 63, 64, 63, 60, 66, 60, 65, 70, 60, 64, 68]
N_ROUNDS_F = 8


class _Grain:
    def __init__(self, t, rf, rp, n=254):
        bits = []

        def put(v, w):
            bits.extend(int(b) for b in bin(v)[2:].zfill(w))
        put(1, 2)
        put(0, 4)
        put(n, 12)
        put(t, 12)
        put(rf, 10)
        put(rp, 10)
        bits.extend([1] * 30)
        self.s = bits
        for _ in range(160):
            self._step()

    def _step(self):
        s = self.s
        b = s[62] ^ s[51] ^ s[38] ^ s[23] ^ s[13] ^ s[0]
        s.pop(0)
        s.append(b)
        return b

    def bit(self):
        b = self._step()
        while b == 0:
            self._step()
            b = self._step()
        return self._step()

    def bits(self, n):
        v = 0
        for _ in range(n):
            v = (v << 1) | self.bit()
        return v


_CACHE = {}


def constants(t):
    """-> (C: list of (R_F+R_P)*t round constants, M: t x t MDS matrix)"""
    if t in _CACHE:
        return _CACHE[t]
    rf, rp = N_ROUNDS_F, N_ROUNDS_P[t - 2]
    g = _Grain(t, rf, rp)
    C = []
    while len(C) < (rf + rp) * t:
        v = g.bits(254)
        if v < P:
            C.append(v)
    while True:
        rand = [g.bits(254) % P for _ in range(2 * t)]
        while len(set(rand)) != len(rand):
            rand = [g.bits(254) % P for _ in range(2 * t)]
        xs, ys = rand[:t], rand[t:]
        if any((x + y) % P == 0 for x in xs for y in ys):
            continue
        M = [[pow((xs[i] + ys[j]) % P, -1, P) for j in range(t)] for i in range(t)]
        break
    _CACHE[t] = (C, M)
    return C, M


def poseidon_hash(inputs):
    """Plain-Python Poseidon (x^5 s-box), capacity element first: the function the circuit computes."""
    t = len(inputs) + 1
    C, M = constants(t)
    rf, rp = N_ROUNDS_F, N_ROUNDS_P[t - 2]
    state = [0] + [x % P for x in inputs]
    for r in range(rf + rp):
        state = [(s + C[r * t + i]) % P for i, s in enumerate(state)]
        if r < rf // 2 or r >= rf // 2 + rp:
            state = [pow(s, 5, P) for s in state]
        else:
            state[0] = pow(state[0], 5, P)
        state = [sum(M[i][j] * state[j] for j in range(t)) % P for i in range(t)]
    return state[0]


# ---------------------------------------------------------------- circuit templates
def Sigma(T):
    inp = T.input("in")
    out = T.output("out")
    in2 = T.signal("in2")
    in4 = T.signal("in4")
    T.bind(in2, inp * inp)
    T.bind(in4, in2 * in2)
    T.bind(out, in4 * inp)


def Ark(T, t, C):
    """out[i] <== in[i] + C[i]  (the round's slice of constants is the template argument)"""
    t_ = T.param("t", t)
    C_ = T.param("C", list(C))
    inp = T.input("in", (t,))
    out = T.output("out", (t,))
    i = T.var("i")
    with T.for_(i, 0, i < t_):
        T.bind(out[i], inp[i] + C_[i])


def Mix(T, t, M):
    t_ = T.param("t", t)
    M_ = T.param("M", [list(r) for r in M])
    inp = T.input("in", (t,))
    out = T.output("out", (t,))
    lc = T.var("lc")
    i = T.var("i")
    j = T.var("j")
    with T.for_(i, 0, i < t_):
        T.set(lc, 0)
        with T.for_(j, 0, j < t_):
            T.set(lc, lc + M_[i][j] * inp[j])
        T.bind(out[i], lc)


def Poseidon(T, n_inputs):
    t = n_inputs + 1
    C, M = constants(t)
    rf, rp = N_ROUNDS_F, N_ROUNDS_P[t - 2]
    inputs = T.input("inputs", (n_inputs,))
    out = T.output("out")
    mix = T.component("mix", (rf + rp,))
    sigma_f = T.component("sigmaF", (rf, t))
    sigma_p = T.component("sigmaP", (rp,))
    arks = [T.component("ark_%d" % r) for r in range(rf + rp)]
    i = T.var("i")
    j = T.var("j")
    for r in range(rf + rp):
        T.new(arks[r], Ark, t, C[r * t:(r + 1) * t])
    with T.for_(i, 0, i < rf + rp):
        T.new(mix[i], Mix, t, M)
    with T.for_(i, 0, i < rf):
        with T.for_(j, 0, j < t):
            T.new(sigma_f[i][j], Sigma)
    with T.for_(i, 0, i < rp):
        T.new(sigma_p[i], Sigma)
    k = T.var("k")
    fi = 0
    for r in range(rf + rp):
        ark = arks[r]
        # Ark input: initial state or previous Mix output
        if r == 0:
            T.bind(ark.pin("in")[0], 0)
            with T.for_(k, 1, k < t):
                T.bind(ark.pin("in")[k], inputs[k - 1])
        else:
            with T.for_(k, 0, k < t):
                T.bind(ark.pin("in")[k], mix[r - 1].pin("out")[k])
        if r < rf // 2 or r >= rf // 2 + rp:
            with T.for_(k, 0, k < t):
                T.bind(sigma_f[fi][k].pin("in"), ark.pin("out")[k])
                T.bind(mix[r].pin("in")[k], sigma_f[fi][k].pin("out"))
            fi += 1
        else:
            pi = r - rf // 2
            T.bind(sigma_p[pi].pin("in"), ark.pin("out")[0])
            T.bind(mix[r].pin("in")[0], sigma_p[pi].pin("out"))
            with T.for_(k, 1, k < t):
                T.bind(mix[r].pin("in")[k], ark.pin("out")[k])
    T.bind(out, mix[rf + rp - 1].pin("out")[0])


def ArkAt(T, t, C, r):
    """circomlib 0.5.x `template Ark(t, C, r)`: the WHOLE constant table and an offset are the arguments, so every round is
    a different template instance."""
    t_ = T.param("t", t)
    C_ = T.param("C", list(C))
    r_ = T.param("r", r)
    inp = T.input("in", (t,))
    out = T.output("out", (t,))
    i = T.var("i")
    with T.for_(i, 0, i < t_):
        T.bind(out[i], inp[i] + C_[i + r_])


def PoseidonMixed(T, n_inputs):
    """The same permutation written the way circomlib 0.5.x's poseidon.circom writes it: ONE loop over the rounds,
    `ark[i] = Ark(t, C, t*i)` inside it.  `ark` is therefore a MIXED component array (each position a different template
    instance) and every `ark[i].in[j]` / `ark[i].out[j]` is a "mapped" access through the io-map (location_rule.rs:86-171),
    with a component index and a signal index that are loop variables."""
    t = n_inputs + 1
    C, M = constants(t)
    rf, rp = N_ROUNDS_F, N_ROUNDS_P[t - 2]
    inputs = T.input("inputs", (n_inputs,))
    out = T.output("out")
    ark = T.component("ark", (rf + rp,))
    sigma_f = T.component("sigmaF", (rf, t))
    sigma_p = T.component("sigmaP", (rp,))
    mix = T.component("mix", (rf + rp,))
    i = T.var("i")
    j = T.var("j")
    k = T.var("k")
    with T.for_(i, 0, i < rf + rp):
        T.new(ark[i], ArkAt, t, C, t * i)
        with T.for_(j, 0, j < t):
            with T.if_(i.eq(0)):
                with T.if_(j > 0):
                    T.bind(ark[i].pin("in")[j], inputs[j - 1])
                with T.else_():
                    T.bind(ark[i].pin("in")[j], 0)
            with T.else_():
                T.bind(ark[i].pin("in")[j], mix[i - 1].pin("out")[j])
        T.new(mix[i], Mix, t, M)
        with T.if_((i < rf // 2).lor(i >= rp + rf // 2)):
            with T.if_(i < rf // 2):
                T.set(k, i)
            with T.else_():
                T.set(k, i - rp)
            with T.for_(j, 0, j < t):
                T.new(sigma_f[k][j], Sigma)
                T.bind(sigma_f[k][j].pin("in"), ark[i].pin("out")[j])
                T.bind(mix[i].pin("in")[j], sigma_f[k][j].pin("out"))
        with T.else_():
            T.set(k, i - rf // 2)
            T.new(sigma_p[k], Sigma)
            T.bind(sigma_p[k].pin("in"), ark[i].pin("out")[0])
            T.bind(mix[i].pin("in")[0], sigma_p[k].pin("out"))
            with T.for_(j, 1, j < t):
                T.bind(mix[i].pin("in")[j], ark[i].pin("out")[j])
    T.bind(out, mix[rf + rp - 1].pin("out")[0])
