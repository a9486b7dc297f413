"""Baby Jubjub point arithmetic, written from circomlib's public babyjub.circom definitions (circomlib is
not in the reference tree; SURVEY.md 'fixture gap').  Twisted Edwards curve a*x^2 + y^2 = 1 + d*x^2*y^2 over
BN254 Fr with a = 168700, d = 168696."""
from __future__ import annotations

from ..dsl import P

A = 168700
D = 168696


def BabyAdd(T):
    """circomlib babyjub.circom BabyAdd: complete twisted-Edwards addition, two `<--` divisions."""
    x1 = T.input("x1")
    y1 = T.input("y1")
    x2 = T.input("x2")
    y2 = T.input("y2")
    xout = T.output("xout")
    yout = T.output("yout")
    beta = T.signal("beta")
    gamma = T.signal("gamma")
    delta = T.signal("delta")
    tau = T.signal("tau")
    a = T.var("a", init=A)
    d = T.var("d", init=D)
    T.bind(beta, x1 * y2)
    T.bind(gamma, y1 * x2)
    T.bind(delta, (-a * x1 + y1) * (x2 + y2))
    T.bind(tau, beta * gamma)
    T.assign(xout, (beta + gamma) / (1 + d * tau))
    T.constrain((1 + d * tau) * xout, beta + gamma)
    T.assign(yout, (delta + a * beta - gamma) / (1 - d * tau))
    T.constrain((1 - d * tau) * yout, delta + a * beta - gamma)


def BabyDbl(T):
    x = T.input("x")
    y = T.input("y")
    xout = T.output("xout")
    yout = T.output("yout")
    adder = T.component("adder")
    T.new(adder, BabyAdd)
    T.bind(adder.pin("x1"), x)
    T.bind(adder.pin("y1"), y)
    T.bind(adder.pin("x2"), x)
    T.bind(adder.pin("y2"), y)
    T.bind(xout, adder.pin("xout"))
    T.bind(yout, adder.pin("yout"))


def BabyCheck(T):
    x = T.input("x")
    y = T.input("y")
    x2 = T.signal("x2")
    y2 = T.signal("y2")
    a = T.var("a", init=A)
    d = T.var("d", init=D)
    T.bind(x2, x * x)
    T.bind(y2, y * y)
    T.constrain(a * x2 + y2, 1 + d * x2 * y2)


def BabyAddChain(T, n):
    """Fixture: n dependent BabyAdds (p <- p + q, then q <- old p): 2n divisions on n dependency levels,
    two independent ones per level -- exercises the batched inversion of the trace compiler."""
    p = T.input("p", (2,))
    q = T.input("q", (2,))
    out = T.output("out", (2,))
    adders = T.component("adders", (n,))
    i = T.var("i")
    with T.for_(i, 0, i < n):
        T.new(adders[i], BabyAdd)
    T.bind(adders[0].pin("x1"), p[0])
    T.bind(adders[0].pin("y1"), p[1])
    T.bind(adders[0].pin("x2"), q[0])
    T.bind(adders[0].pin("y2"), q[1])
    with T.for_(i, 1, i < n):
        T.bind(adders[i].pin("x1"), adders[i - 1].pin("xout"))
        T.bind(adders[i].pin("y1"), adders[i - 1].pin("yout"))
        T.bind(adders[i].pin("x2"), adders[i - 1].pin("x1"))
        T.bind(adders[i].pin("y2"), adders[i - 1].pin("y1"))
    T.bind(out[0], adders[n - 1].pin("xout"))
    T.bind(out[1], adders[n - 1].pin("yout"))


# ---- plain-integer model (for tests and the host-side EdDSA signer)
def inv(x):
    return pow(x % P, P - 2, P)


def add(p, q):
    x1, y1 = p
    x2, y2 = q
    t = D * x1 * x2 * y1 * y2 % P
    x3 = (x1 * y2 + y1 * x2) * inv(1 + t) % P
    y3 = (y1 * y2 - A * x1 * x2) * inv(1 - t) % P
    return (x3, y3)


def _padd(p, q):
    """projective (X:Y:Z) unified addition (Bernstein-Lange add-2008-bbjlp): no inversion"""
    x1, y1, z1 = p
    x2, y2, z2 = q
    a_ = z1 * z2 % P
    b_ = a_ * a_ % P
    c_ = x1 * x2 % P
    d_ = y1 * y2 % P
    e_ = D * c_ * d_ % P
    f_ = (b_ - e_) % P
    g_ = (b_ + e_) % P
    x3 = a_ * f_ * ((x1 + y1) * (x2 + y2) - c_ - d_) % P
    y3 = a_ * g_ * (d_ - A * c_) % P
    return (x3, y3, f_ * g_ % P)


def mul(k, p):
    r = (0, 1, 1)
    q = (p[0], p[1], 1)
    while k:
        if k & 1:
            r = _padd(r, q)
        q = _padd(q, q)
        k >>= 1
    zi = inv(r[2])
    return (r[0] * zi % P, r[1] * zi % P)


GENERATOR = (995203441582195749578291179787384436505546430278305826713579947235728471134,
             5472060717959818805561601436314318772137091100104008585924551046643952123905)
BASE8 = (5299619240641551281634865583518297030282874472190772894086521144482721001553,
         16950150798460657717958625567821834550301663161624707787222815936182638968203)
SUBORDER = 2736030358979909402780800718157159386076813972158567259200215660948447373041
