"""CPU restatement of the reference's BN254 Fr value semantics (TEST INFRASTRUCTURE ONLY).

Every function follows the value-level behaviour of the reference's
code_producers/src/c_elements/generic/fr.cpp (line ranges cited per function; the asm
twin is bn128/fr.asm).  Values are canonical Python ints in [0, q).  The tagged
short/long/Montgomery representation of the reference (bn128/fr.hpp:12-21) never
reaches the .wtns (common/main.cpp:328 normalises), so only field values are modelled.

Pinned against the reference itself: tests/golden/fr_kat.json is produced by
oracle/gen_fr_kat.py from oracle/_ref/libfr_ref.so (the reference's own generic/fr.cpp
compiled here) and this model must reproduce every vector (tests/test_oracle_fr.py).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
"""

Q = 21888242871839275222246405745257275088548364400416034343698204186575808495617
HALF = Q >> 1                      # generic/fr.cpp:9  (half = q/2)
NBITS = 254                        # qbits, c_code_generator.rs:1031
MASK = (1 << NBITS) - 1            # lboMask applied to the top limb, generic/fr.cpp:293-376
R = 1 << 256                       # Montgomery radix, c_code_generator.rs:560-562
R2 = (R * R) % Q
R_INV = pow(R, -1, Q)
NP32 = (-pow(Q, -1, 1 << 32)) % (1 << 32)


class FrError(Exception):
    """Reference behaviour is abort/undefined for this input."""


def signed(v):
    """generic/fr.cpp:1172-1363 -- order is on v > half ? v - q : v."""
    return v - Q if v > HALF else v


def add(a, b):   # generic/fr.cpp:1017-1084
    return (a + b) % Q


def sub(a, b):   # generic/fr.cpp:827-891
    return (a - b) % Q


def neg(a):      # generic/fr.cpp:1372-1398
    return (-a) % Q


def mul(a, b):   # generic/fr.cpp:559-637
    return (a * b) % Q


def square(a):   # generic/fr.cpp:2309-2349
    return (a * a) % Q


def inv(a):      # bn128/fr.cpp:146-157 Fr_inv: mpz_init(mr); mpz_invert(mr, ma, q); Fr_fromMpz(r, mr)
    # 0 has no inverse: mpz_invert returns 0 and leaves mr untouched, i.e. the 0 of mpz_init, so the reference's
    # Fr_inv(0) is 0 and Fr_div(a, 0) is 0 (observed by running the reference's own code here:
    # tests/test_oracle_fr.py::test_reference_division_by_zero_is_zero).
    if a == 0:
        return 0
    return pow(a, -1, Q)


def div(a, b):   # generic/fr.cpp:2908-2912
    return mul(a, inv(b))


def idiv(a, b):  # generic/fr.cpp:2835-2857 (mpz_fdiv_q on canonical integers)
    if b == 0:
        raise FrError("integer division by zero")
    return a // b


def mod(a, b):   # generic/fr.cpp:2859-2875
    if b == 0:
        raise FrError("modulo by zero")
    return a % b


def pow_(a, b):  # generic/fr.cpp:2877-2893 (mpz_powm(a, b, q), exponent = canonical b)
    return pow(a, b, Q)


def _mask_reduce(t):
    t &= MASK
    return t - Q if t >= Q else t


def band(a, b):  # generic/fr.cpp:293-303, 1799-1988
    return _mask_reduce(a & b)


def bor(a, b):   # generic/fr.cpp:305-315
    return _mask_reduce(a | b)


def bxor(a, b):  # generic/fr.cpp:317-327
    return _mask_reduce(a ^ b)


def bnot(a):     # generic/fr.cpp:366-376, 2730-2755
    return _mask_reduce(~a & ((1 << 256) - 1))


def shl(a, b):   # generic/fr.cpp:329-349, 2019-2099, 2233-2307
    if b < NBITS:
        return _mask_reduce(a << b)
    s = Q - b
    return 0 if s >= NBITS else a >> s


def shr(a, b):   # generic/fr.cpp:351-364, 2101-2231
    if b < NBITS:
        return a >> b
    s = Q - b
    return 0 if s >= NBITS else _mask_reduce(a << s)


def eq(a, b):    # generic/fr.cpp:1400-1467
    return int(a == b)


def neq(a, b):   # generic/fr.cpp:1469-1537
    return int(a != b)


def lt(a, b):    # generic/fr.cpp:1294-1363
    return int(signed(a) < signed(b))


def gt(a, b):    # generic/fr.cpp:1583-1650
    return int(signed(a) > signed(b))


def leq(a, b):   # generic/fr.cpp:1652-1768
    return int(signed(a) <= signed(b))


def geq(a, b):   # generic/fr.cpp:1172-1292
    return int(signed(a) >= signed(b))


def land(a, b):  # generic/fr.cpp:1540-1580
    return int(a != 0 and b != 0)


def lor(a, b):   # generic/fr.cpp:1771-1797
    return int(a != 0 or b != 0)


def lnot(a):     # generic/fr.cpp:1086-1100
    return int(a == 0)


def is_true(a):  # generic/fr.cpp:1086-1100 (Fr_isTrue)
    return int(a != 0)


def to_int(a):   # generic/fr.cpp:1102-1170: value must fit int32 around 0 (mod q)
    if a < (1 << 31):
        return a
    if a >= Q - (1 << 31):
        return a - Q
    raise FrError("Fr_toInt overflow")


def to_mont(a):   # generic/fr.cpp:211-214 (rawToMontgomery = MMul by R2)
    return (a * R) % Q


def from_mont(a):  # generic/fr.cpp:216-255
    return (a * R_INV) % Q


def mont_mul(a, b):  # generic/fr.cpp:110-164 (a*b*R^-1 mod q on raw limbs)
    return (a * b * R_INV) % Q


BINOPS = {
    "add": add, "sub": sub, "mul": mul, "div": div, "idiv": idiv, "mod": mod, "pow": pow_,
    "shl": shl, "shr": shr, "band": band, "bor": bor, "bxor": bxor,
    "eq": eq, "neq": neq, "lt": lt, "gt": gt, "leq": leq, "geq": geq,
    "land": land, "lor": lor,
}
UNOPS = {"neg": neg, "bnot": bnot, "lnot": lnot, "inv": inv, "square": square, "copy": lambda a: a}


def to_le32(v):
    return int(v).to_bytes(32, "little")


def from_le32(b):
    return int.from_bytes(bytes(b), "little")
