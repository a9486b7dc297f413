#!/usr/bin/env python3
"""Generate tests/golden/fr_kat.txt from the REFERENCE's own field arithmetic.

TEST INFRASTRUCTURE ONLY.  Runs in the build container (needs oracle/_ref/libfr_ref.so,
i.e. the reference's generic/fr.cpp compiled by oracle/build_ref.py).  The vectors are
committed so that the GPU box -- which has no /root/reference -- can check both the
oracle restatement (oracle/fr_model.py) and the CUDA kernels against them.

Line format:  <op> <a_form> <b_form> <a_hex> <b_hex|-> <out_hex>
forms: see oracle/fr_ref_driver.cpp (0 parsed, 1 Montgomery, 2 negative-short, 3 both).
Special ops: isTrue / toInt print the int result in decimal in <out_hex>;
rawMMul / rawToMont / rawFromMont act on raw 256-bit limbs.
"""
import ctypes
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import fr_model as M  # noqa: E402

Q = M.Q


def load_lib():
    lib = ctypes.CDLL(os.path.join(HERE, "_ref", "libfr_ref.so"))
    lib.frref_op.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_int,
                             ctypes.c_char_p]
    lib.frref_op.restype = ctypes.c_int
    lib.frref_isTrue.argtypes = [ctypes.c_char_p, ctypes.c_int]
    lib.frref_toInt.argtypes = [ctypes.c_char_p, ctypes.c_int]
    for n in ("frref_rawMMul",):
        getattr(lib, n).argtypes = [ctypes.c_char_p] * 3
    for n in ("frref_rawToMontgomery", "frref_rawFromMontgomery"):
        getattr(lib, n).argtypes = [ctypes.c_char_p] * 2
    return lib


def ref_op(lib, op, a, af, b=None, bf=0):
    out = ctypes.create_string_buffer(32)
    rc = lib.frref_op(op.encode(), M.to_le32(a), af, None if b is None else M.to_le32(b), bf, out)
    assert rc == 0, op
    return M.from_le32(out.raw)


EDGE = [0, 1, 2, 3, 7, 31, 32, 63, 64, 65, 127, 128, 253, 254, 255, 256, 257,
        (1 << 31) - 1, 1 << 31, (1 << 31) + 1, (1 << 32) - 1, 1 << 32, (1 << 63) - 1, 1 << 63, (1 << 64) - 1,
        1 << 64, (1 << 128) - 1, (1 << 253) - 1, 1 << 253, (1 << 253) + 12345,
        M.HALF - 1, M.HALF, M.HALF + 1, M.HALF + 2, Q - 1, Q - 2, Q - 3, Q - 31, Q - 32, Q - 64, Q - 100,
        Q - 253, Q - 254, Q - 255, Q - (1 << 31), Q - (1 << 31) - 1, Q - (1 << 31) + 1, Q - (1 << 32)]


def main():
    lib = load_lib()
    rng = random.Random(0xC1C0F00D)

    def rnd():
        k = rng.random()
        if k < 0.35:
            return rng.choice(EDGE)
        if k < 0.5:
            return rng.randrange(0, 1 << rng.choice([8, 16, 31, 32, 40, 64, 100, 200]))
        if k < 0.6:
            return Q - 1 - rng.randrange(0, 1 << rng.choice([8, 16, 31, 32, 64]))
        return rng.randrange(0, Q)

    lines = []
    mism = 0
    per_op = 110
    for op in sorted(M.BINOPS):
        for _ in range(per_op):
            a, b = rnd(), rnd()
            if op in ("shl", "shr") and rng.random() < 0.7:
                b = rng.choice([rng.randrange(0, 300), Q - rng.randrange(1, 300)])
            if op == "pow" and rng.random() < 0.5:
                b = rng.randrange(0, 1 << 16)
            if op in ("div", "idiv", "mod") and b == 0:
                b = 5
            af, bf = rng.randrange(4), rng.randrange(4)
            out = ref_op(lib, op, a, af, b, bf)
            exp = M.BINOPS[op](a, b)
            if out != exp:
                mism += 1
                print("MISMATCH", op, a, af, b, bf, out, exp)
            lines.append("%s %d %d %x %x %x" % (op, af, bf, a, b, out))
    for op in sorted(M.UNOPS):
        for _ in range(per_op):
            a = rnd()
            if op == "inv" and a == 0:
                a = 3
            af = rng.randrange(4)
            out = ref_op(lib, op, a, af)
            exp = M.UNOPS[op](a)
            if out != exp:
                mism += 1
                print("MISMATCH", op, a, af, out, exp)
            lines.append("%s %d 0 %x - %x" % (op, af, a, out))
    for _ in range(per_op):
        a = rnd()
        af = rng.randrange(4)
        t = lib.frref_isTrue(M.to_le32(a), af)
        if t != M.is_true(a):
            mism += 1
        lines.append("isTrue %d 0 %x - %d" % (af, a, t))
    for _ in range(per_op):
        a = rng.choice([rng.randrange(0, 1 << 31), Q - rng.randrange(1, (1 << 31) + 1), rng.choice([0, 1, (1 << 31) - 1, Q - (1 << 31)])])
        af = rng.randrange(4)
        t = lib.frref_toInt(M.to_le32(a), af)
        if t != M.to_int(a):
            mism += 1
            print("MISMATCH toInt", a, t)
        lines.append("toInt %d 0 %x - %d" % (af, a, t))
    # raw Montgomery primitives (inputs < q as the runtime guarantees)
    for _ in range(2 * per_op):
        a, b = rnd(), rnd()
        out = ctypes.create_string_buffer(32)
        lib.frref_rawMMul(M.to_le32(a), M.to_le32(b), out)
        o = M.from_le32(out.raw)
        if o != M.mont_mul(a, b):
            mism += 1
            print("MISMATCH rawMMul")
        lines.append("rawMMul 0 0 %x %x %x" % (a, b, o))
    for name, fn, model in (("rawToMont", lib.frref_rawToMontgomery, M.to_mont),
                            ("rawFromMont", lib.frref_rawFromMontgomery, M.from_mont)):
        for _ in range(per_op):
            a = rnd()
            out = ctypes.create_string_buffer(32)
            fn(M.to_le32(a), out)
            o = M.from_le32(out.raw)
            if o != model(a):
                mism += 1
                print("MISMATCH", name)
            lines.append("%s 0 0 %x - %x" % (name, a, o))
    dst = os.path.join(os.path.dirname(HERE), "tests", "golden", "fr_kat.txt")
    with open(dst, "w") as f:
        f.write("# generated by oracle/gen_fr_kat.py from the reference's generic/fr.cpp (oracle/_ref/libfr_ref.so)\n")
        f.write("\n".join(lines) + "\n")
    print("wrote %d vectors to %s; model mismatches: %d" % (len(lines), dst, mism))
    return 1 if mism else 0


if __name__ == "__main__":
    sys.exit(main())
