"""Extended GPU fuzz of the R1CS check (not part of the test-suite): random constraint systems over every coefficient
class (+-2^k, small +-, general, constants on wire 0), empty and repeated combinations, carry-stressing values; the
first violated constraint per witness against Python integers.  python tools/fuzz_r1cs_gpu.py [n_seeds]"""
import os
import random
import sys
import tempfile

sys.path.insert(0, '.')
from circom_cvm_b200 import engine as E, formats

Q = 21888242871839275222246405745257275088548364400416034343698204186575808495617
TOP = (Q >> 224) - 1
EXT = [Q - 1, Q - 2, (TOP << 224) | ((1 << 224) - 1), (1 << 253) - 1, (1 << 224) - 1, (1 << 32) - 1, (1 << 64) - 1, 1, 0, 2,
       Q - (1 << 32), Q - (1 << 224), (Q - 1) // 2, (Q + 1) // 2]


def coef(rng):
    k = rng.randrange(10)
    if k < 3:
        return rng.choice([1, Q - 1])
    if k < 4:
        return rng.choice([2, 4, 8, Q - 2, Q - 4, Q - 8])
    if k < 6:
        c = rng.choice([(1 << 32) - 1, 3, 5, 16, rng.randrange(9, 1 << 32)])
        return c if rng.random() < 0.5 else Q - c
    if k < 7:
        return rng.choice(EXT[:7] + EXT[10:])
    return rng.randrange(1 << 33, Q - (1 << 33))


def lc(rng, n_free, max_terms):
    n = rng.choice([0, 1, 1, 2, 3, 5, 8, 16, 17, rng.randrange(1, max_terms)])
    out = {}
    for _ in range(n):
        out[rng.randrange(0, n_free)] = coef(rng)       # wire 0 included: constants
    return out


def main(n_seeds):
    n_cases = bad_total = 0
    tmp = tempfile.mkdtemp()
    for seed in range(n_seeds):
        rng = random.Random(90000 + seed)
        n_free = rng.randrange(4, 60)
        cons = []
        for _ in range(rng.randrange(1, 40)):
            kind = rng.randrange(6)
            a = lc(rng, n_free, 40)
            b = dict(a) if kind == 0 else ({} if kind == 1 else lc(rng, n_free, 40))
            if kind == 2:
                a = {}
            c = lc(rng, n_free, 24)
            out_wire = n_free + len(cons)
            c[out_wire] = rng.choice([1, Q - 1, 2, 7, rng.randrange(1 << 40, Q)])   # solved below
            cons.append((a, b, c, out_wire))
        n_wires = n_free + len(cons)
        path = os.path.join(tmp, "f.r1cs")
        formats.write_r1cs(path, [(a, b, c) for a, b, c, _ in cons], n_wires, 0, 0, n_free - 1, list(range(n_wires)))
        r = E.R1cs(path)
        ev = lambda l, w: sum(v * w[k] for k, v in l.items()) % Q
        B = 64
        rows, expect = [], []
        for bidx in range(B):
            w = [1] + [rng.choice(EXT) if rng.random() < 0.4 else rng.randrange(Q) for _ in range(n_free - 1)] + [0] * len(cons)
            first = E.NO_BAD
            for ci, (a, b, c, ow) in enumerate(cons):
                rest = ev({k: v for k, v in c.items() if k != ow}, w)
                prod = ev(a, w) * ev(b, w) % Q if a and b else 0
                w[ow] = (prod - rest) * pow(c[ow], -1, Q) % Q
                if rng.random() < 0.02:                      # violate this one
                    w[ow] = (w[ow] + 1 + rng.randrange(5)) % Q
                    if first == E.NO_BAD:
                        first = ci
            rows.append(w)
            expect.append(first)
        got = r.check(E.ints_to_le(rows, n_wires).reshape(B, n_wires, 32))
        for bidx in range(B):
            n_cases += 1
            if int(got[bidx]) != expect[bidx]:
                bad_total += 1
                print("MISMATCH seed", seed, "witness", bidx, "got", int(got[bidx]), "expected", expect[bidx])
        r.close()
    print("checked", n_cases, "witness/system pairs, bad", bad_total)
    return n_cases, bad_total


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 300)
