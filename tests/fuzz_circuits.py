"""Random straight-line circuits over the whole operator set (test helper, used by the CPU and the GPU suites).

A template with 4 inputs and 12 outputs; a seeded sequence of `var = op(x, y)` statements whose operands are earlier
values, inputs or constants (small, bit-like, near q), data-dependent `if/else` blocks, shifts and bit extractions by
constants and by values, divisions (also by zero), comparisons feeding products (the value-range typing of the trace
compiler), sums of products by constants (the dot-product fusion) and a few `===` that random inputs may violate."""
import random

from tools.circuitgen.circuits.basic import FABS, FCLAMP, FIRST_GE
from tools.circuitgen.dsl import P

BIN = ["add", "sub", "mul", "div", "idiv", "mod", "pow", "shl", "shr", "band", "bor", "bxor", "lt", "leq", "gt", "geq",
       "eq", "neq", "land", "lor"]


def make_circuit(seed, n_stmts=40, bits=False):
    """bits=True adds the shapes the typed paths of the tape compiler live on: sums of 0/1 values times powers of two that
    are taken apart again bit by bit (circomlib BinSum / Num2Bits: small-integer typing, T_ISUM, T_IBITG), boolean
    functions written as field polynomials (a + b - 2ab, Ch, Maj: T_LUT / T_LUTG), bit values feeding field arithmetic and
    constraints over them (the check's integer and truth-table paths)."""
    rng = random.Random(seed)

    def tmpl(T):
        inp = T.input("in", (4,))
        out = T.output("out", (12,))
        vals = [inp[k] for k in range(4)]
        vars_ = []

        def const():
            c = rng.choice([0, 1, 2, 3, 5, 31, 32, 200, 253, 254, 255, (1 << 64) - 1, 1 << 200, P - 1, P - 2, P - 200,
                            (P - 1) // 2, (P + 1) // 2, rng.randrange(P)])
            return c

        def operand():
            return rng.choice(vals) if rng.random() < 0.8 else const()

        def expr():
            r = rng.random()
            x, y = operand(), operand()
            if isinstance(x, int) and isinstance(y, int):
                x = rng.choice(vals)
            if r < 0.08:
                return (x >> rng.randrange(0, 254)) & 1            # bit extraction
            if r < 0.14:
                c1, c2, c3 = rng.randrange(P), rng.randrange(P), rng.randrange(P)
                return c1 * rng.choice(vals) + c2 * rng.choice(vals) + c3 * rng.choice(vals) + const()   # dot product
            if r < 0.20:
                return (x < y) * rng.choice(vals)                  # 0/1 factor
            if r < 0.24:
                return -x if not isinstance(x, int) else ~y if not isinstance(y, int) else x
            if r < 0.28:
                return ~x if not isinstance(x, int) else x
            if r < 0.32:
                return x.lnot() if not isinstance(x, int) else x
            op = rng.choice(BIN)
            if isinstance(x, int):
                x, y = y, x                                        # keep the expression object on the left
            if op == "pow":
                y = rng.choice([0, 1, 2, 3, 5, 7, 300])
            if op in ("shl", "shr") and rng.random() < 0.7:
                y = rng.choice([0, 1, 7, 31, 32, 33, 100, 253, 254, 300, P - 3])
            return x._b(op, y)

        bools = []

        def a_bool():
            if bools and rng.random() < 0.7:
                return rng.choice(bools)
            x = rng.choice(vals)
            r = rng.random()
            if r < 0.4:
                return (x >> rng.randrange(0, 40)) & 1
            if r < 0.7:
                return x < rng.choice(vals)
            return x.eq(rng.choice([0, 1, 2]))

        for k in range(n_stmts):
            v = T.var("v%d" % k)
            if bits and rng.random() < 0.45:
                r = rng.random()
                if r < 0.35:
                    # lin = sum_j b_j * 2^(k_j) (several operands overlap, like BinSum), then its bits
                    n_terms = rng.choice([2, 3, 8, 20, 33, 40])
                    acc = None
                    for j in range(n_terms):
                        t = a_bool() * (1 << rng.choice([j % 34, j % 34, rng.randrange(0, 36), 61 if rng.random() < 0.03 else 0]))
                        acc = t if acc is None else acc + t
                    T.set(v, acc)
                    vals.append(v)
                    vars_.append(v)
                    lo = rng.randrange(0, 8)
                    for j in range(rng.choice([1, 3, 9, 36])):
                        b = T.var("v%d_b%d" % (k, j))
                        T.set(b, (v >> (lo + j)) & 1)
                        bools.append(b)
                        vars_.append(b)
                    continue
                x, y, z = a_bool(), a_bool(), a_bool()
                if r < 0.55:
                    T.set(v, x + y - 2 * x * y)                          # xor
                elif r < 0.7:
                    T.set(v, x * (y - z) + z)                            # Ch
                elif r < 0.8:
                    T.set(v, x * y + z * (x + y - 2 * x * y))            # Maj
                elif r < 0.9:
                    T.set(v, 1 - x)
                else:
                    T.set(v, x * rng.choice(vals) + y * 5)               # bits into field arithmetic
                    vals.append(v)
                    vars_.append(v)
                    continue
                bools.append(v)
                vars_.append(v)
                continue
            if rng.random() < 0.15 and len(vals) > 4:
                cond = rng.choice(vals)
                with T.if_(cond if rng.random() < 0.5 else cond.ne(rng.choice([0, 1]))):
                    T.set(v, expr())
                with T.else_():
                    T.set(v, expr())
            elif rng.random() < 0.12:
                # circom functions with early returns under data-dependent conditions (also from inside a loop)
                fn = rng.choice([FABS, FCLAMP, FIRST_GE])
                nargs = {id(FABS): 1, id(FCLAMP): 3, id(FIRST_GE): 2}[id(fn)]
                T.set(v, T.call(fn, *[rng.choice(vals) for _ in range(nargs)]))
            else:
                T.set(v, expr())
            vals.append(v)
            vars_.append(v)
        for k in range(12):
            T.assign(out[k], rng.choice(vars_))
        # constraints must be quadratic in SIGNALS (as in circom): the outputs.  Odd seeds get one that random inputs
        # usually violate (status parity), all get a trivially true one
        T.constrain(out[0] * 1, out[0])
        if seed % 2 == 1:
            T.constrain(out[rng.randrange(12)] * out[rng.randrange(12)], out[rng.randrange(12)])
        if bits:
            # constraints over 0/1 outputs with small coefficients (truth-table / integer paths of the check), some of them
            # false for some inputs
            for _ in range(4):
                a, b, c = (out[rng.randrange(12)] for _ in range(3))
                T.constrain((a * rng.choice([1, 2, -1]) + rng.choice([0, 1])) * (b - rng.choice([0, 1])), c * rng.choice([0, 1, 3]))
    tmpl.__name__ = "Fuzz%d" % seed
    return tmpl


def inputs_for(seed, n):
    rng = random.Random(seed * 7919 + 1)
    rows = []
    for _ in range(n):
        rows.append([rng.choice([0, 1, 2, P - 1, rng.randrange(P), rng.randrange(1 << 40), rng.randrange(256)]) for _ in range(4)])
    return rows
