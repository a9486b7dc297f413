"""Emit the eight row functions of fr.cuh's mont_sqr_chain (one PTX asm statement per chain).

Row i of a word-serial Montgomery squaring adds a_i * (a_i, 2*a_{i+1}, limbs i+2.. of 2a) at word positions i..7 of the
running value T = X + pend + 2^32 * Y (fr.cuh MontAcc): even positions land on X's register pairs, odd positions on Y's.
The carry of `x[0] += pend` has the weight of Y's first word and ripples through the Y words below the row's first pair.

    python tools/gen_sqr_rows.py > /tmp/rows.inc     # pasted between the GENERATED markers of fr.cuh
"""


def operand(i, p):
    if p == i:
        return "a[%d]" % i
    if p == i + 1:
        return "(a[%d] << 1)" % (i + 1)
    return "d[%d]" % p


def emit_row(i):
    odd = [p for p in range(i, 8) if p % 2 == 1]
    even = [p for p in range(i, 8) if p % 2 == 0]
    out = []
    out.append("__device__ __forceinline__ void sqr_row_%d(MontAcc &t, const uint32_t *a, const uint32_t *d) {" % i)
    # ---- Y chain: operands: y[0..7] (%0-%7), x[0] (%8), pend (%9), multiplicands (%10..), multiplier a[i]
    k0 = (odd[0] - 1) // 2
    lines = ["add.cc.u32 %8, %8, %9;"]
    for w in range(2 * k0):
        lines.append("addc.cc.u32 %%%d, %%%d, 0;" % (w, w))
    nm = len(odd)
    mult = 10 + nm
    for n, p in enumerate(odd):
        k = (p - 1) // 2
        last = n == nm - 1
        lines.append("madc.lo.cc.u32 %%%d, %%%d, %%%d, %%%d;" % (2 * k, 10 + n, mult, 2 * k))
        lines.append("madc.hi%s.u32 %%%d, %%%d, %%%d, %%%d;" % ("" if last and k == 3 else ".cc", 2 * k + 1, 10 + n, mult, 2 * k + 1))
    # odd positions always run up to 7 (pair 3), so the chain ends at the top of Y: no carry out (Y < 2^256)
    assert odd[-1] == 7
    asm = "\\n\\t\"\n        \"".join(lines)
    out.append("    asm(\"%s\"" % asm)
    out.append("        : \"+r\"(t.y[0]), \"+r\"(t.y[1]), \"+r\"(t.y[2]), \"+r\"(t.y[3]), \"+r\"(t.y[4]), \"+r\"(t.y[5]), \"+r\"(t.y[6]), \"+r\"(t.y[7]), \"+r\"(t.x[0])")
    ins = ["\"r\"(t.pend)"] + ["\"r\"(%s)" % operand(i, p) for p in odd] + ["\"r\"(a[%d])" % i]
    out.append("        : %s);" % ", ".join(ins))
    # ---- X chain: x[0..7] (%0-%7), x[8] (%8), multiplicands (%9..), multiplier
    if even:
        nm = len(even)
        mult = 9 + nm
        lines = []
        for n, p in enumerate(even):
            k = p // 2
            lines.append("%s.lo.cc.u32 %%%d, %%%d, %%%d, %%%d;" % ("mad" if n == 0 else "madc", 2 * k, 9 + n, mult, 2 * k))
            lines.append("madc.hi.cc.u32 %%%d, %%%d, %%%d, %%%d;" % (2 * k + 1, 9 + n, mult, 2 * k + 1))
        assert even[-1] == 6
        lines.append("addc.u32 %8, %8, 0;")
        asm = "\\n\\t\"\n        \"".join(lines)
        out.append("    asm(\"%s\"" % asm)
        out.append("        : \"+r\"(t.x[0]), \"+r\"(t.x[1]), \"+r\"(t.x[2]), \"+r\"(t.x[3]), \"+r\"(t.x[4]), \"+r\"(t.x[5]), \"+r\"(t.x[6]), \"+r\"(t.x[7]), \"+r\"(t.x[8])")
        ins = ["\"r\"(%s)" % operand(i, p) for p in even] + ["\"r\"(a[%d])" % i]
        out.append("        : %s);" % ", ".join(ins))
    out.append("}")
    return "\n".join(out)


if __name__ == "__main__":
    print("\n".join(emit_row(i) for i in range(8)))
