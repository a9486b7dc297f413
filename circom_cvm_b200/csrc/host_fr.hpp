// Host-side field operations on CANONICAL values, built from the same limb routines the
// device uses (fr.cuh).  Used by the tracer for constant folding, and exported through the C ABI
// (cvmgpu_fr_host_op) so that the limb algorithms can be checked against the reference-generated
// KATs without a GPU.  Semantics: SURVEY.md App. B / generic/fr.cpp (line refs in fr.cuh).
#pragma once
#include <string>

#include "fr.cuh"

namespace hostfr {

using fr::Fr;

enum FfOp {
    F_ADD, F_SUB, F_MUL, F_DIV, F_IDIV, F_MOD, F_POW, F_SHL, F_SHR, F_BAND, F_BOR, F_BXOR, F_BNOT,
    F_LT, F_LE, F_GT, F_GE, F_EQ, F_NEQ, F_LAND, F_LOR, F_EQZ, F_NEG, F_INV, F_SQUARE, F_COPY, F_NONE
};

inline Fr from_u64(uint64_t x) {
    Fr r = fr::zero();
    r.v[0] = (uint32_t)x;
    r.v[1] = (uint32_t)(x >> 32);
    return r;
}
inline Fr from_i64(int64_t x) {
    if (x >= 0) return from_u64((uint64_t)x);
    Fr m = from_u64((uint64_t)(-(x + 1)) + 1);   // |x| without overflow on INT64_MIN
    return fr::neg(m);
}
inline Fr boolean(bool b) { return from_u64(b ? 1 : 0); }

inline Fr mul(const Fr &a, const Fr &b) { return fr::mont_mul(fr::mont_mul(a, b), fr::r2_mont()); }

// returns false when the reference aborts (integer division / modulo by zero: GMP raises)
inline bool apply(FfOp op, const Fr &a, const Fr &b, Fr &out) {
    switch (op) {
        case F_ADD: out = fr::add(a, b); return true;
        case F_SUB: out = fr::sub(a, b); return true;
        case F_NEG: out = fr::neg(a); return true;
        case F_MUL: out = mul(a, b); return true;
        case F_SQUARE: out = mul(a, a); return true;
        // 0 has no inverse: the reference's Fr_inv hands back the 0 its mpz_init wrote (bn128/fr.cpp:146-157), so
        // Fr_inv(0) = 0 and Fr_div(a, 0) = 0; mont_inv (a^(q-2)) gives the same
        case F_INV: {
            out = fr::from_mont(fr::mont_inv(fr::to_mont(a)));
            return true;
        }
        case F_DIV: {
            Fr bi = fr::mont_inv(fr::to_mont(b));     // Montgomery form of b^-1
            out = fr::mont_mul(a, bi);                // a * b^-1 * R * R^-1
            return true;
        }
        case F_IDIV: {
            if (fr::is_zero(b)) return false;
            Fr q, r;
            fr::divmod(a, b, q, r);
            out = q;
            return true;
        }
        case F_MOD: {
            if (fr::is_zero(b)) return false;
            Fr q, r;
            fr::divmod(a, b, q, r);
            out = r;
            return true;
        }
        case F_POW: out = fr::from_mont(fr::mont_pow_var(fr::to_mont(a), b)); return true;
        case F_SHL: out = fr::shl(a, b); return true;
        case F_SHR: out = fr::shr(a, b); return true;
        case F_BAND: out = fr::band(a, b); return true;
        case F_BOR: out = fr::bor(a, b); return true;
        case F_BXOR: out = fr::bxor(a, b); return true;
        case F_BNOT: out = fr::bnot(a); return true;
        case F_LT: out = boolean(fr::lt_signed(a, b)); return true;
        case F_GT: out = boolean(fr::lt_signed(b, a)); return true;
        case F_LE: out = boolean(!fr::lt_signed(b, a)); return true;
        case F_GE: out = boolean(!fr::lt_signed(a, b)); return true;
        case F_EQ: out = boolean(fr::equal(a, b)); return true;
        case F_NEQ: out = boolean(!fr::equal(a, b)); return true;
        case F_LAND: out = boolean(!fr::is_zero(a) && !fr::is_zero(b)); return true;
        case F_LOR: out = boolean(!fr::is_zero(a) || !fr::is_zero(b)); return true;
        case F_EQZ: out = boolean(fr::is_zero(a)); return true;
        case F_COPY: out = a; return true;
        default: return false;
    }
}

inline FfOp op_from_name(const std::string &n) {
    static const struct { const char *name; FfOp op; } tab[] = {
        {"add", F_ADD}, {"sub", F_SUB}, {"mul", F_MUL}, {"div", F_DIV}, {"idiv", F_IDIV}, {"mod", F_MOD},
        {"pow", F_POW}, {"shl", F_SHL}, {"shr", F_SHR}, {"band", F_BAND}, {"bor", F_BOR}, {"bxor", F_BXOR},
        {"bnot", F_BNOT}, {"lt", F_LT}, {"leq", F_LE}, {"gt", F_GT}, {"geq", F_GE}, {"eq", F_EQ}, {"neq", F_NEQ},
        {"land", F_LAND}, {"lor", F_LOR}, {"lnot", F_EQZ}, {"neg", F_NEG}, {"inv", F_INV}, {"square", F_SQUARE},
        {"copy", F_COPY}};
    for (auto &e : tab)
        if (n == e.name) return e.op;
    return F_NONE;
}

}  // namespace hostfr
