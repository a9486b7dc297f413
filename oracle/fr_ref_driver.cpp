// TEST INFRASTRUCTURE ONLY -- C-callable driver around the REFERENCE's own field
// arithmetic (code_producers/src/c_elements/generic/fr.cpp rendered for BN254 by
// oracle/build_ref.py).  Linked into oracle/_ref/libfr_ref.so; used to generate and
// re-check the known-answer vectors under tests/golden/.  Never linked by the product.
//
// Operand "forms" let the KATs reach every representation the reference runtime can
// hold (bn128/fr.hpp:12-21): 0 = as parsed (short if it fits int32, else long normal),
// 1 = Montgomery (Fr_toMontgomery in place: short+Montgomery / long+Montgomery),
// 2 = negative short (value >= q-2^31 stored as shortVal = v-q), 3 = form 2 then
// Montgomery.
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include "fr.hpp"

void Fr_square(PFrElement r, PFrElement a);
void Fr_rawMMul(FrRawElement r, const FrRawElement a, const FrRawElement b);
void Fr_rawMSquare(FrRawElement r, const FrRawElement a);
void Fr_rawToMontgomery(FrRawElement r, const FrRawElement a);
void Fr_rawFromMontgomery(FrRawElement r, const FrRawElement a);

static void load_elem(FrElement *e, const uint8_t *le32, int form) {
    char hex[65];
    for (int i = 0; i < 32; i++) sprintf(hex + 2 * i, "%02x", le32[31 - i]);
    hex[64] = 0;
    Fr_str2element(e, hex, 16);
    if (form & 2) {
        // negative-short representation when the value is within 2^31 of q
        FrElement t, m;
        Fr_neg(&t, e);                  // t = q - v (short if small)
        if (!(t.type & Fr_LONG) && t.shortVal > 0) {
            m.type = Fr_SHORT;
            m.shortVal = -t.shortVal;
            memset(m.longVal, 0, sizeof(m.longVal));
            *e = m;
        }
    }
    if (form & 1) Fr_toMontgomery(e, e);
}

static void store_elem(uint8_t *le32, FrElement *e) {
    FrElement t;
    Fr_toLongNormal(&t, e);
    memcpy(le32, t.longVal, 32);
}

extern "C" {

// returns 0 on success, -1 unknown op
int frref_op(const char *op, const uint8_t *a, int af, const uint8_t *b, int bf, uint8_t *out) {
    FrElement ea, eb, r;
    memset(&r, 0, sizeof(r));
    load_elem(&ea, a, af);
    if (b) load_elem(&eb, b, bf);
#define BIN(name, fn) if (!strcmp(op, name)) { fn(&r, &ea, &eb); store_elem(out, &r); return 0; }
#define UN(name, fn)  if (!strcmp(op, name)) { fn(&r, &ea); store_elem(out, &r); return 0; }
    BIN("add", Fr_add) BIN("sub", Fr_sub) BIN("mul", Fr_mul) BIN("div", Fr_div)
    BIN("idiv", Fr_idiv) BIN("mod", Fr_mod) BIN("pow", Fr_pow)
    BIN("shl", Fr_shl) BIN("shr", Fr_shr) BIN("band", Fr_band) BIN("bor", Fr_bor) BIN("bxor", Fr_bxor)
    BIN("eq", Fr_eq) BIN("neq", Fr_neq) BIN("lt", Fr_lt) BIN("gt", Fr_gt) BIN("leq", Fr_leq) BIN("geq", Fr_geq)
    BIN("land", Fr_land) BIN("lor", Fr_lor)
    UN("neg", Fr_neg) UN("bnot", Fr_bnot) UN("lnot", Fr_lnot) UN("inv", Fr_inv) UN("square", Fr_square)
    UN("copy", Fr_copy)
#undef BIN
#undef UN
    return -1;
}

int frref_isTrue(const uint8_t *a, int af) {
    FrElement ea;
    load_elem(&ea, a, af);
    return Fr_isTrue(&ea);
}

// Fr_toInt asserts on overflow in the reference; the caller pre-filters to the valid domain.
int frref_toInt(const uint8_t *a, int af) {
    FrElement ea;
    load_elem(&ea, a, af);
    return Fr_toInt(&ea);
}

// raw Montgomery primitives on 4x64 LE limbs (R = 2^256)
void frref_rawMMul(const uint8_t *a, const uint8_t *b, uint8_t *out) {
    FrRawElement ra, rb, rr;
    memcpy(ra, a, 32); memcpy(rb, b, 32);
    Fr_rawMMul(rr, ra, rb);
    memcpy(out, rr, 32);
}
void frref_rawToMontgomery(const uint8_t *a, uint8_t *out) {
    FrRawElement ra, rr;
    memcpy(ra, a, 32);
    Fr_rawToMontgomery(rr, ra);
    memcpy(out, rr, 32);
}
void frref_rawFromMontgomery(const uint8_t *a, uint8_t *out) {
    FrRawElement ra, rr;
    memcpy(ra, a, 32);
    Fr_rawFromMontgomery(rr, ra);
    memcpy(out, rr, 32);
}
void frref_q(uint8_t *out) { memcpy(out, Fr_q.longVal, 32); }

}  // extern "C"
