// BN254 scalar field (Fr) arithmetic on 8x32-bit limbs, Montgomery form with R = 2^256.
//
// Replaces, for the GPU path, the reference's x86-64 field library
//   code_producers/src/c_elements/bn128/fr.asm   (Fr_rawMMul :365, rawMSquare :533, add/sub/neg :1110-2300,
//                                                  band/bor/bxor/bnot :2728-6415, shr/shl :7069-7420,
//                                                  comparisons :7421-8380, constants :8779-8793)
//   code_producers/src/c_elements/generic/fr.cpp  (the readable twin; value semantics in SURVEY.md App. B)
// Same R as the reference, so constants are in the same Montgomery form as its .dat files
// (c_code_generator.rs:560-612).
//
// Values on the device are ALWAYS long+Montgomery (no short/long tag dispatch); operations that look
// at the integer (comparisons, shifts, bit ops, idiv/mod, pow exponent, toInt) leave Montgomery form
// first, exactly like generic/fr.cpp does.
//
// The header is also compilable by a host C++ compiler (FR_HD expands to nothing) so that the
// limb-level algorithms can be unit-tested on CPU against the reference-generated KATs.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FR_HD __host__ __device__ __forceinline__
#define FR_D __device__ __forceinline__
#else
#define FR_HD inline
#define FR_D inline
#endif

namespace fr {

struct Fr {
    uint32_t v[8];
};

// q, little-endian 32-bit limbs (program_structure/src/utils/constants.rs:3-4)
#define FR_Q0 0xf0000001u
#define FR_Q1 0x43e1f593u
#define FR_Q2 0x79b97091u
#define FR_Q3 0x2833e848u
#define FR_Q4 0x8181585du
#define FR_Q5 0xb85045b6u
#define FR_Q6 0xe131a029u
#define FR_Q7 0x30644e72u
#define FR_NP0 0xefffffffu  // -q^-1 mod 2^32 (low word of bn128/fr.asm:8793 np)

FR_HD uint32_t qlimb(int i) {
    switch (i) {
        case 0: return FR_Q0; case 1: return FR_Q1; case 2: return FR_Q2; case 3: return FR_Q3;
        case 4: return FR_Q4; case 5: return FR_Q5; case 6: return FR_Q6; default: return FR_Q7;
    }
}

// R mod q   (Montgomery form of 1)
FR_HD Fr one_mont() {
    Fr r = {{0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u, 0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u}};
    return r;
}
// q - (R mod q)   (Montgomery form of -1)
FR_HD Fr minus_one_mont() {
    Fr r = {{0xa0000006u, 0x974bc177u, 0xda58a367u, 0xf13771b2u, 0x0908122eu, 0x51e1a247u, 0x4729c0fau, 0x2259d6b1u}};
    return r;
}
// R^2 mod q  (bn128/fr.asm:8789 R2)
FR_HD Fr r2_mont() {
    Fr r = {{0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u, 0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u}};
    return r;
}
FR_HD Fr zero() {
    Fr r = {{0, 0, 0, 0, 0, 0, 0, 0}};
    return r;
}
FR_HD Fr modulus() {
    Fr r = {{FR_Q0, FR_Q1, FR_Q2, FR_Q3, FR_Q4, FR_Q5, FR_Q6, FR_Q7}};
    return r;
}
// (q-1)/2  (bn128/fr.asm:8787 half)
FR_HD Fr half_q() {
    Fr r = {{0xf8000000u, 0xa1f0fac9u, 0x3cdcb848u, 0x9419f424u, 0x40c0ac2eu, 0xdc2822dbu, 0x7098d014u, 0x18322739u}};
    return r;
}

FR_HD bool is_zero(const Fr &a) {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) o |= a.v[i];
    return o == 0;
}
FR_HD bool equal(const Fr &a, const Fr &b) {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) o |= a.v[i] ^ b.v[i];
    return o == 0;
}
// a >= b on raw 256-bit integers
FR_HD bool geq_raw(const Fr &a, const Fr &b) {
    uint64_t borrow = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t d = (uint64_t)a.v[i] - b.v[i] - borrow;
        borrow = (d >> 32) & 1;
    }
    return borrow == 0;
}
FR_HD bool gt_raw(const Fr &a, const Fr &b) { return !geq_raw(b, a); }

// r = a - b (raw), returns borrow
FR_HD uint32_t sub_raw(Fr &r, const Fr &a, const Fr &b) {
    uint64_t borrow = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t d = (uint64_t)a.v[i] - b.v[i] - borrow;
        r.v[i] = (uint32_t)d;
        borrow = (d >> 32) & 1;
    }
    return (uint32_t)borrow;
}
// r = a + b (raw), returns carry
FR_HD uint32_t add_raw(Fr &r, const Fr &a, const Fr &b) {
    uint64_t carry = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t s = (uint64_t)a.v[i] + b.v[i] + carry;
        r.v[i] = (uint32_t)s;
        carry = s >> 32;
    }
    return (uint32_t)carry;
}
#if defined(__CUDA_ARCH__)
// Device formulations: one carry chain per 256-bit add/subtract (IADD3.X), the borrow materialised as a mask.
// r = a - b, returns 0xffffffff when the subtraction borrowed, else 0
__device__ __forceinline__ uint32_t sub_mask(Fr &r, const Fr &a, const Fr &b) {
    uint32_t m;
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]), "=r"(m)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
    return m;
}
// r = t - q, returns the borrow mask
__device__ __forceinline__ uint32_t sub_q_mask(Fr &r, const Fr &t) {
    uint32_t m;
    asm("sub.cc.u32 %0, %9, 0xf0000001;\n\t"
        "subc.cc.u32 %1, %10, 0x43e1f593;\n\t"
        "subc.cc.u32 %2, %11, 0x79b97091;\n\t"
        "subc.cc.u32 %3, %12, 0x2833e848;\n\t"
        "subc.cc.u32 %4, %13, 0x8181585d;\n\t"
        "subc.cc.u32 %5, %14, 0xb85045b6;\n\t"
        "subc.cc.u32 %6, %15, 0xe131a029;\n\t"
        "subc.cc.u32 %7, %16, 0x30644e72;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]), "=r"(m)
        : "r"(t.v[0]), "r"(t.v[1]), "r"(t.v[2]), "r"(t.v[3]), "r"(t.v[4]), "r"(t.v[5]), "r"(t.v[6]), "r"(t.v[7]));
    return m;
}
#endif

// branch-free: r = (t >= q) ? t - q : t     (one conditional subtraction)
FR_HD Fr reduce_once(const Fr &t) {
    Fr d, r;
#if defined(__CUDA_ARCH__)
    const bool keep = sub_q_mask(d, t) != 0;     // t < q
#else
    const bool keep = sub_raw(d, t, modulus()) != 0;
#endif
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = keep ? t.v[i] : d.v[i];
    return r;
}

// ---- field add / sub / neg  (generic/fr.cpp:19-98 rawAdd/rawSub/rawNeg; valid in either representation)
FR_HD Fr add(const Fr &a, const Fr &b) {
    Fr t;
#if defined(__CUDA_ARCH__)
    asm("add.cc.u32 %0, %8, %16;\n\t"
        "addc.cc.u32 %1, %9, %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32 %7, %15, %23;"
        : "=r"(t.v[0]), "=r"(t.v[1]), "=r"(t.v[2]), "=r"(t.v[3]), "=r"(t.v[4]), "=r"(t.v[5]), "=r"(t.v[6]), "=r"(t.v[7])
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
#else
    add_raw(t, a, b);
#endif
    return reduce_once(t);       // a, b < q < 2^254: no carry out of 256 bits
}
FR_HD Fr sub(const Fr &a, const Fr &b) {
    Fr d, r;
#if defined(__CUDA_ARCH__)
    const uint32_t m = sub_mask(d, a, b);        // borrowed: add q back
    asm("add.cc.u32 %0, %8, %16;\n\t"
        "addc.cc.u32 %1, %9, %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32 %7, %15, %23;"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
        : "r"(d.v[0]), "r"(d.v[1]), "r"(d.v[2]), "r"(d.v[3]), "r"(d.v[4]), "r"(d.v[5]), "r"(d.v[6]), "r"(d.v[7]),
          "r"(FR_Q0 & m), "r"(FR_Q1 & m), "r"(FR_Q2 & m), "r"(FR_Q3 & m), "r"(FR_Q4 & m), "r"(FR_Q5 & m), "r"(FR_Q6 & m), "r"(FR_Q7 & m));
#else
    Fr e;
    uint32_t borrow = sub_raw(d, a, b);
    add_raw(e, d, modulus());
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = borrow ? e.v[i] : d.v[i];
#endif
    return r;
}
FR_HD Fr neg(const Fr &a) {
    Fr d;
    sub_raw(d, modulus(), a);
    bool z = is_zero(a);
    Fr r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = z ? 0u : d.v[i];
    return r;
}

// ---- Montgomery multiplication, CIOS on 8x32 limbs  (generic/fr.cpp:110-164 Fr_rawMMul; bn128/fr.asm:365)
// 8*8 (a*b) + 8*8 (m*q) + 8 (m) = 136 32x32->64 multiply-accumulates.
// Portable formulation: 64-bit row accumulators; nvcc lowers each step to IMAD.WIDE.U32 + carry adds.
FR_HD Fr mont_mul_portable(const Fr &a, const Fr &b) {
    uint32_t t[10];
#pragma unroll
    for (int i = 0; i < 10; i++) t[i] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t carry = 0;
        uint32_t bi = b.v[i];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            uint64_t p = (uint64_t)a.v[j] * bi + t[j] + carry;
            t[j] = (uint32_t)p;
            carry = p >> 32;
        }
        uint64_t s = (uint64_t)t[8] + carry;
        t[8] = (uint32_t)s;
        t[9] = (uint32_t)(s >> 32);
        uint32_t m = t[0] * FR_NP0;
        uint64_t p = (uint64_t)m * FR_Q0 + t[0];
        carry = p >> 32;
#pragma unroll
        for (int j = 1; j < 8; j++) {
            p = (uint64_t)m * qlimb(j) + t[j] + carry;
            t[j - 1] = (uint32_t)p;
            carry = p >> 32;
        }
        s = (uint64_t)t[8] + carry;
        t[7] = (uint32_t)s;
        t[8] = t[9] + (uint32_t)(s >> 32);
    }
    Fr r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = t[i];
    // result < 2q and q < 2^254 so t[8] == 0 here; one conditional subtraction finishes
    return reduce_once(r);
}

#if defined(__CUDACC__)
// First device formulation of the round, kept as the comparison variant of cvmgpu_mul_peak (the kernels use
// mont_mul_chain below).  Per b-limb ("row"): the eight 32x32->64 products a[j]*b_i come from mul.wide.u32
// (IMAD.WIDE.U32 with a zero addend: no register-pair set-up moves).  The products of the EVEN limbs occupy
// disjoint 64-bit windows (columns 2k, 2k+1), so they form one 256-bit number that is added to T with a single
// add.cc/addc.cc chain; the products of the ODD limbs form a second number shifted by 32 bits.  Hence
// 8 IMAD.WIDE + 17 carry-chain adds per 8 multiply-accumulates, with the multiplies on the FMA pipe and the
// adds on the ALU pipe.  The reduction row (m*q) has the same shape with q as immediates.
__device__ __forceinline__ void mac_row(uint32_t *t, const uint32_t *x, uint32_t y) {
    asm(
        "{\n\t"
        ".reg .u64 e0, e1, e2, e3, o0, o1, o2, o3;\n\t"
        ".reg .u32 el0, eh0, el1, eh1, el2, eh2, el3, eh3, ol0, oh0, ol1, oh1, ol2, oh2, ol3, oh3;\n\t"
        "mul.wide.u32 e0, %9, %17;\n\t"
        "mul.wide.u32 o0, %10, %17;\n\t"
        "mul.wide.u32 e1, %11, %17;\n\t"
        "mul.wide.u32 o1, %12, %17;\n\t"
        "mul.wide.u32 e2, %13, %17;\n\t"
        "mul.wide.u32 o2, %14, %17;\n\t"
        "mul.wide.u32 e3, %15, %17;\n\t"
        "mul.wide.u32 o3, %16, %17;\n\t"
        "mov.b64 {el0, eh0}, e0;\n\t"
        "mov.b64 {el1, eh1}, e1;\n\t"
        "mov.b64 {el2, eh2}, e2;\n\t"
        "mov.b64 {el3, eh3}, e3;\n\t"
        "mov.b64 {ol0, oh0}, o0;\n\t"
        "mov.b64 {ol1, oh1}, o1;\n\t"
        "mov.b64 {ol2, oh2}, o2;\n\t"
        "mov.b64 {ol3, oh3}, o3;\n\t"
        "add.cc.u32 %0, %0, el0;\n\t"
        "addc.cc.u32 %1, %1, eh0;\n\t"
        "addc.cc.u32 %2, %2, el1;\n\t"
        "addc.cc.u32 %3, %3, eh1;\n\t"
        "addc.cc.u32 %4, %4, el2;\n\t"
        "addc.cc.u32 %5, %5, eh2;\n\t"
        "addc.cc.u32 %6, %6, el3;\n\t"
        "addc.cc.u32 %7, %7, eh3;\n\t"
        "addc.u32 %8, %8, 0;\n\t"
        "add.cc.u32 %1, %1, ol0;\n\t"
        "addc.cc.u32 %2, %2, oh0;\n\t"
        "addc.cc.u32 %3, %3, ol1;\n\t"
        "addc.cc.u32 %4, %4, oh1;\n\t"
        "addc.cc.u32 %5, %5, ol2;\n\t"
        "addc.cc.u32 %6, %6, oh2;\n\t"
        "addc.cc.u32 %7, %7, ol3;\n\t"
        "addc.u32 %8, %8, oh3;\n\t"
        "}"
        : "+r"(t[0]), "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6]), "+r"(t[7]), "+r"(t[8])
        : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(x[4]), "r"(x[5]), "r"(x[6]), "r"(x[7]), "r"(y));
}
// T += m * q with q as immediates
__device__ __forceinline__ void mac_row_q(uint32_t *t, uint32_t m) {
    asm(
        "{\n\t"
        ".reg .u64 e0, e1, e2, e3, o0, o1, o2, o3;\n\t"
        ".reg .u32 el0, eh0, el1, eh1, el2, eh2, el3, eh3, ol0, oh0, ol1, oh1, ol2, oh2, ol3, oh3;\n\t"
        "mul.wide.u32 e0, %9, 0xf0000001;\n\t"
        "mul.wide.u32 o0, %9, 0x43e1f593;\n\t"
        "mul.wide.u32 e1, %9, 0x79b97091;\n\t"
        "mul.wide.u32 o1, %9, 0x2833e848;\n\t"
        "mul.wide.u32 e2, %9, 0x8181585d;\n\t"
        "mul.wide.u32 o2, %9, 0xb85045b6;\n\t"
        "mul.wide.u32 e3, %9, 0xe131a029;\n\t"
        "mul.wide.u32 o3, %9, 0x30644e72;\n\t"
        "mov.b64 {el0, eh0}, e0;\n\t"
        "mov.b64 {el1, eh1}, e1;\n\t"
        "mov.b64 {el2, eh2}, e2;\n\t"
        "mov.b64 {el3, eh3}, e3;\n\t"
        "mov.b64 {ol0, oh0}, o0;\n\t"
        "mov.b64 {ol1, oh1}, o1;\n\t"
        "mov.b64 {ol2, oh2}, o2;\n\t"
        "mov.b64 {ol3, oh3}, o3;\n\t"
        "add.cc.u32 %0, %0, el0;\n\t"
        "addc.cc.u32 %1, %1, eh0;\n\t"
        "addc.cc.u32 %2, %2, el1;\n\t"
        "addc.cc.u32 %3, %3, eh1;\n\t"
        "addc.cc.u32 %4, %4, el2;\n\t"
        "addc.cc.u32 %5, %5, eh2;\n\t"
        "addc.cc.u32 %6, %6, el3;\n\t"
        "addc.cc.u32 %7, %7, eh3;\n\t"
        "addc.u32 %8, %8, 0;\n\t"
        "add.cc.u32 %1, %1, ol0;\n\t"
        "addc.cc.u32 %2, %2, oh0;\n\t"
        "addc.cc.u32 %3, %3, ol1;\n\t"
        "addc.cc.u32 %4, %4, oh1;\n\t"
        "addc.cc.u32 %5, %5, ol2;\n\t"
        "addc.cc.u32 %6, %6, oh2;\n\t"
        "addc.cc.u32 %7, %7, ol3;\n\t"
        "addc.u32 %8, %8, oh3;\n\t"
        "}"
        : "+r"(t[0]), "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6]), "+r"(t[7]), "+r"(t[8])
        : "r"(m));
}

__device__ __forceinline__ Fr mont_mul_wide(const Fr &a, const Fr &b) {
    uint32_t t[9];
#pragma unroll
    for (int i = 0; i < 9; i++) t[i] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        mac_row(t, a.v, b.v[i]);            // T += a * b_i          (T < 2^287)
        uint32_t m = t[0] * FR_NP0;
        mac_row_q(t, m);                    // T += m * q, T[0] == 0  (T < 2^288)
#pragma unroll
        for (int j = 0; j < 8; j++) t[j] = t[j + 1];   // T >>= 32 (register renaming after unrolling)
        t[8] = 0;
    }
    Fr r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = t[i];
    return reduce_once(r);                  // T < 2q
}

// ---- carry-chained multiply-accumulate rows ----------------------------------------------------------------
// ptxas fuses a  mad.lo.cc r[2k], x, y, r[2k] ; madc.hi.cc r[2k+1], x, y, r[2k+1]  pair into ONE
// IMAD.WIDE.U32(.X) whose carry travels in a predicate: a 32x32->64 multiply-accumulate with carry-in and
// carry-out is a single instruction when the 64-bit window is a register pair.  The products of limbs 0,2,4,6 of
// an operand tile eight consecutive words, those of limbs 1,3,5,7 tile the eight words one position higher, so a
// number is kept as TWO word arrays, T = E + 2^32 * O, and each row of a product is two 4-instruction chains.
// (mont_mul_wide above - mul.wide + add chains on one array - needs 407 instructions per product; this form 210.)

// acc[0..7] += (x0, x1, x2, x3) * y, x_k * y landing on acc[2k], acc[2k+1]; the carry out is counted in top
__device__ __forceinline__ void mad4(uint32_t *acc, uint32_t &top, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t y) {
    asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]), "+r"(top)
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(y));
}
// the same when the sum is known to fit the eight words
__device__ __forceinline__ void mad4_fit(uint32_t *acc, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t y) {
    asm("mad.lo.cc.u32 %0, %8, %12, %0;\n\t"
        "madc.hi.cc.u32 %1, %8, %12, %1;\n\t"
        "madc.lo.cc.u32 %2, %9, %12, %2;\n\t"
        "madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
        "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
        "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
        "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
        "madc.hi.u32 %7, %11, %12, %7;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7])
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(y));
}
// low += pend (one word below acc's alignment); acc[0..7] += (x0..x3) * y + that carry; the sum fits
__device__ __forceinline__ void mad4_pend(uint32_t *acc, uint32_t &low, uint32_t pend, uint32_t x0, uint32_t x1, uint32_t x2,
                                          uint32_t x3, uint32_t y) {
    asm("add.cc.u32 %8, %8, %9;\n\t"
        "madc.lo.cc.u32 %0, %10, %14, %0;\n\t"
        "madc.hi.cc.u32 %1, %10, %14, %1;\n\t"
        "madc.lo.cc.u32 %2, %11, %14, %2;\n\t"
        "madc.hi.cc.u32 %3, %11, %14, %3;\n\t"
        "madc.lo.cc.u32 %4, %12, %14, %4;\n\t"
        "madc.hi.cc.u32 %5, %12, %14, %5;\n\t"
        "madc.lo.cc.u32 %6, %13, %14, %6;\n\t"
        "madc.hi.u32 %7, %13, %14, %7;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]), "+r"(low)
        : "r"(pend), "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(y));
}
// acc[0..7] += (q0, q2, q4, q6) * m, carry out counted in top
__device__ __forceinline__ void mad4_q_even(uint32_t *acc, uint32_t &top, uint32_t m) {
    asm("mad.lo.cc.u32 %0, %9, 0xf0000001, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, 0xf0000001, %1;\n\t"
        "madc.lo.cc.u32 %2, %9, 0x79b97091, %2;\n\t"
        "madc.hi.cc.u32 %3, %9, 0x79b97091, %3;\n\t"
        "madc.lo.cc.u32 %4, %9, 0x8181585d, %4;\n\t"
        "madc.hi.cc.u32 %5, %9, 0x8181585d, %5;\n\t"
        "madc.lo.cc.u32 %6, %9, 0xe131a029, %6;\n\t"
        "madc.hi.cc.u32 %7, %9, 0xe131a029, %7;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]), "+r"(top)
        : "r"(m));
}
// acc[0..7] += (q1, q3, q5, q7) * m; the sum fits
__device__ __forceinline__ void mad4_q_odd(uint32_t *acc, uint32_t m) {
    asm("mad.lo.cc.u32 %0, %8, 0x43e1f593, %0;\n\t"
        "madc.hi.cc.u32 %1, %8, 0x43e1f593, %1;\n\t"
        "madc.lo.cc.u32 %2, %8, 0x2833e848, %2;\n\t"
        "madc.hi.cc.u32 %3, %8, 0x2833e848, %3;\n\t"
        "madc.lo.cc.u32 %4, %8, 0xb85045b6, %4;\n\t"
        "madc.hi.cc.u32 %5, %8, 0xb85045b6, %5;\n\t"
        "madc.lo.cc.u32 %6, %8, 0x30644e72, %6;\n\t"
        "madc.hi.u32 %7, %8, 0x30644e72, %7;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7])
        : "r"(m));
}
// low += pend; m = low * (-1/q); acc[0..7] += (q1, q3, q5, q7) * m + that carry (mul.lo leaves the carry flag alone)
__device__ __forceinline__ uint32_t mad4_q_odd_pend(uint32_t *acc, uint32_t &low, uint32_t pend) {
    uint32_t m;
    asm("add.cc.u32 %8, %8, %10;\n\t"
        "mul.lo.u32 %9, %8, 0xefffffff;\n\t"
        "madc.lo.cc.u32 %0, %9, 0x43e1f593, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, 0x43e1f593, %1;\n\t"
        "madc.lo.cc.u32 %2, %9, 0x2833e848, %2;\n\t"
        "madc.hi.cc.u32 %3, %9, 0x2833e848, %3;\n\t"
        "madc.lo.cc.u32 %4, %9, 0xb85045b6, %4;\n\t"
        "madc.hi.cc.u32 %5, %9, 0xb85045b6, %5;\n\t"
        "madc.lo.cc.u32 %6, %9, 0x30644e72, %6;\n\t"
        "madc.hi.u32 %7, %9, 0x30644e72, %7;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]), "+r"(low),
          "=&r"(m)
        : "r"(pend));
    return m;
}

// Running value of a word-serial Montgomery pass: T = X + pend + 2^32 * Y, X aligned with word 0 (nine words),
// Y with word 1 (eight words).  One pass step adds a row (for a product), makes word 0 vanish with m * q and
// divides by 2^32.  T < 2^288 throughout (T < 2q before a step, the two rows add < 2^33 q), hence Y < 2^256
// never carries out and X needs one overflow word.  The division is a renaming: the new X is Y, the new Y is
// X[2..8], and X[1] - one word below the new Y - becomes "pend": it is added to the new X[0] at the start of the
// next step, and that carry has exactly the weight of the new Y's first word, so it enters Y's next chain.
struct MontAcc {
    uint32_t x[9], y[8], pend;
};
__device__ __forceinline__ void mont_acc_zero(MontAcc &t) {
#pragma unroll
    for (int i = 0; i < 9; i++) t.x[i] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) t.y[i] = 0;
    t.pend = 0;
}
__device__ __forceinline__ void mont_acc_shift(MontAcc &t) {
    const uint32_t p = t.x[1];
    uint32_t ny[8];
#pragma unroll
    for (int i = 0; i < 7; i++) ny[i] = t.x[i + 2];
    ny[7] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) t.x[i] = t.y[i];
    t.x[8] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) t.y[i] = ny[i];
    t.pend = p;
}
// T = (T + a * s + m * q) / 2^32
__device__ __forceinline__ void mont_acc_step(MontAcc &t, const uint32_t *a, uint32_t s) {
    mad4_pend(t.y, t.x[0], t.pend, a[1], a[3], a[5], a[7], s);
    mad4(t.x, t.x[8], a[0], a[2], a[4], a[6], s);
    const uint32_t m = t.x[0] * FR_NP0;
    mad4_q_even(t.x, t.x[8], m);
    mad4_q_odd(t.y, m);
    mont_acc_shift(t);
}
// T = (T + m * q) / 2^32
__device__ __forceinline__ void mont_acc_redc_step(MontAcc &t) {
    const uint32_t m = mad4_q_odd_pend(t.y, t.x[0], t.pend);
    mad4_q_even(t.x, t.x[8], m);
    mont_acc_shift(t);
}
// the value as eight words (it is < 2^256 when the caller's bound says so)
__device__ __forceinline__ Fr mont_acc_value(const MontAcc &t) {
    Fr r;
    asm("add.cc.u32 %0, %8, %16;\n\t"
        "addc.cc.u32 %1, %9, %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32 %7, %15, %23;"
        : "=&r"(r.v[0]), "=&r"(r.v[1]), "=&r"(r.v[2]), "=&r"(r.v[3]), "=&r"(r.v[4]), "=&r"(r.v[5]), "=&r"(r.v[6]), "=&r"(r.v[7])
        : "r"(t.x[0]), "r"(t.x[1]), "r"(t.x[2]), "r"(t.x[3]), "r"(t.x[4]), "r"(t.x[5]), "r"(t.x[6]), "r"(t.x[7]),
          "r"(t.pend), "r"(t.y[0]), "r"(t.y[1]), "r"(t.y[2]), "r"(t.y[3]), "r"(t.y[4]), "r"(t.y[5]), "r"(t.y[6]));
    return r;
}

// a * b / 2^256 mod q: 128 IMAD.WIDE + 8 IMAD + ~50 adds
__device__ __forceinline__ Fr mont_mul_chain(const Fr &a, const Fr &b) {
    MontAcc t;
    mont_acc_zero(t);
#pragma unroll
    for (int i = 0; i < 8; i++) mont_acc_step(t, a.v, b.v[i]);
    return reduce_once(mont_acc_value(t));   // < 2q
}

// ---- squaring: 36 + 64 multiply-accumulates instead of 128 --------------------------------------------------------
// a^2 = sum_i a_i 2^(32 i) * (a_i 2^(32 i) + 2 * sum_{j > i} a_j 2^(32 j)): row i of the word-serial pass multiplies a_i
// by the words (a_i, 2 a_{i+1} mod 2^32, limbs i+2.. of the number 2a) at positions i..7 of the running value, so the rows
// get shorter instead of repeating the symmetric products.  The running value is bounded by 3.01 q before a row (< 2^288
// after it), so MontAcc's word counts hold; the final value (a^2 + M q) / 2^256 is < 2q like a product's.
// >>> GENERATED by tools/gen_sqr_rows.py (one asm statement per carry chain; do not edit by hand)
__device__ __forceinline__ void sqr_row_0(MontAcc &t, const uint32_t *a, const uint32_t *d) {
    asm("add.cc.u32 %8, %8, %9;\n\t"
        "madc.lo.cc.u32 %0, %10, %14, %0;\n\t"
        "madc.hi.cc.u32 %1, %10, %14, %1;\n\t"
        "madc.lo.cc.u32 %2, %11, %14, %2;\n\t"
        "madc.hi.cc.u32 %3, %11, %14, %3;\n\t"
        "madc.lo.cc.u32 %4, %12, %14, %4;\n\t"
        "madc.hi.cc.u32 %5, %12, %14, %5;\n\t"
        "madc.lo.cc.u32 %6, %13, %14, %6;\n\t"
        "madc.hi.u32 %7, %13, %14, %7;"
        : "+r"(t.y[0]), "+r"(t.y[1]), "+r"(t.y[2]), "+r"(t.y[3]), "+r"(t.y[4]), "+r"(t.y[5]), "+r"(t.y[6]), "+r"(t.y[7]), "+r"(t.x[0])
        : "r"(t.pend), "r"((a[1] << 1)), "r"(d[3]), "r"(d[5]), "r"(d[7]), "r"(a[0]));
    asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(t.x[0]), "+r"(t.x[1]), "+r"(t.x[2]), "+r"(t.x[3]), "+r"(t.x[4]), "+r"(t.x[5]), "+r"(t.x[6]), "+r"(t.x[7]), "+r"(t.x[8])
        : "r"(a[0]), "r"(d[2]), "r"(d[4]), "r"(d[6]), "r"(a[0]));
}
__device__ __forceinline__ void sqr_row_1(MontAcc &t, const uint32_t *a, const uint32_t *d) {
    asm("add.cc.u32 %8, %8, %9;\n\t"
        "madc.lo.cc.u32 %0, %10, %14, %0;\n\t"
        "madc.hi.cc.u32 %1, %10, %14, %1;\n\t"
        "madc.lo.cc.u32 %2, %11, %14, %2;\n\t"
        "madc.hi.cc.u32 %3, %11, %14, %3;\n\t"
        "madc.lo.cc.u32 %4, %12, %14, %4;\n\t"
        "madc.hi.cc.u32 %5, %12, %14, %5;\n\t"
        "madc.lo.cc.u32 %6, %13, %14, %6;\n\t"
        "madc.hi.u32 %7, %13, %14, %7;"
        : "+r"(t.y[0]), "+r"(t.y[1]), "+r"(t.y[2]), "+r"(t.y[3]), "+r"(t.y[4]), "+r"(t.y[5]), "+r"(t.y[6]), "+r"(t.y[7]), "+r"(t.x[0])
        : "r"(t.pend), "r"(a[1]), "r"(d[3]), "r"(d[5]), "r"(d[7]), "r"(a[1]));
    asm("mad.lo.cc.u32 %2, %9, %12, %2;\n\t"
        "madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
        "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
        "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
        "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
        "madc.hi.cc.u32 %7, %11, %12, %7;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(t.x[0]), "+r"(t.x[1]), "+r"(t.x[2]), "+r"(t.x[3]), "+r"(t.x[4]), "+r"(t.x[5]), "+r"(t.x[6]), "+r"(t.x[7]), "+r"(t.x[8])
        : "r"((a[2] << 1)), "r"(d[4]), "r"(d[6]), "r"(a[1]));
}
__device__ __forceinline__ void sqr_row_2(MontAcc &t, const uint32_t *a, const uint32_t *d) {
    asm("add.cc.u32 %8, %8, %9;\n\t"
        "addc.cc.u32 %0, %0, 0;\n\t"
        "addc.cc.u32 %1, %1, 0;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
        "madc.hi.u32 %7, %12, %13, %7;"
        : "+r"(t.y[0]), "+r"(t.y[1]), "+r"(t.y[2]), "+r"(t.y[3]), "+r"(t.y[4]), "+r"(t.y[5]), "+r"(t.y[6]), "+r"(t.y[7]), "+r"(t.x[0])
        : "r"(t.pend), "r"((a[3] << 1)), "r"(d[5]), "r"(d[7]), "r"(a[2]));
    asm("mad.lo.cc.u32 %2, %9, %12, %2;\n\t"
        "madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
        "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
        "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
        "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
        "madc.hi.cc.u32 %7, %11, %12, %7;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(t.x[0]), "+r"(t.x[1]), "+r"(t.x[2]), "+r"(t.x[3]), "+r"(t.x[4]), "+r"(t.x[5]), "+r"(t.x[6]), "+r"(t.x[7]), "+r"(t.x[8])
        : "r"(a[2]), "r"(d[4]), "r"(d[6]), "r"(a[2]));
}
__device__ __forceinline__ void sqr_row_3(MontAcc &t, const uint32_t *a, const uint32_t *d) {
    asm("add.cc.u32 %8, %8, %9;\n\t"
        "addc.cc.u32 %0, %0, 0;\n\t"
        "addc.cc.u32 %1, %1, 0;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
        "madc.hi.u32 %7, %12, %13, %7;"
        : "+r"(t.y[0]), "+r"(t.y[1]), "+r"(t.y[2]), "+r"(t.y[3]), "+r"(t.y[4]), "+r"(t.y[5]), "+r"(t.y[6]), "+r"(t.y[7]), "+r"(t.x[0])
        : "r"(t.pend), "r"(a[3]), "r"(d[5]), "r"(d[7]), "r"(a[3]));
    asm("mad.lo.cc.u32 %4, %9, %11, %4;\n\t"
        "madc.hi.cc.u32 %5, %9, %11, %5;\n\t"
        "madc.lo.cc.u32 %6, %10, %11, %6;\n\t"
        "madc.hi.cc.u32 %7, %10, %11, %7;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(t.x[0]), "+r"(t.x[1]), "+r"(t.x[2]), "+r"(t.x[3]), "+r"(t.x[4]), "+r"(t.x[5]), "+r"(t.x[6]), "+r"(t.x[7]), "+r"(t.x[8])
        : "r"((a[4] << 1)), "r"(d[6]), "r"(a[3]));
}
__device__ __forceinline__ void sqr_row_4(MontAcc &t, const uint32_t *a, const uint32_t *d) {
    asm("add.cc.u32 %8, %8, %9;\n\t"
        "addc.cc.u32 %0, %0, 0;\n\t"
        "addc.cc.u32 %1, %1, 0;\n\t"
        "addc.cc.u32 %2, %2, 0;\n\t"
        "addc.cc.u32 %3, %3, 0;\n\t"
        "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
        "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
        "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
        "madc.hi.u32 %7, %11, %12, %7;"
        : "+r"(t.y[0]), "+r"(t.y[1]), "+r"(t.y[2]), "+r"(t.y[3]), "+r"(t.y[4]), "+r"(t.y[5]), "+r"(t.y[6]), "+r"(t.y[7]), "+r"(t.x[0])
        : "r"(t.pend), "r"((a[5] << 1)), "r"(d[7]), "r"(a[4]));
    asm("mad.lo.cc.u32 %4, %9, %11, %4;\n\t"
        "madc.hi.cc.u32 %5, %9, %11, %5;\n\t"
        "madc.lo.cc.u32 %6, %10, %11, %6;\n\t"
        "madc.hi.cc.u32 %7, %10, %11, %7;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(t.x[0]), "+r"(t.x[1]), "+r"(t.x[2]), "+r"(t.x[3]), "+r"(t.x[4]), "+r"(t.x[5]), "+r"(t.x[6]), "+r"(t.x[7]), "+r"(t.x[8])
        : "r"(a[4]), "r"(d[6]), "r"(a[4]));
}
__device__ __forceinline__ void sqr_row_5(MontAcc &t, const uint32_t *a, const uint32_t *d) {
    asm("add.cc.u32 %8, %8, %9;\n\t"
        "addc.cc.u32 %0, %0, 0;\n\t"
        "addc.cc.u32 %1, %1, 0;\n\t"
        "addc.cc.u32 %2, %2, 0;\n\t"
        "addc.cc.u32 %3, %3, 0;\n\t"
        "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
        "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
        "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
        "madc.hi.u32 %7, %11, %12, %7;"
        : "+r"(t.y[0]), "+r"(t.y[1]), "+r"(t.y[2]), "+r"(t.y[3]), "+r"(t.y[4]), "+r"(t.y[5]), "+r"(t.y[6]), "+r"(t.y[7]), "+r"(t.x[0])
        : "r"(t.pend), "r"(a[5]), "r"(d[7]), "r"(a[5]));
    asm("mad.lo.cc.u32 %6, %9, %10, %6;\n\t"
        "madc.hi.cc.u32 %7, %9, %10, %7;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(t.x[0]), "+r"(t.x[1]), "+r"(t.x[2]), "+r"(t.x[3]), "+r"(t.x[4]), "+r"(t.x[5]), "+r"(t.x[6]), "+r"(t.x[7]), "+r"(t.x[8])
        : "r"((a[6] << 1)), "r"(a[5]));
}
__device__ __forceinline__ void sqr_row_6(MontAcc &t, const uint32_t *a, const uint32_t *d) {
    asm("add.cc.u32 %8, %8, %9;\n\t"
        "addc.cc.u32 %0, %0, 0;\n\t"
        "addc.cc.u32 %1, %1, 0;\n\t"
        "addc.cc.u32 %2, %2, 0;\n\t"
        "addc.cc.u32 %3, %3, 0;\n\t"
        "addc.cc.u32 %4, %4, 0;\n\t"
        "addc.cc.u32 %5, %5, 0;\n\t"
        "madc.lo.cc.u32 %6, %10, %11, %6;\n\t"
        "madc.hi.u32 %7, %10, %11, %7;"
        : "+r"(t.y[0]), "+r"(t.y[1]), "+r"(t.y[2]), "+r"(t.y[3]), "+r"(t.y[4]), "+r"(t.y[5]), "+r"(t.y[6]), "+r"(t.y[7]), "+r"(t.x[0])
        : "r"(t.pend), "r"((a[7] << 1)), "r"(a[6]));
    asm("mad.lo.cc.u32 %6, %9, %10, %6;\n\t"
        "madc.hi.cc.u32 %7, %9, %10, %7;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(t.x[0]), "+r"(t.x[1]), "+r"(t.x[2]), "+r"(t.x[3]), "+r"(t.x[4]), "+r"(t.x[5]), "+r"(t.x[6]), "+r"(t.x[7]), "+r"(t.x[8])
        : "r"(a[6]), "r"(a[6]));
}
__device__ __forceinline__ void sqr_row_7(MontAcc &t, const uint32_t *a, const uint32_t *d) {
    asm("add.cc.u32 %8, %8, %9;\n\t"
        "addc.cc.u32 %0, %0, 0;\n\t"
        "addc.cc.u32 %1, %1, 0;\n\t"
        "addc.cc.u32 %2, %2, 0;\n\t"
        "addc.cc.u32 %3, %3, 0;\n\t"
        "addc.cc.u32 %4, %4, 0;\n\t"
        "addc.cc.u32 %5, %5, 0;\n\t"
        "madc.lo.cc.u32 %6, %10, %11, %6;\n\t"
        "madc.hi.u32 %7, %10, %11, %7;"
        : "+r"(t.y[0]), "+r"(t.y[1]), "+r"(t.y[2]), "+r"(t.y[3]), "+r"(t.y[4]), "+r"(t.y[5]), "+r"(t.y[6]), "+r"(t.y[7]), "+r"(t.x[0])
        : "r"(t.pend), "r"(a[7]), "r"(a[7]));
}
// <<< GENERATED
__device__ __forceinline__ Fr mont_sqr_chain(const Fr &a) {
    uint32_t d[8];   // limbs of 2a (a < 2^254: no ninth limb)
    d[0] = a.v[0] << 1;
#pragma unroll
    for (int j = 1; j < 8; j++) d[j] = __funnelshift_l(a.v[j - 1], a.v[j], 1);
    MontAcc t;
    mont_acc_zero(t);
#define FR_SQR_STEP(i)                                  \
    sqr_row_##i(t, a.v, d);                             \
    {                                                   \
        const uint32_t m = t.x[0] * FR_NP0;             \
        mad4_q_even(t.x, t.x[8], m);                    \
        mad4_q_odd(t.y, m);                             \
        mont_acc_shift(t);                              \
    }
    FR_SQR_STEP(0) FR_SQR_STEP(1) FR_SQR_STEP(2) FR_SQR_STEP(3) FR_SQR_STEP(4) FR_SQR_STEP(5) FR_SQR_STEP(6) FR_SQR_STEP(7)
#undef FR_SQR_STEP
    return reduce_once(mont_acc_value(t));
}

// ---- lazy-reduction dot products ---------------------------------------------------------------------
// sum_i c_i * v_i (all Montgomery) with ONE Montgomery reduction per <= 16 terms: each term costs the 64
// multiply-accumulates of the plain 8x8 product instead of the 136 of a full Montgomery multiplication.
// Accumulator T = E + 2^32 * O + counters: E holds the partial products c_j * v_i with i + j even (words 0..15), O those
// with i + j odd (words 1..14).  A row chain covers eight words of its array; its carry out goes to a counter of that
// window (ce[k]: word 2k + 8, co[k]: word 2k + 9), folded in before the reduction.
// acc[0..3] (four 64-bit windows = eight words) += (x0, x1, x2, x3) * y, carry out counted in top.  The accumulator is
// held as 64-bit registers so that it stays in aligned register pairs across loop iterations: with 32-bit variables
// ptxas re-packs the pairs around every IMAD.WIDE of a loop-carried accumulator (one IMAD.MOV per word and term).
__device__ __forceinline__ void mad4_pairs(uint64_t *acc, uint32_t &top, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t y) {
    asm("{\n\t"
        ".reg .u32 l0, h0, l1, h1, l2, h2, l3, h3;\n\t"
        "mov.b64 {l0, h0}, %0;\n\t"
        "mov.b64 {l1, h1}, %1;\n\t"
        "mov.b64 {l2, h2}, %2;\n\t"
        "mov.b64 {l3, h3}, %3;\n\t"
        "mad.lo.cc.u32 l0, %5, %9, l0;\n\t"
        "madc.hi.cc.u32 h0, %5, %9, h0;\n\t"
        "madc.lo.cc.u32 l1, %6, %9, l1;\n\t"
        "madc.hi.cc.u32 h1, %6, %9, h1;\n\t"
        "madc.lo.cc.u32 l2, %7, %9, l2;\n\t"
        "madc.hi.cc.u32 h2, %7, %9, h2;\n\t"
        "madc.lo.cc.u32 l3, %8, %9, l3;\n\t"
        "madc.hi.cc.u32 h3, %8, %9, h3;\n\t"
        "addc.u32 %4, %4, 0;\n\t"
        "mov.b64 %0, {l0, h0};\n\t"
        "mov.b64 %1, {l1, h1};\n\t"
        "mov.b64 %2, {l2, h2};\n\t"
        "mov.b64 %3, {l3, h3};\n\t"
        "}"
        : "+l"(acc[0]), "+l"(acc[1]), "+l"(acc[2]), "+l"(acc[3]), "+r"(top)
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(y));
}
__device__ __forceinline__ uint32_t lo32(uint64_t v) { return (uint32_t)v; }
__device__ __forceinline__ uint32_t hi32(uint64_t v) { return (uint32_t)(v >> 32); }

struct Wide {
    uint64_t e[8], o[7];   // e[k] = words 2k, 2k+1;  o[k] = words 2k+1, 2k+2
    uint32_t ce[5], co[4];
};
__device__ __forceinline__ void wide_zero(Wide &T) {
#pragma unroll
    for (int i = 0; i < 8; i++) T.e[i] = 0;
#pragma unroll
    for (int i = 0; i < 7; i++) T.o[i] = 0;
#pragma unroll
    for (int i = 0; i < 5; i++) T.ce[i] = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) T.co[i] = 0;
}
// T += c * v   (c, v < q; at most 16 terms between reductions so that T < 2^4 * q^2 < 2^512)
__device__ __forceinline__ void wide_mac(Wide &T, const Fr &c, const Fr &v) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
        // row 2i: even limbs of c land on words 2i.., odd limbs on words 2i + 1..
        mad4_pairs(T.e + i, T.ce[i], c.v[0], c.v[2], c.v[4], c.v[6], v.v[2 * i]);
        mad4_pairs(T.o + i, T.co[i], c.v[1], c.v[3], c.v[5], c.v[7], v.v[2 * i]);
        // row 2i + 1: even limbs on words 2i + 1.. (O), odd limbs on words 2i + 2.. (E)
        mad4_pairs(T.o + i, T.co[i], c.v[0], c.v[2], c.v[4], c.v[6], v.v[2 * i + 1]);
        mad4_pairs(T.e + i + 1, T.ce[i + 1], c.v[1], c.v[3], c.v[5], c.v[7], v.v[2 * i + 1]);
    }
}
// Montgomery reduction of T = sum of n products (n <= 16): (T + m*q) / 2^256 < q * (n * q / 2^256 + 1)
// with q / 2^256 < 0.19, so ceil(0.19 n) conditional subtractions finish (1 for n <= 5).
__device__ __forceinline__ Fr wide_reduce(Wide &T, uint32_t n_terms) {
    // plain words: w[0..16] = E + (O << 32) + counters
    uint32_t w[17];
    w[0] = lo32(T.e[0]);
    asm("add.cc.u32 %0, %15, %30;\n\t"
        "addc.cc.u32 %1, %16, %31;\n\t"
        "addc.cc.u32 %2, %17, %32;\n\t"
        "addc.cc.u32 %3, %18, %33;\n\t"
        "addc.cc.u32 %4, %19, %34;\n\t"
        "addc.cc.u32 %5, %20, %35;\n\t"
        "addc.cc.u32 %6, %21, %36;\n\t"
        "addc.cc.u32 %7, %22, %37;\n\t"
        "addc.cc.u32 %8, %23, %38;\n\t"
        "addc.cc.u32 %9, %24, %39;\n\t"
        "addc.cc.u32 %10, %25, %40;\n\t"
        "addc.cc.u32 %11, %26, %41;\n\t"
        "addc.cc.u32 %12, %27, %42;\n\t"
        "addc.cc.u32 %13, %28, %43;\n\t"
        "addc.u32 %14, %29, 0;"
        : "=&r"(w[1]), "=&r"(w[2]), "=&r"(w[3]), "=&r"(w[4]), "=&r"(w[5]), "=&r"(w[6]), "=&r"(w[7]), "=&r"(w[8]), "=&r"(w[9]),
          "=&r"(w[10]), "=&r"(w[11]), "=&r"(w[12]), "=&r"(w[13]), "=&r"(w[14]), "=&r"(w[15])
        : "r"(hi32(T.e[0])), "r"(lo32(T.e[1])), "r"(hi32(T.e[1])), "r"(lo32(T.e[2])), "r"(hi32(T.e[2])), "r"(lo32(T.e[3])),
          "r"(hi32(T.e[3])), "r"(lo32(T.e[4])), "r"(hi32(T.e[4])), "r"(lo32(T.e[5])), "r"(hi32(T.e[5])), "r"(lo32(T.e[6])),
          "r"(hi32(T.e[6])), "r"(lo32(T.e[7])), "r"(hi32(T.e[7])),
          "r"(lo32(T.o[0])), "r"(hi32(T.o[0])), "r"(lo32(T.o[1])), "r"(hi32(T.o[1])), "r"(lo32(T.o[2])), "r"(hi32(T.o[2])),
          "r"(lo32(T.o[3])), "r"(hi32(T.o[3])), "r"(lo32(T.o[4])), "r"(hi32(T.o[4])), "r"(lo32(T.o[5])), "r"(hi32(T.o[5])),
          "r"(lo32(T.o[6])), "r"(hi32(T.o[6])));
    // (the carry out of word 15 is impossible here: E + (O << 32) alone is below the full sum < 2^512)
    asm("add.cc.u32 %0, %0, %9;\n\t"
        "addc.cc.u32 %1, %1, %10;\n\t"
        "addc.cc.u32 %2, %2, %11;\n\t"
        "addc.cc.u32 %3, %3, %12;\n\t"
        "addc.cc.u32 %4, %4, %13;\n\t"
        "addc.cc.u32 %5, %5, %14;\n\t"
        "addc.cc.u32 %6, %6, %15;\n\t"
        "addc.cc.u32 %7, %7, %16;\n\t"
        "addc.u32 %8, %17, 0;"
        : "+r"(w[8]), "+r"(w[9]), "+r"(w[10]), "+r"(w[11]), "+r"(w[12]), "+r"(w[13]), "+r"(w[14]), "+r"(w[15]), "=r"(w[16])
        : "r"(T.ce[0]), "r"(T.co[0]), "r"(T.ce[1]), "r"(T.co[1]), "r"(T.ce[2]), "r"(T.co[2]), "r"(T.ce[3]), "r"(T.co[3]), "r"(T.ce[4]));
    // word-serial reduction of the low half; the high half is added to what is left
    MontAcc t;
#pragma unroll
    for (int i = 0; i < 8; i++) t.x[i] = w[i];
    t.x[8] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) t.y[i] = 0;
    t.pend = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) mont_acc_redc_step(t);
    Fr r = mont_acc_value(t);   // (low + m*q) / 2^256 <= q
    asm("add.cc.u32 %0, %0, %8;\n\t"
        "addc.cc.u32 %1, %1, %9;\n\t"
        "addc.cc.u32 %2, %2, %10;\n\t"
        "addc.cc.u32 %3, %3, %11;\n\t"
        "addc.cc.u32 %4, %4, %12;\n\t"
        "addc.cc.u32 %5, %5, %13;\n\t"
        "addc.cc.u32 %6, %6, %14;\n\t"
        "addc.u32 %7, %7, %15;"
        : "+r"(r.v[0]), "+r"(r.v[1]), "+r"(r.v[2]), "+r"(r.v[3]), "+r"(r.v[4]), "+r"(r.v[5]), "+r"(r.v[6]), "+r"(r.v[7])
        : "r"(w[8]), "r"(w[9]), "r"(w[10]), "r"(w[11]), "r"(w[12]), "r"(w[13]), "r"(w[14]), "r"(w[15]));
    // w[16] == 0 and no carry: the value is < 4.03 q < 2^256
    r = reduce_once(r);
    if (n_terms > 5) {
        r = reduce_once(r);
        if (n_terms > 10) {
            r = reduce_once(r);
            r = reduce_once(r);
        }
    }
    return r;
}
// ---- small-scalar sums -----------------------------------------------------------------------------------
// X = sum_k m_k * v_k with 32-bit scalars m_k is a plain integer (10 limbs, < 2^320), kept as E + 2^32 * O + counters
// like Wide; one word-serial Montgomery pass with the constant 2^320 mod q brings it back:
// sum_i X_i * C * 2^(32 i) / 2^320 = X (mod q), result < 2q.
struct Small {
    uint64_t e[4], o[4];
    uint32_t ce, co;
};
__device__ __forceinline__ void small_zero(Small &X) {
#pragma unroll
    for (int i = 0; i < 4; i++) X.e[i] = X.o[i] = 0;
    X.ce = X.co = 0;
}
__device__ __forceinline__ void small_mac(Small &X, uint32_t m, const Fr &v) {
    mad4_pairs(X.e, X.ce, v.v[0], v.v[2], v.v[4], v.v[6], m);
    mad4_pairs(X.o, X.co, v.v[1], v.v[3], v.v[5], v.v[7], m);
}
__device__ __forceinline__ Fr small_reduce(const Small &X) {
    const uint32_t C[8] = {0x7c5fb586u, 0xb4c6edf9u, 0xbfeb93beu, 0x708c8d50u, 0x04f7e0efu, 0x9ffd1de4u, 0x9a392866u, 0x215b02acu};
    uint32_t w[10];
    w[0] = lo32(X.e[0]);
    asm("add.cc.u32 %0, %9, %16;\n\t"
        "addc.cc.u32 %1, %10, %17;\n\t"
        "addc.cc.u32 %2, %11, %18;\n\t"
        "addc.cc.u32 %3, %12, %19;\n\t"
        "addc.cc.u32 %4, %13, %20;\n\t"
        "addc.cc.u32 %5, %14, %21;\n\t"
        "addc.cc.u32 %6, %15, %22;\n\t"
        "addc.cc.u32 %7, %24, %23;\n\t"
        "addc.u32 %8, %25, 0;"
        : "=&r"(w[1]), "=&r"(w[2]), "=&r"(w[3]), "=&r"(w[4]), "=&r"(w[5]), "=&r"(w[6]), "=&r"(w[7]), "=&r"(w[8]), "=&r"(w[9])
        : "r"(hi32(X.e[0])), "r"(lo32(X.e[1])), "r"(hi32(X.e[1])), "r"(lo32(X.e[2])), "r"(hi32(X.e[2])), "r"(lo32(X.e[3])),
          "r"(hi32(X.e[3])),
          "r"(lo32(X.o[0])), "r"(hi32(X.o[0])), "r"(lo32(X.o[1])), "r"(hi32(X.o[1])), "r"(lo32(X.o[2])), "r"(hi32(X.o[2])),
          "r"(lo32(X.o[3])), "r"(hi32(X.o[3])),
          "r"(X.ce), "r"(X.co));
    MontAcc t;
    mont_acc_zero(t);
#pragma unroll
    for (int i = 0; i < 10; i++) mont_acc_step(t, C, w[i]);
    return reduce_once(mont_acc_value(t));
}
#endif

FR_HD Fr mont_mul(const Fr &a, const Fr &b) {
#if defined(__CUDA_ARCH__) && !defined(FR_PORTABLE_MUL)
    return mont_mul_chain(a, b);
#else
    return mont_mul_portable(a, b);
#endif
}
// (the reference's rawMSquare just multiplies, fr.cpp:166-169; the result is the same number)
FR_HD Fr mont_sqr(const Fr &a) {
#if defined(__CUDA_ARCH__) && !defined(FR_PORTABLE_MUL)
    return mont_sqr_chain(a);
#else
    return mont_mul(a, a);
#endif
}

// canonical <-> Montgomery (generic/fr.cpp:211-255)
FR_HD Fr to_mont(const Fr &a) { return mont_mul(a, r2_mont()); }
// leaving Montgomery form is a pure reduction: a / 2^256 mod q = eight word-serial REDC steps (72 multiply-accumulates
// instead of the 136 of a product by 1); (a + m*q) / 2^256 <= q, so one conditional subtraction finishes
FR_HD Fr from_mont(const Fr &a) {
#if defined(__CUDA_ARCH__) && !defined(FR_PORTABLE_MUL)
    MontAcc t;
#pragma unroll
    for (int i = 0; i < 8; i++) t.x[i] = a.v[i];
    t.x[8] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) t.y[i] = 0;
    t.pend = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) mont_acc_redc_step(t);
    return reduce_once(mont_acc_value(t));
#else
    Fr one = {{1, 0, 0, 0, 0, 0, 0, 0}};
    return mont_mul(a, one);
#endif
}

// a^e for a fixed public exponent given as 8 limbs (square-and-multiply, MSB first); a in Montgomery form
FR_HD Fr mont_pow(const Fr &a, const Fr &e) {
    Fr r = one_mont();
    bool started = false;
    for (int i = 7; i >= 0; i--) {
        for (int bit = 31; bit >= 0; bit--) {
            if (started) r = mont_sqr(r);
            if ((e.v[i] >> bit) & 1) {
                r = started ? mont_mul(r, a) : a;
                started = true;
            }
        }
    }
    return r;
}
// inverse by Fermat: a^(q-2); inverse of 0 is 0 (kept as the cross-check of mont_inv below)
FR_HD Fr mont_inv_fermat(const Fr &a) {
    Fr e = modulus();
    e.v[0] -= 2;  // q-2 (no borrow: low limb is 0xf0000001)
    return mont_pow(a, e);
}

// ---- modular inversion by Bernstein-Yang "safegcd" division steps ---------------------------------------
// (D. J. Bernstein, B.-Y. Yang, "Fast constant-time gcd computation and modular inversion", TCHES 2019; the
// 30-bit-limb formulation widely used for 256-bit fields.)  Replaces the reference's mpz_invert
// (bn128/fr.cpp:146-157).  20 rounds of 30 division steps on the low words of (f, g), each round followed by
// the application of its 2x2 transition matrix to (f, g) exactly and to (d, e) modulo q: ~16 K simple integer
// instructions with NO data-dependent control flow (every witness of a warp runs the same stream), against ~160 K
// for the 381 Montgomery multiplications of a^(q-2).  inv(0) = 0 falls out of the algorithm (g stays 0, d stays 0),
// which is also what the reference returns.
struct S30 {
    int32_t v[9];   // signed 30-bit limbs
};
// 30 division steps on the low words of (f, g); t = the transition matrix (u v; q r), scaled by 2^30.  The matrix is
// tracked in three runs of 10 steps with its rows PACKED two entries to a word (P = u + v * 2^16, Q = q + r * 2^16,
// |entries| <= 2^10 within a run): conditional negation, masked addition and doubling are linear, so they act on both
// entries at once - 20 instead of 29 operations per step - and the three 2x2 matrices are multiplied together at the end
// of each run (the inversions are 59 % of the EdDSA verifier tape's instructions, profiles/r01_summary.md).
FR_HD int32_t by_divsteps_30(int32_t zeta, uint32_t f0, uint32_t g0, int32_t *t) {
    uint32_t f = f0, g = g0;
    int32_t m0 = 1, m1 = 0, m2 = 0, m3 = 1;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int part = 0; part < 3; ++part) {
        uint32_t P = 1u, Q = 1u << 16;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int i = 0; i < 10; ++i) {
            uint32_t c1 = (uint32_t)(zeta >> 31);      // zeta < 0
            const uint32_t c2 = 0u - (g & 1u);         // g odd
            const uint32_t x = (f ^ c1) - c1;          // conditionally negated f and (u, v)
            const uint32_t y = (P ^ c1) - c1;
            g += x & c2;
            Q += y & c2;
            c1 &= c2;
            zeta = (int32_t)((uint32_t)zeta ^ c1) - 1;
            f += g & c1;
            P += Q & c1;
            g >>= 1;
            P <<= 1;
        }
        const int32_t u = (int32_t)(P << 16) >> 16, v = (int32_t)(P - (uint32_t)u) >> 16;
        const int32_t q = (int32_t)(Q << 16) >> 16, r = (int32_t)(Q - (uint32_t)q) >> 16;
        // (u v; q r) * (m0 m1; m2 m3): entries stay within 2^30
        const int32_t n0 = u * m0 + v * m2, n1 = u * m1 + v * m3, n2 = q * m0 + r * m2, n3 = q * m1 + r * m3;
        m0 = n0; m1 = n1; m2 = n2; m3 = n3;
    }
    t[0] = m0; t[1] = m1; t[2] = m2; t[3] = m3;
    return zeta;
}
FR_HD int32_t by_q30(int i) {
    switch (i) {
        case 0: return 0x30000001; case 1: return 0x0f87d64f; case 2: return 0x1b970914; case 3: return 0x0cfa121e;
        case 4: return 0x01585d28; case 5: return 0x0116da06; case 6: return 0x1a029b85; case 7: return 0x139cb84c;
        default: return 0x00003064;
    }
}
// (d, e) <- t * (d, e) / 2^30 mod q, kept in (-2q, q)
FR_HD void by_update_de(S30 &d, S30 &e, const int32_t *t) {
    const int32_t M30 = 0x3fffffff;
    const uint32_t QINV30 = 0x10000001u;   // q^-1 mod 2^30
    const int32_t u = t[0], v = t[1], q = t[2], r = t[3];
    const int32_t sd = d.v[8] >> 31, se = e.v[8] >> 31;
    int32_t md = (u & sd) + (v & se);
    int32_t me = (q & sd) + (r & se);
    int32_t di = d.v[0], ei = e.v[0];
    int64_t cd = (int64_t)u * di + (int64_t)v * ei;
    int64_t ce = (int64_t)q * di + (int64_t)r * ei;
    md -= (int32_t)((QINV30 * (uint32_t)cd + (uint32_t)md) & (uint32_t)M30);
    me -= (int32_t)((QINV30 * (uint32_t)ce + (uint32_t)me) & (uint32_t)M30);
    cd += (int64_t)by_q30(0) * md;
    ce += (int64_t)by_q30(0) * me;
    cd >>= 30;
    ce >>= 30;
#pragma unroll
    for (int i = 1; i < 9; ++i) {
        di = d.v[i];
        ei = e.v[i];
        cd += (int64_t)u * di + (int64_t)v * ei;
        ce += (int64_t)q * di + (int64_t)r * ei;
        cd += (int64_t)by_q30(i) * md;
        ce += (int64_t)by_q30(i) * me;
        d.v[i - 1] = (int32_t)cd & M30;
        cd >>= 30;
        e.v[i - 1] = (int32_t)ce & M30;
        ce >>= 30;
    }
    d.v[8] = (int32_t)cd;
    e.v[8] = (int32_t)ce;
}
// (f, g) <- t * (f, g) / 2^30 (exact)
FR_HD void by_update_fg(S30 &f, S30 &g, const int32_t *t) {
    const int32_t M30 = 0x3fffffff;
    const int32_t u = t[0], v = t[1], q = t[2], r = t[3];
    int32_t fi = f.v[0], gi = g.v[0];
    int64_t cf = (int64_t)u * fi + (int64_t)v * gi;
    int64_t cg = (int64_t)q * fi + (int64_t)r * gi;
    cf >>= 30;
    cg >>= 30;
#pragma unroll
    for (int i = 1; i < 9; ++i) {
        fi = f.v[i];
        gi = g.v[i];
        cf += (int64_t)u * fi + (int64_t)v * gi;
        cg += (int64_t)q * fi + (int64_t)r * gi;
        f.v[i - 1] = (int32_t)cf & M30;
        cf >>= 30;
        g.v[i - 1] = (int32_t)cg & M30;
        cg >>= 30;
    }
    f.v[8] = (int32_t)cf;
    g.v[8] = (int32_t)cg;
}
// raw integer x in [0, q) -> x^-1 mod q in [0, q)   (0 -> 0)
FR_HD Fr inv_raw(const Fr &x) {
    const int32_t M30 = 0x3fffffff;
    S30 d, e, f, g;
#pragma unroll
    for (int i = 0; i < 9; i++) {
        d.v[i] = 0;
        e.v[i] = 0;
        f.v[i] = by_q30(i);
        // bits [30 i, 30 i + 30) of the 256-bit x
        const int lo = (30 * i) >> 5, sh = (30 * i) & 31;
        uint64_t w = x.v[lo];
        if (lo + 1 < 8) w |= (uint64_t)x.v[lo + 1] << 32;
        g.v[i] = (int32_t)((uint32_t)(w >> sh) & (uint32_t)M30);
    }
    e.v[0] = 1;
    int32_t zeta = -1;
    for (int it = 0; it < 20; ++it) {    // 600 division steps (590 suffice for 256-bit inputs)
        int32_t t[4];
        zeta = by_divsteps_30(zeta, (uint32_t)f.v[0], (uint32_t)g.v[0], t);
        by_update_de(d, e, t);
        by_update_fg(f, g, t);
    }
    // f = +-1; d = +-x^-1 in (-2q, q): add q if negative, negate if f < 0, carry, add q again if still negative
    int32_t r[9];
    int32_t cond_add = d.v[8] >> 31;
    const int32_t cond_neg = f.v[8] >> 31;
#pragma unroll
    for (int i = 0; i < 9; i++) {
        r[i] = d.v[i] + (by_q30(i) & cond_add);
        r[i] = (r[i] ^ cond_neg) - cond_neg;
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
        r[i + 1] += r[i] >> 30;
        r[i] &= M30;
    }
    cond_add = r[8] >> 31;
#pragma unroll
    for (int i = 0; i < 9; i++) r[i] += by_q30(i) & cond_add;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        r[i + 1] += r[i] >> 30;
        r[i] &= M30;
    }
    Fr out;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        // bits [32 k, 32 k + 32) from the 30-bit limbs
        const int li = (32 * k) / 30, sh = (32 * k) % 30;
        uint64_t w = (uint64_t)(uint32_t)r[li] | ((uint64_t)(uint32_t)r[li + 1] << 30);
        if (li + 2 < 9) w |= (uint64_t)(uint32_t)r[li + 2] << 60;
        out.v[k] = (uint32_t)(w >> sh);
    }
    return out;
}
// R^3 mod q  (bn128/fr.asm:8791 R3)
FR_HD Fr r3_mont() {
    Fr r = {{0xb4bf0040u, 0x5e94d8e1u, 0x1cfbb6b8u, 0x2a489cbeu, 0xa19fcfedu, 0x893cc664u, 0x7fcc657cu, 0x0cf8594bu}};
    return r;
}
// Montgomery form in, Montgomery form out: the limbs X = a R; X^-1 * R^3 / R = a^-1 R
FR_HD Fr mont_inv(const Fr &a) { return mont_mul(inv_raw(a), r3_mont()); }

// ---- integer-view operations on CANONICAL values ------------------------------
// signed comparison around half = (q-1)/2: v > half means v - q   (generic/fr.cpp:1172-1363)
FR_HD bool is_neg(const Fr &a) { return gt_raw(a, half_q()); }
FR_HD bool lt_signed(const Fr &a, const Fr &b) {
    bool na = is_neg(a), nb = is_neg(b);
    if (na != nb) return na;           // negative < non-negative
    return gt_raw(b, a);               // same sign: order of canonical values
}

// mask to 254 bits then one conditional subtraction (generic/fr.cpp:293-376)
FR_HD Fr mask_reduce(Fr t) {
    t.v[7] &= 0x3fffffffu;
    return reduce_once(t);
}
FR_HD Fr band(const Fr &a, const Fr &b) { Fr t; for (int i = 0; i < 8; i++) t.v[i] = a.v[i] & b.v[i]; return mask_reduce(t); }
FR_HD Fr bor(const Fr &a, const Fr &b) { Fr t; for (int i = 0; i < 8; i++) t.v[i] = a.v[i] | b.v[i]; return mask_reduce(t); }
FR_HD Fr bxor(const Fr &a, const Fr &b) { Fr t; for (int i = 0; i < 8; i++) t.v[i] = a.v[i] ^ b.v[i]; return mask_reduce(t); }
FR_HD Fr bnot(const Fr &a) { Fr t; for (int i = 0; i < 8; i++) t.v[i] = ~a.v[i]; return mask_reduce(t); }

// raw shifts by 0 <= s < 256
FR_HD Fr shl_raw(const Fr &a, uint32_t s) {
    Fr r;
    uint32_t w = s >> 5, b = s & 31;
    for (int i = 7; i >= 0; i--) {
        uint32_t lo = (i - (int)w >= 0) ? a.v[i - w] : 0u;
        uint32_t lo2 = (i - (int)w - 1 >= 0) ? a.v[i - w - 1] : 0u;
        r.v[i] = b ? ((lo << b) | (lo2 >> (32 - b))) : lo;
    }
    return r;
}
FR_HD Fr shr_raw(const Fr &a, uint32_t s) {
    Fr r;
    uint32_t w = s >> 5, b = s & 31;
    for (int i = 0; i < 8; i++) {
        uint32_t lo = (i + w < 8) ? a.v[i + w] : 0u;
        uint32_t hi = (i + w + 1 < 8) ? a.v[i + w + 1] : 0u;
        r.v[i] = b ? ((lo >> b) | (hi << (32 - b))) : lo;
    }
    return r;
}
// does the canonical value fit in [0, 254)?  returns the amount, else 0xffffffff
FR_HD uint32_t small_amount(const Fr &b) {
    uint32_t hi = 0;
    for (int i = 1; i < 8; i++) hi |= b.v[i];
    return (hi == 0 && b.v[0] < 254u) ? b.v[0] : 0xffffffffu;
}
// shl / shr with the reference's "negative shift" rule (generic/fr.cpp:1995-2307)
FR_HD Fr shr(const Fr &a, const Fr &b);
FR_HD Fr shl(const Fr &a, const Fr &b) {
    uint32_t s = small_amount(b);
    if (s != 0xffffffffu) return mask_reduce(shl_raw(a, s));
    Fr nb;
    sub_raw(nb, modulus(), b);          // q - b
    uint32_t t = small_amount(nb);
    if (t == 0xffffffffu) return zero();
    return shr_raw(a, t);
}
FR_HD Fr shr(const Fr &a, const Fr &b) {
    uint32_t s = small_amount(b);
    if (s != 0xffffffffu) return shr_raw(a, s);
    Fr nb;
    sub_raw(nb, modulus(), b);
    uint32_t t = small_amount(nb);
    if (t == 0xffffffffu) return zero();
    return mask_reduce(shl_raw(a, t));
}

// number of significant bits of a raw 256-bit integer
FR_HD int bit_length(const Fr &a) {
    for (int i = 7; i >= 0; i--) {
        if (a.v[i]) {
            int n = 0;
            uint32_t x = a.v[i];
            while (x) { n++; x >>= 1; }
            return 32 * i + n;
        }
    }
    return 0;
}
// floor division and remainder of canonical integers (generic/fr.cpp:2835-2875, mpz_fdiv_q / mpz_fdiv_r);
// shift-subtract long division.  d must be non-zero.
FR_HD void divmod(const Fr &n, const Fr &d, Fr &qt, Fr &rem) {
    qt = zero();
    rem = n;
    int shift = bit_length(n) - bit_length(d);
    if (shift < 0) return;
    Fr ds = shl_raw(d, (uint32_t)shift);
    for (int s = shift; s >= 0; s--) {
        if (geq_raw(rem, ds)) {
            Fr t;
            sub_raw(t, rem, ds);
            rem = t;
            qt.v[s >> 5] |= 1u << (s & 31);
        }
        ds = shr_raw(ds, 1);
    }
}

// a (Montgomery) ^ e (canonical, data-dependent)  (generic/fr.cpp:2877-2893 mpz_powm)
FR_HD Fr mont_pow_var(const Fr &a, const Fr &e) {
    Fr r = one_mont();
    int n = bit_length(e);
    for (int i = n - 1; i >= 0; i--) {
        r = mont_sqr(r);
        if ((e.v[i >> 5] >> (i & 31)) & 1) r = mont_mul(r, a);
    }
    return r;
}

// Fr_toInt domain check on a canonical value (generic/fr.cpp:1102-1170): fits int32 around 0 (mod q)
FR_HD bool to_int(const Fr &a, int32_t &out) {
    uint32_t hi = 0;
    for (int i = 1; i < 8; i++) hi |= a.v[i];
    if (hi == 0 && a.v[0] < 0x80000000u) { out = (int32_t)a.v[0]; return true; }
    Fr d;
    sub_raw(d, modulus(), a);           // q - a
    hi = 0;
    for (int i = 1; i < 8; i++) hi |= d.v[i];
    if (hi == 0 && d.v[0] <= 0x80000000u && d.v[0] != 0) { out = (int32_t)(0u - d.v[0]); return true; }
    return false;
}

}  // namespace fr
