// CUDA kernels (sm_100a): tape interpreter, witness export/import, R1CS check, Fr self-test, IMAD probe.
//
// Data layout in HBM ("value store"): structure-of-arrays at 128-bit granularity,
//     row r (witness wire or spill row), half h in {0,1}, witness w  ->  uint4 at ((r*2+h)*bstride + w)
// so the 32 lanes of a warp (32 consecutive witnesses) read/write 512 contiguous bytes per half with
// LDG.128/STG.128.  Values are in Montgomery form.
//
// On-chip: every witness (thread) owns n_slots 32-byte slots in shared memory, stored as
//     smem[(slot*2+h)*NT + tid]  (uint4)  -- conflict-free LDS.128/STS.128.
// This is the GPU counterpart of the reference's per-call scratch `FrElement expaux[..], lvar[..]`
// on the C stack (template.rs:343-344) and of `signalValues[]` (calcwit.cpp:33).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fr.cuh"
#include "tape.hpp"

namespace kern {

// Address of row r of one witness's column: rows are 2 * bstride uint4 apart; the host guarantees bstride < 2^27 so that the
// byte stride fits 32 bits and the address is ONE IMAD.WIDE.U32 (row * stride + base) instead of a 64-bit multiply.
__device__ __forceinline__ uint4 *row_ptr(const uint4 *wbase, uint32_t row, uint64_t bstride) {
    const uint32_t row_bytes = (uint32_t)bstride * 32u;
    return (uint4 *)((char *)wbase + (uint64_t)row * row_bytes);
}

using fr::Fr;

#define TAPE_PAD 96   // no-op instructions after the last one on the device copy of a tape (fetch / prefetch without bound checks)
#define CVM_NT 128   // threads (= witnesses) per CTA of the tape kernel for large batches (64 / 32 for small ones)

__device__ __forceinline__ Fr unpack(const uint4 &lo, const uint4 &hi) {
    Fr r;
    r.v[0] = lo.x; r.v[1] = lo.y; r.v[2] = lo.z; r.v[3] = lo.w;
    r.v[4] = hi.x; r.v[5] = hi.y; r.v[6] = hi.z; r.v[7] = hi.w;
    return r;
}
__device__ __forceinline__ void pack(const Fr &a, uint4 &lo, uint4 &hi) {
    lo = make_uint4(a.v[0], a.v[1], a.v[2], a.v[3]);
    hi = make_uint4(a.v[4], a.v[5], a.v[6], a.v[7]);
}

struct TapeParams {
    const tape::TapeIns *tape;
    uint32_t n_ins;
    const uint4 *consts;     // Montgomery, 2 x uint4 per constant
    uint4 *store;            // field rows of the value store
    uint32_t *bits;          // bit rows: word (w >> 5) * n_brows + row holds the row's bit of the 32 witnesses of warp w >> 5
    uint64_t bstride;
    uint32_t n_brows;
    uint32_t n_bslots;       // bit slots per warp
    const uint4 *inputs;     // B x n_inputs x 2 uint4, canonical
    uint32_t n_inputs;
    uint32_t *status;
    uint64_t B;
    const unsigned long long *iconsts;   // constants of the integer operations
    uint32_t ring_off;       // uint4 offset of the reload ring inside the dynamic shared memory (after the field slots)
    uint32_t bslot_off;      // uint4 offset of the bit-slot file (after the ring)
    uint32_t *first_bad;     // tapes with a fused R1CS check (T_RNE): first violated constraint per witness, else nullptr
    uint32_t in_row_bytes;   // != 0: the inputs are PACKED BITS, in_row_bytes per witness, input k = bit k & 7 of byte k >> 3
                             // (only tapes that read their inputs with T_INPUT_BIT)
};

__device__ __forceinline__ Fr mont_bool(bool b) { return b ? fr::one_mont() : fr::zero(); }

// canonical input (possibly >= q, like a decimal string fed to Fr_str2element) -> Montgomery
__device__ __forceinline__ Fr op_input(Fr v) {
    for (int k = 0; k < 6; k++) {
        Fr d;
        uint32_t borrow = fr::sub_raw(d, v, fr::modulus());
        if (!borrow) v = d;
    }
    return fr::to_mont(v);
}

// operand fetch: constant table (uniform address, L1-resident), the thread's field slot in shared memory, or -- for a
// value typed 0/1 -- the thread's bit of the warp's word in the bit-slot file (bw = this warp's file)
template <int NT, bool BITS>
__device__ __forceinline__ Fr tape_operand(const uint4 *slots, const uint32_t *bw, const uint4 *consts, uint32_t idx, bool is_const,
                                           uint32_t tid) {
    if (BITS && !is_const && (idx & tape::BSLOT)) return mont_bool((bw[idx & 0xffffu] >> (tid & 31u)) & 1u);
    uint4 lo, hi;
    if (is_const) {
        lo = __ldg(consts + 2 * (uint64_t)idx);
        hi = __ldg(consts + 2 * (uint64_t)idx + 1);
    } else {
        lo = slots[(idx * 2) * NT + tid];
        hi = slots[(idx * 2 + 1) * NT + tid];
    }
    return unpack(lo, hi);
}
// truth value of an operand (select conditions): the bit itself when the value is typed 0/1
template <int NT, bool BITS>
__device__ __forceinline__ bool tape_truth(const uint4 *slots, const uint32_t *bw, const uint4 *consts, uint32_t idx, bool is_const,
                                           uint32_t tid) {
    if (BITS && !is_const && (idx & tape::BSLOT)) return (bw[idx & 0xffffu] >> (tid & 31u)) & 1u;
    return !fr::is_zero(tape_operand<NT, BITS>(slots, bw, consts, idx, is_const, tid));
}

// operand of an integer operation (tape.hpp type_ints): a raw 64-bit integer in the low words of a field slot, a 0/1
// value of the bit file, or an integer constant
template <int NT, bool BITS>
__device__ __forceinline__ unsigned long long tape_int(const uint4 *slots, const uint32_t *bw, const unsigned long long *iconsts,
                                                       uint32_t idx, bool is_const, uint32_t tid) {
    if (is_const) return __ldg(iconsts + idx);
    if (BITS && (idx & tape::BSLOT)) return (bw[idx & 0xffffu] >> (tid & 31u)) & 1u;
    const uint2 v = *reinterpret_cast<const uint2 *>(slots + (idx * 2) * NT + tid);
    return (unsigned long long)v.x | ((unsigned long long)v.y << 32);
}

// integer-view operations, division, logic: shared by the tape's slow path and the device self-test.
// Operations that look at the integer leave Montgomery form first, as generic/fr.cpp does (one conversion site each
// way, to keep the code small).
__device__ __forceinline__ Fr slow_compute(uint32_t op, const Fr &a, const Fr &b, uint32_t &status) {
    const bool logic = (op == tape::T_LAND || op == tape::T_LOR);
    if (op == tape::T_INV) return fr::mont_inv(a);
    if (op == tape::T_DIV) return fr::mont_mul(a, fr::mont_inv(b));
    if (logic) {
        bool x = !fr::is_zero(a), y = !fr::is_zero(b);
        return mont_bool(op == tape::T_LAND ? (x && y) : (x || y));
    }
    const Fr cb = fr::from_mont(b);
    if (op == tape::T_POW) return fr::mont_pow_var(a, cb);
    const Fr ca = fr::from_mont(a);
    Fr t = fr::zero();
    switch (op) {
        case tape::T_IDIV: case tape::T_MOD:
            // a zero divisor leaves 0; the failure itself is raised by the explicit check the tracer emits in front of
            // the operation (tracer.hpp), which knows whether the statement executes at all
            if (!fr::is_zero(cb)) {
                Fr qt, rem;
                fr::divmod(ca, cb, qt, rem);
                t = (op == tape::T_MOD) ? rem : qt;
            }
            break;
        case tape::T_SHL: t = fr::shl(ca, cb); break;
        case tape::T_SHR: t = fr::shr(ca, cb); break;
        case tape::T_BAND: t = fr::band(ca, cb); break;
        case tape::T_BOR: t = fr::bor(ca, cb); break;
        case tape::T_BXOR: t = fr::bxor(ca, cb); break;
        case tape::T_BNOT: t = fr::bnot(ca); break;
        case tape::T_LT: return mont_bool(fr::lt_signed(ca, cb));
        case tape::T_GT: return mont_bool(fr::lt_signed(cb, ca));
        case tape::T_LE: return mont_bool(!fr::lt_signed(cb, ca));
        case tape::T_GE: return mont_bool(!fr::lt_signed(ca, cb));
        default: return fr::zero();
    }
    return fr::to_mont(t);
}

// Everything that is not on the fast path of the tape loop (integer-view operations, division, inputs).  Out of
// line, operands and result go through the slots, so that the hot loop keeps its working set in registers.  A result
// typed 0/1 is returned as `truth` (the caller packs the warp's bits); a field result is written to its slot here.
template <int NT, bool BITS>
__device__ __noinline__ uint32_t tape_slow_op(uint4 cur, uint4 *slots, const uint32_t *bw, const uint4 *consts, const uint4 *inputs,
                                              uint32_t n_inputs, uint64_t w, uint32_t status, uint32_t tid, uint32_t &truth) {
    const uint32_t op = cur.x & 0xffu, flags = (cur.x >> 8) & 0xffu, dst = cur.x >> 16;
    Fr r;
    if (op == tape::T_INPUT) {
        const uint4 *src = inputs + (w * n_inputs + cur.y) * 2;
        r = op_input(unpack(src[0], src[1]));
    } else {
        Fr a = tape_operand<NT, BITS>(slots, bw, consts, cur.y, flags & 1u, tid);
        Fr b = fr::zero();
        if (op != tape::T_BNOT && op != tape::T_INV) b = tape_operand<NT, BITS>(slots, bw, consts, cur.z, flags & 2u, tid);
        r = slow_compute(op, a, b, status);
    }
    truth = r.v[0] != 0u;   // a 0/1 value in Montgomery form is 0 or R mod q, whose low word is not 0
    if (!BITS || !(dst & tape::BSLOT_DST)) {
        uint4 lo, hi;
        pack(r, lo, hi);
        slots[(dst * 2) * NT + tid] = lo;
        slots[(dst * 2 + 1) * NT + tid] = hi;
    }
    return status;
}

// ADD / SUB / MUL of the tape (shared by the dispatch fast path of field-only programs and the switch)
template <int NT, bool BITS>
__device__ __forceinline__ Fr tape_arith(uint32_t op, uint32_t flags, const uint4 &cur, const uint4 *slots, const uint32_t *bw,
                                         const uint4 *consts, uint32_t tid) {
    Fr r;
    const Fr a = tape_operand<NT, BITS>(slots, bw, consts, cur.y, flags & 1u, tid);
    const Fr b = tape_operand<NT, BITS>(slots, bw, consts, cur.z, flags & 2u, tid);
    if (op == tape::T_MUL) {
        // bit-heavy programs (F_TRIVIAL, set by the tape builder): factors that are 0 or 1 at run time without
        // being provably so (input bits of a hash) need no product; decided per warp to keep control flow uniform
        bool cheap = false;
        if (flags & tape::F_TRIVIAL) {
            const Fr one = fr::one_mont();
            const bool triv = fr::is_zero(a) || fr::is_zero(b) || fr::equal(a, one) || fr::equal(b, one);
            cheap = __all_sync(0xffffffffu, triv);
            if (cheap) {
                const bool z = fr::is_zero(a) || fr::is_zero(b);
                const bool a1 = fr::equal(a, one);
#pragma unroll
                for (int i = 0; i < 8; i++) r.v[i] = z ? 0u : (a1 ? b.v[i] : a.v[i]);
            }
        }
        if (!cheap) {
            // same operand twice (x^2, x^4 of an S-box): the squaring needs 100 instead of 128 IMAD.WIDE
            if (cur.y == cur.z && (flags & 3u) == 0) r = fr::mont_sqr(a);
            else r = fr::mont_mul(a, b);
        }
    } else if (op == tape::T_ADD) {
        r = fr::add(a, b);
    } else {
        r = fr::sub(a, b);
    }
    return r;
}

// One tape pass per witness.  Control flow is uniform (one instruction stream per circuit), so the branches on the
// opcode never diverge.  Fast path, inlined once each: MUL / ADD / SUB, DOT, SEL, EQ / NEQ / EQZ, BITC, LUT, the failure
// checks and the value-store moves; the result of a producing instruction can be written to its witness wire by
// the same instruction (flag bit 3).
//
// Typed values (tape.hpp build_tape): a result the trace compiler proved 0/1 goes to the BIT file -- the 32 lanes of a
// warp pack their bits into one word (__ballot_sync), every lane stores that same word to the warp's slot (so each lane
// later reads its own store: no fence needed), and a bit row of the value store is that word per warp: 4 bytes where a
// field row costs 1 KiB.
// (programs without 0/1-typed values are bound by the multiplier pipe and want 20 resident warps per SM: 96 registers;
// the bit-file instantiation gets 128: the register file is split between the four schedulers, so 129..168 registers mean 3
// warps per scheduler = 12 per SM, and a 64 K batch -- 13.8 one-warp CTAs per SM -- no longer fits one wave)
// records that follow an instruction (tape.hpp extra_records) as (a * M + A) >> S; entry = M | A << 8 | S << 16
__constant__ uint32_t c_ext_table[256];
inline void fill_ext_table(uint32_t *t) {
    for (int i = 0; i < 256; i++) t[i] = 0;
    t[tape::T_DOT] = 1u | (1u << 8) | (1u << 16);
    t[tape::T_ISUM] = 1u | (3u << 8) | (2u << 16);
    t[tape::T_ISUMT] = 8u;
    t[tape::T_LUTG] = t[tape::T_IBITG] = t[tape::T_INBITG] = 1u;
}

template <int NT, bool BITS>
__global__ void __launch_bounds__(NT, BITS ? 512 / NT : 640 / NT) tape_kernel(TapeParams p) {
    extern __shared__ uint4 slots[];
    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    uint64_t w = (uint64_t)blockIdx.x * NT + tid;
    const bool active = w < p.B;
    // a warp whose first witness exists owns its word of every bit row (its padding lanes carry copies of the last witness)
    const bool warp_active = (w - lane) < p.B;
    if (!active) w = p.B - 1;   // keep the warp converged; results of padding lanes are discarded
    uint32_t status = 0;
    uint32_t first_bad = 0xffffffffu;
    uint4 *const wbase = p.store + w;
    const uint64_t bstride = p.bstride;
    const uint4 *const consts = p.consts;
    uint32_t *const bw = reinterpret_cast<uint32_t *>(slots + p.bslot_off) + (tid >> 5) * p.n_bslots;
    uint32_t *const brow = p.bits + (warp_active ? (((uint64_t)blockIdx.x * NT + tid) >> 5) : 0ull) * p.n_brows;

    // The tape is padded with TAPE_PAD no-ops on the device (cvmgpu.cu upload_program): the fetch of the next instruction
    // and the prefetch of the lines ahead never need a bound check.
    const uint4 *tp = reinterpret_cast<const uint4 *>(p.tape);
    const uint32_t n_ins = p.n_ins;
    uint4 raw = __ldg(tp);
    uint4 lrec = BITS ? __ldg(tp + 1 + lane) : make_uint4(0u, 0u, 0u, 0u);
    uint32_t pf = 0;   // first tape word (8 per 128-byte line) not yet prefetched into L1
    for (uint32_t pc = 0; pc < n_ins; pc++) {
        const uint4 cur = raw;
        // Bit-file programs are bound by latency, and most of what they run are group instructions whose lane m works on
        // record m: as soon as an instruction is known, the NEXT one (which follows this one's records) and the 32 words
        // after it -- one per lane: its records, if it is a group -- are requested, so that they arrive while this one
        // executes.  myrec = the word after this instruction that belongs to this lane.
        const uint4 myrec = lrec;
        if (BITS) {
            // records behind this instruction: (a * M + A) >> S with per-opcode constants (one packed word from constant memory)
            const uint32_t ek = c_ext_table[cur.x & 0xffu];
            const uint32_t ext = (cur.y * (ek & 0xffu) + ((ek >> 8) & 0xffu)) >> (ek >> 16);
            raw = __ldg(tp + pc + 1 + ext);
            lrec = __ldg(tp + pc + 2 + ext + lane);
        } else
            raw = __ldg(tp + pc + 1);   // the next instruction (re-done after the records of a DOT)
        // tape lines ahead.  With group instructions pc jumps over up to 33 words at a time: the window follows pc.
        if (BITS) {
#pragma unroll 1
            while (pf < pc + 64u) {
                asm volatile("prefetch.global.L1 [%0];" ::"l"(tp + pf));
                pf += 8u;
            }
        } else if ((pc & 7u) == 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(tp + pc + 32));
        const uint32_t op = cur.x & 0xffu;
        const uint32_t flags = (cur.x >> 8) & 0xffu;
        const uint32_t dst = cur.x >> 16;
        Fr r;
        uint32_t rb = 0;        // result of an instruction that produces a truth value
        bool is_rb = false;
        // field-only programs (Poseidon: 240 products + 194 additions + 195 dot products in 1 025 instructions) test for
        // their arithmetic first; everything else goes through the jump table
        if (!BITS && op >= tape::T_ADD && op <= tape::T_MUL) {
            r = tape_arith<NT, BITS>(op, flags, cur, slots, bw, consts, tid);
        } else
        switch (op) {
        case tape::T_ADD: case tape::T_SUB: case tape::T_MUL: {
            r = tape_arith<NT, BITS>(op, flags, cur, slots, bw, consts, tid);
            break;
        }
        case tape::T_DOT: {
            // sum_j c_j * x_j (+ addend): cur.y terms in the following ceil(n/2) records of (constant, slot) pairs
            const uint32_t n = cur.y;
            fr::Wide T;
            fr::wide_zero(T);
            for (uint32_t j = 0; j < n; j++) {
                const uint4 rec = __ldg(tp + pc + 1 + (j >> 1));
                const uint32_t cidx = (j & 1u) ? rec.z : rec.x, slot = (j & 1u) ? rec.w : rec.y;
                const Fr c = unpack(__ldg(consts + 2 * (uint64_t)cidx), __ldg(consts + 2 * (uint64_t)cidx + 1));
                const Fr x = tape_operand<NT, BITS>(slots, bw, consts, slot, false, tid);
                fr::wide_mac(T, c, x);
            }
            r = fr::wide_reduce(T, n);
            if (flags & tape::F_ADDEND) r = fr::add(r, tape_operand<NT, BITS>(slots, bw, consts, cur.z, flags & 2u, tid));
            pc += (n + 1) >> 1;
            if (!BITS) raw = __ldg(tp + pc + 1);
            break;
        }
        case tape::T_SEL: {
            const bool t = tape_truth<NT, BITS>(slots, bw, consts, cur.y, flags & 1u, tid);
            // only the selected operand is fetched; with F_CZERO the "else" value is the constant 0
            const bool isc = t ? (flags & 2u) : (flags & 4u);
            const uint32_t idx = t ? cur.z : cur.w;
            r = fr::zero();
            if (t || !(flags & tape::F_CZERO)) r = tape_operand<NT, BITS>(slots, bw, consts, idx, isc, tid);
            break;
        }
        case tape::T_CADD: {
            // a + (b != 0 ? constant c : 0)
            const Fr a = tape_operand<NT, BITS>(slots, bw, consts, cur.y, flags & 1u, tid);
            const bool t = tape_truth<NT, BITS>(slots, bw, consts, cur.z, false, tid);
            const Fr c = unpack(__ldg(consts + 2 * (uint64_t)cur.w), __ldg(consts + 2 * (uint64_t)cur.w + 1));
            const Fr sum = fr::add(a, c);
#pragma unroll
            for (int i = 0; i < 8; i++) r.v[i] = t ? sum.v[i] : a.v[i];
            break;
        }
        case tape::T_ISUM: {
            // small-integer arithmetic: addend + sum_j (bit_j << shift_j), four (bit slot, shift) terms per record
            const uint32_t n = cur.y;
            unsigned long long v = 0;
            if (flags & tape::F_ADDEND) v = tape_int<NT, BITS>(slots, bw, p.iconsts, cur.z, flags & 2u, tid);
            if (flags & tape::F_ADDEND2) v += tape_int<NT, BITS>(slots, bw, p.iconsts, cur.w, flags & 4u, tid);
            // record 0 is what the loop fetched as "the next instruction"; each iteration fetches the record after the one it
            // works on, and the last one thereby fetches the instruction that follows the records
            uint4 rec = BITS ? __ldg(tp + pc + 1) : raw;
            for (uint32_t j = 0; j < n; j += 4) {
                const uint4 nxt = __ldg(tp + pc + 2 + (j >> 2));
                v += (unsigned long long)((bw[rec.x & 0xffffu] >> lane) & 1u) << (rec.x >> 16);
                if (j + 1 < n) v += (unsigned long long)((bw[rec.y & 0xffffu] >> lane) & 1u) << (rec.y >> 16);
                if (j + 2 < n) v += (unsigned long long)((bw[rec.z & 0xffffu] >> lane) & 1u) << (rec.z >> 16);
                if (j + 3 < n) v += (unsigned long long)((bw[rec.w & 0xffffu] >> lane) & 1u) << (rec.w >> 16);
                rec = nxt;
            }
            slots[(dst * 2) * NT + tid] = make_uint4((uint32_t)v, (uint32_t)(v >> 32), 0u, 0u);
            pc += (n + 3) >> 2;
            if (!BITS) raw = rec;
            continue;
        }
        case tape::T_ISUMT: {
            // T_ISUM in transposed form (tape.hpp): cur.y layers; in a layer lane l fetches the packed word of the term whose shift
            // is base + l (bit w = witness w, 0 if none) and the transpose of that 32 x 32 bit matrix hands every lane the
            // integer sum_l bit_l << l of its own witness.
            const uint32_t nl = cur.y;
            unsigned long long v = 0;
            if (flags & tape::F_ADDEND) v = tape_int<NT, BITS>(slots, bw, p.iconsts, cur.z, flags & 2u, tid);
            if (flags & tape::F_ADDEND2) v += tape_int<NT, BITS>(slots, bw, p.iconsts, cur.w, flags & 4u, tid);
            const uint32_t *words = reinterpret_cast<const uint32_t *>(tp + pc + 1);
            uint32_t e = __ldg(words + lane);
            for (uint32_t L = 0; L < nl; L++) {
                const uint32_t nxt = __ldg(words + (L + 1) * 32 + lane);   // next layer (after the last: the following instructions)
                uint32_t x = (e & 0xffffu) == 0xffffu ? 0u : bw[e & 0xffffu];
                // 32 x 32 bit-matrix transpose across the warp: five exchange steps
#pragma unroll
                for (uint32_t k = 16, m = 0x0000ffffu; k >= 1; k >>= 1, m ^= m << k) {
                    const uint32_t y = __shfl_xor_sync(0xffffffffu, x, k);
                    x = (lane & k) ? ((y >> k) & m) | (x & ~m) : (x & m) | ((y & m) << k);
                }
                v += (unsigned long long)x << (e >> 16);
                e = nxt;
            }
            slots[(dst * 2) * NT + tid] = make_uint4((uint32_t)v, (uint32_t)(v >> 32), 0u, 0u);
            pc += nl * 8;
            if (!BITS) raw = __ldg(tp + pc + 1);
            continue;
        }
        case tape::T_LUTG: case tape::T_IBITG: case tape::T_INBITG: {
            // Warp-cooperative groups (tape.hpp group_bit_ops): cur.y independent operations on values typed 0/1, lane m doing
            // member m for all 32 witnesses of the warp at once on their packed words.  Record m = (operand slots, result
            // slot, truth table, bit row or NO_ROW).  Every lane reads its operands before any lane writes a result.
            const uint32_t n = cur.y;
            uint32_t word = 0, dslot = 0, row = tape::NO_ROW;
            if (op == tape::T_LUTG) {
                if (lane < n) {
                    const uint4 rec = BITS ? myrec : __ldg(tp + pc + 1 + lane);
                    const uint32_t nin = (rec.z >> 8) & 0xffu;
                    const uint32_t x0 = bw[rec.x & 0xffffu];
                    const uint32_t x1 = nin > 1 ? bw[rec.x >> 16] : 0u;
                    const uint32_t x2 = nin > 2 ? bw[rec.y & 0xffffu] : 0u;
                    // multiplexer tree over the truth table, bitwise for the 32 witnesses: m_k = all ones if entry k is set
                    const uint32_t t = rec.z;
                    const uint32_t m0 = 0u - (t & 1u), m1 = 0u - ((t >> 1) & 1u), m2 = 0u - ((t >> 2) & 1u), m3 = 0u - ((t >> 3) & 1u),
                                   m4 = 0u - ((t >> 4) & 1u), m5 = 0u - ((t >> 5) & 1u), m6 = 0u - ((t >> 6) & 1u), m7 = 0u - ((t >> 7) & 1u);
                    const uint32_t a0 = (x0 & m1) | (~x0 & m0), a1 = (x0 & m3) | (~x0 & m2), a2 = (x0 & m5) | (~x0 & m4),
                                   a3 = (x0 & m7) | (~x0 & m6);
                    const uint32_t b0 = (x1 & a1) | (~x1 & a0), b1 = (x1 & a3) | (~x1 & a2);
                    word = (x2 & b1) | (~x2 & b0);
                    dslot = rec.y >> 16;
                    row = rec.w;
                }
            } else if (op == tape::T_INBITG) {
                // main inputs cur.w .. cur.w + n - 1 taken as bits (speculative typing): lane m reads input cur.w + m of each of the
                // warp's 32 witnesses -- for a fixed witness the 32 lanes read 1 KB in a row, and the 32 loads of a lane are
                // independent (one T_INPUT_BIT per input waited for a DRAM access each).  Bit i of lane m's word = witness i.
                const uint64_t w_first = (uint64_t)blockIdx.x * NT + (tid & ~31u);
                uint32_t badw = 0;
                if (lane < n) {
                    const uint32_t k = cur.w + lane;
                    if (p.in_row_bytes) {
                        const unsigned char *base = reinterpret_cast<const unsigned char *>(p.inputs);
#pragma unroll 8
                        for (uint32_t i = 0; i < 32; i++) {
                            const uint64_t wi = min(w_first + i, p.B - 1);
                            word |= (((uint32_t)base[wi * p.in_row_bytes + (k >> 3)] >> (k & 7u)) & 1u) << i;
                        }
                    } else {
#pragma unroll 8
                        for (uint32_t i = 0; i < 32; i++) {
                            const uint64_t wi = min(w_first + i, p.B - 1);
                            const uint4 *src = p.inputs + (wi * p.n_inputs + k) * 2;
                            const uint4 lo = __ldg(src), hi = __ldg(src + 1);
                            const bool bad = lo.x > 1u || (lo.y | lo.z | lo.w | hi.x | hi.y | hi.z | hi.w) != 0u;
                            word |= (lo.x & 1u) << i;
                            badw |= (bad ? 1u : 0u) << i;
                        }
                    }
                    const uint4 rec = BITS ? myrec : __ldg(tp + pc + 1 + lane);
                    dslot = rec.x & 0xffffu;
                    row = rec.w;
                }
                // witness i is flagged when any of the n inputs read for it is not a bit
                if ((__reduce_or_sync(0xffffffffu, badw) >> lane) & 1u) status = tape::ST_SPECULATION;
            } else {
                // bits cur.w .. cur.w + n - 1 of the integer in slot cur.z: the 32 x 32 bit matrix (lane = witness, bit = position)
                // transposed across the warp -- lane m ends up with the word of bit cur.w + m
                const unsigned long long a = tape_int<NT, BITS>(slots, bw, p.iconsts, cur.z, false, tid);
                uint32_t x = cur.w < 64u ? (uint32_t)(a >> cur.w) : 0u;
#pragma unroll
                for (uint32_t k = 16, m = 0x0000ffffu; k >= 1; k >>= 1, m ^= m << k) {
                    const uint32_t y = __shfl_xor_sync(0xffffffffu, x, k);
                    x = (lane & k) ? ((y >> k) & m) | (x & ~m) : (x & m) | ((y & m) << k);
                }
                word = x;
                if (lane < n) {
                    const uint4 rec = BITS ? myrec : __ldg(tp + pc + 1 + lane);
                    dslot = rec.x & 0xffffu;
                    row = rec.w;
                }
            }
            __syncwarp();
            if (lane < n) {
                bw[dslot] = word;
                if (row != tape::NO_ROW && warp_active) brow[row & ~tape::ROW_BIT] = word;
            }
            __syncwarp();
            pc += n;
            if (!BITS) raw = __ldg(tp + pc + 1);
            continue;
        }
        case tape::T_FILL: {
            // constant bit rows [cur.w, cur.w + cur.z) = the word cur.y
            if (warp_active)
                for (uint32_t k = lane; k < cur.z; k += 32u) brow[(cur.w & ~tape::ROW_BIT) + k] = cur.y;
            continue;
        }
        case tape::T_ICADD: case tape::T_IADD: case tape::T_ISEL: {
            // raw 64-bit values (sums of 0/1 values times small constants)
            unsigned long long v;
            if (op == tape::T_ICADD) {
                v = tape_int<NT, BITS>(slots, bw, p.iconsts, cur.y, flags & 1u, tid);
                if (tape_truth<NT, BITS>(slots, bw, consts, cur.z, false, tid)) v += __ldg(p.iconsts + cur.w);
            } else if (op == tape::T_IADD) {
                v = tape_int<NT, BITS>(slots, bw, p.iconsts, cur.y, flags & 1u, tid) +
                    tape_int<NT, BITS>(slots, bw, p.iconsts, cur.z, flags & 2u, tid);
            } else {
                const bool t = tape_truth<NT, BITS>(slots, bw, consts, cur.y, flags & 1u, tid);
                v = tape_int<NT, BITS>(slots, bw, p.iconsts, t ? cur.z : cur.w, t ? (flags & 2u) : (flags & 4u), tid);
            }
            slots[(dst * 2) * NT + tid] = make_uint4((uint32_t)v, (uint32_t)(v >> 32), 0u, 0u);
            continue;
        }
        case tape::T_IBIT: {
            const unsigned long long a = tape_int<NT, BITS>(slots, bw, p.iconsts, cur.y, false, tid);
            rb = cur.z < 64u ? (uint32_t)(a >> cur.z) & 1u : 0u;
            is_rb = true;
            break;
        }
        case tape::T_IFAIL_NE: {
            const unsigned long long a = tape_int<NT, BITS>(slots, bw, p.iconsts, cur.y, flags & 1u, tid);
            const unsigned long long b = tape_int<NT, BITS>(slots, bw, p.iconsts, cur.z, flags & 2u, tid);
            if (a != b && status == 0) status = cur.w;
            continue;
        }
        case tape::T_LUT: {
            // boolean function of up to three typed 0/1 values: this lane's bits of the three words index the table
            const uint32_t nin = (cur.z >> 8) & 0xffu;
            uint32_t idx = (bw[cur.y & 0xffffu] >> lane) & 1u;
            if (nin > 1) idx |= ((bw[cur.y >> 16] >> lane) & 1u) << 1;
            if (nin > 2) idx |= ((bw[cur.z >> 16] >> lane) & 1u) << 2;
            rb = (cur.z >> idx) & 1u;
            is_rb = true;
            break;
        }
        case tape::T_INPUT_BIT: {
            // speculative typing: main input cur.y taken as a bit.  Anything but a literal 0 / 1 marks the witness for the
            // program traced without the assumption (the caller recomputes it there); sticky, nothing else overwrites it.
            if (p.in_row_bytes) {   // packed-bit inputs: nothing to check
                const unsigned char *row = reinterpret_cast<const unsigned char *>(p.inputs) + w * p.in_row_bytes;
                rb = ((uint32_t)row[cur.y >> 3] >> (cur.y & 7u)) & 1u;
                is_rb = true;
                break;
            }
            const uint4 *src = p.inputs + (w * p.n_inputs + cur.y) * 2;
            const uint4 lo = src[0], hi = src[1];
            if (lo.x > 1u || (lo.y | lo.z | lo.w | hi.x | hi.y | hi.z | hi.w) != 0u) status = tape::ST_SPECULATION;
            rb = lo.x & 1u;
            is_rb = true;
            break;
        }
        case tape::T_BITC: {
            // bit cur.z of the raw limbs of slot a: one 32-bit shared-memory read
            const uint32_t *s32 = reinterpret_cast<const uint32_t *>(slots);
            const uint32_t limb = cur.z >> 5;
            const uint32_t word = s32[(((cur.y * 2 + (limb >> 2)) * NT + tid) << 2) + (limb & 3u)];
            rb = (word >> (cur.z & 31u)) & 1u;
            is_rb = true;
            break;
        }
        case tape::T_EQ: case tape::T_NEQ: case tape::T_EQZ: case tape::T_FAIL_IF: case tape::T_FAIL_NE: case tape::T_RNE: {
            bool e;
            if (op == tape::T_EQZ || op == tape::T_FAIL_IF) e = !tape_truth<NT, BITS>(slots, bw, consts, cur.y, flags & 1u, tid);
            else e = fr::equal(tape_operand<NT, BITS>(slots, bw, consts, cur.y, flags & 1u, tid),
                               tape_operand<NT, BITS>(slots, bw, consts, cur.z, flags & 2u, tid));
            if (op == tape::T_FAIL_IF || op == tape::T_FAIL_NE) {
                if (!e && status == 0) status = cur.w;
                continue;
            }
            if (op == tape::T_RNE) {   // constraint cur.w of the fused R1CS check (fused.hpp)
                if (!e) first_bad = min(first_bad, cur.w);
                continue;
            }
            rb = (op == tape::T_NEQ) ? !e : e;
            is_rb = true;
            break;
        }
        case tape::T_LD: {
            if (BITS && (dst & tape::BSLOT_DST)) {     // a bit row: the warp's word (written earlier by these same lanes)
                bw[dst & 0x7fffu] = brow[cur.w & ~tape::ROW_BIT];
                continue;
            }
            // reload stream (tape.hpp schedule_reloads): ring entry cur.z holds this value if F_RING; cur.y is the row
            // to request now for the reload LD_RING reloads ahead.  Every streamed reload commits one cp.async group.
            uint4 *ring = slots + p.ring_off + cur.z * 2 * NT;
            if (flags & tape::F_RING) asm volatile("cp.async.wait_group %0;" ::"n"(tape::LD_RING - 1) : "memory");
            uint4 lo, hi;
            if (flags & tape::F_RING) {
                lo = ring[tid];
                hi = ring[NT + tid];
            } else {
                const uint4 *src = row_ptr(wbase, cur.w, bstride);
                lo = src[0];
                hi = src[bstride];
            }
            slots[(dst * 2) * NT + tid] = lo;
            slots[(dst * 2 + 1) * NT + tid] = hi;
            if (cur.y != tape::NO_ROW) {
                const uint4 *nxt = row_ptr(wbase, cur.y, bstride);
                const uint32_t d0 = (uint32_t)__cvta_generic_to_shared(ring + tid);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d0), "l"(nxt) : "memory");
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d0 + NT * 16), "l"(nxt + bstride) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            continue;
        }
        case tape::T_ST: case tape::T_STC: {
            if (cur.w & tape::ROW_BIT) {
                uint32_t word;
                if (op == tape::T_STC) word = __ldg(consts + 2 * (uint64_t)cur.y).x ? 0xffffffffu : 0u;
                else word = BITS ? bw[cur.y & 0xffffu] : 0u;
                if (warp_active) brow[cur.w & ~tape::ROW_BIT] = word;
                continue;
            }
            uint4 lo, hi;
            if (op == tape::T_STC) { lo = __ldg(consts + 2 * (uint64_t)cur.y); hi = __ldg(consts + 2 * (uint64_t)cur.y + 1); }
            else { lo = slots[(cur.y * 2) * NT + tid]; hi = slots[(cur.y * 2 + 1) * NT + tid]; }
            if (active) {
                uint4 *d = row_ptr(wbase, cur.w, bstride);
                d[0] = lo;
                d[bstride] = hi;
            }
            continue;
        }
        default: {
            status = tape_slow_op<NT, BITS>(cur, slots, bw, consts, p.inputs, p.n_inputs, w, status, tid, rb);
            if (!BITS || !(dst & tape::BSLOT_DST)) {
                if ((flags & tape::F_STORE) && active) {
                    uint4 *d = row_ptr(wbase, cur.w, bstride);
                    d[0] = slots[(dst * 2) * NT + tid];
                    d[bstride] = slots[(dst * 2 + 1) * NT + tid];
                }
                continue;
            }
            is_rb = true;
            break;
        }
        }
        if (BITS && (dst & tape::BSLOT_DST)) {
            const uint32_t word = __ballot_sync(0xffffffffu, is_rb ? rb != 0u : r.v[0] != 0u);
            bw[dst & 0x7fffu] = word;
            if ((flags & tape::F_STORE) && warp_active) brow[cur.w & ~tape::ROW_BIT] = word;
        } else {
            if (is_rb) r = mont_bool(rb);
            if (flags & tape::F_CHECK) {
                // last instruction of a constraint of the fused R1CS check (fused.hpp): dst names the slot that holds the other
                // side of the comparison, cur.w the constraint; nothing is written
                if (!fr::equal(r, tape_operand<NT, BITS>(slots, bw, consts, dst, false, tid))) first_bad = min(first_bad, cur.w);
                continue;
            }
            uint4 lo, hi;
            pack(r, lo, hi);
            slots[(dst * 2) * NT + tid] = lo;
            slots[(dst * 2 + 1) * NT + tid] = hi;
            if ((flags & tape::F_STORE) && active) {
                uint4 *d = row_ptr(wbase, cur.w, bstride);
                d[0] = lo;
                d[bstride] = hi;
            }
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (active && p.status) p.status[w] = status;
    if (active && p.first_bad) p.first_bad[w] = first_bad;
}

// ---- typed value store -> .wtns rows (canonical AoS: B x n_sel x 32 B) ----------------------------------
// Fuses Fr_toLongNormal + the 32-byte write of writeBinWitness (common/main.cpp:324-330) with the
// SoA->AoS transpose: tile of 32 witnesses x 32 wires through shared memory so that both sides coalesce.
// A wire is a field row (Montgomery, left with an 8-step REDC) or a bit row (one word per 32 witnesses, expanded to
// the canonical 0 / 1 here); wire_loc == nullptr means the plain layout (row = wire, all field rows).
// Wires [wire0, wire0 + n_sel) are exported (the whole witness, or e.g. only the public outputs and inputs).
struct StoreView {
    const uint4 *store;
    const uint32_t *bits;
    uint64_t bstride;
    uint32_t n_brows;
    const uint32_t *wire_loc;
};
__global__ void __launch_bounds__(256) export_kernel(StoreView sv, uint64_t B, uint32_t wire0, uint32_t n_sel, uint4 *out) {
    // [wire][witness][limb]: 9 words per element and 33 elements per row make BOTH phases conflict-free -- the write phase
    // strides lanes by 9 words, the read phase (lane = wire) by 33 * 9 = 297 = 9 mod 32, and 9 is coprime with the 32
    // banks.  (With 32 elements per row the read stride was 288 = 0 mod 32: a 32-way bank conflict on every read.)
    __shared__ uint32_t tile[32][33][9];
    const uint32_t lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const uint64_t w0 = (uint64_t)blockIdx.x * 32;
    const uint32_t r0 = blockIdx.y * 32;
    for (uint32_t k = wrp; k < 32; k += 8) {
        uint32_t row = r0 + k;
        uint64_t w = w0 + lane;
        if (row < n_sel && w < B) {
            const uint32_t loc = sv.wire_loc ? __ldg(sv.wire_loc + wire0 + row) : wire0 + row;
            Fr v = fr::zero();
            if (loc & tape::ROW_BIT) {
                v.v[0] = (__ldg(sv.bits + (w0 >> 5) * sv.n_brows + (loc & ~tape::ROW_BIT)) >> lane) & 1u;
            } else {
                const uint4 *src = sv.store + ((uint64_t)loc * 2) * sv.bstride + w;
                v = fr::from_mont(unpack(src[0], src[sv.bstride]));
            }
#pragma unroll
            for (int i = 0; i < 8; i++) tile[k][lane][i] = v.v[i];
        }
    }
    __syncthreads();
    for (uint32_t k = wrp; k < 32; k += 8) {   // k = witness in tile, lane = wire in tile
        uint64_t w = w0 + k;
        uint32_t row = r0 + lane;
        if (row < n_sel && w < B) {
            uint4 *d = out + (w * n_sel + row) * 2;
            d[0] = make_uint4(tile[lane][k][0], tile[lane][k][1], tile[lane][k][2], tile[lane][k][3]);
            d[1] = make_uint4(tile[lane][k][4], tile[lane][k][5], tile[lane][k][6], tile[lane][k][7]);
        }
    }
}

// ---- typed value store -> PACKED rows: per witness [field wires x 32 B canonical, in wire order][0/1 wires, in wire order,
// packed LSB-first into 32-bit words].  flist / blist: the field rows / bit rows of those wires in that order.
// blockIdx.y < n_ftiles: field wires (thread = witness: coalesced reads of the row, one 32-byte segment written per witness);
// the other blocks: one warp per 32 witnesses x 32 bit wires -- lane l fetches the word of bit row blist[t * 32 + l] (bit i =
// witness i), 32 ballots transpose the 32 x 32 bit matrix, lane i ends up with the word of witness i.
struct PackedView {
    StoreView sv;
    const uint32_t *flist, *blist;
    uint32_t n_f, n_b, n_ftiles;
    uint64_t row_bytes;
};
__global__ void __launch_bounds__(256) export_packed_kernel(PackedView pv, uint64_t B, unsigned char *out) {
    const uint32_t lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    if (blockIdx.y < pv.n_ftiles) {
        const uint64_t w = (uint64_t)blockIdx.x * 256 + threadIdx.x;
        if (w >= B) return;
        for (uint32_t j = blockIdx.y * 8; j < pv.n_f && j < blockIdx.y * 8 + 8; j++) {
            const uint32_t loc = __ldg(pv.flist + j);
            const uint4 *src = pv.sv.store + ((uint64_t)loc * 2) * pv.sv.bstride + w;
            const Fr v = fr::from_mont(unpack(src[0], src[pv.sv.bstride]));
            uint4 lo, hi;
            pack(v, lo, hi);
            uint4 *d = reinterpret_cast<uint4 *>(out + w * pv.row_bytes + (uint64_t)j * 32);
            d[0] = lo;
            d[1] = hi;
        }
        return;
    }
    const uint64_t wg = (uint64_t)blockIdx.x * 8 + wrp;          // group of 32 witnesses
    const uint32_t t = blockIdx.y - pv.n_ftiles;                 // tile of 32 bit wires
    if (wg * 32 >= B) return;
    const uint32_t e = t * 32 + lane;
    uint32_t word = 0;
    if (e < pv.n_b) word = __ldg(pv.sv.bits + wg * pv.sv.n_brows + __ldg(pv.blist + e));
    uint32_t mine = 0;
#pragma unroll
    for (int i = 0; i < 32; i++) {
        const uint32_t col = __ballot_sync(0xffffffffu, (word >> i) & 1u);   // bit l = wire e(l) of witness i
        if (lane == (uint32_t)i) mine = col;
    }
    const uint64_t w = wg * 32 + lane;
    if (w < B) *reinterpret_cast<uint32_t *>(out + w * pv.row_bytes + (uint64_t)pv.n_f * 32 + (uint64_t)t * 4) = mine;
}

// ---- R1CS: linear constraints over 0/1 wires with coefficients +-2^k (r1cs.hpp Bound::shl).  Warp = 32 witnesses; per layer
// lane l fetches the packed word of the wire whose shift is base + l, the 32 x 32 bit-matrix transpose (five shuffle steps)
// hands every lane the partial sum of its own witness; the constraint holds iff the signed total is 0 (|total| < 2^62).
struct R1csShiftParams {
    const uint4 *shh;        // per constraint: first layer, number of layers, constraint index, 0
    const uint32_t *shl;     // 32 words per layer: bit row or 0xffffffff
    const uint32_t *shm;     // per layer: base | negative << 8
    uint32_t n_cons, per_chunk;
    const uint32_t *bits;
    uint32_t n_brows;
    uint64_t B;
    uint32_t *first_bad;
};
__global__ void __launch_bounds__(128) r1cs_shift_kernel(R1csShiftParams p) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t wg = (uint64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (wg * 32 >= p.B) return;
    const uint32_t *brow = p.bits + wg * p.n_brows;
    const uint32_t c0 = blockIdx.y * p.per_chunk, c1 = min(p.n_cons, c0 + p.per_chunk);
    uint32_t bad = 0xffffffffu;
    for (uint32_t c = c0; c < c1; c++) {
        const uint4 h = __ldg(p.shh + c);
        long long acc = 0;
        uint32_t row = __ldg(p.shl + (uint64_t)h.x * 32 + lane);
        for (uint32_t L = 0; L < h.y; L++) {
            const uint32_t meta = __ldg(p.shm + h.x + L);
            const uint32_t nxt = __ldg(p.shl + (uint64_t)(h.x + L + 1) * 32 + lane);   // (the array is padded by one layer)
            uint32_t x = row == 0xffffffffu ? 0u : __ldg(brow + row);
#pragma unroll
            for (uint32_t k = 16, m = 0x0000ffffu; k >= 1; k >>= 1, m ^= m << k) {
                const uint32_t y = __shfl_xor_sync(0xffffffffu, x, k);
                x = (lane & k) ? ((y >> k) & m) | (x & ~m) : (x & m) | ((y & m) << k);
            }
            const long long term = (long long)((unsigned long long)x << (meta & 0xffu));
            acc += (meta >> 8) & 1u ? -term : term;
            row = nxt;
        }
        if (acc != 0) bad = min(bad, h.z);
    }
    const uint64_t w = wg * 32 + lane;
    if (bad != 0xffffffffu && w < p.B) atomicMin(p.first_bad + w, bad);
}

// canonical AoS -> Montgomery SoA (used by the stand-alone R1CS check on externally produced witnesses)
__global__ void __launch_bounds__(256) import_kernel(const uint4 *in, uint64_t B, uint32_t n_wires, uint4 *store,
                                                     uint64_t bstride) {
    __shared__ uint32_t tile[32][33][9];   // [witness][wire][limb]; the padding as in export_kernel
    const uint32_t lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const uint64_t w0 = (uint64_t)blockIdx.x * 32;
    const uint32_t r0 = blockIdx.y * 32;
    for (uint32_t k = wrp; k < 32; k += 8) {   // k = witness, lane = wire: 1 KiB contiguous per warp
        uint64_t w = w0 + k;
        uint32_t row = r0 + lane;
        if (row < n_wires && w < B) {
            const uint4 *s = in + (w * n_wires + row) * 2;
            uint4 lo = s[0], hi = s[1];
            tile[k][lane][0] = lo.x; tile[k][lane][1] = lo.y; tile[k][lane][2] = lo.z; tile[k][lane][3] = lo.w;
            tile[k][lane][4] = hi.x; tile[k][lane][5] = hi.y; tile[k][lane][6] = hi.z; tile[k][lane][7] = hi.w;
        }
    }
    __syncthreads();
    for (uint32_t k = wrp; k < 32; k += 8) {   // k = wire, lane = witness
        uint32_t row = r0 + k;
        uint64_t w = w0 + lane;
        if (row < n_wires && w < B) {
            Fr v;
#pragma unroll
            for (int i = 0; i < 8; i++) v.v[i] = tile[lane][k][i];
            v = op_input(v);
            uint4 lo, hi;
            pack(v, lo, hi);
            uint4 *d = store + ((uint64_t)row * 2) * bstride + w;
            d[0] = lo;
            d[bstride] = hi;
        }
    }
}

// ---- R1CS satisfiability: (A.w)*(B.w) - C.w == 0 for every constraint (constraints-json.md:17) ------
// thread = one witness; blockIdx.y = chunk of constraints.  All lanes of a warp walk the same CSR rows
// (uniform, broadcast loads); witness values are read from the SoA store (coalesced 128-bit loads).
// Coefficient index 0 is +1 and 1 is -1: those terms are one add/sub; others cost one Montgomery product.
struct R1csParams {
    const uint4 *hdr;           // 3*n_cons+1 linear-combination headers {begin, end of +-1, of small +, of small -}; the end of
                                // an LC is the begin of the next one (the last header only carries the total)
    const uint32_t *cmag;       // per coefficient: 32-bit magnitude of a small coefficient
    const uint2 *terms;         // (wire, coef index)
    const uint4 *coefs;         // Montgomery, 2 x uint4 each
    uint32_t n_cons;
    uint32_t cons_per_chunk;
    const uint4 *store;
    uint64_t bstride;
    uint64_t B;
    uint32_t *first_bad;        // B words, pre-set to 0xffffffff
    // typed stores (a program's layout, tape.hpp): terms on bit rows live in their own CSR.  bhdr[c] = begin of the bit
    // terms of A | mode << 30, of B, of C, of the next constraint's A (n_cons + 1 entries); mode 1 / 2 marks a constraint
    // made only of bit terms with integer coefficients (cint): it is evaluated in 64- / 32-bit integers (r1cs.hpp Bound).
    // bterms = (bit row, coefficient index).
    const uint32_t *bits;
    uint32_t n_brows;
    const uint4 *bhdr;
    const uint2 *bterms;
    const long long *cint;
    // TYPED: the constraints this kernel walks (all but those r1cs_table_kernel takes), n_cons = their number
    const uint32_t *active;
    uint32_t n_all;
};

// One linear combination.  Its terms are ordered by coefficient class (r1cs.hpp):
//   +-2^k, k <= 3    [beg, e0): k modular doublings, then modular add/sub (+-1 is k = 0)
//   small +, small - [e0, e1), [e1, e2): integer sums with 32-bit scalars (8 multiply-accumulates per term), one
//                    reduction per class (fr.cuh small_reduce)
//   general          [e2, end): lazy-reduction dot product (64 multiply-accumulates per term, one Montgomery
//                    reduction per <= 16 terms); a general coefficient on wire 0 comes last and is added as is
//
// Operand stream.  The terms of a chunk of constraints are contiguous in the CSR (A, B, C of constraint c, then
// c+1, ...) and their addresses do not depend on data, so every thread keeps R1CS_STAGES of its own witness's
// values in flight with cp.async (LDGSTS, global -> its private column of a shared-memory ring, L2-only caching)
// and consumes them in order: HBM latency is hidden even when few warps are resident (small batches).
#define R1CS_NT 128
#define R1CS_SAME_AS_A 0xffffffffu   // r1cs.hpp SAME_AS_A: first class boundary of a B combination that repeats A
#define R1CS_STAGES 8

struct TermStream {
    const uint2 *terms;
    const uint4 *wbase;
    uint64_t bstride;
    uint4 *ring;            // [R1CS_STAGES][2][R1CS_NT]
    uint32_t t_end;
    // the stream is position-determined: term t's value is requested R1CS_STAGES takes before it is consumed
    // (measured: a 32-bit row stride / hand-written shared-space addressing here changes nothing -- Poseidon check
    // 17.3 vs 17.6 ms)
    __device__ __forceinline__ void issue(uint32_t t) const {
        if (t < t_end) {
            const uint2 term = __ldg(terms + t);
            const uint4 *src = wbase + ((uint64_t)(term.x & 0x0fffffffu) * 2) * bstride;   // top bits: +-2^k meta (r1cs.hpp)
            uint4 *dst = ring + (t % R1CS_STAGES) * 2 * R1CS_NT + threadIdx.x;
            const uint32_t d0 = (uint32_t)__cvta_generic_to_shared(dst);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d0), "l"(src) : "memory");
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d0 + R1CS_NT * 16), "l"(src + bstride) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    __device__ __forceinline__ void start(uint32_t t_begin) const {
#pragma unroll
        for (int k = 0; k < R1CS_STAGES; k++) issue(t_begin + k);
    }
    // value of term t (terms are consumed strictly in order); refills the stage it frees
    __device__ __forceinline__ Fr take(uint32_t t) const {
        asm volatile("cp.async.wait_group %0;" ::"n"(R1CS_STAGES - 1) : "memory");
        const uint4 *s = ring + (t % R1CS_STAGES) * 2 * R1CS_NT + threadIdx.x;
        const Fr v = unpack(s[0], s[R1CS_NT]);
        issue(t + R1CS_STAGES);
        return v;
    }
};

// called up to three times per constraint (A, B, C in CSR order); hdr = {begin, e0, e1, e2}.  Inlined in the plain kernel
// (Poseidon check 17.3 -> 16.5 ms: the calls were 13 % of its instructions), out of line in the typed one, whose register
// budget is spent on the bit stream (EdDSA check 19.9 vs 20.3 ms inlined).
__device__ __forceinline__ Fr lc_eval(const uint2 *terms, const uint4 *coefs, const uint32_t *cmag, const uint4 *wbase,
                                   uint64_t bstride, uint4 *ring, uint32_t t_end, uint4 hdr, uint32_t end) {
    TermStream ts;
    ts.terms = terms;
    ts.wbase = wbase;
    ts.bstride = bstride;
    ts.ring = ring;
    ts.t_end = t_end;
    const uint32_t e0 = hdr.y, e1 = hdr.z, e2 = hdr.w;
    Fr acc = fr::zero();
    uint32_t t = hdr.x;
    for (; t < e0; t++) {
        const uint32_t meta = __ldg(terms + t).x >> 28;   // sign | k << 1: coefficient +-2^k, k <= 3
        Fr v = ts.take(t);
        for (uint32_t k = meta >> 1; k; k--) v = fr::add(v, v);
        acc = (meta & 1u) ? fr::sub(acc, v) : fr::add(acc, v);
    }
    if (e0 < e2) {
        for (int neg = 0; neg < 2; neg++) {
            const uint32_t stop = neg ? e2 : e1;
            if (t >= stop) continue;
            fr::Small X;
            fr::small_zero(X);
            for (; t < stop; t++) {
                const uint32_t coef = __ldg(terms + t).y;
                const Fr v = ts.take(t);
                fr::small_mac(X, __ldg(cmag + coef), v);
            }
            const Fr sres = fr::small_reduce(X);
            acc = neg ? fr::sub(acc, sres) : fr::add(acc, sres);
        }
    }
    // a general coefficient on wire 0 (the constant 1) is the LC's last term, marked in bit 31: its value is the
    // coefficient itself
    const bool has_const = e2 < end && (__ldg(terms + end - 1).x >> 31);
    const uint32_t gend = end - (has_const ? 1u : 0u);
    while (t < gend) {
        const uint32_t n = min(gend - t, 16u);
        fr::Wide T;
        fr::wide_zero(T);
        for (uint32_t k = 0; k < n; k++, t++) {
            const uint32_t coef = __ldg(terms + t).y;
            const Fr c = unpack(__ldg(coefs + 2 * (uint64_t)coef), __ldg(coefs + 2 * (uint64_t)coef + 1));
            const Fr v = ts.take(t);
            fr::wide_mac(T, c, v);
        }
        acc = fr::add(acc, fr::wide_reduce(T, n));
    }
    if (has_const) {
        const uint32_t coef = __ldg(terms + t).y;
        (void)ts.take(t);   // its slot of the operand stream
        acc = fr::add(acc, unpack(__ldg(coefs + 2 * (uint64_t)coef), __ldg(coefs + 2 * (uint64_t)coef + 1)));
    }
    return acc;
}

__device__ __noinline__ Fr lc_eval_call(const uint2 *terms, const uint4 *coefs, const uint32_t *cmag, const uint4 *wbase,
                                        uint64_t bstride, uint4 *ring, uint32_t t_end, uint4 hdr, uint32_t end) {
    return lc_eval(terms, coefs, cmag, wbase, bstride, ring, t_end, hdr, end);
}
template <bool INL>
__device__ __forceinline__ Fr lc_eval_sel(const uint2 *terms, const uint4 *coefs, const uint32_t *cmag, const uint4 *wbase,
                                          uint64_t bstride, uint4 *ring, uint32_t t_end, uint4 hdr, uint32_t end) {
    if (INL) return lc_eval(terms, coefs, cmag, wbase, bstride, ring, t_end, hdr, end);
    return lc_eval_call(terms, coefs, cmag, wbase, bstride, ring, t_end, hdr, end);
}

// linear combinations made of +-2^k (k <= 3) terms only stay inline
template <bool INL>
__device__ __forceinline__ Fr lc_any(const R1csParams &p, const uint4 *wbase, uint4 *ring, uint32_t t_end, uint4 hdr, uint32_t end) {
    if (hdr.y != end) return lc_eval_sel<INL>(p.terms, p.coefs, p.cmag, wbase, p.bstride, ring, t_end, hdr, end);
    TermStream ts;
    ts.terms = p.terms;
    ts.wbase = wbase;
    ts.bstride = p.bstride;
    ts.ring = ring;
    ts.t_end = t_end;
    // first term: the value itself (most linear combinations of gate-level circuits are a single wire)
    uint32_t t = hdr.x;
    Fr acc;
    {
        const uint32_t meta = __ldg(p.terms + t).x >> 28;   // sign | k << 1: coefficient +-2^k, k <= 3
        acc = ts.take(t);
        for (uint32_t k = meta >> 1; k; k--) acc = fr::add(acc, acc);
        if (meta & 1u) acc = fr::neg(acc);
    }
    for (t++; t < end; t++) {
        const uint32_t meta = __ldg(p.terms + t).x >> 28;
        Fr v = ts.take(t);
        for (uint32_t k = meta >> 1; k; k--) v = fr::add(v, v);
        acc = (meta & 1u) ? fr::sub(acc, v) : fr::add(acc, v);
    }
    return acc;
}

// Stream of the bit-row terms of a chunk of constraints.  They are consumed strictly in CSR order by every lane of a
// warp, and a term's operand is one 32-bit word for the whole warp -- so the warp fetches them cooperatively: lane l
// loads term (base + l), its word of the bit row and its integer coefficient (32 independent gathers in flight at
// once), the batch after the one being consumed waits in registers, and the batch being consumed sits in a 512-byte
// stage of shared memory that every lane reads with broadcast loads.  The DRAM / L2 latency of the gathers is paid
// once per 32 terms, under the arithmetic of the previous batch, instead of once per term on the dependence chain.
struct BitStream {
    const uint2 *bterms;
    const uint32_t *brow;
    const long long *cint;
    uint32_t stage;      // shared-space address of this warp's 32 entries: (word, coefficient lo, hi, coefficient index)
    uint32_t base, t_end;
    uint32_t lane;       // the thread's own lane (which entry of a batch it fetches)
    uint32_t bitlane;    // the lane whose bit of a word this thread reads (a padding thread copies the last witness)
    uint4 nxt;
    __device__ __forceinline__ void load(uint32_t b) {
        const uint32_t t = b + lane;
        nxt = make_uint4(0u, 0u, 0u, 0u);
        if (t < t_end) {
            const uint2 term = __ldg(bterms + t);
            const long long c = __ldg(cint + term.y);
            nxt = make_uint4(__ldg(brow + term.x), (uint32_t)c, (uint32_t)((unsigned long long)c >> 32), term.y);
        }
    }
    __device__ __forceinline__ void put() {
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(stage + lane * 16u), "r"(nxt.x), "r"(nxt.y), "r"(nxt.z), "r"(nxt.w)
                     : "memory");
    }
    __device__ __forceinline__ void start(uint32_t t0) {
        base = t0;
        load(t0);
        put();
        __syncwarp();
        load(t0 + 32);
    }
    // make the batch that holds term t current (uniform: every lane of the warp walks the same terms)
    __device__ __forceinline__ void seek(uint32_t t) {
        while (t >= base + 32) {
            __syncwarp();
            put();
            __syncwarp();
            base += 32;
            load(base + 32);
        }
    }
    __device__ __forceinline__ uint4 at(uint32_t t) {
        seek(t);
        uint4 s;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(s.x), "=r"(s.y), "=r"(s.z), "=r"(s.w) : "r"(stage + (t - base) * 16u));
        return s;
    }
};

// bit-row terms of one linear combination in the field: acc += bit ? coefficient : 0 (no multiplication)
__device__ __forceinline__ Fr lc_bits(const R1csParams &p, BitStream &bs, uint32_t b, uint32_t e, Fr acc) {
    for (uint32_t t = b; t < e; t++) {
        const uint4 s = bs.at(t);
        const bool bit = (s.x >> bs.bitlane) & 1u;
        const uint4 lo = __ldg(p.coefs + 2 * (uint64_t)s.w), hi = __ldg(p.coefs + 2 * (uint64_t)s.w + 1);
        Fr c;
        c.v[0] = bit ? lo.x : 0u; c.v[1] = bit ? lo.y : 0u; c.v[2] = bit ? lo.z : 0u; c.v[3] = bit ? lo.w : 0u;
        c.v[4] = bit ? hi.x : 0u; c.v[5] = bit ? hi.y : 0u; c.v[6] = bit ? hi.z : 0u; c.v[7] = bit ? hi.w : 0u;
        acc = fr::add(acc, c);
    }
    return acc;
}
// the same in plain integers: 64-bit (the host guarantees that no partial sum leaves 62 bits), or 32-bit when every
// partial sum stays below 2^31 -- one multiply-add per term
template <bool I32>
__device__ __forceinline__ long long lc_int(BitStream &bs, uint32_t b, uint32_t e) {
    long long acc = 0;
    int acc32 = 0;
    uint32_t t = b;
    while (t < e) {
        bs.seek(t);
        const uint32_t stop = min(e, bs.base + 32u);
        uint32_t sa = bs.stage + (t - bs.base) * 16u;
        for (; t < stop; t++, sa += 16u) {
            if (I32) {
                uint32_t w, c;
                asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(w), "=r"(c) : "r"(sa));
                acc32 += (int)((w >> bs.bitlane) & 1u) * (int)c;
            } else {
                uint32_t w, c0, c1, ci;
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(w), "=r"(c0), "=r"(c1), "=r"(ci) : "r"(sa));
                const long long c = (long long)((unsigned long long)c0 | ((unsigned long long)c1 << 32));
                acc += ((w >> bs.bitlane) & 1u) ? c : 0ll;
            }
        }
    }
    return I32 ? (long long)acc32 : acc;
}

template <int MINB, bool TYPED>
__global__ void __launch_bounds__(R1CS_NT, MINB) r1cs_kernel(R1csParams p) {
    __shared__ uint4 ring[R1CS_STAGES * 2 * R1CS_NT];
    __shared__ uint4 bstage[TYPED ? R1CS_NT : 1];
    uint64_t w = (uint64_t)blockIdx.x * R1CS_NT + threadIdx.x;
    const bool active = w < p.B;
    // (a padding lane evaluates a copy of the last witness: its bit is the last witness's lane of that warp's word)
    const uint32_t lane = (uint32_t)(active ? w : p.B - 1) & 31u;
    if (!active) w = p.B - 1;
    // chunk of the walk: positions [i0, i1) of the active list (TYPED) or of the constraints themselves; its terms are the
    // CSR range from the first constraint's begin to the begin of the first constraint after the chunk (skipped
    // constraints have no terms)
    const uint32_t i0 = blockIdx.y * p.cons_per_chunk;
    const uint32_t i1 = min(p.n_cons, i0 + p.cons_per_chunk);
    const uint32_t c0 = TYPED ? __ldg(p.active + i0) : i0;
    const uint32_t c1 = TYPED ? (i1 < p.n_cons ? __ldg(p.active + i1) : p.n_all) : i1;
    const uint4 *wbase = p.store + w;
    const uint32_t *brow = TYPED ? p.bits + (w >> 5) * p.n_brows : nullptr;
    const uint32_t t_end = __ldg(&p.hdr[3 * c1].x);
    {
        TermStream ts;
        ts.terms = p.terms;
        ts.wbase = wbase;
        ts.bstride = p.bstride;
        ts.ring = ring;
        ts.t_end = t_end;
        ts.start(__ldg(&p.hdr[3 * c0].x));
    }
    BitStream bs;
    if (TYPED) {
        bs.bterms = p.bterms;
        bs.brow = brow;
        bs.cint = p.cint;
        bs.stage = (uint32_t)__cvta_generic_to_shared(bstage + (threadIdx.x & ~31u));
        bs.lane = threadIdx.x & 31u;
        bs.bitlane = lane;
        bs.t_end = __ldg(&p.bhdr[c1].x) & 0x3fffffffu;
        bs.start(__ldg(&p.bhdr[c0].x) & 0x3fffffffu);
    }
    uint32_t bad = 0xffffffffu;
    uint4 hA = __ldg(p.hdr + 3 * c0);
    for (uint32_t i = i0; i < i1; i++) {
        const uint32_t c = TYPED ? __ldg(p.active + i) : i;
        uint32_t bA = 0, bB = 0, bC = 0, bN = 0;
        if (TYPED) {
            const uint4 bh = __ldg(p.bhdr + c);
            bA = bh.x & 0x3fffffffu; bB = bh.y; bC = bh.z; bN = bh.w;
            const uint32_t mode = bh.x >> 30;
            if (mode == 3) continue;   // a boolean predicate of a few 0/1 wires: r1cs_table_kernel
            if (mode) {
                // every term is a 0/1 wire with a small integer coefficient: |A*B - C| is far below q, so the constraint
                // holds mod q iff it holds in the integers.  (No field-row terms: the operand stream is not touched.)
                bool ok;
                if (mode == 2) {
                    const long long ia = lc_int<true>(bs, bA, bB), ib = lc_int<true>(bs, bB, bC), ic = lc_int<true>(bs, bC, bN);
                    ok = ia * ib == ic;
                } else {
                    const long long ia = lc_int<false>(bs, bA, bB), ib = lc_int<false>(bs, bB, bC), ic = lc_int<false>(bs, bC, bN);
                    ok = ia * ib == ic && __mul64hi(ia, ib) == (ic >> 63);
                }
                if (bad == 0xffffffffu && !ok) bad = c;
                continue;
            }
            hA = __ldg(p.hdr + 3 * c);   // (the chain hA = hN below is broken by integer constraints)
        }
        // headers of A, B, C and of the next constraint's A (its begin is the end of C)
        const uint4 hB = __ldg(p.hdr + 3 * c + 1), hC = __ldg(p.hdr + 3 * c + 2), hN = __ldg(p.hdr + 3 * c + 3);
        const bool hasA = hA.x != hB.x || bA != bB, hasC = hC.x != hN.x || bC != bN;
        const bool hasB = hB.x != hC.x || bB != bC || hB.y == R1CS_SAME_AS_A;
        Fr prod = fr::zero();
        if (hB.y == R1CS_SAME_AS_A) {        // B repeats A (r1cs.hpp): one evaluation, one squaring
            Fr sa = fr::zero();
            if (hA.x != hB.x) sa = lc_any<!TYPED>(p, wbase, ring, t_end, hA, hB.x);
            if (TYPED) sa = lc_bits(p, bs, bA, bB, sa);
            const Fr one = fr::one_mont();
            if (fr::is_zero(sa)) prod = fr::zero();
            else if (sa.v[0] == one.v[0] && fr::equal(sa, one)) prod = one;
            else prod = fr::mont_sqr(sa);
        } else if (hasA && hasB) {   // an empty A or B makes the product 0 (linear constraint, algebra.rs:1052-1054)
            Fr sa = fr::zero(), sb = fr::zero();
            if (hA.x != hB.x) sa = lc_any<!TYPED>(p, wbase, ring, t_end, hA, hB.x);
            if (TYPED) sa = lc_bits(p, bs, bA, bB, sa);
            if (hB.x != hC.x) sb = lc_any<!TYPED>(p, wbase, ring, t_end, hB, hC.x);
            if (TYPED) sb = lc_bits(p, bs, bB, bC, sb);
            // trivial factors need no product: 0, 1 and -1 (bit-valued wires and the +-1 combinations of them that
            // fill hash circuits).  The branch is per lane; a warp pays for the multiplication only if one of its
            // witnesses needs it.
            const Fr one = fr::one_mont();
            const Fr mone = fr::minus_one_mont();
            // (the low limb filters first: general field values fail it with one comparison per case)
            if (fr::is_zero(sa) || fr::is_zero(sb)) prod = fr::zero();
            else if (sa.v[0] == one.v[0] && fr::equal(sa, one)) prod = sb;
            else if (sb.v[0] == one.v[0] && fr::equal(sb, one)) prod = sa;
            else if (sa.v[0] == mone.v[0] && fr::equal(sa, mone)) prod = fr::neg(sb);
            else if (sb.v[0] == mone.v[0] && fr::equal(sb, mone)) prod = fr::neg(sa);
            else if (sa.v[0] == sb.v[0] && fr::equal(sa, sb)) prod = fr::mont_sqr(sa);   // S-box squarings: A and B are the same combination
            else prod = fr::mont_mul(sa, sb);
        } else {
            // the field terms of a lone A or B still occupy the stream: consume them
            if (hA.x != hB.x) (void)lc_eval_call(p.terms, p.coefs, p.cmag, wbase, p.bstride, ring, t_end, hA, hB.x);
            if (hB.x != hC.x) (void)lc_eval_call(p.terms, p.coefs, p.cmag, wbase, p.bstride, ring, t_end, hB, hC.x);
        }
        Fr sc = fr::zero();
        if (hC.x != hN.x) sc = lc_any<!TYPED>(p, wbase, ring, t_end, hC, hN.x);
        if (TYPED && hasC) sc = lc_bits(p, bs, bC, bN, sc);
        if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = c;
        hA = hN;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (active && bad != 0xffffffffu) atomicMin(p.first_bad + w, bad);
}

// Constraints over at most five distinct 0/1 wires (r1cs.hpp Bound::tcons): the constraint is a boolean predicate of
// those wires, tabulated on the host.  A bit row holds the wire for 32 witnesses in one word, so ONE lane evaluates the
// predicate for the warp's 32 witnesses with a multiplexer tree over the 32-entry table (63 bitwise instructions), and
// the 32 lanes take 32 different constraints: ~2.5 instructions per constraint and 32 witnesses, against ~100 per
// constraint and witness on the per-witness path.  Hash circuits are made of such gates (Sha256: 90 % of the constraints).
// thread block = 4 warps = 4 groups of 32 witnesses; blockIdx.y = chunk of the table constraints.
struct R1csTableParams {
    const uint4 *tcons;          // 2 x uint4 per constraint: {row0..row3}, {row4, table, constraint index, 0}
    uint32_t n_tcons, per_chunk;
    const uint32_t *bits;
    uint32_t n_brows;
    uint64_t B;
    uint32_t *first_bad;
};
__global__ void __launch_bounds__(128) r1cs_table_kernel(R1csTableParams p) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t wg = (uint64_t)blockIdx.x * 4 + (threadIdx.x >> 5);   // group of 32 witnesses
    const uint64_t w0 = wg * 32;
    if (w0 >= p.B) return;
    const uint32_t valid = (p.B - w0 >= 32) ? 0xffffffffu : ((1u << (uint32_t)(p.B - w0)) - 1u);
    const uint32_t *brow = p.bits + wg * p.n_brows;
    const uint32_t t0 = blockIdx.y * p.per_chunk, t1 = min(p.n_tcons, t0 + p.per_chunk);
    for (uint32_t base = t0; base < t1; base += 32) {
        const uint32_t i = base + lane;
        uint32_t badw = 0, cidx = 0;
        if (i < t1) {
            const uint4 ra = __ldg(p.tcons + 2 * (uint64_t)i), rb = __ldg(p.tcons + 2 * (uint64_t)i + 1);
            const uint32_t x0 = __ldg(brow + ra.x), x1 = __ldg(brow + ra.y), x2 = __ldg(brow + ra.z), x3 = __ldg(brow + ra.w),
                           x4 = __ldg(brow + rb.x);
            const int t = (int)rb.y;
            cidx = rb.z;
            // multiplexer tree: level 0 selects on x0 between table entries 2j and 2j+1 (as all-ones / all-zero masks)
            uint32_t v[16];
#pragma unroll
            for (int j = 0; j < 16; j++) {
                const uint32_t m0 = (uint32_t)((t << (31 - 2 * j)) >> 31), m1 = (uint32_t)((t << (30 - 2 * j)) >> 31);
                v[j] = (x0 & m1) | (~x0 & m0);
            }
#pragma unroll
            for (int j = 0; j < 8; j++) v[j] = (x1 & v[2 * j + 1]) | (~x1 & v[2 * j]);
#pragma unroll
            for (int j = 0; j < 4; j++) v[j] = (x2 & v[2 * j + 1]) | (~x2 & v[2 * j]);
#pragma unroll
            for (int j = 0; j < 2; j++) v[j] = (x3 & v[2 * j + 1]) | (~x3 & v[2 * j]);
            const uint32_t sat = (x4 & v[1]) | (~x4 & v[0]);
            badw = ~sat & valid;
        }
        if (__any_sync(0xffffffffu, badw != 0u)) {   // rare: some witness violates one of these 32 constraints
            while (badw) {
                const uint32_t j = __ffs(badw) - 1;
                badw &= badw - 1;
                atomicMin(p.first_bad + w0 + j, cidx);
            }
        }
    }
}

// ---- device self-test of the field routines -----------------------------------------------------------
__global__ void fr_op_kernel(int op, const uint4 *a, const uint4 *b, uint4 *out, uint64_t n, uint32_t *undef) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr x = fr::to_mont(unpack(a[2 * i], a[2 * i + 1]));
    Fr y = fr::to_mont(unpack(b[2 * i], b[2 * i + 1]));
    Fr r = fr::zero();
    uint32_t st = 0;
    switch (op) {
        case hostfr::F_ADD: r = fr::add(x, y); break;
        case hostfr::F_SUB: r = fr::sub(x, y); break;
        case hostfr::F_NEG: r = fr::neg(x); break;
        case hostfr::F_MUL: r = fr::mont_mul(x, y); break;
        case hostfr::F_SQUARE: r = fr::mont_sqr(x); break;
        case hostfr::F_INV: r = fr::mont_inv(x); break;
        case hostfr::F_DIV: r = fr::mont_mul(x, fr::mont_inv(y)); break;
        case hostfr::F_IDIV: r = slow_compute(tape::T_IDIV, x, y, st); break;
        case hostfr::F_MOD: r = slow_compute(tape::T_MOD, x, y, st); break;
        case hostfr::F_POW: r = slow_compute(tape::T_POW, x, y, st); break;
        case hostfr::F_SHL: r = slow_compute(tape::T_SHL, x, y, st); break;
        case hostfr::F_SHR: r = slow_compute(tape::T_SHR, x, y, st); break;
        case hostfr::F_BAND: r = slow_compute(tape::T_BAND, x, y, st); break;
        case hostfr::F_BOR: r = slow_compute(tape::T_BOR, x, y, st); break;
        case hostfr::F_BXOR: r = slow_compute(tape::T_BXOR, x, y, st); break;
        case hostfr::F_BNOT: r = slow_compute(tape::T_BNOT, x, y, st); break;
        case hostfr::F_LT: r = slow_compute(tape::T_LT, x, y, st); break;
        case hostfr::F_LE: r = slow_compute(tape::T_LE, x, y, st); break;
        case hostfr::F_GT: r = slow_compute(tape::T_GT, x, y, st); break;
        case hostfr::F_GE: r = slow_compute(tape::T_GE, x, y, st); break;
        case hostfr::F_EQ: r = mont_bool(fr::equal(x, y)); break;
        case hostfr::F_NEQ: r = mont_bool(!fr::equal(x, y)); break;
        case hostfr::F_LAND: r = slow_compute(tape::T_LAND, x, y, st); break;
        case hostfr::F_LOR: r = slow_compute(tape::T_LOR, x, y, st); break;
        case hostfr::F_EQZ: r = mont_bool(fr::is_zero(x)); break;
        default: r = x; break;
    }
    if (st) undef[i] = st;
    r = fr::from_mont(r);
    uint4 lo, hi;
    pack(r, lo, hi);
    out[2 * i] = lo;
    out[2 * i + 1] = hi;
}

// ---- integer-multiply roofline probe -----------------------------------------------------------------
// 8 independent chains per thread; kind 0: mad.lo.u32 + mad.hi.u32 pairs (2 instructions per 32x32->64
// multiply-accumulate), kind 1: mad.wide.u32 (1 instruction, 64-bit accumulate).
template <int KIND>
__global__ void __launch_bounds__(256) imad_kernel(uint32_t *out, uint32_t iters, uint32_t seed) {
    uint32_t x = seed + threadIdx.x, y = seed * 3 + blockIdx.x;
    uint32_t a0 = x, a1 = x + 1, a2 = x + 2, a3 = x + 3, a4 = x + 4, a5 = x + 5, a6 = x + 6, a7 = x + 7;
    uint64_t d0 = x, d1 = x + 1, d2 = x + 2, d3 = x + 3, d4 = x + 4, d5 = x + 5, d6 = x + 6, d7 = x + 7;
    uint32_t b0 = y, b1 = y + 1, b2 = y + 2, b3 = y + 3, b4 = y + 4, b5 = y + 5, b6 = y + 6, b7 = y + 7;
    uint32_t c0 = x ^ y, c1 = c0 + 1, c2 = c0 + 2, c3 = c0 + 3, c4 = c0 + 4, c5 = c0 + 5, c6 = c0 + 6, c7 = c0 + 7;
    uint32_t e0 = x * y, e1 = e0 + 1, e2 = e0 + 2, e3 = e0 + 3, e4 = e0 + 4, e5 = e0 + 5, e6 = e0 + 6, e7 = e0 + 7;
    for (uint32_t i = 0; i < iters; i++) {
        if (KIND == 8) {        // the form the field arithmetic uses (fr.cuh mad4): mad.lo.cc / madc.hi.cc pairs with the SAME
            // multiplicand on a register pair, which ptxas fuses into one IMAD.WIDE.U32.X each; four independent chains of
            // four (16 units per iteration), multiplicands taken from another chain so that nothing is loop-invariant
#define CHAIN(r0, r1, r2, r3, r4, r5, r6, r7, m0, m1, m2, m3)                                                              \
            asm volatile("mad.lo.cc.u32 %0, %8, %12, %0;\n\tmadc.hi.cc.u32 %1, %8, %12, %1;\n\t"                            \
                         "madc.lo.cc.u32 %2, %9, %12, %2;\n\tmadc.hi.cc.u32 %3, %9, %12, %3;\n\t"                           \
                         "madc.lo.cc.u32 %4, %10, %12, %4;\n\tmadc.hi.cc.u32 %5, %10, %12, %5;\n\t"                         \
                         "madc.lo.cc.u32 %6, %11, %12, %6;\n\tmadc.hi.u32 %7, %11, %12, %7;"                                 \
                         : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7)                   \
                         : "r"(m0), "r"(m1), "r"(m2), "r"(m3), "r"(y));
            CHAIN(a0, a1, a2, a3, a4, a5, a6, a7, b0, b2, b4, b6)
            CHAIN(b0, b1, b2, b3, b4, b5, b6, b7, c1, c3, c5, c7)
            CHAIN(c0, c1, c2, c3, c4, c5, c6, c7, e0, e2, e4, e6)
            CHAIN(e0, e1, e2, e3, e4, e5, e6, e7, a1, a3, a5, a7)
#undef CHAIN
        } else if (KIND == 0) {        // mad.lo + mad.hi pair = one 32x32->64 multiply-accumulate
#define STEP(r) asm volatile("mad.lo.u32 %0, %0, %1, %0;\n\tmad.hi.u32 %0, %0, %1, %0;" : "+r"(r) : "r"(y));
            STEP(a0) STEP(a1) STEP(a2) STEP(a3) STEP(a4) STEP(a5) STEP(a6) STEP(a7)
#undef STEP
        } else if (KIND == 1) { // mad.wide.u32
            // the multiplicand is the low word of the chain's own accumulator (a register alias, no instruction): with
            // loop-invariant multiplicands ptxas hoists the product and the loop degenerates into IADD3s
#define STEPW(r) asm volatile("{ .reg .u32 lo; cvt.u32.u64 lo, %0; mad.wide.u32 %0, lo, %1, %0; }" : "+l"(r) : "r"(y));
            STEPW(d0) STEPW(d1) STEPW(d2) STEPW(d3) STEPW(d4) STEPW(d5) STEPW(d6) STEPW(d7)
#undef STEPW
        } else if (KIND == 7) { // mul.wide.u32 with a zero addend (the form mont_mul_wide uses)
            // both halves of the product feed the next multiplicand (a dead high half would let ptxas use a 32-bit IMAD);
            // the add runs on the ALU pipe
#define STEPM(r) asm volatile("{ .reg .u64 e; .reg .u32 el, eh; mul.wide.u32 e, %0, %1; mov.b64 {el, eh}, e; add.u32 %0, el, eh; }" : "+r"(r) : "r"(y));
            STEPM(a0) STEPM(a1) STEPM(a2) STEPM(a3) STEPM(a4) STEPM(a5) STEPM(a6) STEPM(a7)
#undef STEPM
        } else if (KIND == 2) { // mad.lo only (2 per step so that the op count matches kind 0)
#define STEP(r) asm volatile("mad.lo.u32 %0, %0, %1, %0;\n\tmad.lo.u32 %0, %0, %1, %0;" : "+r"(r) : "r"(y));
            STEP(a0) STEP(a1) STEP(a2) STEP(a3) STEP(a4) STEP(a5) STEP(a6) STEP(a7)
#undef STEP
        } else if (KIND == 3) { // mad.hi only
#define STEP(r) asm volatile("mad.hi.u32 %0, %0, %1, %0;\n\tmad.hi.u32 %0, %0, %1, %0;" : "+r"(r) : "r"(y));
            STEP(a0) STEP(a1) STEP(a2) STEP(a3) STEP(a4) STEP(a5) STEP(a6) STEP(a7)
#undef STEP
        } else if (KIND == 4) { // carry chain of 8 adds (IADD3.X)
            asm volatile(
                "add.cc.u32 %0, %0, %8;\n\taddc.cc.u32 %1, %1, %8;\n\taddc.cc.u32 %2, %2, %8;\n\taddc.cc.u32 %3, %3, %8;\n\t"
                "addc.cc.u32 %4, %4, %8;\n\taddc.cc.u32 %5, %5, %8;\n\taddc.cc.u32 %6, %6, %8;\n\taddc.u32 %7, %7, %8;\n\t"
                "add.cc.u32 %0, %0, %8;\n\taddc.cc.u32 %1, %1, %8;\n\taddc.cc.u32 %2, %2, %8;\n\taddc.cc.u32 %3, %3, %8;\n\t"
                "addc.cc.u32 %4, %4, %8;\n\taddc.cc.u32 %5, %5, %8;\n\taddc.cc.u32 %6, %6, %8;\n\taddc.u32 %7, %7, %8;"
                : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(a4), "+r"(a5), "+r"(a6), "+r"(a7) : "r"(y));
        } else if (KIND == 5) { // independent plain adds (IADD3)
#define STEP(r) asm volatile("add.u32 %0, %0, %1;\n\tadd.u32 %0, %0, %1;" : "+r"(r) : "r"(y));
            STEP(a0) STEP(a1) STEP(a2) STEP(a3) STEP(a4) STEP(a5) STEP(a6) STEP(a7)
#undef STEP
        } else {                // mad.lo.cc / madc.hi.cc chain (8 instructions)
            asm volatile(
                "mad.lo.cc.u32 %0, %0, %8, %0;\n\tmadc.hi.cc.u32 %1, %1, %8, %1;\n\tmadc.lo.cc.u32 %2, %2, %8, %2;\n\t"
                "madc.hi.cc.u32 %3, %3, %8, %3;\n\tmadc.lo.cc.u32 %4, %4, %8, %4;\n\tmadc.hi.cc.u32 %5, %5, %8, %5;\n\t"
                "madc.lo.cc.u32 %6, %6, %8, %6;\n\tmadc.hi.u32 %7, %7, %8, %7;\n\t"
                "mad.lo.cc.u32 %0, %0, %8, %0;\n\tmadc.hi.cc.u32 %1, %1, %8, %1;\n\tmadc.lo.cc.u32 %2, %2, %8, %2;\n\t"
                "madc.hi.cc.u32 %3, %3, %8, %3;\n\tmadc.lo.cc.u32 %4, %4, %8, %4;\n\tmadc.hi.cc.u32 %5, %5, %8, %5;\n\t"
                "madc.lo.cc.u32 %6, %6, %8, %6;\n\tmadc.hi.u32 %7, %7, %8, %7;"
                : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(a4), "+r"(a5), "+r"(a6), "+r"(a7) : "r"(y));
        }
    }
    uint32_t acc = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7 ^ b0 ^ b1 ^ b2 ^ b3 ^ b4 ^ b5 ^ b6 ^ b7 ^ c0 ^ c1 ^ c2 ^ c3 ^ c4 ^ c5 ^ c6 ^
                   c7 ^ e0 ^ e1 ^ e2 ^ e3 ^ e4 ^ e5 ^ e6 ^ e7 ^ (uint32_t)(d0 ^ d1 ^ d2 ^ d3 ^ d4 ^ d5 ^ d6 ^ d7) ^
                   (uint32_t)((d0 ^ d1 ^ d2 ^ d3 ^ d4 ^ d5 ^ d6 ^ d7) >> 32);
    if (acc == 0x12345678u) out[0] = acc;   // keep the chains alive
}

// ---- field-multiplication throughput probe: two independent dependent-chains of Montgomery products per thread
template <int VARIANT>
__global__ void __launch_bounds__(128) mulbench_kernel(uint4 *out, uint32_t iters, uint32_t seed) {
    Fr x = fr::one_mont(), y = fr::r2_mont(), z = fr::half_q();
    x.v[0] += threadIdx.x + seed;
    z.v[1] ^= blockIdx.x;
    for (uint32_t i = 0; i < iters; i++) {
        if (VARIANT == 0) { x = fr::mont_mul_portable(x, y); z = fr::mont_mul_portable(z, y); }
        else if (VARIANT == 2) { x = fr::mont_mul_wide(x, y); x = fr::mont_mul_wide(x, z); }   // ONE dependent chain per thread
        else if (VARIANT == 3) { x = fr::mont_mul_chain(x, y); x = fr::mont_mul_chain(x, z); }   // carry-chained rows, one chain
        else if (VARIANT == 4) { x = fr::mont_sqr_chain(x); x = fr::mont_sqr_chain(x); }          // squarings
        else { x = fr::mont_mul_wide(x, y); z = fr::mont_mul_wide(z, y); }
    }
    Fr r = fr::add(x, z);
    if (r.v[0] == 0x12345678u && r.v[7] == 1u) pack(r, out[0], out[1]);
}

}  // namespace kern
