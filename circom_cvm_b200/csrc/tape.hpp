// From the tracer's SSA list to the executable tape: dead-code elimination, then linear-scan
// allocation of the per-witness on-chip slots (shared memory) with furthest-next-use eviction.
// Evicted values that are still needed go to a spill row of the value store in HBM; values bound to
// witness wires are stored to their wire row when produced (they have to be written once anyway,
// common/main.cpp:324-330 reads every witness wire) and are reloaded from there if evicted.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unordered_map>
#include <vector>

#include "tracer.hpp"

namespace tape {

// 16-byte tape instruction
struct TapeIns {
    uint8_t op;
    uint8_t flags;    // bit0: a is a constant index, bit1: b is constant, bit2: c is constant,
                      // bit3: the result is also stored to value-store row c (fused witness-wire store)
    uint16_t dst;     // slot; BSLOT_DST: a slot of the bit file (the value is typed 0/1)
    uint32_t a, b;    // slot (BSLOT: of the bit file) or constant index; T_INPUT: a = input index; T_BITC: b = bit number;
                      // T_LUT: a = bit slots 0 | 1 << 16, b = truth table | number of inputs << 8 | bit slot 2 << 16;
                      // T_LD: a = row to request now for the reload LD_RING reloads ahead (NO_ROW: none), b = ring entry
    uint32_t c;       // T_SEL: third operand; T_LD/T_ST/T_STC: value-store row (ROW_BIT: a bit row); T_FAIL_IF/T_FAIL_NE:
                      // status; with flag bit3: value-store row
};
static const uint8_t F_STORE = 8;
static const uint8_t F_CZERO = 16;
static const uint8_t F_TRIVIAL = 16;  // T_MUL: check at run time whether the factors are 0 / 1 (bit-heavy programs)
static const uint8_t F_ADDEND = 32;
static const uint8_t F_ADDEND2 = 16;  // T_ISUM / T_ISUMT: a second addend in field c (constant: flag bit 2)
static const uint8_t F_CHECK = 128;  // T_ADD / T_SUB / T_MUL / T_DOT of a fused R1CS check: dst = slot to compare the result with, c = constraint
static const uint8_t F_RING = 64;    // T_LD: the value was requested LD_RING reloads ago and sits in ring entry b
static const uint32_t LD_RING = 4;   // reloads in flight per witness (32 B of shared memory each)
static const uint32_t BSLOT = 0x40000000u;     // operand field: the value lives in the bit-slot file (one word per warp; bit = lane)
static const uint16_t BSLOT_DST = 0x8000;      // dst field: idem
static const uint32_t ROW_BIT = 0x80000000u;   // row field: a bit row (one word per warp) instead of a field row
static const uint32_t NO_ROW = 0xffffffffu;  // T_DOT: field b holds an addend (slot, or constant index with bit1)   // T_SEL: the third operand is the constant 0 (field c is free for the fused store)
static_assert(sizeof(TapeIns) == 16, "tape instruction must be 16 bytes");

struct TapeStats {
    uint64_t n_ssa = 0, n_live = 0, n_tape = 0;
    uint64_t n_mul = 0, n_div = 0, n_addsub = 0, n_other = 0, n_inv = 0, n_sel = 0;   // executed per witness
    uint64_t n_ld = 0, n_st = 0, n_spill_st = 0, n_stc = 0, n_input = 0, n_fail = 0, n_rne = 0, n_dot = 0, n_dot_terms = 0, n_ld_streamed = 0, n_lut = 0;
    uint64_t n_groups = 0;         // warp-cooperative group instructions (T_LUTG / T_IBITG)
    uint64_t n_isum_terms = 0;     // conditional adds of bits fused into T_ISUM instructions
    uint64_t n_isum_layers = 0;    // 32 x 32 bit-matrix transposes of the T_ISUMT form
    uint64_t n_int = 0;            // small-integer operations (type_ints): sums of 0/1 values kept as raw 64-bit integers
    uint64_t n_ld_bool = 0, n_spill_st_bool = 0;   // of n_ld / n_spill_st: the value is typed 0/1 (what compact bit rows would shrink)
    uint32_t n_spill_rows = 0;
    uint32_t max_live_field = 0, max_live_bool = 0;   // simultaneously live values of each kind (unlimited slots)
    // 32x32->64 multiply-accumulates the kernel executes per witness: 136 per Montgomery product (also the one that brings
    // an input to Montgomery form and the one after an inversion), 64 per DOT term + 72 per DOT reduction, and the
    // 20 x 90 of the safegcd inversion's matrix updates
    uint64_t macs = 0;
    uint32_t max_live = 0;
};

struct Tape {
    std::vector<TapeIns> ins;
    uint32_t n_slots = 0;
    uint32_t n_wires = 0;
    uint32_t n_rows = 0;      // n_frows + n_brows
    // typed value store: field rows (32 B per witness) hold the field-typed wires [0, n_fwires) and field spills; bit rows
    // (one 32-bit word per warp of witnesses) hold the wires proven 0/1 [0, n_bwires) and bit spills
    uint32_t n_bslots = 0;    // bit slots per warp
    uint32_t n_frows = 0, n_brows = 0, n_fwires = 0, n_bwires = 0;
    std::vector<uint64_t> iconsts;    // constants of the integer operations
    std::vector<uint32_t> wire_loc;   // per witness wire: field row, or ROW_BIT | bit row
    bool use_ring = false;            // field reloads are streamed through the cp.async ring (schedule_reloads)
    uint32_t const_rows[2][2] = {{0, 0}, {0, 0}};   // bit rows of the wires bound to the constant v: [v] = {first, count}
    uint32_t one_brow = 0;            // bit row that holds the constant 1 for every witness (wire 0 itself is a field row)
    TapeStats stats;
};

// ---- batch inversion (Montgomery's trick) over the SSA list ---------------------------------------------
// Every value gets an "inversion level": the largest number of INV operations on a path from the inputs to
// it.  INVs whose operands have the same level cannot depend on each other, so the list is re-ordered by
// stage (level s non-INV operations in their original order, then all INVs fed by level s) and each stage
// with k >= 2 inversions performs ONE field inversion:
//     z_i = (b_i == 0), c_i = z_i ? 1 : b_i, p_i = p_{i-1} * c_i, I = (p_k)^-1,
//     back to front: inv_i = z_i ? 0 : I * p_{i-1},  I = I * c_i
// i.e. 3(k-1) multiplications and 2k selects instead of k inversions (an inversion is ~380 multiplications,
// fr.cuh mont_inv).  0^-1 stays 0, as for the unbatched INV.  Returns the number of inversions left.
struct BatchInvStats {
    uint64_t inv_before = 0, inv_after = 0, stages = 0;
};

inline BatchInvStats batch_inversions(Tracer &tr, uint32_t max_batch = 256) {
    std::vector<SOp> &ops = tr.ops;
    const size_t N = ops.size();
    BatchInvStats st;
    std::vector<uint32_t> lev(N, 0);
    uint32_t max_lev = 0;
    auto lev_of = [&](uint32_t r) -> uint32_t { return (r == NO_REF || (r & CONST_FLAG)) ? 0u : lev[r]; };
    for (size_t i = 0; i < N; i++) {
        const SOp &o = ops[i];
        uint32_t l = std::max(lev_of(o.a), std::max(lev_of(o.b), lev_of(o.c)));
        if (o.op == T_INV) { l++; st.inv_before++; }
        lev[i] = l;
        max_lev = std::max(max_lev, l);
    }
    st.inv_after = st.inv_before;
    if (st.inv_before < 2) return st;
    // stage lists: non-INV ops by their level; INVs by the level of their operand (= lev - 1)
    std::vector<std::vector<uint32_t>> plain(max_lev + 1), invs(max_lev + 1);
    bool any = false;
    for (size_t i = 0; i < N; i++) {
        if (ops[i].op == T_INV) {
            invs[lev[i] - 1].push_back((uint32_t)i);
            if (invs[lev[i] - 1].size() >= 2) any = true;
        } else plain[lev[i]].push_back((uint32_t)i);
    }
    if (!any) return st;
    std::vector<SOp> out;
    std::vector<uint8_t> outbool;
    out.reserve(N + 8 * st.inv_before);
    std::vector<uint32_t> remap(N, NO_REF);
    auto mapref = [&](uint32_t r) -> uint32_t {
        if (r == NO_REF || (r & CONST_FLAG)) return r;
        return remap[r];
    };
    auto push = [&](uint8_t op, uint32_t a, uint32_t b, uint32_t c, uint32_t aux, bool isb) -> uint32_t {
        out.push_back(SOp{op, a, b, c, aux});
        outbool.push_back(isb ? 1 : 0);
        return (uint32_t)out.size() - 1;
    };
    const uint32_t zero = tr.zero_ref(), one = tr.one_ref();
    st.inv_after = 0;
    for (uint32_t s = 0; s <= max_lev; s++) {
        for (uint32_t i : plain[s]) {
            const SOp &o = ops[i];
            remap[i] = push(o.op, mapref(o.a), mapref(o.b), mapref(o.c), o.aux, tr.isbool[i] != 0);
        }
        const std::vector<uint32_t> &g = invs[s];
        if (!g.empty()) st.stages++;
        for (size_t g0 = 0; g0 < g.size(); g0 += max_batch) {
            size_t k = std::min<size_t>(max_batch, g.size() - g0);
            st.inv_after++;
            if (k == 1) {
                const SOp &o = ops[g[g0]];
                remap[g[g0]] = push(T_INV, mapref(o.a), NO_REF, NO_REF, 0, false);
                continue;
            }
            std::vector<uint32_t> z(k), c(k), p(k);
            for (size_t j = 0; j < k; j++) {
                uint32_t b = mapref(ops[g[g0 + j]].a);
                z[j] = push(T_EQZ, b, NO_REF, NO_REF, 0, true);
                c[j] = push(T_SEL, z[j], one, b, 0, false);
                p[j] = j ? push(T_MUL, p[j - 1], c[j], NO_REF, 0, false) : c[j];
            }
            uint32_t I = push(T_INV, p[k - 1], NO_REF, NO_REF, 0, false);
            for (size_t j = k; j-- > 0;) {
                uint32_t inv = j ? push(T_MUL, I, p[j - 1], NO_REF, 0, false) : I;
                remap[g[g0 + j]] = push(T_SEL, z[j], zero, inv, 0, false);
                if (j) I = push(T_MUL, I, c[j], NO_REF, 0, false);
            }
        }
    }
    for (uint32_t &r : tr.witness_ref) r = mapref(r);
    ops.swap(out);
    tr.isbool.swap(outbool);
    return st;
}

// ---- dot-product fusion ------------------------------------------------------------------------------------
// A tree of ff.add whose leaves are single-use products by constants (MDS mixing layers, linear combinations
// computed in vars) becomes one DOT: sum_k c_k * x_k (+ addend) evaluated with ONE Montgomery reduction
// (fr.cuh wide_mac / wide_reduce: 64 multiply-accumulates per term + 72 per reduction instead of 136 per term).
struct XOp {
    uint8_t op;
    uint32_t a, b, c, aux;
    uint32_t t0 = 0, tn = 0;   // T_DOT: terms [t0, t0+tn) of XProg::terms; c = addend ref or NO_REF
    // fused R1CS check (fused.hpp): the result is compared with a value instead of kept -- T_ADD / T_SUB / T_MUL: with c,
    // T_DOT: with a -- and constraint `aux` is recorded in first_bad when they differ
    uint8_t chk = 0;
};
struct XProg {
    std::vector<XOp> ops;
    std::vector<std::pair<uint32_t, uint32_t>> terms;   // (constant ref, value id)
    std::vector<uint32_t> witness_ref;
    std::vector<uint8_t> isbool;   // per op: the tracer proved the value 0/1 (fused results: not typed)
    std::vector<uint8_t> isint;    // per op: a raw 64-bit integer (type_ints), only consumed by integer operations
    // group_bit_ops: ops [i, i + group_len[i]) form one warp-cooperative group instruction (0 / absent: none starts at i)
    std::vector<uint8_t> group_len;
};

inline XProg fuse_dots(const Tracer &tr, uint32_t max_terms, bool enable) {
    const std::vector<SOp> &ops = tr.ops;
    const size_t N = ops.size();
    XProg xp;
    std::vector<uint8_t> live(N, 0);
    for (size_t i = 0; i < N; i++)
        if (ops[i].op == T_FAIL_IF || ops[i].op == T_FAIL_NE) live[i] = 1;
    std::vector<uint32_t> uses(N, 0);
    for (uint32_t r : tr.witness_ref)
        if (!(r & CONST_FLAG)) { live[r] = 1; uses[r] += 2; }   // a wire must exist as a value of its own
    for (size_t i = N; i-- > 0;) {
        if (!live[i]) continue;
        const SOp &o = ops[i];
        uint32_t rs[3] = {o.a, o.b, o.c};
        for (uint32_t r : rs)
            if (r != NO_REF && !(r & CONST_FLAG)) { live[r] = 1; uses[r]++; }
    }
    std::vector<uint8_t> absorbed(N, 0);
    struct Fused {
        std::vector<std::pair<uint32_t, uint32_t>> terms;
        std::vector<uint32_t> addends;
    };
    std::unordered_map<uint32_t, Fused> roots;
    if (enable && max_terms >= 2) {
        std::vector<uint32_t> stack, nodes;
        for (size_t i = N; i-- > 0;) {
            if (!live[i] || absorbed[i] || ops[i].op != T_ADD) continue;
            Fused f;
            nodes.clear();
            stack.assign(1, (uint32_t)i);
            while (!stack.empty()) {
                uint32_t n = stack.back();
                stack.pop_back();
                const SOp &o = ops[n];
                uint32_t ch[2] = {o.a, o.b};
                for (uint32_t r : ch) {
                    if (r & CONST_FLAG) { f.addends.push_back(r); continue; }
                    const SOp &c = ops[r];
                    bool single = uses[r] == 1;
                    if (single && c.op == T_ADD) { nodes.push_back(r); stack.push_back(r); continue; }
                    if (single && c.op == T_MUL && ((c.a & CONST_FLAG) != 0) != ((c.b & CONST_FLAG) != 0)) {
                        nodes.push_back(r);
                        f.terms.emplace_back((c.a & CONST_FLAG) ? c.a : c.b, (c.a & CONST_FLAG) ? c.b : c.a);
                        continue;
                    }
                    f.addends.push_back(r);
                }
            }
            if (f.terms.size() < 2) continue;
            for (uint32_t n : nodes) absorbed[n] = 1;
            roots.emplace((uint32_t)i, std::move(f));
        }
    }
    // acc + (bit ? constant : 0): an ADD (not part of a dot product) whose operand is a single-use select between a
    // constant and 0 -- the `lin += in[k] * 2^k` of circomlib's Bits2Num / BinSum once 0/1 typing turned the product into
    // a select.  cadd[i] = which operand (0: a, 1: b) is the select.
    std::unordered_map<uint32_t, int> cadd;
    if (enable) {
        // The select may have other users (the same bit * 2^k enters several sums of a hash round): every ADD that takes it
        // re-does the select from the bit -- which costs a CADD nothing -- and the select itself disappears once no other
        // kind of user is left.  Otherwise such selects stay live as 32-byte field values across whole rounds.
        auto is_bit_times_const = [&](uint32_t r) {
            if (r == NO_REF || (r & CONST_FLAG) || absorbed[r]) return false;
            const SOp &sel = ops[r];
            return sel.op == T_SEL && !(sel.a & CONST_FLAG) && (sel.b & CONST_FLAG) && (sel.c & CONST_FLAG) &&
                   fr::is_zero(tr.consts[sel.c & ~CONST_FLAG]);
        };
        std::vector<uint32_t> rem(uses);
        for (size_t i = 0; i < N; i++) {
            if (!live[i] || absorbed[i] || ops[i].op != T_ADD || roots.count((uint32_t)i)) continue;
            for (int k = 0; k < 2; k++) {
                uint32_t r = k ? ops[i].b : ops[i].a;
                if (!is_bit_times_const(r)) continue;
                rem[r]--;
                cadd.emplace((uint32_t)i, k);
                break;
            }
        }
        for (size_t r = 0; r < N; r++)
            if (live[r] && rem[r] == 0 && is_bit_times_const((uint32_t)r)) absorbed[r] = 1;
    }
    std::vector<uint32_t> remap(N, NO_REF);
    auto mapref = [&](uint32_t r) -> uint32_t { return (r == NO_REF || (r & CONST_FLAG)) ? r : remap[r]; };
    xp.ops.reserve(N);
    for (size_t i = 0; i < N; i++) {
        if (!live[i] || absorbed[i]) continue;
        const SOp &o = ops[i];
        auto it = roots.find((uint32_t)i);
        if (it == roots.end()) {
            auto ca = cadd.find((uint32_t)i);
            if (ca != cadd.end()) {   // acc + (bit ? constant : 0)
                const SOp &sel = ops[ca->second ? o.b : o.a];
                XOp x{T_CADD, mapref(ca->second ? o.a : o.b), mapref(sel.a), sel.b, 0};
                xp.ops.push_back(x);
                remap[i] = (uint32_t)xp.ops.size() - 1;
                continue;
            }
            XOp x{o.op, mapref(o.a), mapref(o.b), mapref(o.c), o.aux};
            xp.ops.push_back(x);
            remap[i] = (uint32_t)xp.ops.size() - 1;
            continue;
        }
        Fused &f = it->second;
        uint32_t v = NO_REF;
        size_t next_addend = 0;
        // constant addends fold into one constant is not possible here (no field arithmetic on this side of the
        // tracer): they are applied one by one like the others
        for (size_t t0 = 0; t0 < f.terms.size(); t0 += max_terms) {
            size_t n = std::min<size_t>(max_terms, f.terms.size() - t0);
            XOp x{T_DOT, NO_REF, NO_REF, NO_REF, 0};
            x.t0 = (uint32_t)xp.terms.size();
            x.tn = (uint32_t)n;
            for (size_t k = 0; k < n; k++) xp.terms.emplace_back(f.terms[t0 + k].first, mapref(f.terms[t0 + k].second));
            if (v != NO_REF) x.c = v;
            else if (next_addend < f.addends.size()) x.c = mapref(f.addends[next_addend++]);
            xp.ops.push_back(x);
            v = (uint32_t)xp.ops.size() - 1;
        }
        for (; next_addend < f.addends.size(); next_addend++) {
            XOp x{T_ADD, v, mapref(f.addends[next_addend]), NO_REF, 0};
            xp.ops.push_back(x);
            v = (uint32_t)xp.ops.size() - 1;
        }
        remap[i] = v;
    }
    xp.witness_ref.reserve(tr.witness_ref.size());
    for (uint32_t r : tr.witness_ref) xp.witness_ref.push_back(mapref(r));
    xp.isbool.assign(xp.ops.size(), 0);
    for (size_t i = 0; i < N; i++)
        if (remap[i] != NO_REF && roots.find((uint32_t)i) == roots.end() && cadd.find((uint32_t)i) == cadd.end())
            xp.isbool[remap[i]] = tr.ref_is_bool((uint32_t)i) ? 1 : 0;
    return xp;
}

// ---- small-integer typing ---------------------------------------------------------------------------------
// circomlib's BinSum / Bits2Num compute  lin = sum_k bit_k * 2^k  in the field and then take it apart again with
// (lin >> k) & 1.  Every value on that path is a small non-negative integer; as field elements each step costs a
// 256-bit modular addition and each decomposition a Montgomery product (to read the canonical limbs).  This pass finds
// the values that (1) are provably below 2^62 -- sums of 0/1 values times small constants -- and (2) are ONLY consumed
// by other such sums, by bit extractions and by equality asserts between two of them, and retypes them as raw 64-bit
// integers: T_CADD -> T_ICADD, T_ADD -> T_IADD, T_SEL -> T_ISEL, BITC(x * R^-1, k) -> IBIT(x, k) (the product dies),
// FAIL_NE -> IFAIL_NE.  A value with any other consumer (field arithmetic, a witness wire) stays a field element, so
// no conversion is ever needed; the result of every wire is bit-identical.
inline void type_ints(XProg &xp, const Tracer &tr) {
    std::vector<XOp> &ops = xp.ops;
    const size_t N = ops.size();
    xp.isint.assign(N, 0);
    const uint64_t LIM = 1ull << 62;
    auto small_const = [&](uint32_t r, uint64_t &v) -> bool {
        if (r == NO_REF || !(r & CONST_FLAG)) return false;
        const fr::Fr &c = tr.consts[r & ~CONST_FLAG];
        for (int i = 2; i < 8; i++)
            if (c.v[i]) return false;
        v = ((uint64_t)c.v[1] << 32) | c.v[0];
        return v < LIM;
    };
    std::vector<uint8_t> cand(N, 0);
    std::vector<uint64_t> bound(N, 0);
    // integer view of a ref under the current candidates: constants below 2^62, 0/1 values, candidate integers
    auto intable = [&](uint32_t r, uint64_t &b) -> bool {
        if (r == NO_REF) return false;
        if (r & CONST_FLAG) return small_const(r, b);
        if (xp.isbool[r]) { b = 1; return true; }
        if (cand[r]) { b = bound[r]; return true; }
        return false;
    };
    auto eval = [&](size_t i) -> bool {   // can op i be an integer op, given its operands?  sets bound[i]
        const XOp &o = ops[i];
        uint64_t x, y;
        if (xp.isbool[i]) return false;
        switch (o.op) {
            case T_CADD: {
                uint64_t k;
                if (!intable(o.a, x) || !small_const(o.c, k) || x + k >= LIM) return false;
                bound[i] = x + k;
                return true;
            }
            case T_ADD:
                if (!intable(o.a, x) || !intable(o.b, y) || x + y >= LIM) return false;
                bound[i] = x + y;
                return true;
            case T_SEL:
                if (!intable(o.b, x) || !intable(o.c, y)) return false;
                bound[i] = std::max(x, y);
                return true;
            default: return false;
        }
    };
    for (size_t i = 0; i < N; i++) cand[i] = eval(i) ? 1 : 0;
    // uses
    std::vector<uint32_t> use_cnt(N + 1, 0);
    auto each_operand = [&](const XOp &o, auto &&fn) {
        if (o.op == T_DOT)
            for (uint32_t k = 0; k < o.tn; k++) fn(xp.terms[o.t0 + k].second, 3);
        fn(o.a, 0);
        fn(o.b, 1);
        fn(o.c, 2);
    };
    for (size_t i = 0; i < N; i++)
        each_operand(ops[i], [&](uint32_t r, int) { if (r != NO_REF && !(r & CONST_FLAG)) use_cnt[r + 1]++; });
    for (size_t i = 0; i < N; i++) use_cnt[i + 1] += use_cnt[i];
    std::vector<uint32_t> use_op(use_cnt[N]);
    {
        std::vector<uint32_t> fill(use_cnt.begin(), use_cnt.end() - 1);
        for (size_t i = 0; i < N; i++)
            each_operand(ops[i], [&](uint32_t r, int) { if (r != NO_REF && !(r & CONST_FLAG)) use_op[fill[r]++] = (uint32_t)i; });
    }
    std::vector<uint8_t> on_wire(N, 0);
    for (uint32_t r : xp.witness_ref)
        if (!(r & CONST_FLAG)) on_wire[r] = 1;
    const uint32_t rinv = tr.rinv_ref();
    // x * R^-1 whose only consumers are bit extractions: with an integer x those read the integer directly
    auto is_canon_for_bits = [&](uint32_t u, uint32_t v) -> bool {
        const XOp &o = ops[u];
        if (o.op != T_MUL || on_wire[u]) return false;
        if (!((o.a == v && o.b == rinv) || (o.b == v && o.a == rinv))) return false;
        for (uint32_t k = use_cnt[u]; k < use_cnt[u + 1]; k++)
            if (ops[use_op[k]].op != T_BITC || ops[use_op[k]].a != u) return false;
        return use_cnt[u + 1] > use_cnt[u];
    };
    auto uses_ok = [&](uint32_t v) -> bool {
        if (on_wire[v]) return false;
        for (uint32_t k = use_cnt[v]; k < use_cnt[v + 1]; k++) {
            const uint32_t u = use_op[k];
            const XOp &o = ops[u];
            if (cand[u]) {
                // as an integer operand (not as the condition of a select / conditional add)
                if (o.op == T_CADD && o.a == v && o.b != v) continue;
                if (o.op == T_ADD) continue;
                if (o.op == T_SEL && o.a != v) continue;
                return false;
            }
            if (o.op == T_FAIL_NE) {
                uint64_t x, y;
                if (intable(o.a, x) && intable(o.b, y)) continue;
                return false;
            }
            if (is_canon_for_bits(u, v)) continue;
            return false;
        }
        return true;
    };
    for (bool changed = true; changed;) {
        changed = false;
        for (size_t i = 0; i < N; i++)
            if (cand[i] && !(eval(i) && uses_ok((uint32_t)i))) { cand[i] = 0; changed = true; }
        for (size_t i = N; i-- > 0;)
            if (cand[i] && !(eval(i) && uses_ok((uint32_t)i))) { cand[i] = 0; changed = true; }
    }
    // rewrite
    for (size_t i = 0; i < N; i++) {
        XOp &o = ops[i];
        if (cand[i]) {
            o.op = o.op == T_CADD ? T_ICADD : o.op == T_ADD ? T_IADD : T_ISEL;
            xp.isint[i] = 1;
        } else if (o.op == T_BITC && !(o.a & CONST_FLAG)) {
            const XOp &m = ops[o.a];
            if (m.op == T_MUL && (m.a == rinv || m.b == rinv)) {
                const uint32_t x = m.a == rinv ? m.b : m.a;
                if (x != NO_REF && !(x & CONST_FLAG) && cand[x]) {
                    o.op = T_IBIT;
                    o.a = x;
                }
            }
        } else if (o.op == T_FAIL_NE) {
            const bool ia = !(o.a & CONST_FLAG) && cand[o.a], ib = !(o.b & CONST_FLAG) && cand[o.b];
            uint64_t x, y;
            if ((ia || ib) && intable(o.a, x) && intable(o.b, y)) o.op = T_IFAIL_NE;
        }
    }
    // drop what died (the x * R^-1 products of retyped values)
    std::vector<uint8_t> live(N, 0);
    for (uint32_t r : xp.witness_ref)
        if (!(r & CONST_FLAG)) live[r] = 1;
    for (size_t i = N; i-- > 0;) {
        const XOp &o = ops[i];
        if (o.op == T_FAIL_IF || o.op == T_FAIL_NE || o.op == T_IFAIL_NE) live[i] = 1;
        if (!live[i]) continue;
        each_operand(o, [&](uint32_t r, int) { if (r != NO_REF && !(r & CONST_FLAG)) live[r] = 1; });
    }
    std::vector<uint32_t> remap(N, NO_REF);
    auto mapref = [&](uint32_t r) -> uint32_t { return (r == NO_REF || (r & CONST_FLAG)) ? r : remap[r]; };
    std::vector<XOp> out;
    std::vector<uint8_t> ob, oi;
    out.reserve(N);
    for (size_t i = 0; i < N; i++) {
        if (!live[i]) continue;
        XOp o = ops[i];
        if (o.op == T_DOT)
            for (uint32_t k = 0; k < o.tn; k++) xp.terms[o.t0 + k].second = mapref(xp.terms[o.t0 + k].second);
        o.a = mapref(o.a);
        o.b = mapref(o.b);
        o.c = mapref(o.c);
        remap[i] = (uint32_t)out.size();
        out.push_back(o);
        ob.push_back(xp.isbool[i]);
        oi.push_back(xp.isint[i]);
    }
    for (uint32_t &r : xp.witness_ref) r = mapref(r);
    ops.swap(out);
    xp.isbool.swap(ob);
    xp.isint.swap(oi);
}

// ---- integer sums ------------------------------------------------------------------------------------------
// lin = sum_k bit_k * 2^k arrives as a chain of T_ICADD, each result used only by the next link.  One T_ISUM evaluates
// up to ISUM_MAX links: the accumulator stays in a register instead of going through its slot once per bit, and the
// instruction fetch / dispatch of the interpreter is paid once.  Links whose constant is not a power of two, or whose
// condition is not a value typed 0/1, stay T_ICADD.
static const uint32_t ISUM_MAX = 128;

inline void fuse_isums(XProg &xp, const Tracer &tr) {
    std::vector<XOp> &ops = xp.ops;
    const size_t N = ops.size();
    std::vector<uint32_t> uses(N, 0);
    for (uint32_t r : xp.witness_ref)
        if (!(r & CONST_FLAG)) uses[r] += 2;
    for (size_t i = 0; i < N; i++) {
        const XOp &o = ops[i];
        auto use = [&](uint32_t r) { if (r != NO_REF && !(r & CONST_FLAG)) uses[r]++; };
        if (o.op == T_DOT)
            for (uint32_t k = 0; k < o.tn; k++) use(xp.terms[o.t0 + k].second);
        use(o.a); use(o.b); use(o.c);
    }
    auto shift_of = [&](uint32_t cref, uint32_t &sh) -> bool {   // constant = 2^sh, sh < 62
        const fr::Fr &c = tr.consts[cref & ~CONST_FLAG];
        for (int i = 2; i < 8; i++)
            if (c.v[i]) return false;
        const uint64_t v = ((uint64_t)c.v[1] << 32) | c.v[0];
        if (v == 0 || (v & (v - 1)) || v >= (1ull << 62)) return false;
        sh = 0;
        while (!((v >> sh) & 1)) sh++;
        return true;
    };
    auto fusable = [&](size_t i) -> bool {
        const XOp &o = ops[i];
        uint32_t sh;
        return o.op == T_ICADD && !(o.b & CONST_FLAG) && xp.isbool[o.b] && (o.c & CONST_FLAG) && shift_of(o.c, sh);
    };
    std::vector<uint8_t> absorbed(N, 0);
    // walk chains from their last link backwards
    for (size_t i = N; i-- > 0;) {
        if (absorbed[i] || !fusable(i)) continue;
        std::vector<uint32_t> chain;   // links, last first
        uint32_t cur = (uint32_t)i;
        while (chain.size() < ISUM_MAX) {
            chain.push_back(cur);
            const uint32_t prev = ops[cur].a;
            if (prev == NO_REF || (prev & CONST_FLAG) || absorbed[prev] || uses[prev] != 1 || !fusable(prev)) break;
            cur = prev;
        }
        if (chain.size() < 2) continue;
        XOp x{T_ISUM, NO_REF, NO_REF, ops[chain.back()].a, 0};
        x.t0 = (uint32_t)xp.terms.size();
        x.tn = (uint32_t)chain.size();
        for (size_t k = chain.size(); k-- > 0;) {
            xp.terms.emplace_back(ops[chain[k]].c, ops[chain[k]].b);
            if (k) absorbed[chain[k]] = 1;
        }
        ops[i] = x;
    }
    // Sums of sums: a + b of the 32-bit words of a hash round arrives as T_IADD of two such sums (or of a sum and another
    // integer).  The terms do not care which instruction adds them: a single-use T_ISUM operand is merged into the addition,
    // which becomes one T_ISUM over all the terms (the other operand, if it is not a sum, is its addend).  Cascades: the merged
    // sum is itself a single-use T_ISUM for the next addition of the chain.
    for (size_t i = 0; i < N; i++) {
        if (absorbed[i] || ops[i].op != T_IADD) continue;
        const XOp o = ops[i];
        auto is_sum = [&](uint32_t r) {
            return r != NO_REF && !(r & CONST_FLAG) && !absorbed[r] && ops[r].op == T_ISUM && uses[r] == 1;
        };
        const bool sa = is_sum(o.a), sb = is_sum(o.b) && o.b != o.a;
        if (!sa && !sb) continue;
        uint32_t addend = NO_REF, addend2 = NO_REF;   // a T_ISUM carries up to two (c and a)
        std::vector<std::pair<uint32_t, uint32_t>> terms;
        bool ok = true;
        auto other = [&](uint32_t r) {
            if (addend == NO_REF) addend = r;
            else if (addend2 == NO_REF) addend2 = r;
            else ok = false;
        };
        auto take = [&](uint32_t r) {
            for (uint32_t k = 0; k < ops[r].tn; k++) terms.push_back(xp.terms[ops[r].t0 + k]);
            for (uint32_t c : {ops[r].c, ops[r].a})
                if (c != NO_REF && !((c & CONST_FLAG) && fr::is_zero(tr.consts[c & ~CONST_FLAG]))) other(c);   // (a chain usually starts from 0)
        };
        if (sa) take(o.a); else other(o.a);
        if (sb) take(o.b); else other(o.b);
        if (!ok || terms.size() > ISUM_MAX || terms.size() < 2) continue;
        XOp x{T_ISUM, addend2, NO_REF, addend, 0};
        x.t0 = (uint32_t)xp.terms.size();
        x.tn = (uint32_t)terms.size();
        for (const auto &t : terms) xp.terms.push_back(t);
        if (sa) absorbed[o.a] = 1;
        if (sb) absorbed[o.b] = 1;
        ops[i] = x;
    }
    // compact
    std::vector<uint32_t> remap(N, NO_REF);
    auto mapref = [&](uint32_t r) -> uint32_t { return (r == NO_REF || (r & CONST_FLAG)) ? r : remap[r]; };
    std::vector<XOp> out;
    std::vector<uint8_t> ob, oi;
    out.reserve(N);
    for (size_t i = 0; i < N; i++) {
        if (absorbed[i]) continue;
        XOp o = ops[i];
        if (o.op == T_DOT || o.op == T_ISUM)
            for (uint32_t k = 0; k < o.tn; k++) xp.terms[o.t0 + k].second = mapref(xp.terms[o.t0 + k].second);
        o.a = mapref(o.a);
        o.b = mapref(o.b);
        o.c = mapref(o.c);
        remap[i] = (uint32_t)out.size();
        out.push_back(o);
        ob.push_back(xp.isbool[i]);
        oi.push_back(xp.isint[i]);
    }
    for (uint32_t &r : xp.witness_ref) r = mapref(r);
    ops.swap(out);
    xp.isbool.swap(ob);
    xp.isint.swap(oi);
}

// ---- warp-cooperative groups ---------------------------------------------------------------------------------
// A value typed 0/1 is one packed word per warp, so a boolean function of such values can be evaluated for the warp's 32
// witnesses by ONE lane with a handful of bitwise instructions -- and 32 independent functions by the 32 lanes at
// once.  This pass defers every T_LUT until something needs one of the deferred results (or 32 are waiting) and emits
// them as one group; likewise runs of T_IBIT that extract consecutive bits of the same integer.  Deferring is safe:
// values are in SSA form, the deferred operations have no side effects, and the group is placed before the first
// consumer of any of its results.  Sha256(512): 32 808 look-ups -> ~1 100 group instructions.
inline void group_bit_ops(XProg &xp) {
    std::vector<XOp> &ops = xp.ops;
    const size_t N = ops.size();
    std::vector<uint32_t> order;
    std::vector<uint8_t> glen;
    order.reserve(N);
    glen.reserve(N);
    std::vector<uint32_t> open_lut, open_bit, open_in;
    std::vector<uint8_t> deferred(N, 0);   // 1: in open_lut, 2: in open_bit, 3: in open_in
    auto flush = [&](std::vector<uint32_t> &open) {
        if (open.empty()) return;
        const size_t at = order.size();
        for (uint32_t i : open) {
            order.push_back(i);
            glen.push_back(0);
            deferred[i] = 0;
        }
        if (open.size() >= 2) glen[at] = (uint8_t)open.size();
        open.clear();
    };
    for (size_t i = 0; i < N; i++) {
        const XOp &o = ops[i];
        bool use_lut = false, use_bit = false, use_in = false;
        auto chk = [&](uint32_t r) {
            if (r == NO_REF || (r & CONST_FLAG)) return;
            if (deferred[r] == 1) use_lut = true;
            if (deferred[r] == 2) use_bit = true;
            if (deferred[r] == 3) use_in = true;
        };
        if (o.op == T_DOT || o.op == T_ISUM)
            for (uint32_t k = 0; k < o.tn; k++) chk(xp.terms[o.t0 + k].second);
        chk(o.a); chk(o.b); chk(o.c);
        if (use_lut) flush(open_lut);
        if (use_bit) flush(open_bit);
        if (use_in) flush(open_in);
        if (o.op == T_INPUT_BIT) {   // runs of consecutive main inputs taken as bits (speculative typing): one coalesced group load
            if (!open_in.empty() && ops[open_in.back()].aux + 1 != o.aux) flush(open_in);
            open_in.push_back((uint32_t)i);
            deferred[i] = 3;
            if (open_in.size() == 32) flush(open_in);
        } else if (o.op == T_LUT) {
            open_lut.push_back((uint32_t)i);
            deferred[i] = 1;
            if (open_lut.size() == 32) flush(open_lut);
        } else if (o.op == T_IBIT) {
            if (!open_bit.empty()) {
                const XOp &l = ops[open_bit.back()];
                if (l.a != o.a || l.aux + 1 != o.aux) flush(open_bit);
            }
            open_bit.push_back((uint32_t)i);
            deferred[i] = 2;
            if (open_bit.size() == 32) flush(open_bit);
        } else {
            order.push_back((uint32_t)i);
            glen.push_back(0);
        }
    }
    flush(open_lut);
    flush(open_bit);
    flush(open_in);
    std::vector<uint32_t> remap(N, NO_REF);
    for (size_t k = 0; k < N; k++) remap[order[k]] = (uint32_t)k;
    auto mapref = [&](uint32_t r) -> uint32_t { return (r == NO_REF || (r & CONST_FLAG)) ? r : remap[r]; };
    std::vector<XOp> out(N);
    std::vector<uint8_t> ob(N), oi(N);
    for (size_t k = 0; k < N; k++) {
        XOp o = ops[order[k]];
        if (o.op == T_DOT || o.op == T_ISUM)
            for (uint32_t t = 0; t < o.tn; t++) xp.terms[o.t0 + t].second = mapref(xp.terms[o.t0 + t].second);
        o.a = mapref(o.a);
        o.b = mapref(o.b);
        o.c = mapref(o.c);
        out[k] = o;
        ob[k] = xp.isbool[order[k]];
        oi[k] = xp.isint[order[k]];
    }
    for (uint32_t &r : xp.witness_ref) r = mapref(r);
    ops.swap(out);
    xp.isbool.swap(ob);
    xp.isint.swap(oi);
    xp.group_len.swap(glen);
}

// records that follow an instruction on the tape (term / member lists)
inline uint32_t extra_records(const TapeIns &in) {
    if (in.op == T_DOT) return (in.a + 1) / 2;
    if (in.op == T_ISUM) return (in.a + 3) / 4;
    if (in.op == T_ISUMT) return in.a * 8;
    if (in.op == T_LUTG || in.op == T_IBITG || in.op == T_INBITG) return in.a;
    return 0;
}

// ---- reload stream ---------------------------------------------------------------------------------------
// The FIELD rows a tape reloads (T_LD into a field slot) and their order are fixed, so the kernel streams them: the
// n-th reload's row is requested (cp.async into a per-witness ring in shared memory) when reload n - LD_RING executes,
// and reload n finds it there -- HBM latency leaves the dependence chain, which is what small batches (few resident
// warps) are bound by.  A reload is only streamed when the store that produced its row precedes the request point
// (spill rows are recycled); the others load directly.  Bit-row reloads (one word per warp) are not streamed.
inline bool is_field_ld(const TapeIns &in) { return in.op == T_LD && !(in.dst & BSLOT_DST); }

inline void schedule_reloads(Tape &t) {
    std::vector<uint32_t> ld_pos, ld_store;
    std::vector<uint32_t> last_store(t.n_frows, NO_ROW);
    for (size_t pc = 0; pc < t.ins.size(); pc++) {
        TapeIns &in = t.ins[pc];
        if (extra_records(in)) {
            if ((in.flags & F_STORE) && !(in.c & ROW_BIT)) last_store[in.c] = (uint32_t)pc;
            pc += extra_records(in);
            continue;
        }
        if (in.op == T_LD) {
            if (!is_field_ld(in)) continue;
            in.a = NO_ROW;
            in.b = (uint32_t)(ld_pos.size() % LD_RING);
            ld_pos.push_back((uint32_t)pc);
            ld_store.push_back(last_store[in.c]);
        } else if (in.op == T_ST || in.op == T_STC || (in.flags & F_STORE)) {
            if (in.op != T_FAIL_IF && in.op != T_FAIL_NE && !(in.c & ROW_BIT)) last_store[in.c] = (uint32_t)pc;
        }
    }
    for (size_t n = LD_RING; n < ld_pos.size(); n++) {
        const uint32_t issue_at = ld_pos[n - LD_RING];
        if (ld_store[n] != NO_ROW && ld_store[n] < issue_at) {
            t.ins[ld_pos[n]].flags |= F_RING;
            t.ins[issue_at].a = t.ins[ld_pos[n]].c;
            t.stats.n_ld_streamed++;
        }
    }
}

// Largest number of simultaneously live values of each kind (0: field, 1: 0/1-typed) in program order: what an
// allocator with unlimited slots would need.  Sizes the bit-slot file and bounds the useful number of field slots.
inline void max_live_by_kind(const XProg &xp, uint32_t out[2]) {
    const size_t N = xp.ops.size();
    std::vector<uint32_t> last(N);
    for (size_t i = 0; i < N; i++) {
        const XOp &o = xp.ops[i];
        last[i] = (uint32_t)i;
        auto use = [&](uint32_t r) {
            if (r != NO_REF && !(r & CONST_FLAG)) last[r] = (uint32_t)i;
        };
        if (o.op == T_DOT || o.op == T_ISUM)
            for (uint32_t k = 0; k < o.tn; k++) use(xp.terms[o.t0 + k].second);
        use(o.a);
        use(o.b);
        use(o.c);
    }
    auto produces = [&](size_t i) {
        const uint8_t op = xp.ops[i].op;
        return op != T_FAIL_IF && op != T_FAIL_NE && op != T_IFAIL_NE && op != T_RNE && !xp.ops[i].chk;
    };
    std::vector<uint32_t> deaths[2];
    deaths[0].assign(N, 0);
    deaths[1].assign(N, 0);
    for (size_t v = 0; v < N; v++)
        if (produces(v)) deaths[xp.isbool[v] ? 1 : 0][last[v]]++;
    uint32_t live[2] = {0, 0};
    out[0] = out[1] = 0;
    size_t peak_at = 0;
    for (size_t i = 0; i < N; i++) {
        if (produces(i)) {
            const int k = xp.isbool[i] ? 1 : 0;
            live[k]++;
            if (k == 0 && live[0] > out[0]) peak_at = i;
            out[k] = std::max(out[k], live[k]);
        }
        live[0] -= deaths[0][i];
        live[1] -= deaths[1][i];
    }
    if (getenv("CVMGPU_DEBUG_LIVE")) {
        uint32_t hist[T_COUNT] = {0};
        for (size_t v = 0; v <= peak_at; v++)
            if (produces(v) && !xp.isbool[v] && last[v] > peak_at) hist[xp.ops[v].op]++;
        fprintf(stderr, "field values live at op %zu of %zu:", peak_at, N);
        for (int o = 0; o < T_COUNT; o++)
            if (hist[o]) fprintf(stderr, " op%d:%u", o, hist[o]);
        fprintf(stderr, "\n");
    }
}

// Typed allocation.  Values the tracer proved 0/1 live in the BIT file: one 32-bit word per warp and slot (bit = lane =
// witness), spilled to / reloaded from bit rows of the value store (one word per warp).  Everything else lives in the
// FIELD file (32-byte slots per witness) and in field rows.  A witness wire is a bit row or a field row according to
// the type of the value bound to it (Tape::wire_loc); consumers of a 0/1 value in field arithmetic convert on fetch.
// fusions and typing that do not depend on the slot files (max_terms: longest fused dot product)
// typed = false: every value a field element (no bit-slot file, no integer typing, no groups): programs with only a
// sprinkling of 0/1 values run the field-only kernel instantiation, which has the registers for 20 warps per SM
inline XProg prepare_program(const Tracer &tr, uint32_t max_terms, bool fuse = true, bool typed = true) {
    XProg xp = fuse_dots(tr, max_terms, fuse);
    if (!typed) {
        xp.isbool.assign(xp.ops.size(), 0);
        xp.isint.assign(xp.ops.size(), 0);
    } else if (fuse) {
        type_ints(xp, tr);
        fuse_isums(xp, tr);
        group_bit_ops(xp);
    } else xp.isint.assign(xp.ops.size(), 0);
    if (xp.group_len.size() != xp.ops.size()) xp.group_len.assign(xp.ops.size(), 0);
    return xp;
}

inline Tape allocate_tape(const std::vector<fr::Fr> &consts, size_t n_ssa, const XProg &xp, uint32_t n_slots, uint32_t max_bslots = 2048) {
    if (n_slots < 4) throw TraceError("need at least 4 slots");
    const std::vector<XOp> &ops = xp.ops;
    const size_t N = ops.size();
    Tape out;
    out.n_slots = n_slots;
    out.n_wires = (uint32_t)xp.witness_ref.size();
    out.stats.n_ssa = n_ssa;
    out.stats.n_live = N;

    auto is_bool_ref = [&](uint32_t r) -> bool {
        if (r == NO_REF) return false;
        if (r & CONST_FLAG) {
            const fr::Fr &c = consts[r & ~CONST_FLAG];
            for (int i = 1; i < 8; i++)
                if (c.v[i]) return false;
            return c.v[0] <= 1;
        }
        return xp.isbool[r] != 0;
    };
    uint32_t ml[2];
    max_live_by_kind(xp, ml);
    out.stats.max_live_field = ml[0];
    out.stats.max_live_bool = ml[1];
    // ---- witness wire -> typed row
    out.wire_loc.resize(out.n_wires);
    {
        // wire 0 is the constant 1 (calcwit.cpp:34): it keeps a field row (linear combinations evaluated in the field add
        // their constant term through it) AND gets a bit row (Tape::one_brow) for constraints evaluated in integers
        // Bit rows: first the wires bound to computed 0/1 values (in wire order), then the wires bound to the constant 0,
        // then those bound to the constant 1 and one_brow -- two contiguous ranges that a T_FILL each writes.
        uint32_t nf = 0, nb = 0;
        auto const_bit = [&](uint32_t w) -> int {   // -1: not a constant 0/1 wire
            const uint32_t r = xp.witness_ref[w];
            // (a program without computed 0/1 values keeps its constant wires as field rows: its constraints then have no
            // bit-row term at all and the check runs the plain kernel)
            if (w == 0 || ml[1] == 0 || !(r & CONST_FLAG) || !is_bool_ref(r)) return -1;
            return (int)consts[r & ~CONST_FLAG].v[0];
        };
        for (uint32_t w = 0; w < out.n_wires; w++) {
            if (const_bit(w) >= 0) continue;
            const bool as_bit = w != 0 && is_bool_ref(xp.witness_ref[w]) && (ml[1] != 0 || !(xp.witness_ref[w] & CONST_FLAG));
            out.wire_loc[w] = as_bit ? (ROW_BIT | nb++) : nf++;
        }
        out.n_fwires = nf;
        for (int v = 0; v < 2; v++) {
            out.const_rows[v][0] = nb;
            for (uint32_t w = 0; w < out.n_wires; w++)
                if (const_bit(w) == v) out.wire_loc[w] = ROW_BIT | nb++;
            if (v == 1) out.one_brow = nb++;
            out.const_rows[v][1] = nb - out.const_rows[v][0];
        }
        out.n_bwires = nb;
    }
    // the bit file: enough for every live 0/1 value when that fits (no bit spills at all), else the cap
    uint32_t n_bslots = ml[1] == 0 ? 0 : std::min<uint32_t>(max_bslots, ((ml[1] + 3 + 31) / 32) * 32);
    if (ml[1] && n_bslots < ISUM_MAX + 64) n_bslots = ISUM_MAX + 64;   // a T_ISUM pins up to ISUM_MAX bit slots at once
    out.n_bslots = n_bslots;
    const uint32_t file_slots[2] = {n_slots, n_bslots};

    auto operands = [&](const XOp &o, std::vector<uint32_t> &rs) {
        rs.clear();
        if (o.op == T_DOT || o.op == T_ISUM) {
            for (uint32_t k = 0; k < o.tn; k++) rs.push_back(xp.terms[o.t0 + k].second);
            rs.push_back(o.c);
            if (o.chk || o.op == T_ISUM) rs.push_back(o.a);   // a checked T_DOT: the value its result is compared with; T_ISUM: a second addend
        } else {
            rs.push_back(o.a);
            rs.push_back(o.b);
            rs.push_back(o.c);
        }
    };
    std::vector<uint32_t> rs;
    // ---- uses (CSR), in program order (fuse_dots already dropped dead code)
    std::vector<uint32_t> use_cnt(N + 1, 0);
    for (size_t i = 0; i < N; i++) {
        operands(ops[i], rs);
        for (uint32_t r : rs)
            if (r != NO_REF && !(r & CONST_FLAG)) use_cnt[r + 1]++;
    }
    for (size_t i = 0; i < N; i++) use_cnt[i + 1] += use_cnt[i];
    std::vector<uint32_t> use_pos(use_cnt[N]);
    {
        std::vector<uint32_t> fill(use_cnt.begin(), use_cnt.end() - 1);
        for (size_t i = 0; i < N; i++) {
            operands(ops[i], rs);
            for (uint32_t r : rs)
                if (r != NO_REF && !(r & CONST_FLAG)) use_pos[fill[r]++] = (uint32_t)i;
        }
    }
    std::vector<uint32_t> use_ptr(use_cnt.begin(), use_cnt.end() - 1);
    auto next_use = [&](uint32_t v, uint32_t pos) -> uint32_t {   // first use strictly after pos, or UINT32_MAX
        uint32_t &p = use_ptr[v];
        while (p < use_cnt[v + 1] && use_pos[p] <= pos) p++;
        return p < use_cnt[v + 1] ? use_pos[p] : 0xffffffffu;
    };
    // ---- witness wires per value
    std::vector<uint32_t> wire_head(N, NO_REF), wire_next(xp.witness_ref.size(), NO_REF);
    for (size_t w = xp.witness_ref.size(); w-- > 0;) {
        uint32_t r = xp.witness_ref[w];
        if (r & CONST_FLAG) continue;
        wire_next[w] = wire_head[r];
        wire_head[r] = (uint32_t)w;
    }
    // constants bound to wires are stored up front: the 0 / 1 wires as two filled ranges of bit rows
    for (int v = 0; v < 2; v++)
        if (out.const_rows[v][1]) {
            out.ins.push_back(TapeIns{T_FILL, 0, 0, v ? 0xffffffffu : 0u, out.const_rows[v][1], ROW_BIT | out.const_rows[v][0]});
            out.stats.n_stc++;
        }
    for (size_t w = 0; w < xp.witness_ref.size(); w++) {
        uint32_t r = xp.witness_ref[w];
        if ((r & CONST_FLAG) && !(out.wire_loc[w] & ROW_BIT)) {
            out.ins.push_back(TapeIns{T_STC, 1, 0, r & ~CONST_FLAG, 0, out.wire_loc[w]});
            out.stats.n_stc++;
        }
    }
    // ---- linear scan over two slot files (0: field, 1: bit)
    std::vector<int32_t> slot_val[2];
    std::vector<uint32_t> free_slots[2], free_spill[2];
    uint32_t spill_rows[2] = {0, 0};
    const uint32_t wire_rows[2] = {out.n_fwires, out.n_bwires};
    for (int k = 0; k < 2; k++) {
        slot_val[k].assign(file_slots[k], -1);
        for (uint32_t s = file_slots[k]; s-- > 0;) free_slots[k].push_back(s);
    }
    std::vector<int32_t> val_slot(N, -1);
    std::vector<uint32_t> val_home(N, NO_REF);      // typed row (ROW_BIT for bit rows)
    std::vector<uint8_t> home_is_spill(N, 0);
    uint32_t live_now = 0;
    auto kind_of = [&](uint32_t v) -> int { return xp.isbool[v] ? 1 : 0; };
    auto slot_code = [&](uint32_t v) -> uint32_t { return (uint32_t)val_slot[v] | (xp.isbool[v] ? BSLOT : 0u); };

    auto release_value = [&](uint32_t v) {
        const int k = kind_of(v);
        if (val_slot[v] >= 0) {
            slot_val[k][(size_t)val_slot[v]] = -1;
            free_slots[k].push_back((uint32_t)val_slot[v]);
            val_slot[v] = -1;
        }
        if (home_is_spill[v]) {
            free_spill[k].push_back((val_home[v] & ~ROW_BIT) - wire_rows[k]);
            home_is_spill[v] = 0;
        }
        live_now--;
    };
    // pinned: slot codes (with BSLOT for the bit file) that must not be evicted
    auto alloc_slot = [&](int k, uint32_t pos, const std::vector<uint32_t> &pinned) -> uint32_t {
        if (!free_slots[k].empty()) {
            uint32_t s = free_slots[k].back();
            free_slots[k].pop_back();
            return s;
        }
        const uint32_t tag = k ? BSLOT : 0u;
        uint32_t best = 0xffffffffu, best_use = 0;
        for (uint32_t s = 0; s < file_slots[k]; s++) {
            bool pin = false;
            for (uint32_t q : pinned)
                if (q == (s | tag)) pin = true;
            if (pin) continue;
            uint32_t v = (uint32_t)slot_val[k][s];
            uint32_t nu = next_use(v, pos);
            if (best == 0xffffffffu || nu > best_use) { best = s; best_use = nu; }
            if (nu == 0xffffffffu) break;
        }
        if (best == 0xffffffffu) throw TraceError("slot allocator: all slots pinned");
        uint32_t v = (uint32_t)slot_val[k][best];
        if (val_home[v] == NO_REF) {   // not yet in HBM: spill
            uint32_t row;
            if (!free_spill[k].empty()) { row = free_spill[k].back(); free_spill[k].pop_back(); }
            else row = spill_rows[k]++;
            val_home[v] = (wire_rows[k] + row) | (k ? ROW_BIT : 0u);
            home_is_spill[v] = 1;
            out.ins.push_back(TapeIns{T_ST, 0, 0, best | tag, 0, val_home[v]});
            out.stats.n_st++;
            out.stats.n_spill_st++;
            out.stats.n_spill_st_bool += (uint64_t)k;
        }
        val_slot[v] = -1;
        slot_val[k][best] = -1;
        return best;
    };

    std::vector<uint32_t> pinned, still;
    std::vector<uint32_t> enc;
    std::unordered_map<uint64_t, uint32_t> iconst_index;
    for (size_t i = 0; i < N; i++) {
        if (xp.group_len[i] >= 2) {
            // ---- warp-cooperative group (group_bit_ops): members i .. i+n-1 execute as ONE instruction, lane m doing member m.
            // All operands must be resident when it starts; results may take the slots of operands that die here (the
            // kernel reads every operand before any lane writes).
            const uint32_t n = xp.group_len[i], last = (uint32_t)i + n - 1;
            const bool lutg = ops[i].op == T_LUT, ing = ops[i].op == T_INPUT_BIT;
            std::vector<uint32_t> grs;
            for (uint32_t m = 0; m < n; m++) {
                const XOp &g = ops[i + m];
                const uint32_t refs[3] = {g.a, lutg ? g.b : NO_REF, lutg ? g.c : NO_REF};
                for (uint32_t r : refs)
                    if (r != NO_REF) {
                        if (r & CONST_FLAG) throw TraceError("group member takes a constant");
                        grs.push_back(r);
                    }
            }
            pinned.clear();
            for (uint32_t r : grs)
                if (val_slot[r] >= 0) pinned.push_back(slot_code(r));
            for (uint32_t r : grs) {
                if (val_slot[r] >= 0) continue;
                if (val_home[r] == NO_REF) throw TraceError("slot allocator: value lost");
                const int kd = kind_of(r);
                uint32_t sl = alloc_slot(kd, (uint32_t)i, pinned);
                out.ins.push_back(TapeIns{T_LD, 0, (uint16_t)(sl | (kd ? BSLOT_DST : 0)), 0, 0, val_home[r]});
                out.stats.n_ld++;
                out.stats.n_ld_bool += (uint64_t)kd;
                val_slot[r] = (int32_t)sl;
                slot_val[kd][sl] = (int32_t)r;
                pinned.push_back(slot_code(r));
            }
            std::vector<uint32_t> recs(4 * (size_t)n, 0);
            for (uint32_t m = 0; m < n; m++) {
                const XOp &g = ops[i + m];
                if (lutg) {
                    const uint32_t s0 = slot_code(g.a) & 0xffffu, s1 = g.b != NO_REF ? slot_code(g.b) & 0xffffu : 0u,
                                   s2 = g.c != NO_REF ? slot_code(g.c) & 0xffffu : 0u;
                    recs[4 * m] = s0 | (s1 << 16);
                    recs[4 * m + 1] = s2;
                    recs[4 * m + 2] = g.aux & 0xffffu;
                }
            }
            const uint32_t src_code = (lutg || ing) ? 0u : slot_code(ops[i].a);
            for (uint32_t r : grs)
                if (val_slot[r] >= 0 && next_use(r, last) == 0xffffffffu) release_value(r);
            still.clear();
            for (uint32_t r : grs)
                if (val_slot[r] >= 0) still.push_back(slot_code(r));
            for (uint32_t m = 0; m < n; m++) {
                const uint32_t v = (uint32_t)i + m;
                if (!xp.isbool[v]) throw TraceError("group member is not typed 0/1");
                const uint32_t d = alloc_slot(1, last, still);
                val_slot[v] = (int32_t)d;
                slot_val[1][d] = (int32_t)v;
                still.push_back(d | BSLOT);
                live_now++;
                out.stats.max_live = std::max(out.stats.max_live, live_now);
                uint32_t row = NO_ROW;
                if (wire_head[v] != NO_REF) {
                    row = out.wire_loc[wire_head[v]];
                    val_home[v] = row;
                    out.stats.n_st++;
                }
                if (lutg) recs[4 * m + 1] |= d << 16;
                else recs[4 * m] = d;
                recs[4 * m + 3] = row;
            }
            out.ins.push_back(TapeIns{(uint8_t)(lutg ? T_LUTG : ing ? T_INBITG : T_IBITG), 0, 0, n, src_code, lutg ? 0u : ops[i].aux});
            for (uint32_t m = 0; m < n; m++) {
                TapeIns raw;
                memcpy(&raw, &recs[4 * m], 16);
                out.ins.push_back(raw);
            }
            if (lutg) out.stats.n_lut += n;
            else out.stats.n_other += n;
            out.stats.n_groups++;
            for (uint32_t m = 0; m < n; m++) {
                const uint32_t v = (uint32_t)i + m;
                if (wire_head[v] != NO_REF)
                    for (uint32_t w = wire_next[wire_head[v]]; w != NO_REF; w = wire_next[w]) {
                        out.ins.push_back(TapeIns{T_ST, 0, 0, (uint32_t)val_slot[v] | BSLOT, 0, out.wire_loc[w]});
                        out.stats.n_st++;
                    }
                if (next_use(v, last) == 0xffffffffu) release_value(v);
            }
            i = last;
            continue;
        }
        const XOp &o = ops[i];
        uint32_t pos = (uint32_t)i;
        operands(o, rs);
        pinned.clear();
        enc.assign(rs.size(), 0);
        std::vector<uint8_t> isc(rs.size(), 0);
        // resident operands are pinned first so that loading one operand cannot evict another
        for (uint32_t r : rs) {
            if (r == NO_REF || (r & CONST_FLAG)) continue;
            if (val_slot[r] >= 0) pinned.push_back(slot_code(r));
        }
        for (size_t k = 0; k < rs.size(); k++) {
            uint32_t r = rs[k];
            if (r == NO_REF) continue;
            if (r & CONST_FLAG) { isc[k] = 1; enc[k] = r & ~CONST_FLAG; continue; }
            if (val_slot[r] < 0) {
                if (val_home[r] == NO_REF) throw TraceError("slot allocator: value lost");
                const int kd = kind_of(r);
                uint32_t s = alloc_slot(kd, pos, pinned);
                out.ins.push_back(TapeIns{T_LD, 0, (uint16_t)(s | (kd ? BSLOT_DST : 0)), 0, 0, val_home[r]});
                out.stats.n_ld++;
                out.stats.n_ld_bool += (uint64_t)kd;
                val_slot[r] = (int32_t)s;
                slot_val[kd][s] = (int32_t)r;
                pinned.push_back(slot_code(r));
            }
            enc[k] = slot_code(r);
        }
        if (o.op == T_LUT)
            for (size_t k = 0; k < 3; k++)
                if (rs[k] != NO_REF && ((rs[k] & CONST_FLAG) || !xp.isbool[rs[k]])) throw TraceError("T_LUT input is not a typed 0/1 value");
        // operands that die here free their slots before the destination is chosen
        for (size_t k = 0; k < rs.size(); k++) {
            uint32_t r = rs[k];
            if (r == NO_REF || (r & CONST_FLAG)) continue;
            if (val_slot[r] >= 0 && next_use(r, pos) == 0xffffffffu) release_value(r);
        }
        uint8_t flags = 0;
        if (o.op != T_DOT && o.op != T_ISUM)
            for (size_t k = 0; k < 3; k++)
                if (isc[k]) flags |= (uint8_t)(1u << k);
        if (o.op == T_ICADD || o.op == T_IADD || o.op == T_ISEL || o.op == T_IFAIL_NE) {
            // integer operations take their constants from the table of raw 64-bit integers
            for (size_t k = (o.op == T_ISEL ? 1 : 0); k < 3; k++) {
                if (!isc[k]) continue;
                const fr::Fr &cv = consts[enc[k]];
                const uint64_t v = ((uint64_t)cv.v[1] << 32) | cv.v[0];
                auto it = iconst_index.find(v);
                if (it == iconst_index.end()) {
                    it = iconst_index.emplace(v, (uint32_t)out.iconsts.size()).first;
                    out.iconsts.push_back(v);
                }
                enc[k] = it->second;
            }
            if (o.op != T_IFAIL_NE) out.stats.n_int++;
        }
        if (o.op == T_FAIL_IF || o.op == T_FAIL_NE || o.op == T_IFAIL_NE || o.op == T_RNE) {
            out.ins.push_back(TapeIns{o.op, flags, 0, enc[0], enc[1], o.aux});
            if (o.op == T_RNE) out.stats.n_rne++;
            else out.stats.n_fail++;
            continue;
        }
        if (o.chk) {
            // checked instruction of a fused R1CS check: the comparison's other side is an operand like the others (resident
            // now); nothing is written
            if (o.op == T_DOT) {
                const bool has_add = o.c != NO_REF;
                if (has_add) flags |= F_ADDEND;
                if (has_add && isc[o.tn]) flags |= 2;
                out.ins.push_back(TapeIns{T_DOT, (uint8_t)(flags | F_CHECK), (uint16_t)enc[o.tn + 1], o.tn, has_add ? enc[o.tn] : 0u, o.aux});
                for (uint32_t k = 0; k < o.tn; k += 2) {
                    uint32_t rec[4] = {xp.terms[o.t0 + k].first & ~CONST_FLAG, enc[k], 0, 0};
                    if (k + 1 < o.tn) { rec[2] = xp.terms[o.t0 + k + 1].first & ~CONST_FLAG; rec[3] = enc[k + 1]; }
                    TapeIns raw;
                    memcpy(&raw, rec, sizeof(rec));
                    out.ins.push_back(raw);
                }
                out.stats.n_dot++;
                out.stats.n_dot_terms += o.tn;
            } else {
                if (isc[2]) throw TraceError("checked instruction compares with a constant");
                out.ins.push_back(TapeIns{o.op, (uint8_t)((flags & 3u) | F_CHECK), (uint16_t)enc[2], enc[0], enc[1], o.aux});
                if (o.op == T_MUL) out.stats.n_mul++;
                else out.stats.n_addsub++;
            }
            out.stats.n_rne++;
            continue;
        }
        // destination (may reuse the slot of an operand that died; ops read all operands before writing)
        still.clear();
        for (uint32_t r : rs) {
            if (r == NO_REF || (r & CONST_FLAG)) continue;
            if (val_slot[r] >= 0) still.push_back(slot_code(r));
        }
        const int dk = kind_of((uint32_t)i);
        uint32_t d = alloc_slot(dk, pos, still);
        val_slot[i] = (int32_t)d;
        slot_val[dk][d] = (int32_t)i;
        const uint16_t dcode = (uint16_t)(d | (dk ? BSLOT_DST : 0));
        live_now++;
        out.stats.max_live = std::max(out.stats.max_live, live_now);
        uint32_t w0 = wire_head[i];
        if (o.op == T_ISUM) {
            // header: a = number of terms, b = addend (integer slot, 0/1 slot, or integer constant with bit1); then ceil(n/4)
            // records of four terms: bit slot | shift << 16
            const bool has_add = o.c != NO_REF;
            uint32_t addend = 0;
            if (has_add) {
                flags |= F_ADDEND;
                addend = enc[o.tn];
                if (isc[o.tn]) {
                    flags |= 2;
                    const fr::Fr &cv = consts[enc[o.tn]];
                    const uint64_t v = ((uint64_t)cv.v[1] << 32) | cv.v[0];
                    auto it = iconst_index.find(v);
                    if (it == iconst_index.end()) {
                        it = iconst_index.emplace(v, (uint32_t)out.iconsts.size()).first;
                        out.iconsts.push_back(v);
                    }
                    addend = it->second;
                }
            }
            // a second addend (sums of sums merged by fuse_isums): field c, flag F_ADDEND2, constant flag bit 2
            uint32_t addend2 = 0;
            if (o.a != NO_REF) {
                flags |= F_ADDEND2;
                addend2 = enc[o.tn + 1];
                if (isc[o.tn + 1]) {
                    flags |= 4;
                    const fr::Fr &cv = consts[enc[o.tn + 1]];
                    const uint64_t v = ((uint64_t)cv.v[1] << 32) | cv.v[0];
                    auto it = iconst_index.find(v);
                    if (it == iconst_index.end()) {
                        it = iconst_index.emplace(v, (uint32_t)out.iconsts.size()).first;
                        out.iconsts.push_back(v);
                    }
                    addend2 = it->second;
                }
            }
            // Transposed form (T_ISUMT).  The terms are (bit slot, shift) pairs and the sum does not care about their order: they
            // are dealt into LAYERS in which every shift occurs at most once and all shifts lie within a 32-wide window.  A
            // layer is a 32 x 32 bit matrix -- lane l holds the word of the term whose shift is base + l (bit w = witness w) --
            // whose transpose gives every lane the integer  sum_j bit_j << (shift_j - base)  of ITS witness in five shuffle
            // steps, instead of one shared-memory read, mask, shift and 64-bit add per term and lane.
            {
                struct Term { uint32_t sh, slot; };
                std::vector<Term> terms;
                for (uint32_t k = 0; k < o.tn; k++) {
                    const fr::Fr &cv = consts[xp.terms[o.t0 + k].first & ~CONST_FLAG];
                    const uint64_t v = ((uint64_t)cv.v[1] << 32) | cv.v[0];
                    uint32_t sh = 0;
                    while (!((v >> sh) & 1)) sh++;
                    terms.push_back(Term{sh, enc[k] & 0xffffu});
                }
                std::stable_sort(terms.begin(), terms.end(), [](const Term &x, const Term &y) { return x.sh < y.sh; });
                struct Layer { uint32_t base; uint32_t slot[32]; };
                std::vector<Layer> layers;
                for (const Term &t : terms) {
                    bool placed = false;
                    for (Layer &L : layers)
                        if (t.sh >= L.base && t.sh - L.base < 32 && L.slot[t.sh - L.base] == 0xffffu) {
                            L.slot[t.sh - L.base] = t.slot;
                            placed = true;
                            break;
                        }
                    if (!placed) {
                        Layer L;
                        L.base = t.sh;
                        for (uint32_t &x : L.slot) x = 0xffffu;
                        L.slot[0] = t.slot;
                        layers.push_back(L);
                    }
                }
                // a layer costs about as much as five terms of the serial form
                if (layers.size() * 5 < o.tn && layers.size() < 256) {
                    out.ins.push_back(TapeIns{T_ISUMT, flags, dcode, (uint32_t)layers.size(), addend, addend2});
                    for (const Layer &L : layers)
                        for (uint32_t k = 0; k < 32; k += 4) {
                            uint32_t rec[4];
                            for (uint32_t j = 0; j < 4; j++) rec[j] = L.slot[k + j] | (L.base << 16);
                            TapeIns raw;
                            memcpy(&raw, rec, sizeof(rec));
                            out.ins.push_back(raw);
                        }
                    out.stats.n_int++;
                    out.stats.n_isum_terms += o.tn;
                    out.stats.n_isum_layers += layers.size();
                    goto isum_done;
                }
            }
            out.ins.push_back(TapeIns{T_ISUM, flags, dcode, o.tn, addend, addend2});
            for (uint32_t k = 0; k < o.tn; k += 4) {
                uint32_t rec[4] = {0, 0, 0, 0};
                for (uint32_t j = 0; j < 4 && k + j < o.tn; j++) {
                    const fr::Fr &cv = consts[xp.terms[o.t0 + k + j].first & ~CONST_FLAG];
                    const uint64_t v = ((uint64_t)cv.v[1] << 32) | cv.v[0];
                    uint32_t sh = 0;
                    while (!((v >> sh) & 1)) sh++;
                    rec[j] = (enc[k + j] & 0xffffu) | (sh << 16);
                }
                TapeIns raw;
                memcpy(&raw, rec, sizeof(rec));
                out.ins.push_back(raw);
            }
            out.stats.n_int++;
            out.stats.n_isum_terms += o.tn;
        isum_done:;
        } else if (o.op == T_DOT) {
            // header: a = number of terms, b = addend (slot or constant index), c = wire row with F_STORE;
            // then ceil(n/2) records of (constant index, slot) x 2
            const bool has_add = o.c != NO_REF;
            if (has_add) flags |= F_ADDEND;
            if (has_add && isc[o.tn]) flags |= 2;
            uint32_t row = 0;
            if (w0 != NO_REF) {
                flags |= F_STORE;
                row = out.wire_loc[w0];
                out.stats.n_st++;
                val_home[i] = row;
                w0 = wire_next[w0];
            }
            out.ins.push_back(TapeIns{T_DOT, flags, dcode, o.tn, has_add ? enc[o.tn] : 0u, row});
            for (uint32_t k = 0; k < o.tn; k += 2) {
                uint32_t rec[4] = {xp.terms[o.t0 + k].first & ~CONST_FLAG, enc[k], 0, 0};
                if (k + 1 < o.tn) { rec[2] = xp.terms[o.t0 + k + 1].first & ~CONST_FLAG; rec[3] = enc[k + 1]; }
                TapeIns raw;
                static_assert(sizeof(raw) == sizeof(rec), "record size");
                memcpy(&raw, rec, sizeof(rec));
                out.ins.push_back(raw);
            }
            out.stats.n_dot++;
            out.stats.n_dot_terms += o.tn;
        } else {
            uint32_t e0 = enc[0], e1 = enc[1], e2 = enc[2];
            if (o.op == T_INPUT || o.op == T_INPUT_BIT) e0 = o.aux;
            if (o.op == T_BITC || o.op == T_IBIT) e1 = o.aux;
            if (o.op == T_LUT) {   // a = bit slots 0 and 1 (16 bits each); b = table | k << 8 | bit slot 2 << 16; c is free for the fused store
                e0 = (enc[0] & 0xffffu) | ((enc[1] & 0xffffu) << 16);
                e1 = (o.aux & 0xffffu) | ((enc[2] & 0xffffu) << 16);
                e2 = 0;
                out.stats.n_lut++;
            }
            if (o.op == T_SEL && (o.c != NO_REF) && (o.c & CONST_FLAG) && fr::is_zero(consts[o.c & ~CONST_FLAG])) {
                flags = (uint8_t)((flags & ~4u) | F_CZERO);
                e2 = 0;
            }
            if (w0 != NO_REF && (o.op != T_SEL || (flags & F_CZERO)) && o.op != T_CADD) {   // the first wire of the value is written by the producing instruction
                flags |= F_STORE;
                e2 = out.wire_loc[w0];
                out.stats.n_st++;
                val_home[i] = e2;
                w0 = wire_next[w0];
            }
            out.ins.push_back(TapeIns{o.op, flags, dcode, e0, e1, e2});
            switch (o.op) {
                case T_MUL: out.stats.n_mul++; break;
                case T_DIV: out.stats.n_div++; break;
                case T_INV: out.stats.n_inv++; break;
                case T_SEL: out.stats.n_sel++; break;
                case T_CADD: out.stats.n_addsub++; break;
                case T_ADD: case T_SUB: out.stats.n_addsub++; break;
                case T_ICADD: case T_IADD: case T_ISEL: break;   // counted in n_int
                case T_INPUT: out.stats.n_input++; break;
                default: out.stats.n_other++; break;
            }
        }
        for (uint32_t w = w0; w != NO_REF; w = wire_next[w]) {
            out.ins.push_back(TapeIns{T_ST, 0, 0, (uint32_t)d | (dk ? BSLOT : 0u), 0, out.wire_loc[w]});
            out.stats.n_st++;
            if (val_home[i] == NO_REF) val_home[i] = out.wire_loc[w];
        }
        if (next_use((uint32_t)i, pos) == 0xffffffffu) release_value((uint32_t)i);
    }
    out.n_frows = out.n_fwires + spill_rows[0];
    out.n_brows = out.n_bwires + spill_rows[1];
    out.n_rows = out.n_frows + out.n_brows;
    out.stats.n_spill_rows = spill_rows[0] + spill_rows[1];
    out.stats.n_tape = out.ins.size();
    // bit-heavy program (more selects / bit extractions than products): its remaining products mostly have factors that
    // are 0 or 1 at run time (unconstrained input bits); let the kernel look before it multiplies
    if (out.stats.n_sel + out.stats.n_other > 2 * out.stats.n_mul) {
        for (size_t pc = 0; pc < out.ins.size(); pc++) {
            TapeIns &in = out.ins[pc];
            if (extra_records(in)) { pc += extra_records(in); continue; }
            if (in.op == T_MUL) in.flags |= F_TRIVIAL;
        }
    }
    // the reload ring costs 128 B of shared memory per witness: only worth it when the tape reloads field rows often
    out.use_ring = (uint64_t)(out.stats.n_ld - out.stats.n_ld_bool) * 50 >= out.ins.size();
    if (out.use_ring) schedule_reloads(out);
    else
        for (size_t pc = 0; pc < out.ins.size(); pc++) {
            TapeIns &in = out.ins[pc];
            if (extra_records(in)) { pc += extra_records(in); continue; }
            if (in.op == T_LD) in.a = NO_ROW;   // nothing to request: there is no ring
        }
    out.stats.macs = 136 * (out.stats.n_mul + out.stats.n_input + out.stats.n_inv + 2 * out.stats.n_div) + 64 * out.stats.n_dot_terms +
                     72 * out.stats.n_dot + 1800 * (out.stats.n_inv + out.stats.n_div);
    return out;
}

inline Tape allocate_tape(const Tracer &tr, const XProg &xp, uint32_t n_slots, uint32_t max_bslots = 2048) {
    return allocate_tape(tr.consts, tr.ops.size(), xp, n_slots, max_bslots);
}

inline Tape build_tape(const Tracer &tr, uint32_t n_slots, bool fuse = true, uint32_t max_bslots = 2048) {
    if (n_slots < 4) throw TraceError("need at least 4 slots");
    return allocate_tape(tr, prepare_program(tr, std::min<uint32_t>(16, n_slots - 2), fuse), n_slots, max_bslots);
}

}  // namespace tape
