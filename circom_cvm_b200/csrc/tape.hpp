// From the tracer's SSA list to the executable tape: dead-code elimination, then linear-scan
// allocation of the per-witness on-chip slots (shared memory) with furthest-next-use eviction.
// Evicted values that are still needed go to a spill row of the value store in HBM; values bound to
// witness wires are stored to their wire row when produced (they have to be written once anyway,
// common/main.cpp:324-330 reads every witness wire) and are reloaded from there if evicted.
#pragma once
#include <algorithm>
#include <cstdint>
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
    uint16_t dst;     // slot
    uint32_t a, b;    // slot or constant index; T_INPUT: a = input index; T_BITC: b = bit number;
                      // T_LUT: a = three input slots (one byte each), b = truth table | number of inputs << 8;
                      // T_LD: a = row to request now for the reload LD_RING reloads ahead (NO_ROW: none), b = ring entry
    uint32_t c;       // T_SEL: third operand; T_LD/T_ST/T_STC: value-store row; T_FAIL_IF/T_FAIL_NE: status;
                      // with flag bit3: value-store row
};
static const uint8_t F_STORE = 8;
static const uint8_t F_CZERO = 16;
static const uint8_t F_TRIVIAL = 16;  // T_MUL: check at run time whether the factors are 0 / 1 (bit-heavy programs)
static const uint8_t F_ADDEND = 32;
static const uint8_t F_RING = 64;    // T_LD: the value was requested LD_RING reloads ago and sits in ring entry b
static const uint32_t LD_RING = 4;   // reloads in flight per witness (32 B of shared memory each)
static const uint32_t NO_ROW = 0xffffffffu;  // T_DOT: field b holds an addend (slot, or constant index with bit1)   // T_SEL: the third operand is the constant 0 (field c is free for the fused store)
static_assert(sizeof(TapeIns) == 16, "tape instruction must be 16 bytes");

struct TapeStats {
    uint64_t n_ssa = 0, n_live = 0, n_tape = 0;
    uint64_t n_mul = 0, n_div = 0, n_addsub = 0, n_other = 0, n_inv = 0, n_sel = 0;   // executed per witness
    uint64_t n_ld = 0, n_st = 0, n_spill_st = 0, n_stc = 0, n_input = 0, n_fail = 0, n_dot = 0, n_dot_terms = 0, n_ld_streamed = 0, n_lut = 0;
    uint64_t n_ld_bool = 0, n_spill_st_bool = 0;   // of n_ld / n_spill_st: the value is typed 0/1 (what compact bit rows would shrink)
    uint32_t n_spill_rows = 0;
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
    uint32_t n_rows = 0;      // wires + spill rows
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
};
struct XProg {
    std::vector<XOp> ops;
    std::vector<std::pair<uint32_t, uint32_t>> terms;   // (constant ref, value id)
    std::vector<uint32_t> witness_ref;
    std::vector<uint8_t> isbool;   // per op: the tracer proved the value 0/1 (fused results: not typed)
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
        for (size_t i = 0; i < N; i++) {
            if (!live[i] || absorbed[i] || ops[i].op != T_ADD || roots.count((uint32_t)i)) continue;
            for (int k = 0; k < 2; k++) {
                uint32_t r = k ? ops[i].b : ops[i].a;
                if (r == NO_REF || (r & CONST_FLAG) || uses[r] != 1 || absorbed[r]) continue;
                const SOp &sel = ops[r];
                if (sel.op == T_SEL && !(sel.a & CONST_FLAG) && (sel.b & CONST_FLAG) && (sel.c & CONST_FLAG) &&
                    fr::is_zero(tr.consts[sel.c & ~CONST_FLAG])) {
                    absorbed[r] = 1;
                    cadd.emplace((uint32_t)i, k);
                    break;
                }
            }
        }
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

// ---- reload stream ---------------------------------------------------------------------------------------
// The rows a tape reloads (T_LD) and their order are fixed, so the kernel streams them: the n-th reload's row is
// requested (cp.async into a per-witness ring in shared memory) when reload n - LD_RING executes, and reload n finds
// it there -- HBM latency leaves the dependence chain, which is what small batches (few resident warps) are bound by.
// A reload is only streamed when the store that produced its row precedes the request point (spill rows are
// recycled); the others load directly, as before.
inline void schedule_reloads(Tape &t) {
    std::vector<uint32_t> ld_pos, ld_store;
    std::vector<uint32_t> last_store(t.n_rows, NO_ROW);
    for (size_t pc = 0; pc < t.ins.size(); pc++) {
        TapeIns &in = t.ins[pc];
        if (in.op == T_DOT) {
            if (in.flags & F_STORE) last_store[in.c] = (uint32_t)pc;
            pc += (in.a + 1) / 2;
            continue;
        }
        if (in.op == T_LD) {
            in.a = NO_ROW;
            in.b = (uint32_t)(ld_pos.size() % LD_RING);
            ld_pos.push_back((uint32_t)pc);
            ld_store.push_back(last_store[in.c]);
        } else if (in.op == T_ST || in.op == T_STC || (in.flags & F_STORE)) {
            if (in.op != T_FAIL_IF && in.op != T_FAIL_NE) last_store[in.c] = (uint32_t)pc;
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

inline Tape build_tape(const Tracer &tr, uint32_t n_slots, bool fuse = true) {
    if (n_slots < 4) throw TraceError("need at least 4 slots");
    const uint32_t max_terms = std::min<uint32_t>(16, n_slots - 2);
    const XProg xp = fuse_dots(tr, max_terms, fuse);
    const std::vector<XOp> &ops = xp.ops;
    const size_t N = ops.size();
    Tape out;
    out.n_slots = n_slots;
    out.n_wires = (uint32_t)xp.witness_ref.size();
    out.stats.n_ssa = tr.ops.size();
    out.stats.n_live = N;

    auto operands = [&](const XOp &o, std::vector<uint32_t> &rs) {
        rs.clear();
        if (o.op == T_DOT) {
            for (uint32_t k = 0; k < o.tn; k++) rs.push_back(xp.terms[o.t0 + k].second);
            rs.push_back(o.c);
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
    // constants bound to wires are stored up front
    for (size_t w = 0; w < xp.witness_ref.size(); w++) {
        uint32_t r = xp.witness_ref[w];
        if (r & CONST_FLAG) {
            out.ins.push_back(TapeIns{T_STC, 1, 0, r & ~CONST_FLAG, 0, (uint32_t)w});
            out.stats.n_stc++;
        }
    }
    // ---- linear scan
    std::vector<int32_t> slot_val(n_slots, -1);
    std::vector<int32_t> val_slot(N, -1);
    std::vector<uint32_t> val_home(N, NO_REF);
    std::vector<uint8_t> home_is_spill(N, 0);
    std::vector<uint32_t> free_slots, free_spill;
    for (uint32_t s = n_slots; s-- > 0;) free_slots.push_back(s);
    uint32_t spill_rows = 0;
    uint32_t live_now = 0;

    auto release_value = [&](uint32_t v) {
        if (val_slot[v] >= 0) {
            slot_val[(size_t)val_slot[v]] = -1;
            free_slots.push_back((uint32_t)val_slot[v]);
            val_slot[v] = -1;
        }
        if (home_is_spill[v]) {
            free_spill.push_back(val_home[v] - out.n_wires);
            home_is_spill[v] = 0;
        }
        live_now--;
    };
    auto alloc_slot = [&](uint32_t pos, const std::vector<int32_t> &pinned) -> uint32_t {
        if (!free_slots.empty()) {
            uint32_t s = free_slots.back();
            free_slots.pop_back();
            return s;
        }
        uint32_t best = 0xffffffffu, best_use = 0;
        for (uint32_t s = 0; s < n_slots; s++) {
            bool pin = false;
            for (int32_t q : pinned)
                if (q == (int32_t)s) pin = true;
            if (pin) continue;
            uint32_t v = (uint32_t)slot_val[s];
            uint32_t nu = next_use(v, pos);
            if (best == 0xffffffffu || nu > best_use) { best = s; best_use = nu; }
        }
        if (best == 0xffffffffu) throw TraceError("slot allocator: all slots pinned");
        uint32_t v = (uint32_t)slot_val[best];
        if (val_home[v] == NO_REF) {   // not yet in HBM: spill
            uint32_t row;
            if (!free_spill.empty()) { row = free_spill.back(); free_spill.pop_back(); }
            else row = spill_rows++;
            val_home[v] = out.n_wires + row;
            home_is_spill[v] = 1;
            out.ins.push_back(TapeIns{T_ST, 0, 0, best, 0, val_home[v]});
            out.stats.n_st++;
            out.stats.n_spill_st++;
            out.stats.n_spill_st_bool += xp.isbool[v];
        }
        val_slot[v] = -1;
        slot_val[best] = -1;
        return best;
    };

    std::vector<int32_t> pinned, still;
    std::vector<uint32_t> enc;
    for (size_t i = 0; i < N; i++) {
        const XOp &o = ops[i];
        uint32_t pos = (uint32_t)i;
        operands(o, rs);
        pinned.clear();
        enc.assign(rs.size(), 0);
        std::vector<uint8_t> isc(rs.size(), 0);
        // resident operands are pinned first so that loading one operand cannot evict another
        for (uint32_t r : rs) {
            if (r == NO_REF || (r & CONST_FLAG)) continue;
            if (val_slot[r] >= 0) pinned.push_back(val_slot[r]);
        }
        for (size_t k = 0; k < rs.size(); k++) {
            uint32_t r = rs[k];
            if (r == NO_REF) continue;
            if (r & CONST_FLAG) { isc[k] = 1; enc[k] = r & ~CONST_FLAG; continue; }
            if (val_slot[r] < 0) {
                if (val_home[r] == NO_REF) throw TraceError("slot allocator: value lost");
                uint32_t s = alloc_slot(pos, pinned);
                out.ins.push_back(TapeIns{T_LD, 0, (uint16_t)s, 0, 0, val_home[r]});
                out.stats.n_ld++;
                out.stats.n_ld_bool += xp.isbool[r];
                val_slot[r] = (int32_t)s;
                slot_val[s] = (int32_t)r;
                pinned.push_back((int32_t)s);
            }
            enc[k] = (uint32_t)val_slot[r];
        }
        // operands that die here free their slots before the destination is chosen
        for (size_t k = 0; k < rs.size(); k++) {
            uint32_t r = rs[k];
            if (r == NO_REF || (r & CONST_FLAG)) continue;
            if (val_slot[r] >= 0 && next_use(r, pos) == 0xffffffffu) release_value(r);
        }
        uint8_t flags = 0;
        if (o.op != T_DOT)
            for (size_t k = 0; k < 3; k++)
                if (isc[k]) flags |= (uint8_t)(1u << k);
        if (o.op == T_FAIL_IF || o.op == T_FAIL_NE) {
            out.ins.push_back(TapeIns{o.op, flags, 0, enc[0], enc[1], o.aux});
            out.stats.n_fail++;
            continue;
        }
        // destination (may reuse the slot of an operand that died; ops read all operands before writing)
        still.clear();
        for (uint32_t r : rs) {
            if (r == NO_REF || (r & CONST_FLAG)) continue;
            if (val_slot[r] >= 0) still.push_back(val_slot[r]);
        }
        uint32_t d = alloc_slot(pos, still);
        val_slot[i] = (int32_t)d;
        slot_val[d] = (int32_t)i;
        live_now++;
        out.stats.max_live = std::max(out.stats.max_live, live_now);
        uint32_t w0 = wire_head[i];
        if (o.op == T_DOT) {
            // header: a = number of terms, b = addend (slot or constant index), c = wire row with F_STORE;
            // then ceil(n/2) records of (constant index, slot) x 2
            const bool has_add = o.c != NO_REF;
            if (has_add) flags |= F_ADDEND;
            if (has_add && isc[o.tn]) flags |= 2;
            uint32_t row = 0;
            if (w0 != NO_REF) {
                flags |= F_STORE;
                row = w0;
                out.stats.n_st++;
                val_home[i] = w0;
                w0 = wire_next[w0];
            }
            out.ins.push_back(TapeIns{T_DOT, flags, (uint16_t)d, o.tn, has_add ? enc[o.tn] : 0u, row});
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
            if (o.op == T_INPUT) e0 = o.aux;
            if (o.op == T_BITC) e1 = o.aux;
            if (o.op == T_LUT) {   // a = the three input slots, one byte each; b = table | k << 8; c is free for the fused store
                e0 = enc[0] | (enc[1] << 8) | (enc[2] << 16);
                e1 = o.aux;
                e2 = 0;
                out.stats.n_lut++;
            }
            if (o.op == T_SEL && (o.c != NO_REF) && (o.c & CONST_FLAG) && fr::is_zero(tr.consts[o.c & ~CONST_FLAG])) {
                flags = (uint8_t)((flags & ~4u) | F_CZERO);
                e2 = 0;
            }
            if (w0 != NO_REF && (o.op != T_SEL || (flags & F_CZERO)) && o.op != T_CADD) {   // the first wire of the value is written by the producing instruction
                flags |= F_STORE;
                e2 = w0;
                out.stats.n_st++;
                val_home[i] = w0;
                w0 = wire_next[w0];
            }
            out.ins.push_back(TapeIns{o.op, flags, (uint16_t)d, e0, e1, e2});
            switch (o.op) {
                case T_MUL: out.stats.n_mul++; break;
                case T_DIV: out.stats.n_div++; break;
                case T_INV: out.stats.n_inv++; break;
                case T_SEL: out.stats.n_sel++; break;
                case T_CADD: out.stats.n_addsub++; break;
                case T_ADD: case T_SUB: out.stats.n_addsub++; break;
                case T_INPUT: out.stats.n_input++; break;
                default: out.stats.n_other++; break;
            }
        }
        for (uint32_t w = w0; w != NO_REF; w = wire_next[w]) {
            out.ins.push_back(TapeIns{T_ST, 0, 0, d, 0, w});
            out.stats.n_st++;
            if (val_home[i] == NO_REF) val_home[i] = w;
        }
        if (next_use((uint32_t)i, pos) == 0xffffffffu) release_value((uint32_t)i);
    }
    out.n_rows = out.n_wires + spill_rows;
    out.stats.n_spill_rows = spill_rows;
    out.stats.n_tape = out.ins.size();
    // bit-heavy program (more selects / bit extractions than products): its remaining products mostly have factors that
    // are 0 or 1 at run time (unconstrained input bits); let the kernel look before it multiplies
    if (out.stats.n_sel + out.stats.n_other > 2 * out.stats.n_mul) {
        for (size_t pc = 0; pc < out.ins.size(); pc++) {
            TapeIns &in = out.ins[pc];
            if (in.op == T_DOT) { pc += (in.a + 1) / 2; continue; }
            if (in.op == T_MUL) in.flags |= F_TRIVIAL;
        }
    }
    schedule_reloads(out);
    out.stats.macs = 136 * (out.stats.n_mul + out.stats.n_input + out.stats.n_inv + 2 * out.stats.n_div) + 64 * out.stats.n_dot_terms +
                     72 * out.stats.n_dot + 1800 * (out.stats.n_inv + out.stats.n_div);
    return out;
}

}  // namespace tape
