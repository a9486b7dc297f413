// From the tracer's SSA list to the executable tape: dead-code elimination, then linear-scan
// allocation of the per-witness on-chip slots (shared memory) with furthest-next-use eviction.
// Evicted values that are still needed go to a spill row of the value store in HBM; values bound to
// witness wires are stored to their wire row when produced (they have to be written once anyway,
// common/main.cpp:324-330 reads every witness wire) and are reloaded from there if evicted.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

#include "tracer.hpp"

namespace tape {

// 16-byte tape instruction
struct TapeIns {
    uint8_t op;
    uint8_t flags;    // bit0: a is a constant index, bit1: b is constant, bit2: c is constant,
                      // bit3: the result is also stored to value-store row c (fused witness-wire store)
    uint16_t dst;     // slot
    uint32_t a, b;    // slot or constant index; T_INPUT: a = input index; T_BITC: b = bit number
    uint32_t c;       // T_SEL: third operand; T_LD/T_ST/T_STC: value-store row; T_FAIL_IF/T_FAIL_NE: status;
                      // with flag bit3: value-store row
};
static const uint8_t F_STORE = 8;
static const uint8_t F_CZERO = 16;   // T_SEL: the third operand is the constant 0 (field c is free for the fused store)
static_assert(sizeof(TapeIns) == 16, "tape instruction must be 16 bytes");

struct TapeStats {
    uint64_t n_ssa = 0, n_live = 0, n_tape = 0;
    uint64_t n_mul = 0, n_div = 0, n_addsub = 0, n_other = 0, n_inv = 0, n_sel = 0;   // executed per witness
    uint64_t n_ld = 0, n_st = 0, n_spill_st = 0, n_stc = 0, n_input = 0, n_fail = 0;
    uint32_t n_spill_rows = 0;
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

inline Tape build_tape(const Tracer &tr, uint32_t n_slots) {
    const std::vector<SOp> &ops = tr.ops;
    const size_t N = ops.size();
    Tape out;
    out.n_slots = n_slots;
    out.n_wires = (uint32_t)tr.witness_ref.size();
    out.stats.n_ssa = N;
    if (n_slots < 4) throw TraceError("need at least 4 slots");

    // ---- liveness (roots: failure checks and witness wires)
    std::vector<uint8_t> live(N, 0);
    for (size_t i = 0; i < N; i++)
        if (ops[i].op == T_FAIL_IF || ops[i].op == T_FAIL_NE) live[i] = 1;
    for (uint32_t r : tr.witness_ref)
        if (!(r & CONST_FLAG)) live[r] = 1;
    for (size_t i = N; i-- > 0;) {
        if (!live[i]) continue;
        const SOp &o = ops[i];
        uint32_t rs[3] = {o.a, o.b, o.c};
        for (uint32_t r : rs)
            if (r != NO_REF && !(r & CONST_FLAG)) live[r] = 1;
    }
    // ---- uses (CSR), in program order
    std::vector<uint32_t> use_cnt(N + 1, 0);
    for (size_t i = 0; i < N; i++) {
        if (!live[i]) continue;
        out.stats.n_live++;
        const SOp &o = ops[i];
        uint32_t rs[3] = {o.a, o.b, o.c};
        for (uint32_t r : rs)
            if (r != NO_REF && !(r & CONST_FLAG)) use_cnt[r + 1]++;
    }
    for (size_t i = 0; i < N; i++) use_cnt[i + 1] += use_cnt[i];
    std::vector<uint32_t> use_pos(use_cnt[N]);
    {
        std::vector<uint32_t> fill(use_cnt.begin(), use_cnt.end() - 1);
        for (size_t i = 0; i < N; i++) {
            if (!live[i]) continue;
            const SOp &o = ops[i];
            uint32_t rs[3] = {o.a, o.b, o.c};
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
    std::vector<uint32_t> wire_head(N, NO_REF), wire_next(tr.witness_ref.size(), NO_REF);
    for (size_t w = tr.witness_ref.size(); w-- > 0;) {
        uint32_t r = tr.witness_ref[w];
        if (r & CONST_FLAG) continue;
        wire_next[w] = wire_head[r];
        wire_head[r] = (uint32_t)w;
    }
    // constants bound to wires are stored up front
    for (size_t w = 0; w < tr.witness_ref.size(); w++) {
        uint32_t r = tr.witness_ref[w];
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
    auto alloc_slot = [&](uint32_t pos, const int32_t *pinned, int npinned) -> uint32_t {
        if (!free_slots.empty()) {
            uint32_t s = free_slots.back();
            free_slots.pop_back();
            return s;
        }
        uint32_t best = 0xffffffffu, best_use = 0;
        for (uint32_t s = 0; s < n_slots; s++) {
            bool pin = false;
            for (int k = 0; k < npinned; k++)
                if (pinned[k] == (int32_t)s) pin = true;
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
        }
        val_slot[v] = -1;
        slot_val[best] = -1;
        return best;
    };

    for (size_t i = 0; i < N; i++) {
        if (!live[i]) continue;
        const SOp &o = ops[i];
        uint32_t pos = (uint32_t)i;
        uint32_t rs[3] = {o.a, o.b, o.c};
        int32_t pinned[8];
        int npinned = 0;
        uint32_t enc[3] = {0, 0, 0};
        uint8_t flags = 0;
        // resident operands are pinned first so that loading one operand cannot evict another
        for (int k = 0; k < 3; k++) {
            uint32_t r = rs[k];
            if (r == NO_REF || (r & CONST_FLAG)) continue;
            if (val_slot[r] >= 0) pinned[npinned++] = val_slot[r];
        }
        for (int k = 0; k < 3; k++) {
            uint32_t r = rs[k];
            if (r == NO_REF) continue;
            if (r & CONST_FLAG) { flags |= (uint8_t)(1u << k); enc[k] = r & ~CONST_FLAG; continue; }
            if (val_slot[r] < 0) {
                if (val_home[r] == NO_REF) throw TraceError("slot allocator: value lost");
                uint32_t s = alloc_slot(pos, pinned, npinned);
                out.ins.push_back(TapeIns{T_LD, 0, (uint16_t)s, 0, 0, val_home[r]});
                out.stats.n_ld++;
                val_slot[r] = (int32_t)s;
                slot_val[s] = (int32_t)r;
                pinned[npinned++] = (int32_t)s;
            }
            enc[k] = (uint32_t)val_slot[r];
        }
        // operands that die here free their slots before the destination is chosen
        for (int k = 0; k < 3; k++) {
            uint32_t r = rs[k];
            if (r == NO_REF || (r & CONST_FLAG)) continue;
            if (val_slot[r] >= 0 && next_use(r, pos) == 0xffffffffu) {
                bool dup = false;
                for (int j = 0; j < k; j++)
                    if (rs[j] == r) dup = true;
                if (!dup) release_value(r);
            } else if (val_slot[r] < 0 && next_use(r, pos) == 0xffffffffu) {
                // already released through a duplicate operand
            }
        }
        if (o.op == T_FAIL_IF || o.op == T_FAIL_NE) {
            out.ins.push_back(TapeIns{o.op, flags, 0, enc[0], enc[1], o.aux});
            out.stats.n_fail++;
            continue;
        }
        // destination (may reuse the slot of an operand that died; ops read all operands before writing)
        int32_t still[4];
        int nstill = 0;
        for (int k = 0; k < 3; k++) {
            uint32_t r = rs[k];
            if (r == NO_REF || (r & CONST_FLAG)) continue;
            if (val_slot[r] >= 0) still[nstill++] = val_slot[r];
        }
        uint32_t d = alloc_slot(pos, still, nstill);
        val_slot[i] = (int32_t)d;
        slot_val[d] = (int32_t)i;
        live_now++;
        out.stats.max_live = std::max(out.stats.max_live, live_now);
        if (o.op == T_INPUT) enc[0] = o.aux;
        if (o.op == T_BITC) enc[1] = o.aux;
        uint32_t w0 = wire_head[i];
        if (o.op == T_SEL && (o.c & CONST_FLAG) && fr::is_zero(tr.consts[o.c & ~CONST_FLAG])) {
            flags = (uint8_t)((flags & ~4u) | F_CZERO);
            enc[2] = 0;
        }
        if (w0 != NO_REF && (o.op != T_SEL || (flags & F_CZERO))) {   // the first wire of the value is written by the producing instruction
            flags |= F_STORE;
            enc[2] = w0;
            out.stats.n_st++;
            val_home[i] = w0;
            w0 = wire_next[w0];
        }
        out.ins.push_back(TapeIns{o.op, flags, (uint16_t)d, enc[0], enc[1], enc[2]});
        switch (o.op) {
            case T_MUL: out.stats.n_mul++; break;
            case T_DIV: out.stats.n_div++; break;
            case T_INV: out.stats.n_inv++; break;
            case T_SEL: out.stats.n_sel++; break;
            case T_ADD: case T_SUB: out.stats.n_addsub++; break;
            case T_INPUT: out.stats.n_input++; break;
            default: out.stats.n_other++; break;
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
    return out;
}

}  // namespace tape
