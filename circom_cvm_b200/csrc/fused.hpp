// R1CS check scheduled INTO the tape (field-only programs).
//
// The stand-alone check (kernels.cuh r1cs_kernel) re-reads every wire a constraint mentions from the value store -- for
// Poseidon(2) 37 GB per 2^20 witnesses, 1.74 x the store itself, plus a CSR walk and operand addressing per term.  While
// the tape runs, those operands sit in the witness's slots.  fuse_check() evaluates every constraint of an .r1cs file
// right after the last wire it mentions has been produced:
//
//     a = A.w, b = B.w, c = C.w      linear combinations over the values bound to the wires: T_DOT (general coefficients:
//                                    64 multiply-accumulates per term, one reduction), T_ADD / T_SUB (+-1 coefficients);
//                                    a combination that is one wire with coefficient 1 is that wire's slot
//     p = a * b  ==  c ?             the LAST instruction of the evaluation carries F_CHECK: it compares its result with the
//                                    slot of the other side instead of writing it, and records the constraint in first_bad
//
// so x * y === z is one checked T_MUL, out === sum_j M_j in_j one checked T_DOT (the -1 term is the comparison's other
// side), out === in + k one checked T_ADD.  Shapes that do not end in such an instruction use T_RNE p, c, k.
//
// Nothing is shared with the instructions that computed the witness: the check's products and sums are separate
// instructions on the same operands (an .r1cs that does not belong to the program is caught like a wrong witness).
// The semantics are those of r1cs_kernel -- A.w * B.w = C.w over the values the tape stores, first violated constraint
// -- which tests compare with it on every fixture and on random circuits.
// Bit-heavy programs (integer-typed values, warp-cooperative groups, many 0/1 values) keep the separate kernels: their
// constraints over bits are evaluated 32 witnesses at a time by r1cs_table_kernel, which no per-witness instruction can
// match.
#pragma once
#include <algorithm>
#include <unordered_map>
#include <vector>

#include "r1cs.hpp"
#include "tape.hpp"

namespace tape {

// Field programs: no integer-typed values, no warp-cooperative groups, and at most a sprinkling of 0/1-typed values (the
// zero tests of batched inversions, a comparison here and there).
inline bool check_fusable(const XProg &xp) {
    size_t n_bool = 0;
    for (size_t i = 0; i < xp.ops.size(); i++) {
        if (xp.isint[i] || xp.group_len[i]) return false;
        n_bool += xp.isbool[i] ? 1 : 0;
    }
    return n_bool * 8 <= xp.ops.size();
}

// xp: the prepared program (fuse_dots ...); consts: its constant table (canonical values; coefficients are appended);
// max_terms: longest dot product the slot file allows.
inline XProg fuse_check(const XProg &xp, std::vector<fr::Fr> &consts, const r1cs::File &f, uint32_t max_terms) {
    if (f.n_wires != xp.witness_ref.size()) throw TraceError("r1cs and program disagree on the number of wires");
    if (max_terms < 1) max_terms = 1;
    std::unordered_map<fr::Fr, uint32_t, FrHash, FrEq> cindex;
    for (size_t i = 0; i < consts.size(); i++) cindex.emplace(consts[i], (uint32_t)i);
    auto cref = [&](const fr::Fr &v) -> uint32_t {
        auto it = cindex.find(v);
        if (it == cindex.end()) {
            it = cindex.emplace(v, (uint32_t)consts.size()).first;
            consts.push_back(v);
        }
        return CONST_FLAG | it->second;
    };
    const fr::Fr one = hostfr::from_u64(1), minus_one = fr::neg(one);
    const uint32_t N = (uint32_t)xp.ops.size();
    // check instructions, in their own numbering (refs >= N point into `extra`), and where each group goes
    std::vector<XOp> extra;
    std::vector<std::pair<uint32_t, uint32_t>> extra_terms_all;   // terms of the extra DOTs
    struct Group { uint32_t after, begin, end; };                 // extra[begin, end) follows op `after` (N: before everything)
    std::vector<Group> groups;
    auto push = [&](XOp x) -> uint32_t {
        extra.push_back(x);
        return N + (uint32_t)extra.size() - 1;
    };
    struct Lin {
        uint32_t ref;     // value, CONST ref, or NO_REF for the constant 0
        uint32_t last;    // latest program op among its operands (NO_REF: none)
    };
    auto later = [](uint32_t a, uint32_t b) { return a == NO_REF ? b : b == NO_REF ? a : std::max(a, b); };
    struct LcTerms {
        fr::Fr kconst;
        std::vector<std::pair<uint32_t, uint32_t>> general;   // (constant ref, value)
        std::vector<uint32_t> plus, minus;
        uint32_t last = NO_REF;
        size_t count() const { return general.size() + plus.size() + minus.size() + (fr::is_zero(kconst) ? 0 : 1); }
    };
    auto gather = [&](uint32_t j) -> LcTerms {
        LcTerms lt;
        lt.kconst = fr::zero();
        for (uint32_t t = f.ptr[j]; t < f.ptr[j + 1]; t++) {
            const uint32_t wire = f.terms[t].wire & 0x0fffffffu;
            const fr::Fr &coef = f.coefs[f.terms[t].coef];
            const uint32_t r = xp.witness_ref[wire];
            if (r & CONST_FLAG) {   // the wire is bound to a constant (wire 0: the constant 1)
                lt.kconst = fr::add(lt.kconst, hostfr::mul(coef, consts[r & ~CONST_FLAG]));
                continue;
            }
            lt.last = later(lt.last, r);
            if (fr::equal(coef, one)) lt.plus.push_back(r);
            else if (fr::equal(coef, minus_one)) lt.minus.push_back(r);
            else lt.general.emplace_back(cref(coef), r);
        }
        return lt;
    };
    auto eval_lc = [&](const LcTerms &lt) -> Lin {
        const fr::Fr &kconst = lt.kconst;
        const auto &general = lt.general;
        const auto &plus = lt.plus;
        const auto &minus = lt.minus;
        const uint32_t last = lt.last;
        uint32_t v = NO_REF;
        bool const_done = false;
        if (general.size() == 1 && plus.empty() && minus.empty() && fr::is_zero(kconst)) {
            v = push(XOp{T_MUL, general[0].first, general[0].second, NO_REF, 0});
        } else {
            for (size_t t0 = 0; t0 < general.size(); t0 += max_terms) {
                const size_t n = std::min<size_t>(max_terms, general.size() - t0);
                XOp x{T_DOT, NO_REF, NO_REF, NO_REF, 0};
                x.t0 = (uint32_t)extra_terms_all.size();
                x.tn = (uint32_t)n;
                for (size_t k = 0; k < n; k++) extra_terms_all.push_back(general[t0 + k]);
                if (v != NO_REF) x.c = v;
                else if (!fr::is_zero(kconst)) { x.c = cref(kconst); const_done = true; }
                v = push(x);
            }
        }
        for (uint32_t r : plus) v = v == NO_REF ? r : push(XOp{T_ADD, v, r, NO_REF, 0});
        for (uint32_t r : minus) v = push(XOp{T_SUB, v == NO_REF ? cref(fr::zero()) : v, r, NO_REF, 0});
        if (!const_done && !fr::is_zero(kconst)) v = v == NO_REF ? cref(kconst) : push(XOp{T_ADD, v, cref(kconst), NO_REF, 0});
        return Lin{v, last};
    };
    for (uint32_t c = 0; c < f.n_constraints; c++) {
        const uint32_t begin = (uint32_t)extra.size();
        const size_t j = 3 * (size_t)c;
        const bool same_b = f.split[3 * (j + 1)] == r1cs::SAME_AS_A;
        const bool a_empty = f.ptr[j] == f.ptr[j + 1], b_empty = !same_b && f.ptr[j + 1] == f.ptr[j + 2];
        uint32_t last = NO_REF, p = NO_REF;   // p: the product (NO_REF: 0)
        const uint32_t zero = cref(fr::zero());
        // the last instruction of [begin, ...) that produced `lhs` becomes the checked one when the other side is a value
        auto finish = [&](uint32_t lhs, uint32_t rhs) {
            auto is_value = [](uint32_t r) { return r != NO_REF && !(r & CONST_FLAG); };
            for (int swap = 0; swap < 2; swap++, std::swap(lhs, rhs)) {
                if (!is_value(rhs) || !is_value(lhs) || lhs < N || lhs != N + (uint32_t)extra.size() - 1) continue;
                if (rhs < N && xp.isbool[rhs]) continue;   // a value in the bit file: compared by T_RNE
                XOp &o = extra[lhs - N];
                if (o.op == T_DOT) o.a = rhs;
                else if (o.op == T_ADD || o.op == T_SUB || o.op == T_MUL) o.c = rhs;
                else continue;
                o.chk = 1;
                o.aux = c;
                return;
            }
            push(XOp{T_RNE, lhs == NO_REF ? zero : lhs, rhs == NO_REF ? zero : rhs, NO_REF, c});
        };
        if (!a_empty && !b_empty) {
            Lin a = eval_lc(gather((uint32_t)j));
            Lin b = same_b ? a : eval_lc(gather((uint32_t)j + 1));
            Lin cl = eval_lc(gather((uint32_t)j + 2));     // before the product: the product is the last instruction
            last = later(later(a.last, b.last), cl.last);
            auto is_const = [&](uint32_t r, const fr::Fr &v) { return r != NO_REF && (r & CONST_FLAG) && fr::equal(consts[r & ~CONST_FLAG], v); };
            if (a.ref == NO_REF || b.ref == NO_REF) p = NO_REF;
            else if (is_const(a.ref, one)) p = b.ref;
            else if (is_const(b.ref, one)) p = a.ref;
            else p = push(XOp{T_MUL, a.ref, b.ref, NO_REF, 0});
            finish(p, cl.ref);
        } else {
            // linear: C.w = 0.  One -1 term moves to the other side: (the rest) == that wire
            LcTerms lt = gather((uint32_t)j + 2);
            last = lt.last;
            uint32_t rhs = NO_REF;
            if (lt.minus.empty() && !lt.plus.empty()) {   // C.w = 0  <=>  (-C).w = 0
                for (auto &g : lt.general) g.first = cref(fr::neg(consts[g.first & ~CONST_FLAG]));
                std::swap(lt.plus, lt.minus);
                lt.kconst = fr::neg(lt.kconst);
            }
            if (!lt.minus.empty() && lt.count() >= 2) {
                rhs = lt.minus.back();
                lt.minus.pop_back();
            }
            Lin cl = eval_lc(lt);
            finish(cl.ref, rhs);
        }
        groups.push_back(Group{last == NO_REF ? N : last, begin, (uint32_t)extra.size()});
    }
    // merge: program op i, then the groups whose last operand it is (constraint order is kept among them)
    std::stable_sort(groups.begin(), groups.end(), [&](const Group &x, const Group &y) {
        const uint64_t kx = x.after == N ? 0 : (uint64_t)x.after + 1, ky = y.after == N ? 0 : (uint64_t)y.after + 1;
        return kx < ky;
    });
    XProg out;
    out.ops.reserve(N + extra.size());
    out.terms.reserve(xp.terms.size() + extra_terms_all.size());
    std::vector<uint32_t> remap((size_t)N + extra.size(), NO_REF);
    auto mapref = [&](uint32_t r) -> uint32_t { return (r == NO_REF || (r & CONST_FLAG)) ? r : remap[r]; };
    auto emit = [&](const XOp &o, uint32_t id, const std::vector<std::pair<uint32_t, uint32_t>> &terms) {
        XOp x = o;
        if (o.op == T_DOT || o.op == T_ISUM) {
            x.t0 = (uint32_t)out.terms.size();
            for (uint32_t k = 0; k < o.tn; k++) out.terms.emplace_back(terms[o.t0 + k].first, mapref(terms[o.t0 + k].second));
            x.c = mapref(o.c);
            if (o.chk) x.a = mapref(o.a);
        } else {
            x.a = mapref(o.a);
            x.b = mapref(o.b);
            x.c = mapref(o.c);
        }
        out.ops.push_back(x);
        remap[id] = (uint32_t)out.ops.size() - 1;
    };
    size_t g = 0;
    auto flush_groups = [&](uint32_t after) {
        for (; g < groups.size() && groups[g].after == after; g++)
            for (uint32_t e = groups[g].begin; e < groups[g].end; e++) emit(extra[e], N + e, extra_terms_all);
    };
    flush_groups(N);
    for (uint32_t i = 0; i < N; i++) {
        emit(xp.ops[i], i, xp.terms);
        flush_groups(i);
    }
    out.witness_ref.reserve(xp.witness_ref.size());
    for (uint32_t r : xp.witness_ref) out.witness_ref.push_back(mapref(r));
    out.isbool.assign(out.ops.size(), 0);
    for (uint32_t i = 0; i < N; i++) out.isbool[remap[i]] = xp.isbool[i];
    out.isint.assign(out.ops.size(), 0);
    out.group_len.assign(out.ops.size(), 0);
    return out;
}

}  // namespace tape
