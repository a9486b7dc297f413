// .r1cs loader -> compact CSR for the check kernel.
// Format: constraint_writers/src/r1cs_writer.rs (header :245-269, constraints :282-308 / :49-91,
// wire2label :328-341); sections are located by scanning the section table because the writer emits
// constraints BEFORE the header (constraint_list/src/r1cs_porting.rs:19-53), as the reference's own
// (unused) reader does (constraint_writers/src/r1cs_reader.rs:459-476).
// Coefficients are interned like the compiler does (circom_algebra/src/constraint_storage/logic.rs:4-12):
// table index 0 is +1, index 1 is -1 (q-1), so the kernel can special-case both as add/sub.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "fr.cuh"
#include "tracer.hpp"   // FrHash / FrEq

namespace r1cs {

static const uint32_t SAME_AS_A = 0xffffffffu;   // first class boundary of a B combination that repeats A

struct Term {
    uint32_t wire;
    uint32_t coef;   // index into coefs
};

struct File {
    uint32_t field_size = 0;
    uint32_t n_wires = 0, n_pub_out = 0, n_pub_in = 0, n_prv_in = 0, n_constraints = 0;
    uint64_t n_labels = 0;
    std::vector<uint32_t> ptr;       // 3*n_constraints + 1 offsets into terms: LC k of constraint c at ptr[3c+k]
    // per LC j, terms are ordered [+-2^k | small positive | small negative | general | constant]; split[3j+k] = end of
    // class k.  "constant": a general coefficient on wire 0 (the constant 1, r1cs_writer.rs wire numbering): the term's
    // value is the coefficient itself, so the kernel adds it instead of multiplying (bit 31 of the wire word marks it;
    // it is the last term of its LC).
    // "small": the coefficient c (positive) or q - c (negative) fits 32 bits; such terms cost 8 multiply-accumulates
    // instead of 64 (kernels.cuh lc_eval).  A class is only used when it has >= 4 members, else they count as general.
    std::vector<uint32_t> split;
    std::vector<uint32_t> cmag;      // per coefficient: the 32-bit magnitude when small, else 0
    std::vector<Term> terms;
    std::vector<fr::Fr> coefs;       // canonical
    std::vector<uint64_t> wire2label;
    uint64_t nnz = 0;                // non-zeros of the file (terms may hold fewer: the B side of squares is dropped)
    uint64_t nnz_pm1 = 0, nnz_small = 0, nnz_const = 0, n_squares = 0, nnz_general = 0;
    // what the check kernel executes per witness: 64 multiply-accumulates per general term + 72 per reduction of a dot
    // product (<= 16 terms), 8 per small term + 170 per small-class reduction, 136 per quadratic constraint
    uint64_t macs = 0, n_quadratic = 0, n_linear = 0;
    // per coefficient: its value as a signed 64-bit integer when |c| < 2^62 (c or c - q), and whether it has one
    std::vector<int64_t> cint;
    std::vector<uint8_t> cint_ok;
};

// The CSR bound to a typed value store (tape.hpp: a wire is a field row or a bit row).  Field-row terms keep the
// layout above (hdr = {begin, end of +-2^k, of small +, of small -} per LC, rows instead of wires); bit-row terms go to a
// second CSR (bhdr / bterms).  A constraint whose terms are ALL bit rows with 64-bit integer coefficients (partial sums
// below 2^62) is evaluated in plain integers: (A*B - C) = 0 mod q iff = 0 in Z then (mode field of bhdr).
struct Bound {
    std::vector<uint32_t> hdr;      // 4 * (3 * n_constraints + 1)
    std::vector<Term> fterms;
    // 4 words per constraint (+ one closing entry): begin of the bit terms of A | mode << 30, of B, of C, of the next
    // constraint's A.  mode 0: evaluated in the field; 1: in 64-bit integers; 2: in 32-bit integers (every partial sum
    // below 2^31, the product in 64 bits)
    std::vector<uint32_t> bhdr;
    std::vector<Term> bterms;       // (bit row, coefficient index)
    // mode 3: a constraint over at most TABLE_VARS distinct 0/1 wires IS a boolean predicate of them; its truth table is
    // computed here and the kernel evaluates it bitwise for 32 witnesses at a time (r1cs_table_kernel): 8 words per
    // constraint = {bit rows of the variables (unused: row of the constant 1), truth table, constraint index, 0}.
    // Such constraints have no terms in either CSR.
    std::vector<uint32_t> tcons;
    uint64_t n_table_constraints = 0;
    uint64_t n_tautologies = 0;      // constraints over 0/1 wires that hold for every assignment (not evaluated)
    // mode 3 as well: LINEAR constraints over 0/1 wires whose coefficients are +-2^k (the sums of a hash round: sum_k 2^k
    // (a_k + b_k ...) = sum_k 2^k out_k).  Their terms are dealt into layers of distinct shifts within a 32-wide window; a
    // layer is a 32 x 32 bit matrix (lane = shift, bit = witness) whose transpose gives every lane the partial sum of its
    // own witness (r1cs_shift_kernel).  shl: 32 words per layer (bit row, 0xffffffff = none); shm: per layer base | negative
    // << 8; shh: per constraint {first layer, number of layers, constraint index, 0}.
    std::vector<uint32_t> shl, shm, shh;
    uint64_t n_shift_constraints = 0;
    // the constraints the per-witness kernel has to walk (every mode but 3), in order
    std::vector<uint32_t> active;
    uint64_t n_int_constraints = 0, n_field_constraints = 0;
    // 32x32->64 multiply-accumulates the check executes per witness on this layout (upper bound: products with a 0 / +-1
    // factor are skipped at run time), and its executed field additions of bit-row terms
    uint64_t macs = 0, bit_adds = 0;
};

static const uint32_t LOC_BIT = 0x80000000u;   // tape::ROW_BIT
static const int TABLE_VARS = 5;

// wire_loc == nullptr: the plain layout (row = wire, no bit rows).  one_brow: the bit row that holds the constant 1 (wire 0
// itself is a field row): terms on wire 0 use it in constraints evaluated in integers.
// Every term is evaluated against the STORED row of its wire -- also for wires the program binds to constants: the check
// is an independent evaluation of the .r1cs on the witness as stored, it does not trust the tape (const_rows is unused).
inline Bound bind(const File &f, const uint32_t *wire_loc, uint32_t one_brow = 0, const uint32_t (*const_rows)[2] = nullptr) {
    Bound b;
    (void)const_rows;
    auto fold = [&](uint32_t wire) -> uint32_t { return wire; };
    const size_t n_lc = f.ptr.size() - 1;
    b.hdr.assign(4 * (n_lc + 1), 0);
    b.bhdr.assign(4 * ((size_t)f.n_constraints + 1), 0);
    b.fterms.reserve(f.terms.size());
    auto loc_of = [&](uint32_t wire) -> uint32_t { return wire_loc ? wire_loc[wire] : wire; };
    for (uint32_t c = 0; c < f.n_constraints; c++) {
        // source range of each LC (a B that repeats A has no terms of its own in File)
        uint32_t sb[3], se[3];
        bool same_b = false;
        for (int k = 0; k < 3; k++) {
            const size_t j = 3 * (size_t)c + k;
            sb[k] = f.ptr[j];
            se[k] = f.ptr[j + 1];
            if (k == 1 && f.split[3 * j] == SAME_AS_A) same_b = true;
        }
        bool intok = wire_loc != nullptr;
        uint64_t max_sumabs = 0;
        for (int k = 0; k < 3 && intok; k++) {
            const int src = (k == 1 && same_b) ? 0 : k;
            uint64_t sumabs = 0;
            for (uint32_t t = sb[src]; t < se[src] && intok; t++) {
                const Term &tm = f.terms[t];
                const uint32_t wire = fold(tm.wire & 0x0fffffffu);
                if (wire == 0xffffffffu) continue;
                if ((wire != 0 && !(loc_of(wire) & LOC_BIT)) || !f.cint_ok[tm.coef]) { intok = false; break; }
                const int64_t v = f.cint[tm.coef];
                sumabs += (uint64_t)(v < 0 ? -v : v);
                if (sumabs >> 62) intok = false;
            }
            max_sumabs = std::max(max_sumabs, sumabs);
        }
        uint32_t mode = !intok ? 0u : (max_sumabs >> 31) ? 1u : 2u;
        if (intok) {
            // distinct variables (bit rows other than the constant 1)
            uint32_t vars[TABLE_VARS];
            int nv = 0;
            bool fits = true;
            for (int k = 0; k < 3 && fits; k++) {
                const int src = (k == 1 && same_b) ? 0 : k;
                for (uint32_t t = sb[src]; t < se[src] && fits; t++) {
                    const uint32_t wire = fold(f.terms[t].wire & 0x0fffffffu);
                    if (wire == 0 || wire == 0xffffffffu) continue;
                    const uint32_t row = loc_of(wire) & ~LOC_BIT;
                    bool seen = false;
                    for (int q = 0; q < nv; q++) seen = seen || vars[q] == row;
                    if (seen) continue;
                    if (nv == TABLE_VARS) fits = false;
                    else vars[nv++] = row;
                }
            }
            if (fits && nv == 0) {
                // no variable at all (an empty constraint, or constants only): decided here once and for all
                __int128 lc[3] = {0, 0, 0};
                for (int k = 0; k < 3; k++) {
                    const int src = (k == 1 && same_b) ? 0 : k;
                    for (uint32_t t = sb[src]; t < se[src]; t++)
                        if (fold(f.terms[t].wire & 0x0fffffffu) == 0) lc[k] += f.cint[f.terms[t].coef];
                }
                if (lc[0] * lc[1] == lc[2]) mode = 3;   // always satisfied: nothing to evaluate
                fits = false;
            }
            if (fits) {
                uint32_t table = 0;
                for (uint32_t asg = 0; asg < (1u << TABLE_VARS); asg++) {
                    __int128 lc[3] = {0, 0, 0};
                    for (int k = 0; k < 3; k++) {
                        const int src = (k == 1 && same_b) ? 0 : k;
                        for (uint32_t t = sb[src]; t < se[src]; t++) {
                            const uint32_t wire = fold(f.terms[t].wire & 0x0fffffffu);
                            if (wire == 0xffffffffu) continue;
                            int64_t bit = 1;
                            if (wire != 0) {
                                const uint32_t row = loc_of(wire) & ~LOC_BIT;
                                for (int q = 0; q < nv; q++)
                                    if (vars[q] == row) bit = (asg >> q) & 1;
                            }
                            lc[k] += (__int128)f.cint[f.terms[t].coef] * bit;
                        }
                    }
                    if (lc[0] * lc[1] == lc[2]) table |= 1u << asg;
                }
                if (table == 0xffffffffu) {
                    // true for EVERY assignment of its 0/1 wires (x * (x - 1) = 0 on a wire stored as a bit): a bit row cannot
                    // hold anything that violates it -- nothing to evaluate
                    b.n_tautologies++;
                } else {
                    for (int q = 0; q < TABLE_VARS; q++) b.tcons.push_back(q < nv ? vars[q] : one_brow);
                    b.tcons.push_back(table);
                    b.tcons.push_back(c);
                    b.tcons.push_back(0);
                    b.n_table_constraints++;
                }
                mode = 3;
            }
        }
        if (intok && mode != 3 && wire_loc != nullptr) {
            // linear, every coefficient +-2^k: a shift-sum constraint
            const bool a_none = sb[0] == se[0], b_none = !same_b && sb[1] == se[1];
            bool ok = a_none || b_none;
            struct T { uint32_t sh, row; bool neg; };
            std::vector<T> ts;
            for (uint32_t t = sb[2]; t < se[2] && ok; t++) {
                const Term &tm = f.terms[t];
                const uint32_t wire = fold(tm.wire & 0x0fffffffu);
                if (wire == 0xffffffffu) continue;
                const int64_t v = f.cint[tm.coef];
                const uint64_t m = (uint64_t)(v < 0 ? -v : v);
                if (m == 0 || (m & (m - 1))) { ok = false; break; }
                uint32_t sh = 0;
                while (!((m >> sh) & 1)) sh++;
                ts.push_back(T{sh, wire == 0 ? one_brow : (loc_of(wire) & ~LOC_BIT), v < 0});
            }
            if (ok && !ts.empty()) {
                std::stable_sort(ts.begin(), ts.end(), [](const T &x, const T &y) { return x.neg != y.neg ? x.neg < y.neg : x.sh < y.sh; });
                struct L { uint32_t base; bool neg; uint32_t row[32]; };
                std::vector<L> layers;
                for (const T &t : ts) {
                    bool placed = false;
                    for (L &l : layers)
                        if (l.neg == t.neg && t.sh >= l.base && t.sh - l.base < 32 && l.row[t.sh - l.base] == 0xffffffffu) {
                            l.row[t.sh - l.base] = t.row;
                            placed = true;
                            break;
                        }
                    if (!placed) {
                        L l;
                        l.base = t.sh;
                        l.neg = t.neg;
                        for (uint32_t &x : l.row) x = 0xffffffffu;
                        l.row[0] = t.row;
                        layers.push_back(l);
                    }
                }
                if (layers.size() * 6 < ts.size() && layers.size() < 65536) {   // a layer costs about as much as six terms
                    b.shh.push_back((uint32_t)b.shm.size());
                    b.shh.push_back((uint32_t)layers.size());
                    b.shh.push_back(c);
                    b.shh.push_back(0);
                    for (const L &l : layers) {
                        b.shm.push_back(l.base | (l.neg ? 1u << 8 : 0u));
                        for (uint32_t x : l.row) b.shl.push_back(x);
                    }
                    b.n_shift_constraints++;
                    mode = 3;
                }
            }
        }
        if (intok) b.n_int_constraints++; else b.n_field_constraints++;
        if (mode != 3) b.active.push_back(c);
        bool has[3] = {false, false, false};
        for (int k = 0; k < 3; k++) {
            const size_t j = 3 * (size_t)c + k;
            const uint32_t fb = (uint32_t)b.fterms.size();
            b.hdr[4 * j] = fb;
            if (b.bterms.size() >> 30) throw std::runtime_error("more than 2^30 terms on bit rows");
            b.bhdr[4 * (size_t)c + k] = (uint32_t)b.bterms.size() | (k == 0 ? mode << 30 : 0u);
            if (c > 0 && k == 0) b.bhdr[4 * (size_t)c - 1] = (uint32_t)b.bterms.size();
            if (k == 1 && same_b && !intok) {
                b.hdr[4 * j + 1] = SAME_AS_A;
                b.hdr[4 * j + 2] = fb;
                b.hdr[4 * j + 3] = fb;
                has[1] = true;
                continue;
            }
            if (mode == 3) continue;   // evaluated from its truth table: no terms
            const int src = (k == 1 && same_b) ? 0 : k;
            const size_t js = 3 * (size_t)c + src;
            const uint32_t e0 = f.split[3 * js], e1 = f.split[3 * js + 1], e2 = f.split[3 * js + 2];
            uint32_t n[4] = {0, 0, 0, 0}, n_const = 0;
            for (uint32_t t = sb[src]; t < se[src]; t++) {
                const Term &tm = f.terms[t];
                const uint32_t wire = fold(tm.wire & 0x0fffffffu);
                if (wire == 0xffffffffu) continue;   // a wire that is always 0
                uint32_t loc = loc_of(wire);
                if (intok && wire == 0) loc = LOC_BIT | one_brow;
                if (loc & LOC_BIT) {
                    b.bterms.push_back(Term{loc & ~LOC_BIT, tm.coef});
                    if (!intok) b.bit_adds++;
                    continue;
                }
                b.fterms.push_back(Term{loc | (tm.wire & 0xf0000000u), tm.coef});
                const int cls = t < e0 ? 0 : t < e1 ? 1 : t < e2 ? 2 : 3;
                n[cls]++;
                if (cls == 3 && (tm.wire >> 31)) n_const++;
            }
            b.hdr[4 * j + 1] = fb + n[0];
            b.hdr[4 * j + 2] = fb + n[0] + n[1];
            b.hdr[4 * j + 3] = fb + n[0] + n[1] + n[2];
            const uint64_t n3 = n[3] - n_const;
            b.macs += 8 * (uint64_t)(n[1] + n[2]) + 170 * (uint64_t)((n[1] ? 1 : 0) + (n[2] ? 1 : 0)) + 64 * n3 + 72 * ((n3 + 15) / 16);
            has[k] = se[src] > sb[src];
        }
        if (!intok && has[0] && has[1]) b.macs += 136;
    }
    b.hdr[4 * n_lc] = (uint32_t)b.fterms.size();
    if (f.n_constraints) b.bhdr[4 * (size_t)f.n_constraints - 1] = (uint32_t)b.bterms.size();
    b.bhdr[4 * (size_t)f.n_constraints] = (uint32_t)b.bterms.size();
    return b;
}

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

inline File load(const std::string &path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw Error("cannot open " + path);
    std::vector<uint8_t> d((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    auto need = [&](size_t pos, size_t n) {
        if (pos + n > d.size()) throw Error("truncated r1cs file");
    };
    auto u32 = [&](size_t pos) { need(pos, 4); uint32_t v; memcpy(&v, &d[pos], 4); return v; };
    auto u64 = [&](size_t pos) { need(pos, 8); uint64_t v; memcpy(&v, &d[pos], 8); return v; };
    need(0, 12);
    if (memcmp(&d[0], "r1cs", 4) != 0) throw Error("not an r1cs file");
    if (u32(4) != 1) throw Error("unsupported r1cs version");
    uint32_t nsec = u32(8);
    size_t pos = 12;
    size_t sec_pos[6] = {0, 0, 0, 0, 0, 0};
    uint64_t sec_len[6] = {0, 0, 0, 0, 0, 0};
    bool have[6] = {false, false, false, false, false, false};
    for (uint32_t s = 0; s < nsec; s++) {
        uint32_t typ = u32(pos);
        uint64_t len = u64(pos + 4);
        pos += 12;
        need(pos, len);
        if (typ < 6) { sec_pos[typ] = pos; sec_len[typ] = len; have[typ] = true; }
        pos += len;
    }
    if (!have[1] || !have[2]) throw Error("r1cs file lacks header or constraint section");
    File out;
    size_t hp = sec_pos[1];
    out.field_size = u32(hp);
    if (out.field_size != 32) throw Error("only 32-byte fields (bn128) are supported");
    static const uint8_t qle[32] = {0x01, 0x00, 0x00, 0xf0, 0x93, 0xf5, 0xe1, 0x43, 0x91, 0x70, 0xb9, 0x79, 0x48, 0xe8, 0x33, 0x28,
                                    0x5d, 0x58, 0x81, 0x81, 0xb6, 0x45, 0x50, 0xb8, 0x29, 0xa0, 0x31, 0xe1, 0x72, 0x4e, 0x64, 0x30};
    need(hp + 4, 32);
    if (memcmp(&d[hp + 4], qle, 32) != 0) throw Error("r1cs prime is not bn128");
    size_t q = hp + 4 + 32;
    out.n_wires = u32(q);
    out.n_pub_out = u32(q + 4);
    out.n_pub_in = u32(q + 8);
    out.n_prv_in = u32(q + 12);
    out.n_labels = u64(q + 16);
    out.n_constraints = u32(q + 24);
    if (have[4] || have[5]) throw Error("custom-gate sections are not supported");
    if (out.n_wires >> 28) throw Error("more than 2^28 wires");   // the top four bits of a term's wire word carry its class

    std::unordered_map<fr::Fr, uint32_t, tape::FrHash, tape::FrEq> index;
    fr::Fr one = fr::zero();
    one.v[0] = 1;
    fr::Fr minus_one = fr::neg(one);
    out.coefs.push_back(one);
    out.coefs.push_back(minus_one);
    index.emplace(one, 0);
    index.emplace(minus_one, 1);

    size_t p = sec_pos[2], end = sec_pos[2] + sec_len[2];
    out.ptr.reserve((size_t)out.n_constraints * 3 + 1);
    out.ptr.push_back(0);
    for (uint32_t c = 0; c < out.n_constraints; c++) {
        for (int k = 0; k < 3; k++) {
            uint32_t n = u32(p);
            p += 4;
            for (uint32_t t = 0; t < n; t++) {
                uint32_t wire = u32(p);
                need(p + 4, 32);
                fr::Fr v;
                memcpy(v.v, &d[p + 4], 32);
                p += 36;
                if (wire >= out.n_wires) throw Error("r1cs term refers to a wire out of range");
                if (!fr::gt_raw(fr::modulus(), v)) throw Error("r1cs coefficient is not reduced");
                auto it = index.find(v);
                uint32_t ci;
                if (it == index.end()) {
                    ci = (uint32_t)out.coefs.size();
                    out.coefs.push_back(v);
                    index.emplace(v, ci);
                } else ci = it->second;
                if (ci < 2) out.nnz_pm1++;
                if (fr::is_zero(v)) continue;
                out.terms.push_back(Term{wire, ci});
            }
            out.ptr.push_back((uint32_t)out.terms.size());
        }
    }
    if (p > end) throw Error("constraint section overrun");
    // ---- coefficient classes and per-LC term order
    // 0: +-2^k with k <= 3 (k modular doublings and one add/sub; +-1 is k = 0), 1: small positive, 2: small negative,
    // 3: general.  cmag: class 0 -> sign | k << 1, classes 1/2 -> the 32-bit magnitude.
    std::vector<uint8_t> kind(out.coefs.size(), 3);
    out.cmag.assign(out.coefs.size(), 0);
    kind[0] = kind[1] = 0;
    out.cmag[1] = 1;
    for (size_t i = 2; i < out.coefs.size(); i++) {
        const fr::Fr &c = out.coefs[i];
        fr::Fr n = fr::neg(c);
        auto small = [](const fr::Fr &x) {
            for (int k = 1; k < 8; k++)
                if (x.v[k]) return false;
            return true;
        };
        const bool sp = small(c), sn = small(n);
        const uint32_t m = sp ? c.v[0] : n.v[0];
        if ((sp || sn) && (m == 2 || m == 4 || m == 8)) {
            kind[i] = 0;
            out.cmag[i] = (sn ? 1u : 0u) | ((m == 2 ? 1u : m == 4 ? 2u : 3u) << 1);
        } else if (sp) { kind[i] = 1; out.cmag[i] = c.v[0]; }
        else if (sn) { kind[i] = 2; out.cmag[i] = n.v[0]; }
    }
    out.cint.assign(out.coefs.size(), 0);
    out.cint_ok.assign(out.coefs.size(), 0);
    for (size_t i = 0; i < out.coefs.size(); i++) {
        const fr::Fr &c = out.coefs[i];
        const fr::Fr n = fr::neg(c);
        auto fits = [](const fr::Fr &x) {
            for (int k = 2; k < 8; k++)
                if (x.v[k]) return false;
            return (x.v[1] >> 30) == 0;
        };
        if (fits(c)) { out.cint[i] = (int64_t)(((uint64_t)c.v[1] << 32) | c.v[0]); out.cint_ok[i] = 1; }
        else if (fits(n)) { out.cint[i] = -(int64_t)(((uint64_t)n.v[1] << 32) | n.v[0]); out.cint_ok[i] = 1; }
    }
    out.split.resize(3 * (size_t)(out.ptr.size() - 1));
    for (size_t j = 0; j + 1 < out.ptr.size(); j++) {
        auto b = out.terms.begin() + (ptrdiff_t)out.ptr[j], e = out.terms.begin() + (ptrdiff_t)out.ptr[j + 1];
        size_t cnt[4] = {0, 0, 0, 0};
        for (auto it = b; it != e; ++it) cnt[kind[it->coef]]++;
        const bool use1 = cnt[1] >= 4, use2 = cnt[2] >= 4;
        auto cls = [&](const Term &t) -> int {
            int k = kind[t.coef];
            if ((k == 1 && !use1) || (k == 2 && !use2)) k = 3;
            if (k == 3 && (t.wire & 0x0fffffffu) == 0) return 4;
            return k;
        };
        std::stable_sort(b, e, [&](const Term &x, const Term &y) { return cls(x) < cls(y); });
        size_t n0 = cnt[0], n1 = use1 ? cnt[1] : 0, n2 = use2 ? cnt[2] : 0;
        // the +-2^k terms carry their sign | k << 1 in the top four bits of the wire index (one load less per term)
        for (auto it = b; it != b + (ptrdiff_t)n0; ++it) {
            it->wire |= out.cmag[it->coef] << 28;
        }
        out.split[3 * j] = out.ptr[j] + (uint32_t)n0;
        out.split[3 * j + 1] = out.ptr[j] + (uint32_t)(n0 + n1);
        out.split[3 * j + 2] = out.ptr[j] + (uint32_t)(n0 + n1 + n2);
        out.nnz_small += n1 + n2;
        size_t n3 = (size_t)(e - b) - n0 - n1 - n2;
        if (n3 && cls(*(e - 1)) == 4) {   // an LC is a map wire -> coefficient: at most one term on wire 0
            (e - 1)->wire |= 0x80000000u;
            n3--;
            out.nnz_const++;
        }
        out.macs += 8 * (n1 + n2) + 170 * ((n1 ? 1 : 0) + (n2 ? 1 : 0)) + 64 * n3 + 72 * ((n3 + 15) / 16);
        out.nnz_general += n3;
    }
    out.nnz = out.terms.size();
    for (uint32_t c = 0; c < out.n_constraints; c++) {
        bool quad = out.ptr[3 * c] != out.ptr[3 * c + 1] && out.ptr[3 * c + 1] != out.ptr[3 * c + 2];
        if (quad) { out.n_quadratic++; out.macs += 136; }
        else out.n_linear++;
    }
    // ---- squares: when B is the same combination as A (x*x of an S-box) its terms are dropped from the CSR and its
    // header carries SAME_AS_A in the first class boundary; the kernel evaluates A once and squares it.
    {
        std::vector<Term> terms;
        std::vector<uint32_t> ptr, split;
        terms.reserve(out.terms.size());
        ptr.reserve(out.ptr.size());
        split.reserve(out.split.size());
        ptr.push_back(0);
        for (size_t j = 0; j + 1 < out.ptr.size(); j++) {
            const uint32_t b = out.ptr[j], e = out.ptr[j + 1];
            bool same = false;
            if (j % 3 == 1 && e > b) {
                const uint32_t ab = out.ptr[j - 1];
                same = (b - ab) == (e - b);
                for (uint32_t k = 0; same && k < e - b; k++)
                    same = out.terms[ab + k].wire == out.terms[b + k].wire && out.terms[ab + k].coef == out.terms[b + k].coef;
            }
            const uint32_t nb = (uint32_t)terms.size();
            if (same) {
                out.n_squares++;
                // (what B would have cost is no longer executed)
                const size_t n0 = out.split[3 * j] - b, n1 = out.split[3 * j + 1] - out.split[3 * j], n2 = out.split[3 * j + 2] - out.split[3 * j + 1];
                size_t n3 = (e - b) - n0 - n1 - n2;
                if (n3 && (out.terms[e - 1].wire >> 31)) n3--;
                out.macs -= 8 * (n1 + n2) + 170 * ((n1 ? 1 : 0) + (n2 ? 1 : 0)) + 64 * n3 + 72 * ((n3 + 15) / 16);
                split.push_back(SAME_AS_A);
                split.push_back(nb);
                split.push_back(nb);
            } else {
                terms.insert(terms.end(), out.terms.begin() + b, out.terms.begin() + e);
                for (int k = 0; k < 3; k++) split.push_back(out.split[3 * j + k] - b + nb);
            }
            ptr.push_back((uint32_t)terms.size());
        }
        out.terms.swap(terms);
        out.ptr.swap(ptr);
        out.split.swap(split);
    }
    if (have[3]) {
        size_t wp = sec_pos[3];
        for (uint64_t i = 0; i < sec_len[3] / 8; i++) out.wire2label.push_back(u64(wp + 8 * i));
    }
    return out;
}

}  // namespace r1cs
