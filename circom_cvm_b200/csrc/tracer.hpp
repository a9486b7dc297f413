// Trace-compiler: runs the CVM program ONCE on the host with abstract values and leaves a flat,
// address-resolved SSA tape of field operations for the GPU.
//
// Why this is sound: the component tree, signal offsets, loop trip counts and array indices of a
// circom program depend only on template parameters, which the compiler already resolved to
// literals (template.rs:240-290, create_component_bucket.rs:221-350) -- so everything except
// arithmetic on *signal values* is identical for every witness of a batch and is evaluated here.
// What the reference's generated C++ does per witness (generated `T_run` bodies: template.rs:334-407;
// sub-component triggering: store_bucket.rs:662-800; calls: call_bucket.rs:465-534) is replayed
// with three kinds of values: known i64 (addresses/counters), known field constants, and dynamic
// values (SSA ids).  Data-dependent `if`s are if-converted (both arms traced, writes merged with SEL);
// data-dependent loops / returns / addresses are rejected with a clear error.
//
// Asserts: `c = ff.eqz v; if c; error 0; end` (assert_bucket.rs:88-107) becomes FAIL_IF(c).  The
// `===` half of every `<==` recomputes the stored expression (translate.rs:676-734); value numbering
// makes that `ff.eq x x`, which folds to 1, so those asserts vanish -- they can never fail.
#pragma once
#include <algorithm>
#include <cstdint>
#include <deque>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "cvm_parse.hpp"
#include "host_fr.hpp"

namespace tape {

enum TOp : uint8_t {
    T_NOP = 0, T_INPUT, T_ADD, T_SUB, T_MUL, T_DIV, T_IDIV, T_MOD, T_POW, T_SHL, T_SHR, T_BAND, T_BOR, T_BXOR,
    T_BNOT, T_LT, T_LE, T_GT, T_GE, T_EQ, T_NEQ, T_LAND, T_LOR, T_EQZ, T_SEL, T_FAIL_IF,
    // produced by the tracer's peepholes
    T_FAIL_NE,   // status if a != b                (FAIL_IF(NEQ(a, b)): the `===` assert shape)
    T_BITC,      // bit `aux` of the RAW limbs of a  (a = y * R^-1, whose Montgomery limbs are the canonical y)
    T_LUT,       // boolean function of up to three 0/1 values a, b, c: bit (a | b<<1 | c<<2) of the 8-bit table in aux
                 // (aux >> 8 = number of inputs).  Produced by the boolean-cone collapsing below.
    T_INV,       // a^-1, with 0^-1 := 0  (ff.div a b is traced as a * INV(b) so that independent inversions can be batched)
    // inserted by the tape builder
    T_CADD,      // a + (b != 0 ? c : 0) with a constant c: the `acc += bit * 2^k` step of every bit-weighted sum (tape.hpp)
    T_DOT,       // sum_k c_k * x_k (+ addend): fused tree of additions of products by constants (tape.hpp fuse_dots)
    T_LD, T_ST, T_STC,
    // small-integer arithmetic (tape.hpp type_ints): values proven below 2^62 that only feed each other and bit
    // extractions -- the bit-weighted sums of BinSum / Bits2Num -- are kept as raw 64-bit integers in their slot
    T_ICADD,     // a + (b != 0 ? K : 0), K = integer constant c
    T_IADD,      // a + b
    T_ISEL,      // a != 0 ? b : c
    T_IBIT,      // bit `aux` of the integer a (a value typed 0/1)
    T_IFAIL_NE,  // status if the integers a and b differ
    T_ISUM,      // addend + sum_j (bit_j << shift_j): a chain of T_ICADD with power-of-two constants (tape.hpp fuse_isums)
    // warp-cooperative group instructions (tape.hpp group_bit_ops): up to 32 independent operations on values typed 0/1,
    // one per LANE, each lane working on the packed word of the warp's 32 witnesses
    T_LUTG,      // up to 32 T_LUT
    T_IBITG,     // up to 32 consecutive bits of one integer (T_IBIT)
    T_FILL,      // constant bit rows [c, c + b) = the word a (witness wires bound to the constants 0 / 1)
    // R1CS check scheduled into the tape (fused.hpp)
    T_RNE,       // first_bad = min(first_bad, c) if a != b   (a = A.w * B.w, b = C.w of constraint c)
    // speculative typing (Tracer::assume_bit_inputs): main input `a`, which must be literally 0 or 1 -- typed 0/1; any
    // other value raises ST_SPECULATION and the witness is recomputed by the program traced without the assumption
    T_INPUT_BIT,
    // T_ISUM with its terms dealt into layers of distinct shifts: a = number of layers, b = addend; per layer 8 records =
    // 32 words, word l = bit slot | base << 16 of the term whose shift is base + l (slot 0xffff: none).  Produced by the
    // allocator only (tape.hpp), never an SSA operation.
    T_ISUMT,
    T_INBITG,    // up to 32 consecutive T_INPUT_BIT (first input index in c): record m = (bit slot, -, -, bit row or NO_ROW)
    T_COUNT
};

static const uint32_t CONST_FLAG = 0x80000000u;
static const uint32_t NO_REF = 0xffffffffu;

enum Status : uint32_t { ST_OK = 0, ST_ASSERT = 1, ST_TOINT = 2, ST_DIVZERO = 3, ST_INPUT = 4, ST_LOOP = 5, ST_SPECULATION = 6 };

struct SOp {
    uint8_t op;
    uint32_t a, b, c;   // refs: CONST_FLAG | const index, or value id (= index of the producing SOp)
    uint32_t aux;       // T_INPUT: input index; T_FAIL_IF: status code
};

struct TraceError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

struct TraceStats {
    uint64_t cvm_instructions = 0;   // CVM instructions executed on the host
    uint64_t ref_mul = 0;            // dynamic ff.mul (+1 per ff.div) the reference program executes: N_mul
    uint64_t ref_div = 0;
    uint64_t ref_field_ops = 0;      // all ff.* executions of the reference program
    uint64_t folded = 0, cse_hits = 0;
    uint64_t dyn_branches = 0;
    uint64_t dyn_addresses = 0;     // ff.wrap_i64 of a witness value (data-dependent array index)
    uint64_t luts = 0;                  // boolean cones emitted as one T_LUT
    uint64_t unrolled_iterations = 0;   // iterations of data-dependent while loops traced under predicates
    uint64_t unset_signal_reads = 0;
};

struct FrHash {
    size_t operator()(const fr::Fr &f) const {
        uint64_t h = 1469598103934665603ull;
        for (int i = 0; i < 8; i++) h = (h ^ f.v[i]) * 1099511628211ull;
        return (size_t)h;
    }
};
struct FrEq {
    bool operator()(const fr::Fr &a, const fr::Fr &b) const { return fr::equal(a, b); }
};

class Tracer {
  public:
    // AV_DADDR: an address that depends on a witness value: i + scale * toInt(value `ref`) (ff.wrap_i64 of a dynamic value,
    // then i64.add / i64.mul with constants).  Loads and variable stores through it are traced as selections over the
    // addresses it can name (mux_candidates); everything else that needs a static integer rejects it.
    enum { AV_UNDEF = 0, AV_I64, AV_FF, AV_DYN, AV_DADDR };
    struct AV {
        uint8_t kind = AV_UNDEF;
        int64_t i = 0;
        uint32_t ref = 0;   // AV_FF: const index ; AV_DYN, AV_DADDR: value id
        int64_t scale = 0;  // AV_DADDR
    };

    // Trace under the assumption that every main input is 0 or 1 (hash circuits take their message as unconstrained
    // signals: nothing in Sha256(n) proves in[k] a bit, so everything derived from the message before the first bit
    // decomposition is field arithmetic).  The assumption is CHECKED per witness at run time (T_INPUT_BIT).
    bool assume_bit_inputs = false;
    // Emit boolean cones as T_LUT (operands and result in the bit-slot file).  Off for programs that are traced for the
    // untyped tape (a field program with a sprinkling of 0/1 values: everything stays a field element).
    bool use_luts = true;
    std::vector<SOp> ops;
    std::vector<uint8_t> isbool;  // per op: the value is provably 0 or 1 (comparison results, extracted bits, ANDs of those)
    // Boolean-cone collapsing.  For every value that is an arithmetic function of at most three provably-0/1 values, the
    // function is tabulated while tracing (small signed integers; anything else invalidates the cone).  A value whose
    // table only holds 0 and 1 IS a boolean function of its leaves -- XOR written as a+b-2ab, circomlib's Xor3 / Ch / Maj
    // polynomials, ANDs, multiplexers -- and is emitted as one T_LUT instead of its cone of field operations; the cone's
    // intermediate values die unless something else uses them.  Exact: the leaves can only be 0 or 1.
    struct Cone {
        uint8_t k = 0xff;            // number of leaves, 0xff = not tabulated
        uint32_t leaf[3] = {0, 0, 0};
        int8_t tt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    };
    std::vector<Cone> cones;
    std::vector<uint32_t> alias;  // per op: the T_LUT (or constant / leaf) that replaced it, for later CSE hits
    std::vector<fr::Fr> consts;   // canonical
    std::vector<uint32_t> witness_ref;   // per witness wire: ref (const or value)
    TraceStats stats;
    int64_t n_inputs = 0, n_outputs = 0;

    uint32_t rinv_ref() const { return CONST_FLAG | c_rinv; }
    uint32_t zero_ref() const { return CONST_FLAG | c_zero; }
    uint32_t one_ref() const { return CONST_FLAG | c_one; }

    explicit Tracer(const cvm::Program &p) : prog(p) {
        ffmap.resize(p.ffconst.size(), NO_REF);
        c_zero = intern(fr::zero());
        c_one = intern(hostfr::from_u64(1));
        fr::Fr raw_one = hostfr::from_u64(1);
        c_rinv = intern(fr::from_mont(raw_one));            // R^-1 mod q
    }

    void trace() {
        const cvm::Code &mainc = prog.codes[prog.start];
        sig.assign((size_t)prog.n_signals, AV());
        sig[0] = ff_av(c_one);                               // calcwit.cpp:34
        n_inputs = mainc.n_inputs;
        n_outputs = mainc.n_outputs;
        int main_idx = new_comp(prog.start, 1);              // circuit.rs:539: main at signal 1
        for (int64_t k = 0; k < n_inputs; k++) {             // main inputs follow main outputs (App. A.5)
            SOp o{(uint8_t)(assume_bit_inputs ? T_INPUT_BIT : T_INPUT), NO_REF, NO_REF, NO_REF, (uint32_t)k};
            push_op(o);
            AV v;
            v.kind = AV_DYN;
            v.ref = (uint32_t)ops.size() - 1;
            sig[(size_t)(1 + n_outputs + k)] = v;
        }
        run_component(main_idx);
        witness_ref.clear();
        for (int64_t s : prog.witness) {
            if (s < 0 || s >= prog.n_signals) throw TraceError("witness list refers to signal out of range");
            witness_ref.push_back(field_ref(sig[(size_t)s]));
        }
    }

  private:
    struct Frame {
        std::vector<AV> regs;
        std::unordered_map<int64_t, AV> lvar;
    };
    struct Comp {
        int code;
        int64_t start;
        int64_t counter;
        std::vector<int> subs;
    };
    struct LogEntry {
        uint8_t space;   // 0 reg, 1 lvar, 2 signal
        Frame *frame;
        int64_t key;
        AV old;
        bool existed;
    };
    struct RetTarget {
        Frame *frame;
        int64_t addr, size;
        size_t pred_depth;
        // disjunction of the predicates under which a `return` has been traced so far (NO_REF: none): what follows a
        // data-dependent return runs under its negation
        mutable uint32_t returned = NO_REF;
    };

    const cvm::Program &prog;
    std::vector<AV> sig;
    std::deque<Comp> comps;
    std::vector<uint32_t> ffmap;
    std::unordered_map<fr::Fr, uint32_t, FrHash, FrEq> const_index;
    std::unordered_map<uint64_t, std::vector<uint32_t>> cse;
    std::vector<std::vector<LogEntry>> logs;
    std::vector<uint32_t> preds;   // refs of enclosing data-dependent conditions (already "is true" predicates)
    uint32_t c_zero, c_one, c_rinv;
    int depth = 0;
    std::vector<int> loop_stack;      // pc of the open `loop` instructions of the code being traced
    int unroll_depth = 0;
    int last_exit_pc = 0;
  public:
    int max_unroll = 260;             // iterations traced for a data-dependent while loop (254-bit scans fit)
    // the value behind a ref (constant or traced value) is provably 0 or 1 for every input
    bool ref_is_bool(uint32_t r) const { return is_bool(r); }
  private:
    static const int64_t SPR_BASE = (int64_t)1 << 40;

    // ---------------------------------------------------------------- constants / values
    uint32_t intern(const fr::Fr &v) {
        auto it = const_index.find(v);
        if (it != const_index.end()) return it->second;
        uint32_t id = (uint32_t)consts.size();
        consts.push_back(v);
        const_index.emplace(v, id);
        return id;
    }
    static AV i64_av(int64_t x) {
        AV v;
        v.kind = AV_I64;
        v.i = x;
        return v;
    }
    static AV ff_av(uint32_t cidx) {
        AV v;
        v.kind = AV_FF;
        v.ref = cidx;
        return v;
    }
    static AV dyn_av(uint32_t id) {
        AV v;
        v.kind = AV_DYN;
        v.ref = id;
        return v;
    }
    AV ref_av(uint32_t ref) { return (ref & CONST_FLAG) ? ff_av(ref & ~CONST_FLAG) : dyn_av(ref); }

    // field view of an abstract value -> ref
    uint32_t field_ref(const AV &v) {
        switch (v.kind) {
            case AV_I64: return CONST_FLAG | intern(hostfr::from_i64(v.i));
            case AV_FF: return CONST_FLAG | v.ref;
            case AV_DYN: return v.ref;
            default: stats.unset_signal_reads++; return CONST_FLAG | c_zero;
        }
    }
    // integer view (addresses, counters); data-dependent addressing is not traceable
    int64_t int_of(const AV &v, const char *what) {
        if (v.kind == AV_I64) return v.i;
        if (v.kind == AV_FF) {
            int32_t out;
            if (!fr::to_int(consts[v.ref], out))
                throw TraceError(std::string("field constant does not fit an address (") + what + ")");
            return out;
        }
        if (v.kind == AV_DYN || v.kind == AV_DADDR)
            throw TraceError(std::string("data-dependent ") + what + " is not supported by the trace compiler");
        return 0;
    }
    // ---- data-dependent addresses
    static const size_t MAX_MUX = 4096;
    // Upper bound of a value as a non-negative integer when the shape of its computation gives one (masks, remainders,
    // shifts, sums and products of such, truth values), else -1.  Bounds the locations an index can name: without one a
    // variable store through a data-dependent index has to treat every variable of the frame as a possible target.
    std::vector<int64_t> ub_memo;
    int64_t ubound(uint32_t r, int depth = 0) {
        static const int64_t LIM = (int64_t)1 << 40;
        if (r == NO_REF) return -1;
        if (r & CONST_FLAG) {
            const int64_t c = small_const(r);
            return c;
        }
        if (isbool[r]) return 1;
        if (depth > 64) return -1;
        if (ub_memo.size() < ops.size()) ub_memo.resize(ops.size(), -2);
        if (ub_memo[r] != -2) return ub_memo[r];
        const SOp &o = ops[r];
        int64_t res = -1;
        const int64_t a = (o.a != NO_REF) ? ubound(o.a, depth + 1) : -1, b = (o.b != NO_REF) ? ubound(o.b, depth + 1) : -1;
        switch (o.op) {
            case T_BAND: res = a < 0 ? b : b < 0 ? a : std::min(a, b); break;
            case T_MOD: if (b > 0) res = a >= 0 ? std::min(a, b - 1) : b - 1; break;
            case T_IDIV: if (a >= 0 && is_const(o.b) && b > 0) res = a / b; break;
            case T_SHR: if (a >= 0 && is_const(o.b) && b >= 0 && b < 63) res = a >> b; else if (a >= 0) res = a; break;
            case T_ADD: if (a >= 0 && b >= 0 && a + b < LIM) res = a + b; break;
            case T_MUL: if (a >= 0 && b >= 0 && (a == 0 || b < LIM / std::max<int64_t>(a, 1))) res = a * b; break;
            case T_SEL: {
                const int64_t c = (o.c != NO_REF) ? ubound(o.c, depth + 1) : -1;
                if (b >= 0 && c >= 0) res = std::max(b, c);
                break;
            }
            default: break;
        }
        if (res >= LIM) res = -1;
        ub_memo[r] = res;
        return res;
    }
    // the (index k, address) pairs an AV_DADDR can name inside [0, limit), smallest address first
    std::vector<std::pair<int64_t, int64_t>> mux_candidates(const AV &a, int64_t limit, const char *what) {
        std::vector<std::pair<int64_t, int64_t>> out;
        if (a.scale == 0) throw TraceError("degenerate data-dependent address");
        const int64_t kmax = ubound(a.ref);
        for (int64_t k = 0;; k++) {
            const int64_t addr = a.i + a.scale * k;
            if (kmax >= 0 && k > kmax) break;
            if (a.scale > 0 ? addr >= limit : addr < 0) break;
            if (addr >= 0 && addr < limit) out.emplace_back(k, addr);
            if (out.size() > MAX_MUX) throw TraceError(std::string("data-dependent ") + what + " ranges over more than 4096 locations");
            if (k > (int64_t)1 << 31) break;
        }
        return out;
    }
    uint32_t index_is(const AV &a, int64_t k) { return emit(T_EQ, a.ref, CONST_FLAG | intern(hostfr::from_i64(k))); }

    // ---------------------------------------------------------------- SSA emission
    static bool commutative(uint8_t op) {
        return op == T_ADD || op == T_MUL || op == T_EQ || op == T_NEQ || op == T_BAND || op == T_BOR ||
               op == T_BXOR || op == T_LAND || op == T_LOR;
    }
    static hostfr::FfOp host_op(uint8_t op) {
        switch (op) {
            case T_ADD: return hostfr::F_ADD; case T_SUB: return hostfr::F_SUB; case T_MUL: return hostfr::F_MUL;
            case T_DIV: return hostfr::F_DIV; case T_IDIV: return hostfr::F_IDIV; case T_MOD: return hostfr::F_MOD;
            case T_POW: return hostfr::F_POW; case T_SHL: return hostfr::F_SHL; case T_SHR: return hostfr::F_SHR;
            case T_BAND: return hostfr::F_BAND; case T_BOR: return hostfr::F_BOR; case T_BXOR: return hostfr::F_BXOR;
            case T_BNOT: return hostfr::F_BNOT; case T_LT: return hostfr::F_LT; case T_LE: return hostfr::F_LE;
            case T_GT: return hostfr::F_GT; case T_GE: return hostfr::F_GE; case T_EQ: return hostfr::F_EQ;
            case T_NEQ: return hostfr::F_NEQ; case T_LAND: return hostfr::F_LAND; case T_LOR: return hostfr::F_LOR;
            case T_EQZ: return hostfr::F_EQZ; case T_INV: return hostfr::F_INV; default: return hostfr::F_NONE;
        }
    }
    bool is_const(uint32_t r) const { return r != NO_REF && (r & CONST_FLAG); }
    const fr::Fr &cval(uint32_t r) const { return consts[r & ~CONST_FLAG]; }
    uint32_t push_op(const SOp &o) {
        ops.push_back(o);
        bool bl = false;
        switch (o.op) {
            case T_LT: case T_LE: case T_GT: case T_GE: case T_EQ: case T_NEQ: case T_LAND: case T_LOR: case T_EQZ:
            case T_BITC: case T_LUT: case T_INPUT_BIT: bl = true; break;
            case T_SEL: bl = is_bool(o.b) && is_bool(o.c); break;
            default: break;
        }
        isbool.push_back(bl ? 1 : 0);
        cones.emplace_back();
        alias.push_back(NO_REF);
        return (uint32_t)ops.size() - 1;
    }
    // ---- boolean cones
    static const int CONE_MAX = 100;   // |table entry| bound: cones of boolean logic only see tiny integers
    bool cone_of(uint32_t r, Cone &out) const {
        if (r == NO_REF) return false;
        if (r & CONST_FLAG) {
            const fr::Fr &c = cval(r);
            fr::Fr n = fr::neg(c);
            auto small = [](const fr::Fr &x, int &v) {
                for (int i = 1; i < 8; i++)
                    if (x.v[i]) return false;
                if (x.v[0] > (uint32_t)CONE_MAX) return false;
                v = (int)x.v[0];
                return true;
            };
            int v;
            out = Cone();
            out.k = 0;
            if (small(c, v)) { out.tt[0] = (int8_t)v; return true; }
            if (small(n, v)) { out.tt[0] = (int8_t)-v; return true; }
            return false;
        }
        if (cones[r].k != 0xff) { out = cones[r]; return true; }
        if (isbool[r]) {            // a 0/1 value that is not itself tabulated: a leaf
            out = Cone();
            out.k = 1;
            out.leaf[0] = r;
            out.tt[0] = 0;
            out.tt[1] = 1;
            return true;
        }
        return false;
    }
    // table of `op` over the union of the operands' leaves; false when it does not apply
    bool cone_eval(uint8_t op, const uint32_t *refs, int n, Cone &res) const {
        Cone in[3];
        for (int i = 0; i < n; i++)
            if (!cone_of(refs[i], in[i])) return false;
        res = Cone();
        uint32_t u[9];
        int nu = 0;
        for (int i = 0; i < n; i++)
            for (int j = 0; j < in[i].k; j++) {
                bool seen = false;
                for (int q = 0; q < nu; q++) seen = seen || u[q] == in[i].leaf[j];
                if (!seen) u[nu++] = in[i].leaf[j];
            }
        if (nu > 3) return false;
        std::sort(u, u + nu);
        res.k = (uint8_t)nu;
        for (int q = 0; q < nu; q++) res.leaf[q] = u[q];
        for (int r = 0; r < (1 << nu); r++) {
            int v[3] = {0, 0, 0};
            for (int i = 0; i < n; i++) {
                int idx = 0;
                for (int j = 0; j < in[i].k; j++)
                    for (int q = 0; q < nu; q++)
                        if (u[q] == in[i].leaf[j] && ((r >> q) & 1)) idx |= 1 << j;
                v[i] = in[i].tt[idx];
            }
            int o;
            switch (op) {
                case T_ADD: o = v[0] + v[1]; break;
                case T_SUB: o = v[0] - v[1]; break;
                case T_MUL: o = v[0] * v[1]; break;
                case T_SEL: o = v[0] != 0 ? v[1] : v[2]; break;
                case T_EQ: o = v[0] == v[1]; break;
                case T_NEQ: o = v[0] != v[1]; break;
                case T_EQZ: o = v[0] == 0; break;
                case T_LAND: o = v[0] != 0 && v[1] != 0; break;
                case T_LOR: o = v[0] != 0 || v[1] != 0; break;
                default: return false;
            }
            if (o > CONE_MAX || o < -CONE_MAX) return false;
            res.tt[r] = (int8_t)o;
        }
        return true;
    }
    // CSE'd creation of an op without peepholes
    uint32_t make_op(uint8_t op, uint32_t a, uint32_t b, uint32_t c, uint32_t aux) {
        uint64_t key = ((uint64_t)op << 56) ^ ((uint64_t)a * 0x9E3779B97F4A7C15ull) ^ ((uint64_t)b * 0xC2B2AE3D27D4EB4Full) ^
                       ((uint64_t)c * 0x165667B19E3779F9ull) ^ aux;
        for (uint32_t id : cse[key]) {
            const SOp &o = ops[id];
            if (o.op == op && o.a == a && o.b == b && o.c == c && o.aux == aux) {
                stats.cse_hits++;
                return alias[id] != NO_REF ? alias[id] : id;
            }
        }
        SOp o{op, a, b, c, aux};
        uint32_t id = push_op(o);
        cse[key].push_back(id);
        return id;
    }
    // after `id` was created: tabulate it; if it is a boolean function of its leaves, return what replaces it
    uint32_t collapse(uint32_t id) {
        const SOp o = ops[id];
        if (o.op == T_LUT) return id;
        uint32_t refs[3] = {o.a, o.b, o.c};
        int n = (o.op == T_SEL) ? 3 : (o.b == NO_REF ? 1 : 2);
        Cone c;
        if (!cone_eval(o.op, refs, n, c)) return id;
        // drop the leaves the function does not depend on
        for (int q = c.k - 1; q >= 0; q--) {
            bool dep = false;
            for (int r = 0; r < (1 << c.k); r++) dep = dep || c.tt[r] != c.tt[r ^ (1 << q)];
            if (dep) continue;
            Cone d;
            d.k = (uint8_t)(c.k - 1);
            for (int j = 0, w = 0; j < c.k; j++)
                if (j != q) d.leaf[w++] = c.leaf[j];
            for (int r = 0; r < (1 << d.k); r++) {
                int lo = r & ((1 << q) - 1), hi = r >> q;
                d.tt[r] = c.tt[lo | (hi << (q + 1))];
            }
            c = d;
        }
        cones[id] = c;
        bool boolean = true;
        for (int r = 0; r < (1 << c.k); r++) boolean = boolean && (c.tt[r] == 0 || c.tt[r] == 1);
        uint32_t rep = id;
        if (c.k == 0) rep = CONST_FLAG | intern(hostfr::from_i64(c.tt[0]));
        else if (!boolean) return id;
        else if (c.k == 1 && c.tt[0] == 0 && c.tt[1] == 1) rep = c.leaf[0];
        else if (isbool[id] && (o.op == T_EQ || o.op == T_NEQ || o.op == T_EQZ || o.op == T_LAND || o.op == T_LOR) && c.k >= 2 &&
                 !(o.a & CONST_FLAG) && (o.b == NO_REF || !(o.b & CONST_FLAG)) && isbool[o.a] && (o.b == NO_REF || isbool[o.b]))
            return id;   // already one cheap instruction on 0/1 operands
        else if (!use_luts) return id;
        else {
            uint32_t table = 0;
            for (int r = 0; r < (1 << c.k); r++) table |= (uint32_t)c.tt[r] << r;
            rep = make_op(T_LUT, c.leaf[0], c.k > 1 ? c.leaf[1] : NO_REF, c.k > 2 ? c.leaf[2] : NO_REF, table | ((uint32_t)c.k << 8));
            if (cones[rep].k == 0xff) {
                cones[rep] = c;
                stats.luts++;
            }
        }
        if (rep != id) alias[id] = rep;
        return rep;
    }
    // value provably in {0, 1}
    bool is_bool(uint32_t r) const {
        if (r == NO_REF) return false;
        if (r & CONST_FLAG) return (r & ~CONST_FLAG) == c_zero || (r & ~CONST_FLAG) == c_one;
        return isbool[r] != 0;
    }
    // the canonical value of a constant if it is a small non-negative integer, else -1
    int64_t small_const(uint32_t r) const {
        if (!is_const(r)) return -1;
        const fr::Fr &c = cval(r);
        for (int i = 1; i < 8; i++)
            if (c.v[i]) return -1;
        return (int64_t)c.v[0];
    }

    uint32_t emit(uint8_t op, uint32_t a, uint32_t b = NO_REF, uint32_t c = NO_REF, uint32_t aux = 0) {
        bool unary = (b == NO_REF);
        if (op != T_SEL && op != T_FAIL_IF && op != T_FAIL_NE && op != T_BITC && is_const(a) && (unary || is_const(b))) {
            fr::Fr out;
            fr::Fr bb = unary ? fr::zero() : cval(b);
            if (hostfr::apply(host_op(op), cval(a), bb, out)) {
                stats.folded++;
                return CONST_FLAG | intern(out);
            }
            if (op == T_DIV || op == T_INV) return CONST_FLAG | c_zero;   // mpz_invert(0): undefined in the reference; we define 0
            throw TraceError("integer division or modulo by a constant zero");
        }
        uint32_t zero = CONST_FLAG | c_zero, one = CONST_FLAG | c_one;
        switch (op) {
            case T_ADD:
                if (a == zero) return b;
                if (b == zero) return a;
                break;
            case T_SUB:
                if (b == zero) return a;
                if (a == b) return zero;
                break;
            case T_DIV:
                // a / b = a * b^-1 (generic/fr.cpp:2895-2912: Fr_inv then Fr_mul); the inversion is its own value so that
                // equal divisors share it and tape::batch_inversions can apply Montgomery's trick to independent ones
                if (a == zero) return zero;
                return emit(T_MUL, a, emit(T_INV, b));
            case T_MUL: {
                if (a == zero || b == zero) return zero;
                if (a == one) return b;
                if (b == one) return a;
                // value-range typing: a factor that is provably 0/1 turns the product into a select, and
                // x*(x-1) of such an x is 0 (the `out*(out-1) === 0` of every bit decomposition)
                for (int k = 0; k < 2; k++) {
                    uint32_t x = k ? b : a, y = k ? a : b;
                    if (!is_bool(x) || is_const(x)) continue;
                    if (!is_const(y) && ops[y].op == T_SUB && ops[y].a == x && ops[y].b == one) return zero;
                    return emit(T_SEL, x, y, zero);
                }
                break;
            }
            case T_BAND: {
                // (y >> k) & 1 with constant k: read bit k of the canonical y.  y * R^-1 has the canonical y as its
                // Montgomery limbs, is shared by all bits of y through value numbering, and costs one multiplication
                // instead of the four representation changes of SHR + BAND.
                for (int k = 0; k < 2; k++) {
                    uint32_t x = k ? b : a, m = k ? a : b;
                    if (m != one || is_const(x) || ops[x].op != T_SHR) continue;
                    int64_t sh = small_const(ops[x].b);
                    if (sh < 0 || sh >= 254 || is_const(ops[x].a)) continue;
                    uint32_t canon = emit(T_MUL, ops[x].a, CONST_FLAG | c_rinv);
                    return emit(T_BITC, canon, NO_REF, NO_REF, (uint32_t)sh);
                }
                break;
            }
            case T_EQZ:
                if (!is_const(a)) {
                    if (ops[a].op == T_EQ) return emit(T_NEQ, ops[a].a, ops[a].b);
                    if (ops[a].op == T_NEQ) return emit(T_EQ, ops[a].a, ops[a].b);
                }
                break;
            case T_EQ:
                if (a == b) return one;
                break;
            case T_NEQ:
                if (a == b) return zero;
                break;
            case T_SEL:
                if (b == c) return b;
                if (is_const(a)) return fr::is_zero(cval(a)) ? c : b;
                break;
            case T_FAIL_IF:
                if (is_const(a) && fr::is_zero(cval(a))) return NO_REF;
                if (!is_const(a) && ops[a].op == T_NEQ) return emit(T_FAIL_NE, ops[a].a, ops[a].b, NO_REF, aux);
                break;
            default: break;
        }
        if (commutative(op) && a > b) std::swap(a, b);
        uint64_t key = ((uint64_t)op << 56) ^ ((uint64_t)a * 0x9E3779B97F4A7C15ull) ^ ((uint64_t)b * 0xC2B2AE3D27D4EB4Full) ^
                       ((uint64_t)c * 0x165667B19E3779F9ull) ^ aux;
        auto &bucket = cse[key];
        for (uint32_t id : bucket) {
            const SOp &o = ops[id];
            if (o.op == op && o.a == a && o.b == b && o.c == c && o.aux == aux) {
                stats.cse_hits++;
                return alias[id] != NO_REF ? alias[id] : id;
            }
        }
        SOp o{op, a, b, c, aux};
        uint32_t id = push_op(o);
        cse[key].push_back(id);
        if (op == T_ADD || op == T_SUB || op == T_MUL || op == T_SEL || op == T_EQ || op == T_NEQ || op == T_EQZ || op == T_LAND ||
            op == T_LOR)
            return collapse(id);
        return id;
    }

    // ---------------------------------------------------------------- logged state writes
    void log_write(uint8_t space, Frame *f, int64_t key, const AV &old, bool existed) {
        if (!logs.empty()) logs.back().push_back(LogEntry{space, f, key, old, existed});
    }
    void set_reg(Frame &f, int r, const AV &v) {
        log_write(0, &f, r, f.regs[(size_t)r], true);
        f.regs[(size_t)r] = v;
    }
    void set_lvar(Frame &f, int64_t addr, const AV &v) {
        auto it = f.lvar.find(addr);
        if (it == f.lvar.end()) {
            log_write(1, &f, addr, AV(), false);
            f.lvar.emplace(addr, v);
        } else {
            log_write(1, &f, addr, it->second, true);
            it->second = v;
        }
    }
    void set_sig(int64_t idx, const AV &v) {
        if (idx < 0 || idx >= (int64_t)sig.size()) throw TraceError("signal index out of range");
        log_write(2, nullptr, idx, sig[(size_t)idx], true);
        sig[(size_t)idx] = v;
    }
    AV read_loc(const LogEntry &e) {
        if (e.space == 0) return e.frame->regs[(size_t)e.key];
        if (e.space == 1) {
            auto it = e.frame->lvar.find(e.key);
            return it == e.frame->lvar.end() ? AV() : it->second;
        }
        return sig[(size_t)e.key];
    }
    void restore(const LogEntry &e) {
        if (e.space == 0) e.frame->regs[(size_t)e.key] = e.old;
        else if (e.space == 1) {
            if (e.existed) e.frame->lvar[e.key] = e.old;
            else e.frame->lvar.erase(e.key);
        } else sig[(size_t)e.key] = e.old;
    }
    void write_loc(const LogEntry &e, const AV &v) {
        if (e.space == 0) set_reg(*e.frame, (int)e.key, v);
        else if (e.space == 1) set_lvar(*e.frame, e.key, v);
        else set_sig(e.key, v);
    }

    // ---------------------------------------------------------------- operands
    AV val(const Frame &f, const cvm::Operand &o) {
        switch (o.kind) {
            case cvm::K_REG: return f.regs[(size_t)o.val];
            case cvm::K_I64: return i64_av(o.val);
            case cvm::K_FF: {
                uint32_t &m = ffmap[(size_t)o.val];
                if (m == NO_REF) m = intern(prog.ffconst[(size_t)o.val]);
                return ff_av(m);
            }
            case cvm::K_SPR: return i64_av(SPR_BASE);
            default: throw TraceError("call-argument operand used as a value");
        }
    }
    AV simple(const Frame &f, uint8_t kind, int64_t v) {
        cvm::Operand o;
        o.kind = kind;
        o.val = v;
        return val(f, o);
    }
    bool truth_known(const AV &v, bool &t) {
        if (v.kind == AV_I64) { t = v.i != 0; return true; }
        if (v.kind == AV_FF) { t = !fr::is_zero(consts[v.ref]); return true; }
        if (v.kind == AV_UNDEF) { t = false; return true; }
        return false;
    }

    // ---------------------------------------------------------------- components
    int new_comp(int code, int64_t start) {
        Comp c;
        c.code = code;
        c.start = start;
        c.counter = prog.codes[(size_t)code].n_inputs;
        c.subs.assign((size_t)prog.codes[(size_t)code].n_subcmps, -1);
        comps.push_back(c);
        return (int)comps.size() - 1;
    }
    void run_component(int ci) {
        const cvm::Code &code = prog.codes[(size_t)comps[(size_t)ci].code];
        Frame f;
        f.regs.assign((size_t)code.nregs, AV());
        if (++depth > 4000) throw TraceError("component/function nesting too deep");
        run(code, f, ci, 0, (int)code.ins.size(), 0, nullptr);
        depth--;
    }
    Comp &sub_of(int ci, const AV &slotv) {
        int64_t slot = int_of(slotv, "sub-component index");
        Comp &c = comps[(size_t)ci];
        if (slot < 0 || slot >= (int64_t)c.subs.size() || c.subs[(size_t)slot] < 0)
            throw TraceError("access to a sub-component that was never created (slot " + std::to_string(slot) +
                             " of " + prog.codes[(size_t)c.code].header + ")");
        return comps[(size_t)c.subs[(size_t)slot]];
    }

    uint32_t pred_conj() {
        uint32_t p = preds[0];
        for (size_t k = 1; k < preds.size(); k++) p = emit(T_LAND, p, preds[k]);
        return p;
    }

    static uint8_t tape_op(uint16_t op) {
        switch (op) {
            case cvm::OP_FF_ADD: return T_ADD; case cvm::OP_FF_SUB: return T_SUB; case cvm::OP_FF_MUL: return T_MUL;
            case cvm::OP_FF_DIV: return T_DIV; case cvm::OP_FF_IDIV: return T_IDIV; case cvm::OP_FF_REM: return T_MOD;
            case cvm::OP_FF_POW: return T_POW; case cvm::OP_FF_SHL: return T_SHL; case cvm::OP_FF_SHR: return T_SHR;
            case cvm::OP_FF_BAND: return T_BAND; case cvm::OP_FF_BOR: return T_BOR; case cvm::OP_FF_BXOR: return T_BXOR;
            case cvm::OP_FF_BNOT: return T_BNOT; case cvm::OP_FF_LT: return T_LT; case cvm::OP_FF_LE: return T_LE;
            case cvm::OP_FF_GT: return T_GT; case cvm::OP_FF_GE: return T_GE; case cvm::OP_FF_EQ: return T_EQ;
            case cvm::OP_FF_NEQ: return T_NEQ; case cvm::OP_FF_AND: return T_LAND; case cvm::OP_FF_OR: return T_LOR;
            case cvm::OP_FF_EQZ: return T_EQZ; default: return T_NOP;
        }
    }

    // ---------------------------------------------------------------- the interpreter
    // executes ins[pc, stop) of `code`; returns true when a `return` was executed
    bool run(const cvm::Code &code, Frame &f, int ci, int pc, int stop, int range_begin, const RetTarget *ret) {
        using namespace cvm;
        uint32_t applied = ret ? ret->returned : NO_REF;   // the "already returned" predicate this range runs under
        // after a nested construct: did a data-dependent `return` happen inside it?  Then the rest of this range (later
        // iterations of an enclosing static loop included) is one more predicated region, under "not returned".
        auto rest_if_returned = [&](int at) -> bool {
            if (!ret || ret->returned == NO_REF || ret->returned == applied) return false;
            applied = ret->returned;
            region(code, f, ci, emit(T_EQZ, ret->returned), at, stop, range_begin, -1, -1, -1, ret);
            return true;
        };
        while (pc < stop) {
            const Ins &in = code.ins[(size_t)pc];
            stats.cvm_instructions++;
            switch (in.op) {
                case OP_MOV: set_reg(f, in.dst, val(f, in.args[0])); pc++; break;
                case OP_FF_ADD: case OP_FF_SUB: case OP_FF_MUL: case OP_FF_DIV: case OP_FF_IDIV: case OP_FF_REM:
                case OP_FF_POW: case OP_FF_SHL: case OP_FF_SHR: case OP_FF_BAND: case OP_FF_BOR: case OP_FF_BXOR:
                case OP_FF_LT: case OP_FF_LE: case OP_FF_GT: case OP_FF_GE: case OP_FF_EQ: case OP_FF_NEQ:
                case OP_FF_AND: case OP_FF_OR: {
                    stats.ref_field_ops++;
                    if (in.op == OP_FF_MUL) stats.ref_mul++;
                    if (in.op == OP_FF_DIV) { stats.ref_mul++; stats.ref_div++; }
                    uint32_t a = field_ref(val(f, in.args.at(0))), b = field_ref(val(f, in.args.at(1)));
                    if (in.op == OP_FF_IDIV || in.op == OP_FF_REM) {
                        // GMP aborts on a zero divisor (generic/fr.cpp:2835-2875) whether or not the quotient is used, but
                        // only when the statement executes: an explicit check, guarded by the enclosing data-dependent
                        // conditions (both arms of an if-converted branch are traced), raises the per-witness flag
                        if (!is_const(b) || fr::is_zero(cval(b))) {
                            uint32_t z = emit(T_EQZ, b);
                            if (!preds.empty()) z = emit(T_LAND, pred_conj(), z);
                            emit(T_FAIL_IF, z, NO_REF, NO_REF, ST_DIVZERO);
                        }
                        if (is_const(b) && fr::is_zero(cval(b))) {
                            set_reg(f, in.dst, ff_av(c_zero));
                            pc++;
                            break;
                        }
                    }
                    set_reg(f, in.dst, ref_av(emit(tape_op(in.op), a, b)));
                    pc++;
                    break;
                }
                case OP_FF_BNOT: case OP_FF_EQZ: {
                    stats.ref_field_ops++;
                    uint32_t a = field_ref(val(f, in.args.at(0)));
                    set_reg(f, in.dst, ref_av(emit(tape_op(in.op), a)));
                    pc++;
                    break;
                }
                case OP_FF_WRAP_I64: {
                    AV v = val(f, in.args.at(0));
                    if (v.kind == AV_DYN) {
                        // Fr_toInt of a witness value (generic/fr.cpp:1102-1170): defined for -2^31 <= x < 2^31 (around 0 mod q),
                        // the reference asserts otherwise -> ST_TOINT.  The result is only usable as (part of) an address.
                        const uint32_t lo = emit(T_GE, v.ref, CONST_FLAG | intern(hostfr::from_i64(-((int64_t)1 << 31))));
                        const uint32_t hi = emit(T_LT, v.ref, CONST_FLAG | intern(hostfr::from_i64((int64_t)1 << 31)));
                        uint32_t bad = emit(T_EQZ, emit(T_LAND, lo, hi));
                        if (!preds.empty()) bad = emit(T_LAND, pred_conj(), bad);
                        emit(T_FAIL_IF, bad, NO_REF, NO_REF, ST_TOINT);
                        stats.dyn_addresses++;
                        AV a;
                        a.kind = AV_DADDR;
                        a.ref = v.ref;
                        a.i = 0;
                        a.scale = 1;
                        set_reg(f, in.dst, a);
                        pc++;
                        break;
                    }
                    set_reg(f, in.dst, i64_av(int_of(v, "array index (ff.wrap_i64)")));
                    pc++;
                    break;
                }
                case OP_I64_ADD: case OP_I64_SUB: case OP_I64_MUL: case OP_I64_LT: case OP_I64_LE: case OP_I64_GT:
                case OP_I64_GE: case OP_I64_EQ: case OP_I64_NEQ: {
                    {
                        // address arithmetic on a data-dependent index: affine in it (AddAddress / MulAddress with constants)
                        const AV x = val(f, in.args.at(0)), y = val(f, in.args.at(1));
                        if (x.kind == AV_DADDR || y.kind == AV_DADDR) {
                            AV r = x.kind == AV_DADDR ? x : y;
                            const AV &o = x.kind == AV_DADDR ? y : x;
                            if (o.kind == AV_DADDR) throw TraceError("arithmetic on two data-dependent addresses is not supported");
                            const int64_t c = int_of(o, "i64 operand");
                            if (in.op == OP_I64_ADD) r.i += c;
                            else if (in.op == OP_I64_SUB && x.kind == AV_DADDR) r.i -= c;
                            else if (in.op == OP_I64_SUB) { r.i = c - r.i; r.scale = -r.scale; }
                            else if (in.op == OP_I64_MUL) { r.i *= c; r.scale *= c; }
                            else throw TraceError("comparison of a data-dependent address is not supported");
                            if (r.scale == 0) r = i64_av(r.i);
                            set_reg(f, in.dst, r);
                            pc++;
                            break;
                        }
                    }
                    int64_t a = int_of(val(f, in.args.at(0)), "i64 operand"), b = int_of(val(f, in.args.at(1)), "i64 operand");
                    int64_t r = 0;
                    switch (in.op) {
                        case OP_I64_ADD: r = a + b; break; case OP_I64_SUB: r = a - b; break;
                        case OP_I64_MUL: r = a * b; break; case OP_I64_LT: r = a < b; break;
                        case OP_I64_LE: r = a <= b; break; case OP_I64_GT: r = a > b; break;
                        case OP_I64_GE: r = a >= b; break; case OP_I64_EQ: r = a == b; break;
                        default: r = a != b; break;
                    }
                    set_reg(f, in.dst, i64_av(r));
                    pc++;
                    break;
                }
                case OP_FF_LOAD: {
                    if (val(f, in.args.at(0)).kind == AV_DADDR) {
                        // var[index]: the value of whichever written variable the index names (an unwritten one reads 0)
                        const AV a = val(f, in.args.at(0));
                        int64_t limit = code.is_function ? code.local_memory : 0;
                        for (const auto &kv : f.lvar) limit = std::max(limit, kv.first + 1);
                        uint32_t r = CONST_FLAG | c_zero;
                        for (const auto &ka : mux_candidates(a, limit, "variable address")) {
                            auto it = f.lvar.find(ka.second);
                            if (it == f.lvar.end()) continue;
                            r = emit(T_SEL, index_is(a, ka.first), field_ref(it->second), r);
                        }
                        set_reg(f, in.dst, ref_av(r));
                        pc++;
                        break;
                    }
                    int64_t addr = int_of(val(f, in.args.at(0)), "variable address");
                    auto it = f.lvar.find(addr);
                    set_reg(f, in.dst, it == f.lvar.end() ? ff_av(c_zero) : it->second);
                    pc++;
                    break;
                }
                case OP_FF_STORE: {
                    if (val(f, in.args.at(0)).kind == AV_DADDR) {
                        // var[index] = v: every variable the index can name keeps its value unless it is the one
                        const AV a = val(f, in.args.at(0));
                        int64_t limit = code.is_function ? code.local_memory : 0;
                        for (const auto &kv : f.lvar) limit = std::max(limit, kv.first + 1);
                        const uint32_t v = field_ref(val(f, in.args.at(1)));
                        for (const auto &ka : mux_candidates(a, limit, "variable address")) {
                            auto it = f.lvar.find(ka.second);
                            const uint32_t old = it == f.lvar.end() ? (CONST_FLAG | c_zero) : field_ref(it->second);
                            set_lvar(f, ka.second, ref_av(emit(T_SEL, index_is(a, ka.first), v, old)));
                        }
                        pc++;
                        break;
                    }
                    int64_t addr = int_of(val(f, in.args.at(0)), "variable address");
                    set_lvar(f, addr, ref_av(field_ref(val(f, in.args.at(1)))));
                    pc++;
                    break;
                }
                case OP_GET_SIGNAL: {
                    if (val(f, in.args.at(0)).kind == AV_DADDR) {
                        const AV a = val(f, in.args.at(0));
                        const int64_t start = comps[(size_t)ci].start, limit = prog.codes[(size_t)comps[(size_t)ci].code].n_signals;
                        uint32_t r = CONST_FLAG | c_zero;
                        for (const auto &ka : mux_candidates(a, limit, "signal index"))
                            r = emit(T_SEL, index_is(a, ka.first), field_ref(sig[(size_t)(start + ka.second)]), r);
                        set_reg(f, in.dst, ref_av(r));
                        pc++;
                        break;
                    }
                    int64_t idx = comps[(size_t)ci].start + int_of(val(f, in.args.at(0)), "signal index");
                    if (idx < 0 || idx >= (int64_t)sig.size()) throw TraceError("signal index out of range");
                    set_reg(f, in.dst, ref_av(field_ref(sig[(size_t)idx])));
                    pc++;
                    break;
                }
                case OP_SET_SIGNAL: {
                    int64_t idx = comps[(size_t)ci].start + int_of(val(f, in.args.at(0)), "signal index");
                    set_sig(idx, ref_av(field_ref(val(f, in.args.at(1)))));
                    pc++;
                    break;
                }
                case OP_GET_CMP_SIGNAL: {
                    if (val(f, in.args.at(1)).kind == AV_DADDR) {
                        Comp &sc = sub_of(ci, val(f, in.args.at(0)));
                        const AV a = val(f, in.args.at(1));
                        const int64_t limit = prog.codes[(size_t)sc.code].n_signals;
                        uint32_t r = CONST_FLAG | c_zero;
                        for (const auto &ka : mux_candidates(a, limit, "signal index"))
                            r = emit(T_SEL, index_is(a, ka.first), field_ref(sig[(size_t)(sc.start + ka.second)]), r);
                        set_reg(f, in.dst, ref_av(r));
                        pc++;
                        break;
                    }
                    Comp &s = sub_of(ci, val(f, in.args.at(0)));
                    int64_t idx = s.start + int_of(val(f, in.args.at(1)), "signal index");
                    if (idx < 0 || idx >= (int64_t)sig.size()) throw TraceError("signal index out of range");
                    set_reg(f, in.dst, ref_av(field_ref(sig[(size_t)idx])));
                    pc++;
                    break;
                }
                case OP_SET_CMP_INPUT: case OP_SET_CMP_INPUT_CNT: case OP_SET_CMP_INPUT_RUN:
                case OP_SET_CMP_INPUT_CNT_CHECK: {
                    if (!preds.empty())
                        throw TraceError("sub-component input assigned under a data-dependent condition");
                    AV slotv = val(f, in.args.at(0));
                    Comp &s = sub_of(ci, slotv);
                    int sidx = comps[(size_t)ci].subs[(size_t)int_of(slotv, "sub-component index")];
                    set_sig(s.start + int_of(val(f, in.args.at(1)), "signal index"),
                            ref_av(field_ref(val(f, in.args.at(2)))));
                    // inputCounter / run rules: store_bucket.rs:662-800
                    if (in.op == OP_SET_CMP_INPUT_CNT) s.counter -= 1;
                    else if (in.op == OP_SET_CMP_INPUT_RUN) run_component(sidx);
                    else if (in.op == OP_SET_CMP_INPUT_CNT_CHECK) {
                        s.counter -= 1;
                        if (s.counter == 0) run_component(sidx);
                    }
                    pc++;
                    break;
                }
                case OP_CREATE_CMP: {
                    if (!preds.empty()) throw TraceError("component created under a data-dependent condition");
                    // create_component_bucket.rs:206-354: slot k -> signalStart = mine + off + k*jump
                    const cvm::Code &tc = prog.codes[(size_t)in.target];
                    // an array with undefined positions creates only the defined ones; signal and component offsets advance
                    // per CREATED component, the slot is the array position (create_component_bucket.rs:258-275, 339-349)
                    const int64_t n_create = in.positions.empty() ? in.cc[5] : (int64_t)in.positions.size();
                    for (int64_t k = 0; k < n_create; k++) {
                        int sidx = new_comp(in.target, comps[(size_t)ci].start + in.cc[1] + k * in.cc[2]);
                        int64_t slot = in.cc[0] + (in.positions.empty() ? k : in.positions[(size_t)k]);
                        if (slot < 0 || slot >= (int64_t)comps[(size_t)ci].subs.size())
                            throw TraceError("create_cmp slot out of range in " + code.header);
                        comps[(size_t)ci].subs[(size_t)slot] = sidx;
                        if (tc.n_inputs == 0) run_component(sidx);   // template.rs:326-331
                    }
                    pc++;
                    break;
                }
                case OP_IF: {
                    AV c = val(f, in.args.at(0));
                    bool t;
                    int then_end = in.m_else >= 0 ? in.m_else : in.m_end;
                    if (truth_known(c, t)) {
                        pc = t ? pc + 1 : (in.m_else >= 0 ? in.m_else + 1 : in.m_end + 1);
                        break;
                    }
                    // assert shape: if c / error n / end  -> FAIL_IF without opening a region
                    if (in.m_else < 0 && in.m_end == pc + 2 && code.ins[(size_t)pc + 1].op == OP_ERROR) {
                        uint32_t p = c.ref;
                        if (!preds.empty()) p = emit(T_LAND, pred_conj(), p);
                        emit(T_FAIL_IF, p, NO_REF, NO_REF, ST_ASSERT);
                        pc = in.m_end + 1;
                        break;
                    }
                    stats.dyn_branches++;
                    region(code, f, ci, c.ref, pc + 1, then_end, pc + 1, in.m_else >= 0 ? in.m_else + 1 : -1, in.m_end,
                           in.m_else + 1, ret);
                    pc = in.m_end + 1;
                    if (rest_if_returned(pc)) pc = stop;
                    break;
                }
                case OP_ELSE: pc = in.m_end + 1; break;    // fell off the end of a then-branch
                case OP_END: pc++; break;                  // end of if, or falling out of a loop
                case OP_LOOP: {
                    // run the loop as its own region so that data-dependent `continue`s know which loop they belong to
                    loop_stack.push_back(pc);
                    const int lend = in.m_end;
                    bool r = run_loop(code, f, ci, pc, lend, ret);
                    loop_stack.pop_back();
                    if (r) return true;
                    if (last_exit_pc == lend) stats.cvm_instructions++;   // fell onto the loop's `end` (a `break` jumps past it)
                    if (rest_if_returned(lend + 1)) { pc = stop; break; }
                    pc = lend + 1;
                    break;
                }
                case OP_CONTINUE:
                    if (in.m_loop < range_begin) {
                        // `while (c) body` with a data-dependent c is `loop; c; if c; body; continue; end; end`
                        // (loop_bucket.rs:97-120): this `continue` sits in the if-converted then-arm.  The next
                        // iteration is traced right here, nested in the arm (its own condition becomes one more
                        // predicate), up to max_unroll iterations; a witness that would need more raises ST_LOOP.
                        // Only the plain while shape is handled: the loop must be the innermost open one.
                        if (preds.empty() || loop_stack.empty() || loop_stack.back() != in.m_loop)
                            throw TraceError("continue under a data-dependent condition");
                        const int loop_end = code.ins[(size_t)in.m_loop].m_end;
                        if (++unroll_depth > max_unroll) {
                            emit(T_FAIL_IF, pred_conj(), NO_REF, NO_REF, ST_LOOP);
                        } else {
                            stats.unrolled_iterations++;
                            run(code, f, ci, in.m_loop + 1, loop_end, in.m_loop + 1, ret);
                        }
                        unroll_depth--;
                        pc = stop;   // nothing after a `continue` executes in this arm
                        break;
                    }
                    pc = in.m_loop + 1;
                    break;
                case OP_BREAK: {
                    if (in.m_loop < range_begin) throw TraceError("break under a data-dependent condition");
                    pc = code.ins[(size_t)in.m_loop].m_end + 1;
                    break;
                }
                case OP_ERROR: {
                    uint32_t p = preds.empty() ? (CONST_FLAG | c_one) : pred_conj();
                    // an unconditional error is a program that fails for every input; keep it as a flag
                    SOp o{T_FAIL_IF, p, NO_REF, NO_REF, ST_ASSERT};
                    push_op(o);
                    pc++;
                    break;
                }
                case OP_CALL: call(code, f, ci, in); pc++; break;
                case OP_RETURN: {
                    if (!ret) throw TraceError("return outside a function");
                    const bool predicated = preds.size() != ret->pred_depth;
                    if (in.scalar_return) {
                        set_lvar(*ret->frame, ret->addr, ref_av(field_ref(val(f, in.args.at(0)))));
                    } else {
                        // (the emitter passes the loaded first element; the parser kept that load's address: link_returns)
                        int64_t src = int_of(val(f, in.args.at(in.ret_from_load ? 2 : 0)), "return address");
                        int64_t n = std::min(int_of(val(f, in.args.at(1)), "return size"), ret->size);
                        for (int64_t k = 0; k < n; k++) {
                            auto it = f.lvar.find(src + k);
                            set_lvar(*ret->frame, ret->addr + k, it == f.lvar.end() ? ff_av(c_zero) : it->second);
                        }
                    }
                    if (!predicated) return true;
                    // `return` in an if-converted arm: the writes above are merged with the previous contents of the
                    // destination when the arm closes (they are logged like any other write); remember under which
                    // condition the function is over, so that whatever follows runs under its negation
                    {
                        uint32_t p = preds[ret->pred_depth];
                        for (size_t k = ret->pred_depth + 1; k < preds.size(); k++) p = emit(T_LAND, p, preds[k]);
                        if (!is_bool(p)) p = emit(T_NEQ, p, CONST_FLAG | c_zero);
                        ret->returned = ret->returned == NO_REF ? p : emit(T_LOR, ret->returned, p);
                    }
                    pc = stop;
                    break;
                }
                case OP_ARRAY_EQ: {
                    // conjunction of the element-wise equalities (what the C++ twin computes, compute_bucket.rs:375-407)
                    stats.ref_field_ops++;
                    int64_t base[2];
                    for (int k = 0; k < 2; k++) {
                        const uint16_t lop = (uint16_t)in.cc[k];
                        base[k] = int_of(val(f, in.args.at(2 * k)), "array address");
                        if (lop == OP_GET_SIGNAL) base[k] += comps[(size_t)ci].start;
                        else if (lop == OP_GET_CMP_SIGNAL) base[k] += sub_of(ci, val(f, in.args.at(2 * k + 1))).start;
                    }
                    uint32_t acc = NO_REF;
                    for (int64_t j = 0; j < in.cc[2]; j++) {
                        uint32_t v[2];
                        for (int k = 0; k < 2; k++) {
                            if ((uint16_t)in.cc[k] == OP_FF_LOAD) {
                                auto it = f.lvar.find(base[k] + j);
                                v[k] = it == f.lvar.end() ? (CONST_FLAG | c_zero) : field_ref(it->second);
                            } else {
                                if (base[k] + j < 0 || base[k] + j >= (int64_t)sig.size()) throw TraceError("signal index out of range");
                                v[k] = field_ref(sig[(size_t)(base[k] + j)]);
                            }
                        }
                        const uint32_t e = emit(T_EQ, v[0], v[1]);
                        acc = acc == NO_REF ? e : emit(T_LAND, acc, e);
                    }
                    set_reg(f, in.dst, ref_av(acc == NO_REF ? (CONST_FLAG | c_one) : acc));
                    pc++;
                    break;
                }
                case OP_GET_TEMPLATE_ID: {
                    // ctx->componentMemory[mySubcomponents[c]].templateId (load_bucket.rs:262-266)
                    const int64_t id = prog.codes[(size_t)sub_of(ci, val(f, in.args.at(0))).code].template_id;
                    if (id < 0) throw TraceError("template id of a sub-component is unknown (header without _<id> suffix)");
                    set_reg(f, in.dst, i64_av(id));
                    pc++;
                    break;
                }
                case OP_GET_TMPL_SIG_POS: case OP_GET_TMPL_SIG_DIM: case OP_GET_TMPL_SIG_SIZE: {
                    // templateInsId2IOSignalInfo[id].defs[code].{offset, lengths[i-1], size}: constants of the component tree
                    const int64_t id = int_of(val(f, in.args.at(0)), "template id");
                    const int64_t sc = int_of(val(f, in.args.at(1)), "signal code");
                    auto it = prog.io_map.find(id);
                    if (it == prog.io_map.end())
                        throw TraceError("mapped access (mixed component array): no io-map entry for template instance " +
                                         std::to_string(id) + " -- the .cvm file does not carry the io-map (circuit.rs:577-621); "
                                         "load the program with its generated .cpp and .dat (cvmgpu_program_load_files)");
                    if (sc < 0 || sc >= (int64_t)it->second.size()) throw TraceError("signal code outside the io-map entry");
                    const cvm::IoDef &d = it->second[(size_t)sc];
                    int64_t r;
                    if (in.op == OP_GET_TMPL_SIG_POS) r = d.offset;
                    else if (in.op == OP_GET_TMPL_SIG_SIZE) r = d.size;
                    else {
                        const int64_t i = int_of(val(f, in.args.at(2)), "dimension index");
                        if (i < 1 || i > (int64_t)d.lengths_tail.size()) throw TraceError("dimension index outside the io-map entry");
                        r = d.lengths_tail[(size_t)(i - 1)];
                    }
                    set_reg(f, in.dst, i64_av(r));
                    pc++;
                    break;
                }
                case OP_MAPPED_UNSUPPORTED:
                    throw TraceError("bus accesses through the io-map (get_template_signal_type / get_bus_signal_*, "
                                     "location_rule.rs:143-152) are not implemented");
                default: throw TraceError("unhandled CVM instruction");
            }
        }
        last_exit_pc = pc;
        return false;
    }

    // body of `loop` at lpc .. its `end` at lend: iterate while `continue` jumps back (static trip counts), leave on
    // `break` or by falling onto the loop's `end`
    bool run_loop(const cvm::Code &code, Frame &f, int ci, int lpc, int lend, const RetTarget *ret) {
        // the body is executed by the ordinary interpreter loop: `continue` (static) sets pc = lpc + 1, `break` jumps
        // past lend; both stay inside [lpc + 1, lend + 1)
        return run(code, f, ci, lpc + 1, lend, lpc, ret);
    }

    // If-conversion of one data-dependent region: the then-arm [t_begin, t_end) runs under `cond`, the optional else-arm
    // [e_begin, e_end) under its negation, both against an undo log; the locations either of them wrote are merged with SEL.
    void region(const cvm::Code &code, Frame &f, int ci, uint32_t cond, int t_begin, int t_end, int t_range, int e_begin,
                int e_end, int e_range, const RetTarget *ret) {
        // then-arm
        logs.emplace_back();
        preds.push_back(cond);
        run(code, f, ci, t_begin, t_end, t_range, ret);
        preds.pop_back();
        std::vector<LogEntry> logT = std::move(logs.back());
        logs.pop_back();
        // collect final values of the then-arm, keep first-old per location, roll back
        struct Merge { LogEntry first; AV vT; bool inT; AV vE; bool inE; };
        std::vector<Merge> merges;
        std::unordered_map<uint64_t, size_t> index;
        auto keyof = [](const LogEntry &e) {
            return ((uint64_t)e.space << 62) ^ ((uint64_t)(uintptr_t)e.frame * 0x9E3779B97F4A7C15ull) ^ (uint64_t)e.key * 0xC2B2AE3D27D4EB4Full;
        };
        auto same = [](const LogEntry &a, const LogEntry &b) { return a.space == b.space && a.frame == b.frame && a.key == b.key; };
        auto slot_for = [&](const LogEntry &e) -> size_t {
            uint64_t k = keyof(e);
            auto it = index.find(k);
            while (it != index.end() && !same(merges[it->second].first, e)) { k = k * 31 + 7; it = index.find(k); }
            if (it != index.end()) return it->second;
            merges.push_back(Merge{e, AV(), false, AV(), false});
            index.emplace(k, merges.size() - 1);
            return merges.size() - 1;
        };
        for (auto &e : logT) {
            size_t s = slot_for(e);
            if (!merges[s].inT) { merges[s].inT = true; merges[s].vT = read_loc(e); }
        }
        for (size_t k = logT.size(); k-- > 0;) restore(logT[k]);
        // else-arm
        logs.emplace_back();
        if (e_begin >= 0) {
            preds.push_back(emit(T_EQZ, cond));
            run(code, f, ci, e_begin, e_end, e_range, ret);
            preds.pop_back();
        }
        std::vector<LogEntry> logE = std::move(logs.back());
        logs.pop_back();
        for (auto &e : logE) {
            size_t s = slot_for(e);
            if (!merges[s].inE) { merges[s].inE = true; merges[s].vE = read_loc(e); }
        }
        for (size_t k = logE.size(); k-- > 0;) restore(logE[k]);
        // merge (writes are logged by the enclosing region, if any, against the pre-branch values)
        for (auto &m : merges) {
            AV before = read_loc(m.first);
            AV vT = m.inT ? m.vT : before, vE = m.inE ? m.vE : before;
            if (m.first.space == 0 && (vT.kind == AV_I64 || vE.kind == AV_I64 || vT.kind == AV_UNDEF || vE.kind == AV_UNDEF)) {
                // registers are block-local temporaries; integer temporaries that differ cannot be merged, and
                // are dead after the branch in emitted code.  Leave them undefined.
                if (vT.kind == vE.kind && vT.i == vE.i && vT.ref == vE.ref) write_loc(m.first, vT);
                else write_loc(m.first, AV());
                continue;
            }
            uint32_t rT = field_ref(vT), rE = field_ref(vE);
            write_loc(m.first, ref_av(emit(T_SEL, cond, rT, rE)));
        }
    }

    void call(const cvm::Code &code, Frame &f, int ci, const cvm::Ins &in) {
        using namespace cvm;
        const cvm::Code &fn = prog.codes[(size_t)in.target];
        if (!fn.is_function) throw TraceError("ff.call target is not a function");
        RetTarget rt;
        rt.frame = &f;
        rt.addr = int_of(val(f, in.args.at(0)), "call destination");
        rt.size = int_of(val(f, in.args.at(1)), "call destination size");
        rt.pred_depth = preds.size();
        Frame g;
        g.regs.assign((size_t)fn.nregs, AV());
        g.regs[(size_t)fn.reg_destination] = i64_av(rt.addr);
        g.regs[(size_t)fn.reg_destination_size] = i64_av(rt.size);
        int64_t pos = 0;   // arguments are packed consecutively into the callee's lvar (call_bucket.rs:478-515)
        for (size_t k = 2; k < in.args.size(); k++) {
            const Operand &o = in.args[k];
            if (o.kind == K_ARG_MEM) {
                int64_t base = int_of(simple(f, o.akind, o.aval), "argument address");
                for (int64_t j = 0; j < o.n; j++) {
                    auto it = f.lvar.find(base + j);
                    g.lvar[pos + j] = it == f.lvar.end() ? ff_av(c_zero) : it->second;
                }
                pos += o.n;
            } else if (o.kind == K_ARG_SIG || o.kind == K_ARG_SUBSIG) {
                int64_t base;
                if (o.kind == K_ARG_SIG) base = comps[(size_t)ci].start;
                else base = sub_of(ci, simple(f, o.ckind, o.cval)).start;
                base += int_of(simple(f, o.akind, o.aval), "argument address");
                for (int64_t j = 0; j < o.n; j++) {
                    if (base + j < 0 || base + j >= (int64_t)sig.size()) throw TraceError("signal index out of range");
                    g.lvar[pos + j] = ref_av(field_ref(sig[(size_t)(base + j)]));
                }
                pos += o.n;
            } else {
                g.lvar[pos++] = ref_av(field_ref(val(f, o)));
            }
        }
        if (++depth > 4000) throw TraceError("component/function nesting too deep");
        run(fn, g, ci, 0, (int)fn.ins.size(), 0, &rt);
        depth--;
        // the callee frame dies here: drop undo-log entries that point into it
        if (!logs.empty()) {
            auto &L = logs.back();
            L.erase(std::remove_if(L.begin(), L.end(), [&](const LogEntry &e) { return e.frame == &g; }), L.end());
        }
        (void)code;
    }
};

}  // namespace tape
