// C ABI of the engine (see include/cvmgpu.h for the reference interfaces each entry replaces).
// There is NO CPU fallback: every compute entry point fails with CVMGPU_ERR_CUDA when no device works.
#include "../../include/cvmgpu.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <initializer_list>
#include <map>
#include <mutex>
#include <thread>
#include <memory>
#include <string>
#include <vector>

#include <atomic>
#include <chrono>

#include "cvm_parse.hpp"
#include "fused.hpp"
#include "host_fr.hpp"
#include "kernels.cuh"
#include "r1cs.hpp"
#include "tape.hpp"
#include "tracer.hpp"

static thread_local std::string g_err;
static int g_tape_mode = 0;
static int g_r1cs_minb = getenv("CVMGPU_R1CS_MINB") ? atoi(getenv("CVMGPU_R1CS_MINB")) : 0;   // 0 = by circuit
static int g_carveout = getenv("CVMGPU_CARVEOUT") ? atoi(getenv("CVMGPU_CARVEOUT")) : 100;

static int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (expr);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return fail(CVMGPU_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));      \
    } while (0)

struct DevBuf {
    void *p = nullptr;
    size_t n = 0;
    int ensure(size_t bytes) {
        if (bytes <= n) return 0;
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) return fail(CVMGPU_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
        n = bytes;
        return 0;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
};

struct cvmgpu_program {
    tape::Tape tape;
    tape::TraceStats tstats;
    tape::BatchInvStats binv;
    std::vector<fr::Fr> consts_mont;
    std::vector<uint64_t> witness;   // %%witness: signal index of every witness wire
    std::string main_inputs;         // ;;%%main_input lines of the program: "name first_signal size\n" each
    std::vector<uint8_t> wire_bool;  // per witness wire: the value is proven 0/1 by the trace compiler's typing
    uint64_t n_signals = 0;
    uint32_t n_inputs = 0, n_outputs = 0;
    uint64_t layout_id = 0;          // identifies this program's value-store layout (r1cs bindings are cached against it)
    // device copies of the tables, one set per device the program has run on (uploaded on first use there)
    struct Dev {
        DevBuf d_tape, d_consts, d_wire_loc, d_iconsts, d_flist, d_blist;
        bool ready = false;
    };
    std::map<int, Dev> dev;
    std::mutex mu;
    // what the fused R1CS check (fused.hpp) is built from: the prepared program, its constants (canonical), and the tapes
    // already built for the constraint systems this program has been run with (by cvmgpu_r1cs::uid)
    tape::XProg xp;
    std::vector<fr::Fr> consts;
    size_t n_ssa = 0;
    uint32_t max_terms = 0;
    uint32_t explicit_slots = 0;     // the caller's slot count (0: chosen by the cost model), also used for fused tapes
    bool fusable = false;
    bool typed = true;               // false: compiled untyped (a field program with a sprinkling of 0/1 values)
    std::map<uint64_t, std::unique_ptr<cvmgpu_program>> fused;
    std::map<uint64_t, bool> fuse_worth;
    // the same circuit traced under the assumption that every main input is 0 or 1 (see build_program), or null
    std::unique_ptr<cvmgpu_program> spec;
    std::atomic<bool> spec_off{false};   // set when a batch showed that the inputs are mostly not bits
};

// the CSR of an .r1cs bound to one value-store layout (r1cs.hpp bind), on the device
struct BoundDev {
    uint64_t layout_id = ~0ull;      // 0: the plain layout (row = wire)
    bool typed = false;
    uint32_t n_brows = 0;
    uint64_t macs = 0, bit_adds = 0, n_int_constraints = 0, n_bterms = 0, n_fterms = 0, n_tcons = 0;
    uint64_t n_active = 0;
    DevBuf d_hdr, d_terms, d_bhdr, d_bterms, d_tcons, d_active, d_shh, d_shl, d_shm;
    uint64_t n_shift = 0;
    void release() {
        d_hdr.release(); d_terms.release(); d_bhdr.release(); d_bterms.release(); d_tcons.release(); d_active.release();
        d_shh.release(); d_shl.release(); d_shm.release();
        layout_id = ~0ull;
    }
};

static std::atomic<uint64_t> g_next_r1cs_uid{1};

struct cvmgpu_r1cs {
    r1cs::File file;
    const uint64_t uid = g_next_r1cs_uid++;
    // per device: the coefficient tables, the CSR bound to the plain layout and to the last program layout used there
    // (re-bound when it changes), and the buffers of the host entry point
    struct Dev {
        DevBuf d_coefs, d_cmag, d_cint, d_store, d_wtns, d_bad;
        BoundDev plain, typed;
        bool ready = false;
        void release() {
            d_coefs.release(); d_cmag.release(); d_cint.release(); d_store.release(); d_wtns.release(); d_bad.release();
            plain.release(); typed.release();
            ready = false;
        }
    };
    std::map<int, Dev> dev;
    std::mutex mu;
    // counters of the most recent typed binding (cvmgpu_r1cs_info_get)
    uint64_t last_macs = 0, last_bit_adds = 0, last_int = 0, last_bterms = 0, last_fterms = 0, last_tcons = 0;
};

static uint64_t g_next_layout_id = 1;

static void release_pipe_buffers();

extern "C" const char *cvmgpu_last_error(void) { return g_err.c_str(); }

extern "C" int cvmgpu_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

extern "C" int cvmgpu_set_tape_mode(int mode) {
    g_tape_mode = mode;
    return CVMGPU_OK;
}

extern "C" int cvmgpu_set_device(int device) {
    CUDA_TRY(cudaSetDevice(device));
    return CVMGPU_OK;
}

// ------------------------------------------------------------------------------------------ program
// launch-cost estimate of a tape (slot-file search of build_program and of fused_for)
static double tape_cost(const tape::Tape &t) {
    // resident warps per SM at one warp per CTA (what small batches launch): 1 KiB of the SM's 228 KiB is reserved
    // per CTA.  A 64 K batch -- 2 048 warps over 148 SMs, BASELINE configs 3 and 4 -- must fit in ONE wave: at 13
    // resident warps per SM it takes two (measured: Sha256(512) 115 ms against 64 ms), hence the quantisation term.
    const size_t per_warp = (size_t)(t.n_slots + (t.use_ring ? tape::LD_RING : 0)) * 1024u + (size_t)t.n_bslots * 4u + 1024u;
    // (the bit-file instantiation is bounded at 128 registers: 4 warps per scheduler, 16 per SM -- at 136 it was 3 and 12,
    // and a 64 K batch took two waves: 17.5 ms instead of 9.5)
    const double resident = std::min<double>(std::floor(228.0 * 1024.0 / (double)per_warp), t.n_bslots ? 16.0 : 64.0);
    const double waves = 2048.0 / (148.0 * std::max(1.0, resident));
    const double quant = std::ceil(waves) / std::max(waves, 1.0);
    // value-store traffic: field rows move 1 KiB per warp; a bit-row reload is one word, but it is a dependent
    // global load on the tape's critical path (measured on Sha256(512): 256 bit slots with 11 K such reloads
    // 22.6 ms, 512 with 1.7 K 20.1 ms)
    const double work = (double)t.ins.size() + 2.0 * (double)(t.stats.n_ld + t.stats.n_st - t.stats.n_ld_bool - t.stats.n_spill_st_bool) +
                        2.0 * (double)t.stats.n_ld_bool;
    return work * quant / std::max(4.0, std::min(resident, 24.0));
}

static int build_program_impl(cvm::Parser &parser, uint32_t n_slots, bool assume_bit_inputs, cvmgpu_program **out) {
    std::unique_ptr<cvmgpu_program> p(new cvmgpu_program());
    if (n_slots > 55) n_slots = 55;   // 55 * 4 KiB = 220 KiB of the 227 KiB a CTA may use
    try {
        std::unique_ptr<tape::Tracer> trp(new tape::Tracer(parser.prog));
        trp->assume_bit_inputs = assume_bit_inputs;
        trp->trace();
        // A field program with a sprinkling of 0/1 values (comparison results, the bits of a few scalars: the EdDSA verifier
        // has 3 %) is traced again without boolean-function instructions and compiled untyped: the bit-slot file would cost
        // it the field-only kernel (96 registers, 20 warps per SM) and the fused R1CS check for nothing.
        bool typed = true;
        {
            size_t n_bool = 0;
            for (uint8_t b : trp->isbool) n_bool += b;
            static const bool allow = !(getenv("CVMGPU_UNTYPED") && atoi(getenv("CVMGPU_UNTYPED")) == 0);
            if (allow && !assume_bit_inputs && n_bool > 0 && n_bool * 10 < trp->ops.size()) {
                typed = false;
                trp.reset(new tape::Tracer(parser.prog));
                trp->use_luts = false;
                trp->trace();
            }
        }
        tape::Tracer &tr = *trp;
        p->typed = typed;
        p->binv = tape::batch_inversions(tr);
        if (n_slots == 0) {
            // Fewer slots per witness = more resident warps per SM (1 KiB of shared memory per field slot and warp, 4 B per
            // bit slot), but more reloads/spills through the value store.  Pick the candidate with the best
            // (work / resident warps) estimate; the kernel stops gaining from occupancy at about 24 warps per SM.
            // Field-slot candidates beyond what the program keeps live at once cannot differ (max_live_field), and the
            // bit file is first sized to hold every live 0/1 value, then tried smaller for the chosen field file.
            static const uint32_t cand[] = {4, 8, 12, 16, 24, 32};   // (4: programs that are all bits keep the shared memory for the bit file)
            std::map<uint32_t, tape::XProg> prepared;   // by longest dot product: fusion and typing run once each
            auto build = [&](uint32_t c, uint32_t max_bslots) {
                const uint32_t mt = std::min<uint32_t>(16, c - 2);
                auto it = prepared.find(mt);
                if (it == prepared.end()) it = prepared.emplace(mt, tape::prepare_program(tr, mt, true, typed)).first;
                return tape::allocate_tape(tr, it->second, c, max_bslots);
            };
            double best = 0, prev_best_c = 0;
            uint32_t best_mt = 0;
            size_t n_bool_ops = 0;
            for (uint8_t b : tr.isbool) n_bool_ops += b;
            const bool bit_heavy = typed && (2 * n_bool_ops > tr.ops.size() || assume_bit_inputs);
            for (uint32_t c : cand) {
                if (c == 4 && !bit_heavy) continue;   // field programs: measured from 8 up (EdDSA verifier: 8 is best)
                uint32_t live_field = 0, full_bslots = 0;
                double best_c = 0;
                for (uint32_t nb : {2048u, 1024u, 512u, 256u}) {
                    if (nb != 2048u && nb >= full_bslots) continue;   // the file already holds every live 0/1 value
                    tape::Tape t = build(c, nb);
                    if (nb == 2048u) { full_bslots = t.n_bslots; live_field = t.stats.max_live_field; }
                    const double cost = tape_cost(t);
                    if (best_c == 0 || cost < best_c) best_c = cost;
                    if (best == 0 || cost < best) {
                        best = cost;
                        p->tape = std::move(t);
                        best_mt = std::min<uint32_t>(16, c - 2);
                    }
                }
                if (c >= live_field + 2) break;   // every field value already has a slot
                if (prev_best_c != 0 && best_c > prev_best_c) break;   // past the optimum: larger files only cost occupancy
                prev_best_c = best_c;
            }
            // field-only programs keep what a fused R1CS check is built from (fused.hpp)
            auto it = prepared.find(best_mt);
            if (it != prepared.end() && tape::check_fusable(it->second)) {
                p->xp = std::move(it->second);
                p->max_terms = best_mt;
                p->fusable = true;
            }
        } else {
            // explicit slot count (experiments): CVMGPU_BSLOTS caps the bit file
            const uint32_t nb = getenv("CVMGPU_BSLOTS") ? (uint32_t)atoi(getenv("CVMGPU_BSLOTS")) : 2048u;
            p->max_terms = std::min<uint32_t>(16, n_slots - 2);
            p->explicit_slots = n_slots;
            tape::XProg xp = tape::prepare_program(tr, p->max_terms, true, typed);
            p->tape = tape::allocate_tape(tr, xp, n_slots, std::max<uint32_t>(8, nb));
            if (tape::check_fusable(xp)) {
                p->xp = std::move(xp);
                p->fusable = true;
            }
        }
        p->consts = tr.consts;
        p->n_ssa = tr.ops.size();
        p->layout_id = g_next_layout_id++;
        p->tstats = tr.stats;
        p->wire_bool.reserve(tr.witness_ref.size());
        for (uint32_t loc : p->tape.wire_loc) p->wire_bool.push_back((loc & tape::ROW_BIT) ? 1 : 0);
        if (!p->wire_bool.empty()) p->wire_bool[0] = 1;   // the constant 1 (kept as a field row, plus a bit row of ones)
        p->n_signals = (uint64_t)parser.prog.n_signals;
        p->witness.assign(parser.prog.witness.begin(), parser.prog.witness.end());
        for (const cvm::MainInput &mi : parser.prog.main_inputs)
            p->main_inputs += mi.name + " " + std::to_string(mi.start) + " " + std::to_string(mi.size) + "\n";
        p->n_inputs = (uint32_t)tr.n_inputs;
        p->n_outputs = (uint32_t)tr.n_outputs;
        p->consts_mont.reserve(tr.consts.size());
        for (const fr::Fr &c : tr.consts) p->consts_mont.push_back(fr::to_mont(c));
    } catch (const tape::TraceError &e) {
        return fail(CVMGPU_ERR_UNSUPPORTED, e.what());
    } catch (const cvm::ParseError &e) {
        return fail(CVMGPU_ERR_PARSE, e.what());
    }
    *out = p.release();
    return CVMGPU_OK;
}

// Speculative typing.  Hash circuits take their message as unconstrained signals (nothing in Sha256(n) proves in[k] a bit),
// so everything derived from the message before the first bit decomposition is field arithmetic on values that ARE 0 / 1
// in every sensible input: for Sha256(512) 2 400 of the 2 401 field rows, 11 K of the 97 K tape instructions and all the
// Montgomery products.  A bit-heavy program is therefore traced a second time under the assumption that every main input
// is 0 or 1; that tape checks the assumption per witness (T_INPUT_BIT) and raises CVMGPU_ST_SPECULATION where it does not
// hold.  The host-buffer entry points run the speculative tape and recompute the witnesses that raised the flag with the
// general one, so results never depend on the assumption; device-API callers opt in (cvmgpu_program_speculative).
// CVMGPU_SPECULATE=0 in the environment disables it.
static int build_program(cvm::Parser &parser, uint32_t n_slots, cvmgpu_program **out) {
    cvmgpu_program *p = nullptr;
    if (int rc = build_program_impl(parser, n_slots, false, &p)) return rc;
    static const bool enabled = !(getenv("CVMGPU_SPECULATE") && atoi(getenv("CVMGPU_SPECULATE")) == 0);
    const tape::Tape &t = p->tape;
    // (a handful of inputs that feed bit decompositions -- Num2Bits, comparators -- are numbers, not bits)
    if (enabled && p->n_inputs >= 8 && t.n_fwires > 1 && t.n_bwires > 4 * (uint64_t)t.n_fwires) {
        cvmgpu_program *q = nullptr;
        if (build_program_impl(parser, n_slots, true, &q) == CVMGPU_OK) {
            // worth it when it removes at least half of the field rows
            if (q->tape.n_wires == t.n_wires && 2 * (uint64_t)q->tape.n_frows <= t.n_frows) p->spec.reset(q);
            else delete q;
        }
    }
    *out = p;
    return CVMGPU_OK;
}

extern "C" int cvmgpu_program_speculative(cvmgpu_program *p, cvmgpu_program **spec) {
    if (!p || !spec) return fail(CVMGPU_ERR_ARG, "null argument");
    *spec = p->spec.get();
    return CVMGPU_OK;
}

extern "C" int cvmgpu_program_wire_types(const cvmgpu_program *p, const uint8_t **is_bool, uint32_t *n) {
    if (!p || !is_bool || !n) return fail(CVMGPU_ERR_ARG, "null argument");
    *is_bool = p->wire_bool.data();
    *n = (uint32_t)p->wire_bool.size();
    return CVMGPU_OK;
}

extern "C" int cvmgpu_program_load(const char *cvm_path, uint32_t n_slots, cvmgpu_program **out) {
    if (!cvm_path || !out) return fail(CVMGPU_ERR_ARG, "null argument");
    cvm::Parser parser;
    try {
        parser.parse_file(cvm_path);
    } catch (const cvm::ParseError &e) {
        std::string m = e.what();
        return fail(m.rfind("cannot open", 0) == 0 ? CVMGPU_ERR_IO : CVMGPU_ERR_PARSE, m);
    } catch (const std::exception &e) {
        return fail(CVMGPU_ERR_PARSE, e.what());
    }
    return build_program(parser, n_slots, out);
}

extern "C" int cvmgpu_program_load_with_cpp(const char *cvm_path, const char *cpp_path, uint32_t n_slots, cvmgpu_program **out) {
    if (!cvm_path || !out) return fail(CVMGPU_ERR_ARG, "null argument");
    cvm::Parser parser;
    try {
        parser.parse_file(cvm_path);
        if (cpp_path) {
            std::ifstream f(cpp_path);
            if (!f) return fail(CVMGPU_ERR_IO, std::string("cannot open ") + cpp_path);
            std::stringstream ss;
            ss << f.rdbuf();
            parser.recover_creates(ss.str());
        }
    } catch (const cvm::ParseError &e) {
        std::string m = e.what();
        return fail(m.rfind("cannot open", 0) == 0 ? CVMGPU_ERR_IO : CVMGPU_ERR_PARSE, m);
    } catch (const std::exception &e) {
        return fail(CVMGPU_ERR_PARSE, e.what());
    }
    return build_program(parser, n_slots, out);
}

extern "C" int cvmgpu_program_load_text2(const char *cvm_text, size_t len, const char *cpp_text, size_t cpp_len, uint32_t n_slots,
                                         cvmgpu_program **out) {
    if (!cvm_text || !out) return fail(CVMGPU_ERR_ARG, "null argument");
    cvm::Parser parser;
    try {
        parser.parse_text(std::string(cvm_text, len));
        if (cpp_text) parser.recover_creates(std::string(cpp_text, cpp_len));
    } catch (const std::exception &e) {
        return fail(CVMGPU_ERR_PARSE, e.what());
    }
    return build_program(parser, n_slots, out);
}

static bool slurp(const char *path, std::string &out) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    std::stringstream ss;
    ss << f.rdbuf();
    out = ss.str();
    return true;
}

extern "C" int cvmgpu_program_load_files(const char *cvm_path, const char *cpp_path, const char *dat_path, uint32_t n_slots,
                                         cvmgpu_program **out) {
    if (!cvm_path || !out) return fail(CVMGPU_ERR_ARG, "null argument");
    if (dat_path && !cpp_path) return fail(CVMGPU_ERR_ARG, "the .dat io-map needs the section sizes of the generated C++ (cpp_path)");
    cvm::Parser parser;
    try {
        parser.parse_file(cvm_path);
        std::string cpp, dat;
        if (cpp_path) {
            if (!slurp(cpp_path, cpp)) return fail(CVMGPU_ERR_IO, std::string("cannot open ") + cpp_path);
            parser.recover_creates(cpp);
        }
        if (dat_path) {
            if (!slurp(dat_path, dat)) return fail(CVMGPU_ERR_IO, std::string("cannot open ") + dat_path);
            parser.read_dat_io_map(cpp, (const unsigned char *)dat.data(), dat.size());
        }
    } catch (const cvm::ParseError &e) {
        std::string m = e.what();
        return fail(m.rfind("cannot open", 0) == 0 ? CVMGPU_ERR_IO : CVMGPU_ERR_PARSE, m);
    } catch (const std::exception &e) {
        return fail(CVMGPU_ERR_PARSE, e.what());
    }
    return build_program(parser, n_slots, out);
}

extern "C" int cvmgpu_program_load_text3(const char *cvm_text, size_t len, const char *cpp_text, size_t cpp_len,
                                         const void *dat, size_t dat_len, uint32_t n_slots, cvmgpu_program **out) {
    if (!cvm_text || !out) return fail(CVMGPU_ERR_ARG, "null argument");
    if (dat && !cpp_text) return fail(CVMGPU_ERR_ARG, "the .dat io-map needs the section sizes of the generated C++ (cpp_text)");
    cvm::Parser parser;
    try {
        parser.parse_text(std::string(cvm_text, len));
        std::string cpp = cpp_text ? std::string(cpp_text, cpp_len) : std::string();
        if (cpp_text) parser.recover_creates(cpp);
        if (dat) parser.read_dat_io_map(cpp, (const unsigned char *)dat, dat_len);
    } catch (const std::exception &e) {
        return fail(CVMGPU_ERR_PARSE, e.what());
    }
    return build_program(parser, n_slots, out);
}

extern "C" int cvmgpu_program_load_text(const char *cvm_text, size_t len, uint32_t n_slots, cvmgpu_program **out) {
    if (!cvm_text || !out) return fail(CVMGPU_ERR_ARG, "null argument");
    cvm::Parser parser;
    try {
        parser.parse_text(std::string(cvm_text, len));
    } catch (const std::exception &e) {
        return fail(CVMGPU_ERR_PARSE, e.what());
    }
    return build_program(parser, n_slots, out);
}

// info structs carry their size in the first field: the caller sets it to sizeof(its struct), the library fills
// min(that, its own size) bytes -- a binding compiled against an older header keeps working when counters are added
template <class T>
static int fill_info(T *dst, const T &src) {
    uint32_t want = dst->struct_size;
    if (want < sizeof(uint32_t)) return fail(CVMGPU_ERR_ARG, "info.struct_size is not set (set it to sizeof(the struct) before the call)");
    if (want > sizeof(T)) want = (uint32_t)sizeof(T);
    memcpy(dst, &src, want);
    dst->struct_size = want;
    return CVMGPU_OK;
}

extern "C" int cvmgpu_program_info_get(const cvmgpu_program *p, cvmgpu_program_info *out) {
    if (!p || !out) return fail(CVMGPU_ERR_ARG, "null argument");
    cvmgpu_program_info v;
    cvmgpu_program_info *info = &v;
    memset(info, 0, sizeof(*info));
    info->n_signals = p->n_signals;
    info->n_wires = p->tape.n_wires;
    info->n_inputs = p->n_inputs;
    info->n_outputs = p->n_outputs;
    info->n_slots = p->tape.n_slots;
    info->n_rows = p->tape.n_rows;
    info->tape_len = p->tape.ins.size();
    info->ref_mul = p->tstats.ref_mul;
    info->ref_field_ops = p->tstats.ref_field_ops;
    info->cvm_instructions = p->tstats.cvm_instructions;
    info->tape_mul = p->tape.stats.n_mul;
    info->tape_div = p->tape.stats.n_div;
    info->tape_addsub = p->tape.stats.n_addsub;
    info->tape_other = p->tape.stats.n_other;
    info->tape_ld = p->tape.stats.n_ld;
    info->tape_st = p->tape.stats.n_st;
    info->tape_spill_st = p->tape.stats.n_spill_st;
    info->n_consts = (uint32_t)p->consts_mont.size();
    info->dyn_branches = (uint32_t)p->tstats.dyn_branches;
    info->ref_div = p->tstats.ref_div;
    info->tape_inv = p->tape.stats.n_inv;
    info->tape_sel = p->tape.stats.n_sel;
    info->tape_dot = p->tape.stats.n_dot;
    info->tape_dot_terms = p->tape.stats.n_dot_terms;
    info->tape_macs = p->tape.stats.macs;
    info->tape_ld_streamed = p->tape.stats.n_ld_streamed;
    info->unrolled_iterations = p->tstats.unrolled_iterations;
    info->tape_lut = p->tape.stats.n_lut;
    info->tape_ld_bool = p->tape.stats.n_ld_bool;
    info->tape_spill_st_bool = p->tape.stats.n_spill_st_bool;
    info->n_bool_wires = p->tape.n_bwires;
    info->n_bslots = p->tape.n_bslots;
    info->n_frows = p->tape.n_frows;
    info->n_brows = p->tape.n_brows;
    info->max_live_field = p->tape.stats.max_live_field;
    info->max_live_bool = p->tape.stats.max_live_bool;
    info->tape_int = p->tape.stats.n_int;
    return fill_info(out, v);
}

extern "C" int cvmgpu_program_tape(const cvmgpu_program *p, const void **ins, uint64_t *n_ins, const void **consts,
                                   uint32_t *n_consts) {
    if (!p) return fail(CVMGPU_ERR_ARG, "null argument");
    if (ins) *ins = p->tape.ins.data();
    if (n_ins) *n_ins = p->tape.ins.size();
    if (consts) *consts = p->consts_mont.data();
    if (n_consts) *n_consts = (uint32_t)p->consts_mont.size();
    return CVMGPU_OK;
}

extern "C" int cvmgpu_program_iconsts(const cvmgpu_program *p, const uint64_t **iconsts, uint32_t *n) {
    if (!p) return fail(CVMGPU_ERR_ARG, "null argument");
    if (iconsts) *iconsts = p->tape.iconsts.data();
    if (n) *n = (uint32_t)p->tape.iconsts.size();
    return CVMGPU_OK;
}

extern "C" int cvmgpu_program_wire_rows(const cvmgpu_program *p, const uint32_t **wire_loc, uint32_t *n) {
    if (!p) return fail(CVMGPU_ERR_ARG, "null argument");
    if (wire_loc) *wire_loc = p->tape.wire_loc.data();
    if (n) *n = (uint32_t)p->tape.wire_loc.size();
    return CVMGPU_OK;
}

extern "C" int cvmgpu_program_main_inputs(const cvmgpu_program *p, const char **text, size_t *len) {
    if (!p || !text) return fail(CVMGPU_ERR_ARG, "null argument");
    *text = p->main_inputs.c_str();
    if (len) *len = p->main_inputs.size();
    return CVMGPU_OK;
}

extern "C" int cvmgpu_program_witness(const cvmgpu_program *p, const uint64_t **signals, uint32_t *n) {
    if (!p) return fail(CVMGPU_ERR_ARG, "null argument");
    if (signals) *signals = p->witness.data();
    if (n) *n = (uint32_t)p->witness.size();
    return CVMGPU_OK;
}

// device buffers belong to the device they were allocated on: release them there
template <class F>
static void on_device(int device, F &&fn) {
    int cur = -1;
    bool sw = device >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != device && cudaSetDevice(device) == cudaSuccess;
    fn();
    if (sw) cudaSetDevice(cur);
}

static void release_program_tables(cvmgpu_program *p) {
    for (auto &kv : p->dev)
        on_device(kv.first, [&] {
            kv.second.d_tape.release(); kv.second.d_consts.release(); kv.second.d_wire_loc.release(); kv.second.d_iconsts.release();
            kv.second.d_flist.release(); kv.second.d_blist.release();
        });
    for (auto &kv : p->fused)
        if (kv.second) release_program_tables(kv.second.get());
    if (p->spec) release_program_tables(p->spec.get());
}

extern "C" void cvmgpu_program_free(cvmgpu_program *p) {
    if (!p) return;
    release_program_tables(p);
    release_pipe_buffers();
    delete p;
}

// The tape of `p` with the check of `r` scheduled into it (fused.hpp), built on first use and kept per constraint
// system; nullptr when the program is not field-only (its check keeps the separate kernels), when the two do not belong
// together, or with CVMGPU_FUSED=0 in the environment (measurements of the separate kernels).
// 0: never, 1: when it pays (field programs whose constraints are all evaluated in the field), 2: whenever it is possible
static std::atomic<int> g_fused_mode{getenv("CVMGPU_FUSED") ? atoi(getenv("CVMGPU_FUSED")) : 1};

extern "C" int cvmgpu_set_fused_mode(int mode) {
    if (mode < 0 || mode > 2) return fail(CVMGPU_ERR_ARG, "fused mode must be 0, 1 or 2");
    g_fused_mode = mode;
    return CVMGPU_OK;
}

static cvmgpu_program *fused_for(cvmgpu_program *p, cvmgpu_r1cs *r) {
    const int mode = g_fused_mode;
    if (mode == 0 || !p || !r || !p->fusable || r->file.n_wires != p->tape.n_wires) return nullptr;
    std::lock_guard<std::mutex> lock(p->mu);
    if (mode == 1) {
        // A program compiled untyped has 0/1 values the stand-alone check exploits at run time (it skips products with a 0 / 1
        // factor): EdDSAPoseidonVerifier untyped, measured 58.3 ms fused against 36.6 + 17.9 ms.
        if (!p->typed) return nullptr;
        // Constraints over 0/1 wires are evaluated 32 witnesses at a time by the table / integer kernels (r1cs.hpp bind):
        // a program that has them (EdDSAPoseidonVerifier typed: measured 62.2 ms fused against 38.8 + 19.8 ms) keeps those kernels
        auto w = p->fuse_worth.find(r->uid);
        if (w == p->fuse_worth.end()) {
            const r1cs::Bound b = r1cs::bind(r->file, p->tape.wire_loc.data(), p->tape.one_brow, p->tape.const_rows);
            w = p->fuse_worth.emplace(r->uid, b.n_table_constraints == 0 && b.n_int_constraints == 0).first;
        }
        if (!w->second) return nullptr;
    }
    auto it = p->fused.find(r->uid);
    if (it != p->fused.end()) return it->second.get();
    std::unique_ptr<cvmgpu_program> q(new cvmgpu_program());
    try {
        std::vector<fr::Fr> consts = p->consts;
        tape::XProg xp = tape::fuse_check(p->xp, consts, r->file, p->max_terms);
        // the slot file is chosen again: the check's operands stay live a little longer
        double best = 0;
        for (uint32_t c : {8u, 12u, 16u, 24u, 32u}) {
            if (c < p->max_terms + 2 || p->explicit_slots) continue;
            tape::Tape t = tape::allocate_tape(consts, p->n_ssa, xp, c, 2048);
            const double cost = tape_cost(t);
            const bool all_resident = c >= t.stats.max_live_field + 2;
            if (best == 0 || cost < best) {
                best = cost;
                q->tape = std::move(t);
            }
            if (all_resident) break;
        }
        if (best == 0) q->tape = tape::allocate_tape(consts, p->n_ssa, xp, p->tape.n_slots, 2048);   // (explicit slot count)
        q->consts_mont.reserve(consts.size());
        for (const fr::Fr &c : consts) q->consts_mont.push_back(fr::to_mont(c));
    } catch (const std::exception &e) {
        fail(CVMGPU_ERR_UNSUPPORTED, e.what());
        p->fused.emplace(r->uid, nullptr);   // do not try again
        return nullptr;
    }
    // the value store keeps the program's wire rows (wire_loc depends on the wires only); spill rows may differ
    if (q->tape.wire_loc != p->tape.wire_loc) {
        p->fused.emplace(r->uid, nullptr);
        return nullptr;
    }
    q->tstats = p->tstats;
    q->binv = p->binv;
    q->witness = p->witness;
    q->wire_bool = p->wire_bool;
    q->n_signals = p->n_signals;
    q->n_inputs = p->n_inputs;
    q->n_outputs = p->n_outputs;
    q->layout_id = g_next_layout_id++;
    cvmgpu_program *out = q.get();
    p->fused.emplace(r->uid, std::move(q));
    return out;
}

// dynamic shared memory of a tape CTA of nt witnesses: the field slots, the reload ring when the tape reloads
// anything, and the bit-slot file of each warp
static size_t tape_field_smem_per_witness(const cvmgpu_program *p) {
    return ((size_t)p->tape.n_slots + (p->tape.use_ring ? tape::LD_RING : 0)) * 2 * sizeof(uint4);
}
static size_t tape_smem(const cvmgpu_program *p, uint32_t nt) {
    return tape_field_smem_per_witness(p) * nt + (((size_t)p->tape.n_bslots * (nt / 32) * 4 + 15) & ~(size_t)15);
}

// the program's tables on the current device (uploaded on first use there)
static int upload_program(cvmgpu_program *p, cvmgpu_program::Dev **out) {
    int dev = -1;
    CUDA_TRY(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(p->mu);
    cvmgpu_program::Dev &d = p->dev[dev];
    *out = &d;
    if (d.ready) return CVMGPU_OK;
    size_t tb = (p->tape.ins.size() + TAPE_PAD) * sizeof(tape::TapeIns);
    size_t cb = std::max<size_t>(32, p->consts_mont.size() * sizeof(fr::Fr));
    if (int rc = d.d_tape.ensure(tb)) return rc;
    if (int rc = d.d_consts.ensure(cb)) return rc;
    if (int rc = d.d_wire_loc.ensure(std::max<size_t>(4, p->tape.wire_loc.size() * 4))) return rc;
    CUDA_TRY(cudaMemset(d.d_tape.p, 0, tb));   // the padding: T_NOP
    if (!p->tape.ins.empty())
        CUDA_TRY(cudaMemcpy(d.d_tape.p, p->tape.ins.data(), p->tape.ins.size() * sizeof(tape::TapeIns), cudaMemcpyHostToDevice));
    if (!p->consts_mont.empty())
        CUDA_TRY(cudaMemcpy(d.d_consts.p, p->consts_mont.data(), p->consts_mont.size() * sizeof(fr::Fr), cudaMemcpyHostToDevice));
    if (int rc = d.d_iconsts.ensure(std::max<size_t>(8, p->tape.iconsts.size() * 8))) return rc;
    if (!p->tape.iconsts.empty())
        CUDA_TRY(cudaMemcpy(d.d_iconsts.p, p->tape.iconsts.data(), p->tape.iconsts.size() * 8, cudaMemcpyHostToDevice));
    if (!p->tape.wire_loc.empty())
        CUDA_TRY(cudaMemcpy(d.d_wire_loc.p, p->tape.wire_loc.data(), p->tape.wire_loc.size() * 4, cudaMemcpyHostToDevice));
    {
        uint32_t ext[256];
        kern::fill_ext_table(ext);
        CUDA_TRY(cudaMemcpyToSymbol(kern::c_ext_table, ext, sizeof(ext)));
    }
    // the tables are read by kernels on non-blocking streams, which do not order themselves after the copies above
    // (pageable-memory copies may return once the data is staged)
    CUDA_TRY(cudaDeviceSynchronize());
    d.ready = true;
    return CVMGPU_OK;
}

// Typed value store of `bstride` witnesses: n_frows field rows (2 x 16 B x bstride each), then n_brows bit rows as one
// 32-bit word per warp of witnesses and row, warp-major: word (w >> 5) * n_brows + row.
static size_t store_field_bytes(const cvmgpu_program *p, uint64_t bstride) { return (size_t)p->tape.n_frows * 2 * sizeof(uint4) * bstride; }
static size_t store_bit_bytes(const cvmgpu_program *p, uint64_t bstride) {
    return (((size_t)((bstride + 31) / 32) * p->tape.n_brows * 4) + 15) & ~(size_t)15;
}
extern "C" size_t cvmgpu_store_bytes(const cvmgpu_program *p, uint64_t bstride) {
    if (!p) return 0;
    return std::max<size_t>(16, store_field_bytes(p, bstride) + store_bit_bytes(p, bstride));
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is per kernel and device, and several programs may be live: only ever raise it
template <int NT, bool BITS>
static int launch_tape(const kern::TapeParams &tp, unsigned grid, size_t smem, cudaStream_t st) {
    static size_t allowed[64] = {0};
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 64 && smem > allowed[dev]) {
        CUDA_TRY(cudaFuncSetAttribute(kern::tape_kernel<NT, BITS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        // the slots want the whole carve-out of the SM (more resident CTAs), nothing here relies on L1
        CUDA_TRY(cudaFuncSetAttribute(kern::tape_kernel<NT, BITS>, cudaFuncAttributePreferredSharedMemoryCarveout, g_carveout));
        allowed[dev] = smem;
    }
    kern::tape_kernel<NT, BITS><<<grid, NT, smem, st>>>(tp);
    CUDA_TRY(cudaGetLastError());
    return CVMGPU_OK;
}

static int run_tape(cvmgpu_program *p, const void *d_inputs, uint64_t B, uint64_t bstride, void *d_store, void *d_status,
                    void *d_first_bad, void *stream, uint32_t in_row_bytes = 0);

extern "C" int cvmgpu_witness_batch_dev(cvmgpu_program *p, const void *d_inputs, uint64_t B, uint64_t bstride,
                                        void *d_store, void *d_status, void *stream) {
    return run_tape(p, d_inputs, B, bstride, d_store, d_status, nullptr, stream);
}

static int run_tape(cvmgpu_program *p, const void *d_inputs, uint64_t B, uint64_t bstride, void *d_store, void *d_status,
                    void *d_first_bad, void *stream, uint32_t in_row_bytes) {
    if (!p || !d_store || (!d_inputs && p->n_inputs)) return fail(CVMGPU_ERR_ARG, "null argument");
    if (B == 0) return CVMGPU_OK;
    if (bstride < B) return fail(CVMGPU_ERR_ARG, "bstride < B");
    if (bstride >> 27) return fail(CVMGPU_ERR_ARG, "bstride must be below 2^27 witnesses (32-bit row stride)");
    cvmgpu_program::Dev *pd = nullptr;
    if (int rc = upload_program(p, &pd)) return rc;
    kern::TapeParams tp;
    tp.tape = (const tape::TapeIns *)pd->d_tape.p;
    tp.n_ins = (uint32_t)p->tape.ins.size();
    tp.consts = (const uint4 *)pd->d_consts.p;
    tp.store = (uint4 *)d_store;
    tp.bits = (uint32_t *)((char *)d_store + store_field_bytes(p, bstride));
    tp.bstride = bstride;
    tp.n_brows = p->tape.n_brows;
    tp.n_bslots = p->tape.n_bslots;
    tp.iconsts = (const unsigned long long *)pd->d_iconsts.p;
    tp.inputs = (const uint4 *)d_inputs;
    tp.n_inputs = p->n_inputs;
    tp.status = (uint32_t *)d_status;
    tp.first_bad = (uint32_t *)d_first_bad;
    tp.in_row_bytes = in_row_bytes;
    tp.B = B;
    // CTA size: 128 witnesses for large batches; smaller CTAs when the batch would leave SMs with an uneven number of
    // CTAs (a 64 K batch is 512 CTAs of 128 for 148 SMs: 3.46 per SM)
    int sms = 148;
    {
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const uint64_t want = (uint64_t)sms * 8;
    const size_t per_w = tape_field_smem_per_witness(p);
    cudaStream_t st = (cudaStream_t)stream;
    const uint32_t nt = (B / 128 >= want) ? 128u : (B / 64 >= want) ? 64u : 32u;
    uint64_t grid = (B + nt - 1) / nt;
    if (grid > 0x7fffffffull) return fail(CVMGPU_ERR_ARG, "batch too large for one launch");
    tp.ring_off = p->tape.n_slots * 2 * nt;
    tp.bslot_off = (uint32_t)(per_w * nt / sizeof(uint4));
    const size_t smem = tape_smem(p, nt);
    // programs without values typed 0/1 run the instantiation that has no bit-slot file
    if (p->tape.n_bslots) {
        if (nt == 128) return launch_tape<128, true>(tp, (unsigned)grid, smem, st);
        if (nt == 64) return launch_tape<64, true>(tp, (unsigned)grid, smem, st);
        return launch_tape<32, true>(tp, (unsigned)grid, smem, st);
    }
    if (nt == 128) return launch_tape<128, false>(tp, (unsigned)grid, smem, st);
    if (nt == 64) return launch_tape<64, false>(tp, (unsigned)grid, smem, st);
    return launch_tape<32, false>(tp, (unsigned)grid, smem, st);
}

// Witness generation + R1CS check on device buffers.  Field-only programs run one kernel (the check is part of the tape,
// fused.hpp); the others run the tape and then the check kernels on the store.  The store must hold
// cvmgpu_store_bytes_checked(p, r, bstride) bytes.
extern "C" int cvmgpu_witness_batch_checked_dev(cvmgpu_program *p, cvmgpu_r1cs *r, const void *d_inputs, uint64_t B, uint64_t bstride,
                                                void *d_store, void *d_status, void *d_first_bad, void *stream) {
    if (!p || !r || !d_first_bad) return fail(CVMGPU_ERR_ARG, "null argument");
    if (r->file.n_wires != p->tape.n_wires) return fail(CVMGPU_ERR_ARG, "r1cs and program disagree on the number of wires");
    if (cvmgpu_program *q = fused_for(p, r)) return run_tape(q, d_inputs, B, bstride, d_store, d_status, d_first_bad, stream);
    if (int rc = run_tape(p, d_inputs, B, bstride, d_store, d_status, nullptr, stream)) return rc;
    if (B == 0) return CVMGPU_OK;
    return cvmgpu_r1cs_check_store_dev(r, p, d_store, B, bstride, d_first_bad, stream);
}

extern "C" size_t cvmgpu_store_bytes_checked(cvmgpu_program *p, cvmgpu_r1cs *r, uint64_t bstride) {
    if (!p) return 0;
    size_t n = cvmgpu_store_bytes(p, bstride);
    if (cvmgpu_program *q = fused_for(p, r)) n = std::max(n, cvmgpu_store_bytes(q, bstride));
    return n;
}

// the tape that cvmgpu_witness_batch_checked[_dev] runs for this pair when it is a fused one (CVMGPU_ERR_UNSUPPORTED if not)
extern "C" int cvmgpu_program_fused_info_get(cvmgpu_program *p, cvmgpu_r1cs *r, cvmgpu_program_info *out) {
    if (!p || !r || !out) return fail(CVMGPU_ERR_ARG, "null argument");
    cvmgpu_program *q = fused_for(p, r);
    if (!q) return fail(CVMGPU_ERR_UNSUPPORTED, "no fused tape for this program / constraint system (not field-only, or disabled)");
    return cvmgpu_program_info_get(q, out);
}

extern "C" int cvmgpu_program_fused_tape(cvmgpu_program *p, cvmgpu_r1cs *r, const void **ins, uint64_t *n_ins, const void **consts,
                                         uint32_t *n_consts) {
    if (!p || !r) return fail(CVMGPU_ERR_ARG, "null argument");
    cvmgpu_program *q = fused_for(p, r);
    if (!q) return fail(CVMGPU_ERR_UNSUPPORTED, "no fused tape for this program / constraint system (not field-only, or disabled)");
    return cvmgpu_program_tape(q, ins, n_ins, consts, n_consts);
}

static kern::StoreView store_view(const cvmgpu_program *p, const cvmgpu_program::Dev *pd, const void *d_store, uint64_t bstride) {
    kern::StoreView sv;
    sv.store = (const uint4 *)d_store;
    sv.bits = (const uint32_t *)((const char *)d_store + store_field_bytes(p, bstride));
    sv.bstride = bstride;
    sv.n_brows = p->tape.n_brows;
    sv.wire_loc = (const uint32_t *)pd->d_wire_loc.p;
    return sv;
}

extern "C" int cvmgpu_witness_export_range_dev(cvmgpu_program *p, const void *d_store, uint64_t B, uint64_t bstride,
                                               uint32_t wire0, uint32_t n_sel, void *d_out, void *stream) {
    if (!p || !d_store || !d_out) return fail(CVMGPU_ERR_ARG, "null argument");
    if ((uint64_t)wire0 + n_sel > p->tape.n_wires) return fail(CVMGPU_ERR_ARG, "wire range exceeds the witness");
    if (B == 0 || n_sel == 0) return CVMGPU_OK;
    cvmgpu_program::Dev *pd = nullptr;
    if (int rc = upload_program(p, &pd)) return rc;
    dim3 grid((unsigned)((B + 31) / 32), (n_sel + 31) / 32);
    kern::export_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(store_view(p, pd, d_store, bstride), B, wire0, n_sel, (uint4 *)d_out);
    CUDA_TRY(cudaGetLastError());
    return CVMGPU_OK;
}

extern "C" int cvmgpu_witness_export_dev(cvmgpu_program *p, const void *d_store, uint64_t B, uint64_t bstride, void *d_wtns,
                                         void *stream) {
    if (!p) return fail(CVMGPU_ERR_ARG, "null argument");
    return cvmgpu_witness_export_range_dev(p, d_store, B, bstride, 0, p->tape.n_wires, d_wtns, stream);
}

// ---- packed rows: [field wires x 32 B][0/1 wires as bits], the witness in the types the program proved
static void packed_lists(const cvmgpu_program *p, std::vector<uint32_t> &fl, std::vector<uint32_t> &bl) {
    for (uint32_t loc : p->tape.wire_loc) {
        if (loc & tape::ROW_BIT) bl.push_back(loc & ~tape::ROW_BIT);
        else fl.push_back(loc);
    }
}
static size_t packed_row_bytes(const cvmgpu_program *p) {
    const size_t nf = p->tape.n_fwires, nb = p->tape.n_wires - p->tape.n_fwires;
    return nf * 32 + ((nb + 127) / 128) * 16;   // rows stay 16-byte aligned
}

extern "C" int cvmgpu_witness_export_packed_dev(cvmgpu_program *p, const void *d_store, uint64_t B, uint64_t bstride, void *d_out,
                                                void *stream) {
    if (!p || !d_store || !d_out) return fail(CVMGPU_ERR_ARG, "null argument");
    if (B == 0) return CVMGPU_OK;
    cvmgpu_program::Dev *pd = nullptr;
    if (int rc = upload_program(p, &pd)) return rc;
    {
        std::lock_guard<std::mutex> lock(p->mu);
        if (!pd->d_flist.p) {   // the row lists of this device, built once
            std::vector<uint32_t> fl, bl;
            packed_lists(p, fl, bl);
            if (int rc = pd->d_flist.ensure(std::max<size_t>(4, fl.size() * 4))) return rc;
            if (int rc = pd->d_blist.ensure(std::max<size_t>(4, bl.size() * 4))) return rc;
            if (!fl.empty()) CUDA_TRY(cudaMemcpy(pd->d_flist.p, fl.data(), fl.size() * 4, cudaMemcpyHostToDevice));
            if (!bl.empty()) CUDA_TRY(cudaMemcpy(pd->d_blist.p, bl.data(), bl.size() * 4, cudaMemcpyHostToDevice));
            CUDA_TRY(cudaDeviceSynchronize());
        }
    }
    kern::PackedView pv;
    pv.sv = store_view(p, pd, d_store, bstride);
    pv.flist = (const uint32_t *)pd->d_flist.p;
    pv.blist = (const uint32_t *)pd->d_blist.p;
    pv.n_f = p->tape.n_fwires;
    pv.n_b = p->tape.n_wires - p->tape.n_fwires;
    pv.n_ftiles = (pv.n_f + 7) / 8;
    pv.row_bytes = packed_row_bytes(p);
    const uint32_t n_btiles = (pv.n_b + 31) / 32;
    if (pv.n_ftiles + n_btiles == 0) return CVMGPU_OK;
    if (pv.n_ftiles + n_btiles > 65535) return fail(CVMGPU_ERR_UNSUPPORTED, "witness too wide for the packed export");
    dim3 grid((unsigned)((B + 255) / 256), pv.n_ftiles + n_btiles);
    kern::export_packed_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pv, B, (unsigned char *)d_out);
    CUDA_TRY(cudaGetLastError());
    return CVMGPU_OK;
}

// row layout of the packed export of `p` (or of its bit-input tape): n_field wires x 32 bytes, then n_bits wires as bits;
// cvmgpu_program_wire_rows of the same handle says which wire is which (ROW_BIT = a 0/1 wire), both in wire order
extern "C" int cvmgpu_packed_layout(cvmgpu_program *p, int bit_input_tape, uint64_t *row_bytes, uint32_t *n_field, uint32_t *n_bits,
                                    const uint32_t **wire_rows) {
    if (!p) return fail(CVMGPU_ERR_ARG, "null argument");
    if (bit_input_tape && !p->spec) return fail(CVMGPU_ERR_UNSUPPORTED, "this program has no bit-input tape (its inputs are field elements)");
    cvmgpu_program *q = bit_input_tape ? p->spec.get() : p;
    uint32_t nf = 0, nb = 0;
    for (uint32_t loc : q->tape.wire_loc) ((loc & tape::ROW_BIT) ? nb : nf)++;
    if (row_bytes) *row_bytes = packed_row_bytes(q);
    if (n_field) *n_field = nf;
    if (n_bits) *n_bits = nb;
    if (wire_rows) *wire_rows = q->tape.wire_loc.data();
    return CVMGPU_OK;
}

static int upload_r1cs(cvmgpu_r1cs *r, cvmgpu_r1cs::Dev **out);

// largest chunk of witnesses whose buffers fit in the free device memory
static uint64_t pick_chunk(uint64_t B, size_t bytes_per_witness, size_t already_held = 0) {
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return 0;
    uint64_t cap = (uint64_t)((double)(free_b + already_held) * 0.85 / (double)std::max<size_t>(1, bytes_per_witness));
    cap = cap / 1024 * 1024;
    if (cap < 128) cap = 128;
    return std::min<uint64_t>(B, cap);
}

// Host-buffer pipeline: the batch is cut into chunks that alternate between two streams, each with its own
// device buffers, so that (with pinned host memory) the H2D copy / kernels / D2H copy of consecutive chunks
// overlap.  Optional R1CS check runs on the value store before it is exported.
struct PipeBufs {
    DevBuf store, inputs, status, wtns, bad;
    cudaStream_t stream = nullptr;
    void release() {
        store.release(); inputs.release(); status.release(); wtns.release(); bad.release();
        if (stream) cudaStreamDestroy(stream);
        stream = nullptr;
    }
};
// one set per device, shared by the host threads (a call holds the device's lock while it uses the buffers); released by
// cvmgpu_program_free / cvmgpu_release_buffers
struct PipeSet {
    PipeBufs pipe[2];
    std::mutex mu;
};
static std::mutex g_pipes_mu;
static std::map<int, PipeSet> g_pipes;

static void release_pipe_buffers() {
    std::lock_guard<std::mutex> lock(g_pipes_mu);
    for (auto &kv : g_pipes) {
        std::lock_guard<std::mutex> l2(kv.second.mu);
        on_device(kv.first, [&] { for (auto &pb : kv.second.pipe) pb.release(); });
    }
}

extern "C" void cvmgpu_release_buffers(void) { release_pipe_buffers(); }

static int batch_select_impl(cvmgpu_program *p, cvmgpu_r1cs *r, const uint8_t *inputs, uint64_t B, uint32_t wire0, uint32_t n_sel,
                             uint8_t *wtns_out, uint32_t *status, uint32_t *first_bad, uint32_t in_row_bytes = 0, bool packed = false);

// The whole witness in the types the program proved (cvmgpu_packed_layout): field wires as 32-byte values, 0/1 wires as
// bits.  For a hash circuit that is what there is to say about a witness -- Sha256(512): 8.6 KB instead of 2.2 MB -- and
// what a device-to-host link can carry at the rate the kernels produce it.  inputs_are_bits: the inputs are packed bits
// (cvmgpu_witness_batch_bits) and the row layout is that of the bit-input tape; else 32-byte field elements and the
// layout of the program itself (computed by the general tape: a fixed layout cannot depend on what the inputs turn out to be).
extern "C" int cvmgpu_witness_batch_packed(cvmgpu_program *p, cvmgpu_r1cs *r, const uint8_t *inputs, int inputs_are_bits, uint64_t B,
                                           uint8_t *packed_out, uint32_t *status, uint32_t *first_bad) {
    if (!p || !packed_out) return fail(CVMGPU_ERR_ARG, "null argument");
    if (inputs_are_bits && !p->spec) return fail(CVMGPU_ERR_UNSUPPORTED, "this program has no bit-input tape (its inputs are field elements)");
    cvmgpu_program *q = inputs_are_bits ? p->spec.get() : p;
    return batch_select_impl(q, r, inputs, B, 0, q->tape.n_wires, packed_out, status, first_bad,
                             inputs_are_bits ? (p->n_inputs + 7) / 8 : 0, true);
}

// Inputs as PACKED BITS (row = ceil(n_inputs / 8) bytes per witness, input k = bit k & 7 of byte k >> 3): for programs whose
// inputs are bits -- those that have a bit-input tape (cvmgpu_program_speculative) -- 256 times less to upload than one
// 32-byte field element per bit, and nothing to speculate about.  Same outputs as cvmgpu_witness_batch_select.
extern "C" int cvmgpu_witness_batch_bits(cvmgpu_program *p, cvmgpu_r1cs *r, const uint8_t *input_bits, uint64_t B, uint32_t wire0,
                                         uint32_t n_sel, uint8_t *wtns_out, uint32_t *status, uint32_t *first_bad) {
    if (!p) return fail(CVMGPU_ERR_ARG, "null argument");
    if (!p->spec) return fail(CVMGPU_ERR_UNSUPPORTED, "this program has no bit-input tape (its inputs are field elements)");
    return batch_select_impl(p->spec.get(), r, input_bits, B, wire0, n_sel, wtns_out, status, first_bad, (p->n_inputs + 7) / 8);
}

extern "C" int cvmgpu_witness_batch_select(cvmgpu_program *p, cvmgpu_r1cs *r, const uint8_t *inputs, uint64_t B,
                                           uint32_t wire0, uint32_t n_sel, uint8_t *wtns_out, uint32_t *status,
                                           uint32_t *first_bad) {
    if (!p) return fail(CVMGPU_ERR_ARG, "null argument");
    if (!p->spec || p->spec_off || B == 0) return batch_select_impl(p, r, inputs, B, wire0, n_sel, wtns_out, status, first_bad);
    // Speculative typing (build_program): the tape traced under "every main input is 0 or 1" runs the batch; the witnesses
    // whose inputs are not (CVMGPU_ST_SPECULATION) are recomputed with the general tape and put in their places.
    std::vector<uint32_t> st_local;
    uint32_t *st = status;
    if (!st) {
        st_local.resize(B);
        st = st_local.data();
    }
    if (int rc = batch_select_impl(p->spec.get(), r, inputs, B, wire0, n_sel, wtns_out, st, first_bad)) return rc;
    std::vector<uint64_t> redo;
    for (uint64_t b = 0; b < B; b++)
        if (st[b] == tape::ST_SPECULATION) redo.push_back(b);
    if (redo.empty()) return CVMGPU_OK;
    // inputs that are mostly not bits: this circuit's inputs are numbers after all -- stop speculating on it
    if (redo.size() * 2 > B && B >= 16) p->spec_off = true;
    if (!wtns_out) n_sel = 0;
    const size_t in_row = (size_t)p->n_inputs * 32, out_row = (size_t)n_sel * 32, n = redo.size();
    std::vector<uint8_t> t_in(n * in_row), t_wt(n * out_row);
    std::vector<uint32_t> t_st(n), t_bad(n);
    for (size_t k = 0; k < n; k++) memcpy(t_in.data() + k * in_row, inputs + redo[k] * in_row, in_row);
    if (int rc = batch_select_impl(p, r, t_in.data(), n, wire0, n_sel, out_row ? t_wt.data() : nullptr, t_st.data(),
                                   r ? t_bad.data() : nullptr))
        return rc;
    for (size_t k = 0; k < n; k++) {
        if (out_row) memcpy(wtns_out + redo[k] * out_row, t_wt.data() + k * out_row, out_row);
        st[redo[k]] = t_st[k];
        if (r && first_bad) first_bad[redo[k]] = t_bad[k];
    }
    return CVMGPU_OK;
}

static int batch_select_impl(cvmgpu_program *p, cvmgpu_r1cs *r, const uint8_t *inputs, uint64_t B, uint32_t wire0, uint32_t n_sel,
                             uint8_t *wtns_out, uint32_t *status, uint32_t *first_bad, uint32_t in_row_bytes, bool packed) {
    if (!p) return fail(CVMGPU_ERR_ARG, "null argument");
    if (B == 0) return CVMGPU_OK;
    if (!inputs && p->n_inputs) return fail(CVMGPU_ERR_ARG, "null argument");
    if (r && !first_bad) return fail(CVMGPU_ERR_ARG, "first_bad is required with an r1cs handle");
    if (r && r->file.n_wires != p->tape.n_wires) return fail(CVMGPU_ERR_ARG, "r1cs and program disagree on the number of wires");
    if (wtns_out && (uint64_t)wire0 + n_sel > p->tape.n_wires) return fail(CVMGPU_ERR_ARG, "wire range exceeds the witness");
    if (cvmgpu_device_count() <= 0) return fail(CVMGPU_ERR_CUDA, "no CUDA device available (there is no CPU fallback)");
    if (!wtns_out) n_sel = 0;
    // field-only programs run ONE kernel: the tape with the check of `r` scheduled into it
    const bool fused = r && fused_for(p, r) != nullptr;
    if (fused) p = fused_for(p, r);
    {
        cvmgpu_program::Dev *pd = nullptr;
        if (int rc = upload_program(p, &pd)) return rc;
        cvmgpu_r1cs::Dev *rd = nullptr;
        if (r && !fused)
            if (int rc = upload_r1cs(r, &rd)) return rc;
    }
    const size_t in_row = in_row_bytes ? (size_t)in_row_bytes : (size_t)p->n_inputs * 32;
    const size_t out_row = packed ? packed_row_bytes(p) : (size_t)n_sel * 32;
    // chunks keep the D2H of one chunk under the kernels of the next: about 1 GiB of exported rows per chunk.  With little
    // or nothing to download there is nothing to overlap, and one large launch fills the GPU best (and pays the host-side
    // cost of a chunk once).
    uint64_t chunk = 1u << 20;
    if (out_row > 256) chunk = std::min<uint64_t>(262144, std::max<uint64_t>(4096, ((1ull << 30) / out_row) / 1024 * 1024));
    if (const char *e = getenv("CVMGPU_CHUNK")) chunk = std::max<uint64_t>(1024, strtoull(e, nullptr, 10));   // experiments
    if (B <= chunk) chunk = B;
    int dev = -1;
    CUDA_TRY(cudaGetDevice(&dev));
    PipeSet *ps;
    {
        std::lock_guard<std::mutex> lock(g_pipes_mu);
        ps = &g_pipes[dev];
    }
    std::lock_guard<std::mutex> pipe_lock(ps->mu);
    PipeBufs *pipe = ps->pipe;
    int nbuf = (B > chunk) ? 2 : 1;
    if (const char *e = getenv("CVMGPU_NBUF")) nbuf = std::max(1, std::min(2, atoi(e)));   // experiments
    {
        // The free-memory query (cudaMemGetInfo) takes driver-wide locks: on a host whose other GPUs are busy allocating it was
        // seen to block for 50-90 ms (traced: the whole of the "slow" calls of the full-row leg, none of it in the pipeline).
        // It is only needed when the buffers have to grow.
        const uint64_t cs = (chunk + 31) / 32 * 32;
        bool have = true;
        for (int k = 0; k < nbuf; k++) {
            const PipeBufs &pb = pipe[k];
            have = have && pb.stream && pb.store.n >= cvmgpu_store_bytes(p, cs) && pb.inputs.n >= std::max<size_t>(32, in_row * chunk) &&
                   pb.status.n >= 4 * chunk && pb.bad.n >= 4 * chunk && (!out_row || pb.wtns.n >= out_row * chunk);
        }
        if (!have) {
            const size_t per_w = (cvmgpu_store_bytes(p, 1024) + 1023) / 1024 + in_row + out_row + 8;
            size_t held = 0;   // what these buffers already hold counts as available
            for (int k = 0; k < 2; k++) held += pipe[k].store.n + pipe[k].inputs.n + pipe[k].wtns.n;
            uint64_t fit = pick_chunk(B, 2 * per_w, held);
            if (fit == 0) return fail(CVMGPU_ERR_CUDA, "cudaMemGetInfo failed");
            if (fit < chunk) {
                chunk = fit;
                nbuf = (B > chunk) ? 2 : 1;
            }
        }
    }
    const uint64_t cstride = (chunk + 31) / 32 * 32;
    for (int k = 0; k < nbuf; k++) {
        PipeBufs &pb = pipe[k];
        if (!pb.stream) CUDA_TRY(cudaStreamCreateWithFlags(&pb.stream, cudaStreamNonBlocking));
        if (int rc = pb.store.ensure(cvmgpu_store_bytes(p, cstride))) return rc;
        if (int rc = pb.inputs.ensure(std::max<size_t>(32, in_row * chunk))) return rc;
        if (int rc = pb.status.ensure(4 * chunk)) return rc;
        if (int rc = pb.bad.ensure(4 * chunk)) return rc;
        if (out_row)
            if (int rc = pb.wtns.ensure(out_row * chunk)) return rc;
    }
    int rc = CVMGPU_OK;
    uint64_t idx = 0;
    // CVMGPU_TRACE=1: per-chunk device timestamps (events) of a call, printed when it took more than 1.5 x its best so far
    static const bool trace = getenv("CVMGPU_TRACE") && atoi(getenv("CVMGPU_TRACE"));
    std::vector<cudaEvent_t> tev;
    auto mark = [&](cudaStream_t st) {
        if (!trace) return;
        cudaEvent_t ev;
        cudaEventCreate(&ev);
        cudaEventRecord(ev, st);
        tev.push_back(ev);
    };
    const auto host_t0 = std::chrono::steady_clock::now();
    for (uint64_t b0 = 0; b0 < B && rc == CVMGPU_OK; b0 += chunk, idx++) {
        PipeBufs &pb = pipe[idx % nbuf];
        cudaStream_t s = pb.stream;
        uint64_t n = std::min<uint64_t>(chunk, B - b0);
        cudaError_t e = cudaSuccess;
        mark(s);
        // the stream's previous chunk must have left its buffers (stream order guarantees it)
        if (in_row) e = cudaMemcpyAsync(pb.inputs.p, inputs + b0 * in_row, n * in_row, cudaMemcpyHostToDevice, s);
        mark(s);
        if (e == cudaSuccess) rc = run_tape(p, pb.inputs.p, n, cstride, pb.store.p, pb.status.p, fused ? pb.bad.p : nullptr, s, in_row_bytes);
        mark(s);
        if (e == cudaSuccess && rc == CVMGPU_OK && r) {
            if (!fused) rc = cvmgpu_r1cs_check_store_dev(r, p, pb.store.p, n, cstride, pb.bad.p, s);
            if (rc == CVMGPU_OK) e = cudaMemcpyAsync(first_bad + b0, pb.bad.p, n * 4, cudaMemcpyDeviceToHost, s);
        }
        if (e == cudaSuccess && rc == CVMGPU_OK && out_row) {
            rc = packed ? cvmgpu_witness_export_packed_dev(p, pb.store.p, n, cstride, pb.wtns.p, s)
                        : cvmgpu_witness_export_range_dev(p, pb.store.p, n, cstride, wire0, n_sel, pb.wtns.p, s);
            mark(s);
            if (rc == CVMGPU_OK) e = cudaMemcpyAsync(wtns_out + b0 * out_row, pb.wtns.p, n * out_row, cudaMemcpyDeviceToHost, s);
            mark(s);
        }
        if (e == cudaSuccess && rc == CVMGPU_OK && status) e = cudaMemcpyAsync(status + b0, pb.status.p, n * 4, cudaMemcpyDeviceToHost, s);
        if (e != cudaSuccess) rc = fail(CVMGPU_ERR_CUDA, std::string("host-buffer pipeline: ") + cudaGetErrorString(e));
    }
    const auto host_t1 = std::chrono::steady_clock::now();
    // also on the error path: copies into the caller's buffers must not be in flight when this returns
    for (int k = 0; k < nbuf; k++) {
        cudaError_t e = cudaStreamSynchronize(pipe[k].stream);
        if (e != cudaSuccess && rc == CVMGPU_OK) rc = fail(CVMGPU_ERR_CUDA, std::string("cudaStreamSynchronize: ") + cudaGetErrorString(e));
    }
    if (trace && !tev.empty()) {
        static double best = 1e30;
        const double total = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count();
        best = std::min(best, total);
        if (total > 1.5 * best) {
            fprintf(stderr, "[cvmgpu trace] call %.1f ms (best %.1f), enqueue %.1f ms; per chunk [start, h2d, tape, export, d2h] ms since the first event:",
                    total, best, std::chrono::duration<double, std::milli>(host_t1 - host_t0).count());
            for (size_t k = 0; k < tev.size(); k++) {
                float ms = 0;
                cudaEventElapsedTime(&ms, tev[0], tev[k]);
                fprintf(stderr, "%s%.1f", k % 5 == 0 ? " |" : " ", ms);
            }
            fprintf(stderr, "\n");
        }
        for (cudaEvent_t ev : tev) cudaEventDestroy(ev);
    }
    return rc;
}

// Single-process multi-device driver: witnesses are independent, so the batch is cut into one contiguous slice per device
// of the mask and each slice runs cvmgpu_witness_batch_select on its device from its own host thread (own streams, own
// pipeline buffers, the program / constraint tables replicated per device).  No inter-device traffic at all: the
// per-witness flags land in the caller's arrays.
extern "C" int cvmgpu_witness_batch_multi(cvmgpu_program *p, cvmgpu_r1cs *r, const uint8_t *inputs, uint64_t B, uint32_t device_mask,
                                          uint32_t wire0, uint32_t n_sel, uint8_t *wtns_out, uint32_t *status, uint32_t *first_bad) {
    if (!p) return fail(CVMGPU_ERR_ARG, "null argument");
    const int n_dev = cvmgpu_device_count();
    if (n_dev <= 0) return fail(CVMGPU_ERR_CUDA, "no CUDA device available (there is no CPU fallback)");
    std::vector<int> devs;
    for (int d = 0; d < 32; d++)
        if ((device_mask >> d) & 1u) {
            if (d >= n_dev) return fail(CVMGPU_ERR_ARG, "device_mask names device " + std::to_string(d) + " but only " + std::to_string(n_dev) + " are visible");
            devs.push_back(d);
        }
    if (devs.empty()) return fail(CVMGPU_ERR_ARG, "empty device_mask");
    if (B == 0) return CVMGPU_OK;
    if (!wtns_out) n_sel = 0;
    const size_t in_row = (size_t)p->n_inputs * 32, out_row = (size_t)n_sel * 32;
    const uint64_t per = (B + devs.size() - 1) / devs.size();
    std::vector<int> rcs(devs.size(), CVMGPU_OK);
    std::vector<std::string> errs(devs.size());
    std::vector<std::thread> threads;
    int caller_dev = 0;
    cudaGetDevice(&caller_dev);
    for (size_t k = 0; k < devs.size(); k++) {
        const uint64_t b0 = std::min<uint64_t>(B, k * per), n = std::min<uint64_t>(per, B - b0);
        if (n == 0) continue;
        threads.emplace_back([&, k, b0, n] {
            if (cudaSetDevice(devs[k]) != cudaSuccess) {
                rcs[k] = CVMGPU_ERR_CUDA;
                errs[k] = "cudaSetDevice(" + std::to_string(devs[k]) + ") failed";
                return;
            }
            rcs[k] = cvmgpu_witness_batch_select(p, r, inputs ? inputs + b0 * in_row : nullptr, n, wire0, n_sel,
                                                 wtns_out ? wtns_out + b0 * out_row : nullptr, status ? status + b0 : nullptr,
                                                 first_bad ? first_bad + b0 : nullptr);
            if (rcs[k] != CVMGPU_OK) errs[k] = "device " + std::to_string(devs[k]) + ": " + cvmgpu_last_error();
        });
    }
    for (auto &t : threads) t.join();
    cudaSetDevice(caller_dev);
    for (size_t k = 0; k < devs.size(); k++)
        if (rcs[k] != CVMGPU_OK) return fail(rcs[k], errs[k]);
    return CVMGPU_OK;
}

extern "C" int cvmgpu_witness_batch_checked(cvmgpu_program *p, cvmgpu_r1cs *r, const uint8_t *inputs, uint64_t B,
                                            uint8_t *wtns_out, uint32_t *status, uint32_t *first_bad) {
    if (!p) return fail(CVMGPU_ERR_ARG, "null argument");
    return cvmgpu_witness_batch_select(p, r, inputs, B, 0, p->tape.n_wires, wtns_out, status, first_bad);
}

extern "C" int cvmgpu_witness_batch(cvmgpu_program *p, const uint8_t *inputs, uint64_t B, uint8_t *wtns_out, uint32_t *status) {
    return cvmgpu_witness_batch_checked(p, nullptr, inputs, B, wtns_out, status, nullptr);
}

extern "C" int cvmgpu_wtns_write(const char *path, const uint8_t *witness, uint32_t n_wires) {
    if (!path || !witness) return fail(CVMGPU_ERR_ARG, "null argument");
    FILE *f = fopen(path, "wb");
    if (!f) return fail(CVMGPU_ERR_IO, std::string("cannot create ") + path);
    // common/main.cpp:286-332
    static const uint8_t qle[32] = {0x01, 0x00, 0x00, 0xf0, 0x93, 0xf5, 0xe1, 0x43, 0x91, 0x70, 0xb9, 0x79, 0x48, 0xe8, 0x33, 0x28,
                                    0x5d, 0x58, 0x81, 0x81, 0xb6, 0x45, 0x50, 0xb8, 0x29, 0xa0, 0x31, 0xe1, 0x72, 0x4e, 0x64, 0x30};
    uint32_t version = 2, nsec = 2, id1 = 1, n8 = 32, id2 = 2;
    uint64_t len1 = 8 + n8, len2 = (uint64_t)n8 * n_wires;
    bool ok = fwrite("wtns", 4, 1, f) == 1 && fwrite(&version, 4, 1, f) == 1 && fwrite(&nsec, 4, 1, f) == 1 &&
              fwrite(&id1, 4, 1, f) == 1 && fwrite(&len1, 8, 1, f) == 1 && fwrite(&n8, 4, 1, f) == 1 &&
              fwrite(qle, 32, 1, f) == 1 && fwrite(&n_wires, 4, 1, f) == 1 && fwrite(&id2, 4, 1, f) == 1 &&
              fwrite(&len2, 8, 1, f) == 1 && (n_wires == 0 || fwrite(witness, (size_t)len2, 1, f) == 1);
    fclose(f);
    return ok ? CVMGPU_OK : fail(CVMGPU_ERR_IO, "short write");
}

// ------------------------------------------------------------------------------------------ r1cs
extern "C" int cvmgpu_r1cs_load(const char *path, cvmgpu_r1cs **out) {
    if (!path || !out) return fail(CVMGPU_ERR_ARG, "null argument");
    std::unique_ptr<cvmgpu_r1cs> r(new cvmgpu_r1cs());
    try {
        r->file = r1cs::load(path);
    } catch (const r1cs::Error &e) {
        std::string m = e.what();
        return fail(m.rfind("cannot open", 0) == 0 ? CVMGPU_ERR_IO : CVMGPU_ERR_PARSE, m);
    } catch (const std::exception &e) {
        return fail(CVMGPU_ERR_PARSE, e.what());
    }
    *out = r.release();
    return CVMGPU_OK;
}

extern "C" int cvmgpu_r1cs_info_get(const cvmgpu_r1cs *r, cvmgpu_r1cs_info *out) {
    if (!r || !out) return fail(CVMGPU_ERR_ARG, "null argument");
    cvmgpu_r1cs_info v;
    cvmgpu_r1cs_info *info = &v;
    memset(info, 0, sizeof(*info));
    info->n_wires = r->file.n_wires;
    info->n_pub_out = r->file.n_pub_out;
    info->n_pub_in = r->file.n_pub_in;
    info->n_prv_in = r->file.n_prv_in;
    info->n_constraints = r->file.n_constraints;
    info->n_labels = r->file.n_labels;
    info->nnz = r->file.nnz;
    info->nnz_pm1 = r->file.nnz_pm1;
    info->nnz_small = r->file.nnz_small;
    info->macs = r->file.macs;
    info->n_quadratic = r->file.n_quadratic;
    info->nnz_const = r->file.nnz_const;
    info->n_squares = r->file.n_squares;
    static_assert(r1cs::SAME_AS_A == R1CS_SAME_AS_A, "marker mismatch");
    static_assert(r1cs::LOC_BIT == tape::ROW_BIT, "row type bit mismatch");
    info->n_coefs = (uint32_t)r->file.coefs.size();
    // the binding to the last program layout this handle checked (cvmgpu_r1cs_check_store_dev / witness_batch_checked)
    info->bound_int_constraints = r->last_int;
    info->bound_bit_terms = r->last_bterms;
    info->bound_field_terms = r->last_fterms;
    info->bound_macs = r->last_macs;
    info->bound_bit_adds = r->last_bit_adds;
    info->bound_table_constraints = r->last_tcons;
    return fill_info(out, v);
}

extern "C" int cvmgpu_r1cs_bind_info(const cvmgpu_r1cs *r, const cvmgpu_program *p, cvmgpu_r1cs_info *out) {
    if (!r || !p || !out) return fail(CVMGPU_ERR_ARG, "null argument");
    if (r->file.n_wires != p->tape.n_wires) return fail(CVMGPU_ERR_ARG, "r1cs and program disagree on the number of wires");
    cvmgpu_r1cs_info v;
    v.struct_size = sizeof(v);
    if (int rc = cvmgpu_r1cs_info_get(r, &v)) return rc;
    const r1cs::Bound b = r1cs::bind(r->file, p->tape.wire_loc.data(), p->tape.one_brow, p->tape.const_rows);
    v.bound_int_constraints = b.n_int_constraints;
    v.bound_bit_terms = b.bterms.size();
    v.bound_field_terms = b.fterms.size();
    v.bound_macs = b.macs;
    v.bound_bit_adds = b.bit_adds;
    v.bound_table_constraints = b.n_table_constraints;
    return fill_info(out, v);
}

extern "C" void cvmgpu_r1cs_free(cvmgpu_r1cs *r) {
    if (!r) return;
    for (auto &kv : r->dev) on_device(kv.first, [&] { kv.second.release(); });
    delete r;
}

static int upload_bound(const r1cs::Bound &b, BoundDev &d) {
    if (int rc = d.d_hdr.ensure(b.hdr.size() * 4)) return rc;
    if (int rc = d.d_terms.ensure(std::max<size_t>(16, b.fterms.size() * 8))) return rc;
    if (int rc = d.d_bhdr.ensure(b.bhdr.size() * 4)) return rc;
    if (int rc = d.d_bterms.ensure(std::max<size_t>(16, b.bterms.size() * 8))) return rc;
    CUDA_TRY(cudaMemcpy(d.d_hdr.p, b.hdr.data(), b.hdr.size() * 4, cudaMemcpyHostToDevice));
    if (!b.fterms.empty()) CUDA_TRY(cudaMemcpy(d.d_terms.p, b.fterms.data(), b.fterms.size() * 8, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d.d_bhdr.p, b.bhdr.data(), b.bhdr.size() * 4, cudaMemcpyHostToDevice));
    if (!b.bterms.empty()) CUDA_TRY(cudaMemcpy(d.d_bterms.p, b.bterms.data(), b.bterms.size() * 8, cudaMemcpyHostToDevice));
    if (int rc = d.d_tcons.ensure(std::max<size_t>(32, b.tcons.size() * 4))) return rc;
    if (!b.tcons.empty()) CUDA_TRY(cudaMemcpy(d.d_tcons.p, b.tcons.data(), b.tcons.size() * 4, cudaMemcpyHostToDevice));
    d.n_tcons = b.n_table_constraints;
    if (int rc = d.d_active.ensure(std::max<size_t>(16, b.active.size() * 4))) return rc;
    if (!b.active.empty()) CUDA_TRY(cudaMemcpy(d.d_active.p, b.active.data(), b.active.size() * 4, cudaMemcpyHostToDevice));
    d.n_active = b.active.size();
    d.n_shift = b.n_shift_constraints;
    if (b.n_shift_constraints) {
        if (int rc = d.d_shh.ensure(b.shh.size() * 4)) return rc;
        if (int rc = d.d_shl.ensure((b.shl.size() + 32) * 4)) return rc;   // + one layer: the kernel fetches one ahead
        if (int rc = d.d_shm.ensure(b.shm.size() * 4)) return rc;
        CUDA_TRY(cudaMemcpy(d.d_shh.p, b.shh.data(), b.shh.size() * 4, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemset((char *)d.d_shl.p + b.shl.size() * 4, 0xff, 32 * 4));
        CUDA_TRY(cudaMemcpy(d.d_shl.p, b.shl.data(), b.shl.size() * 4, cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(d.d_shm.p, b.shm.data(), b.shm.size() * 4, cudaMemcpyHostToDevice));
    }
    d.macs = b.macs;
    d.bit_adds = b.bit_adds;
    d.n_int_constraints = b.n_int_constraints;
    d.n_bterms = b.bterms.size();
    d.n_fterms = b.fterms.size();
    CUDA_TRY(cudaDeviceSynchronize());
    return CVMGPU_OK;
}

// the coefficient tables on the current device (uploaded on first use there)
static int upload_r1cs(cvmgpu_r1cs *r, cvmgpu_r1cs::Dev **out) {
    int dev = -1;
    CUDA_TRY(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(r->mu);
    cvmgpu_r1cs::Dev &d = r->dev[dev];
    *out = &d;
    if (d.ready) return CVMGPU_OK;
    const r1cs::File &f = r->file;
    std::vector<fr::Fr> cm;
    cm.reserve(f.coefs.size());
    for (const fr::Fr &c : f.coefs) cm.push_back(fr::to_mont(c));
    if (int rc = d.d_coefs.ensure(cm.size() * 32)) return rc;
    if (int rc = d.d_cmag.ensure(f.cmag.size() * 4)) return rc;
    if (int rc = d.d_cint.ensure(f.cint.size() * 8)) return rc;
    CUDA_TRY(cudaMemcpy(d.d_cmag.p, f.cmag.data(), f.cmag.size() * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d.d_cint.p, f.cint.data(), f.cint.size() * 8, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d.d_coefs.p, cm.data(), cm.size() * 32, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaDeviceSynchronize());
    d.ready = true;
    return CVMGPU_OK;
}

// bd: the CSR bound to the layout of d_store; bits / n_brows: its bit rows (nullptr / 0 for the plain layout)
static int launch_check(cvmgpu_r1cs *r, const cvmgpu_r1cs::Dev &rd, const BoundDev &bd, const void *d_store, const uint32_t *bits, uint64_t B,
                        uint64_t bstride, void *d_first_bad, cudaStream_t s) {
    CUDA_TRY(cudaMemsetAsync(d_first_bad, 0xff, B * 4, s));
    if (r->file.n_constraints == 0) return CVMGPU_OK;
    const bool walk = !bd.typed || bd.n_active != 0;
    uint64_t gx = (B + R1CS_NT - 1) / R1CS_NT;
    // the per-witness kernel walks every constraint (plain layout) or those the table kernel does not take (typed)
    const uint64_t n_walk = bd.typed ? bd.n_active : r->file.n_constraints;
    // enough CTAs to fill 148 SMs several times over even for small batches (config 5: B = 1K, 1.5M constraints)
    uint64_t want = 148ull * 16;
    uint64_t chunks = std::max<uint64_t>(1, (want + gx - 1) / gx);
    {
        // ... in a whole number of waves: 512 x 5 CTAs at 5 resident per SM are 3.46 waves, i.e. four rounds of which the last
        // is half empty; 512 x 13 are 8.99
        const uint64_t slots = 148ull * 5;
        double best_eff = 0;
        uint64_t best_c = chunks;
        for (uint64_t c = chunks; c <= 4 * chunks && c * 32 <= std::max<uint64_t>(n_walk, 32); c++) {
            const double waves = (double)(gx * c) / (double)slots;
            const double eff = waves / std::ceil(waves);
            if (eff > best_eff + 0.02) { best_eff = eff; best_c = c; }
        }
        chunks = best_c;
    }
    uint32_t per = (uint32_t)std::max<uint64_t>(32, (n_walk + chunks - 1) / chunks);
    chunks = (n_walk + per - 1) / per;
    if (chunks > 65535) {
        per = (uint32_t)((n_walk + 65534) / 65535);
        chunks = (n_walk + per - 1) / per;
    }
    kern::R1csParams rp;
    rp.hdr = (const uint4 *)bd.d_hdr.p;
    rp.terms = (const uint2 *)bd.d_terms.p;
    rp.coefs = (const uint4 *)rd.d_coefs.p;
    rp.cmag = (const uint32_t *)rd.d_cmag.p;
    rp.n_cons = (uint32_t)n_walk;
    rp.n_all = r->file.n_constraints;
    rp.active = (const uint32_t *)bd.d_active.p;
    rp.cons_per_chunk = per;
    rp.store = (const uint4 *)d_store;
    rp.bstride = bstride;
    rp.B = B;
    rp.first_bad = (uint32_t *)d_first_bad;
    rp.bits = bits;
    rp.n_brows = bd.n_brows;
    rp.bhdr = (const uint4 *)bd.d_bhdr.p;
    rp.bterms = (const uint2 *)bd.d_bterms.p;
    rp.cint = (const long long *)rd.d_cint.p;
    dim3 grid((unsigned)gx, (unsigned)chunks);
    // resident CTAs per SM: 4 (128 registers) when the check is multiplier-bound (Poseidon: 28 % of the terms have
    // full-size coefficients, EdDSA 12 %), 5 (96 registers) when it is mostly +-1 / small coefficients and latency-bound
    int minb = g_r1cs_minb ? g_r1cs_minb : (20 * r->file.nnz_general > r->file.nnz ? 4 : 5);
    if (!walk) {
        // every constraint is a truth-table constraint
    } else if (bd.typed) {
        if (minb == 5) kern::r1cs_kernel<5, true><<<grid, R1CS_NT, 0, s>>>(rp);
        else kern::r1cs_kernel<4, true><<<grid, R1CS_NT, 0, s>>>(rp);
    } else {
        if (minb == 5) kern::r1cs_kernel<5, false><<<grid, R1CS_NT, 0, s>>>(rp);
        else kern::r1cs_kernel<4, false><<<grid, R1CS_NT, 0, s>>>(rp);
    }
    CUDA_TRY(cudaGetLastError());
    if (bd.typed && bd.n_tcons) {
        // the constraints that are boolean predicates of a few 0/1 wires: bitwise, 32 witnesses per lane
        kern::R1csTableParams tp;
        tp.tcons = (const uint4 *)bd.d_tcons.p;
        tp.n_tcons = (uint32_t)bd.n_tcons;
        const uint64_t tgx = (B + 127) / 128;
        uint64_t tchunks = std::max<uint64_t>(1, (148ull * 16 + tgx - 1) / tgx);
        uint32_t tper = (uint32_t)std::max<uint64_t>(32, ((bd.n_tcons + tchunks - 1) / tchunks + 31) / 32 * 32);
        tchunks = std::min<uint64_t>(65535, (bd.n_tcons + tper - 1) / tper);
        tper = (uint32_t)(((bd.n_tcons + tchunks - 1) / tchunks + 31) / 32 * 32);
        tp.per_chunk = tper;
        tp.bits = bits;
        tp.n_brows = bd.n_brows;
        tp.B = B;
        tp.first_bad = (uint32_t *)d_first_bad;
        kern::r1cs_table_kernel<<<dim3((unsigned)tgx, (unsigned)((bd.n_tcons + tper - 1) / tper)), 128, 0, s>>>(tp);
        CUDA_TRY(cudaGetLastError());
    }
    if (bd.typed && bd.n_shift) {
        // the linear constraints over 0/1 wires with +-2^k coefficients: bit-matrix transposes, 32 witnesses per warp
        kern::R1csShiftParams sp;
        sp.shh = (const uint4 *)bd.d_shh.p;
        sp.shl = (const uint32_t *)bd.d_shl.p;
        sp.shm = (const uint32_t *)bd.d_shm.p;
        sp.n_cons = (uint32_t)bd.n_shift;
        const uint64_t sgx = (B + 127) / 128;
        uint64_t schunks = std::max<uint64_t>(1, std::min<uint64_t>(bd.n_shift, (148ull * 16 + sgx - 1) / sgx));
        sp.per_chunk = (uint32_t)((bd.n_shift + schunks - 1) / schunks);
        schunks = (bd.n_shift + sp.per_chunk - 1) / sp.per_chunk;
        sp.bits = bits;
        sp.n_brows = bd.n_brows;
        sp.B = B;
        sp.first_bad = (uint32_t *)d_first_bad;
        kern::r1cs_shift_kernel<<<dim3((unsigned)sgx, (unsigned)std::min<uint64_t>(65535, schunks)), 128, 0, s>>>(sp);
        CUDA_TRY(cudaGetLastError());
    }
    return CVMGPU_OK;
}

extern "C" int cvmgpu_r1cs_check_dev(cvmgpu_r1cs *r, const void *d_store, uint64_t B, uint64_t bstride, void *d_first_bad,
                                     void *stream) {
    if (!r || !d_store || !d_first_bad) return fail(CVMGPU_ERR_ARG, "null argument");
    if (B == 0) return CVMGPU_OK;
    if (bstride < B) return fail(CVMGPU_ERR_ARG, "bstride < B");
    if (bstride >> 27) return fail(CVMGPU_ERR_ARG, "bstride must be below 2^27 witnesses (32-bit row stride)");
    cvmgpu_r1cs::Dev *rd = nullptr;
    if (int rc = upload_r1cs(r, &rd)) return rc;
    {
        std::lock_guard<std::mutex> lock(r->mu);
        if (rd->plain.layout_id != 0) {
            if (int rc = upload_bound(r1cs::bind(r->file, nullptr), rd->plain)) return rc;
            rd->plain.layout_id = 0;
            rd->plain.typed = false;
        }
    }
    return launch_check(r, *rd, rd->plain, d_store, nullptr, B, bstride, d_first_bad, (cudaStream_t)stream);
}

extern "C" int cvmgpu_r1cs_check_store_dev(cvmgpu_r1cs *r, cvmgpu_program *p, const void *d_store, uint64_t B, uint64_t bstride,
                                           void *d_first_bad, void *stream) {
    if (!r || !p || !d_store || !d_first_bad) return fail(CVMGPU_ERR_ARG, "null argument");
    if (r->file.n_wires != p->tape.n_wires) return fail(CVMGPU_ERR_ARG, "r1cs and program disagree on the number of wires");
    if (B == 0) return CVMGPU_OK;
    if (bstride < B) return fail(CVMGPU_ERR_ARG, "bstride < B");
    if (bstride >> 27) return fail(CVMGPU_ERR_ARG, "bstride must be below 2^27 witnesses (32-bit row stride)");
    cvmgpu_r1cs::Dev *rd = nullptr;
    if (int rc = upload_r1cs(r, &rd)) return rc;
    {
        std::lock_guard<std::mutex> lock(r->mu);
        if (rd->typed.layout_id != p->layout_id) {
            if (int rc = upload_bound(r1cs::bind(r->file, p->tape.wire_loc.data(), p->tape.one_brow, p->tape.const_rows), rd->typed)) return rc;
            rd->typed.layout_id = p->layout_id;
            rd->typed.typed = rd->typed.n_bterms != 0 || rd->typed.n_tcons != 0;   // no term on a bit row: the plain kernel runs on the field rows
            rd->typed.n_brows = p->tape.n_brows;
            r->last_macs = rd->typed.macs; r->last_bit_adds = rd->typed.bit_adds; r->last_int = rd->typed.n_int_constraints;
            r->last_bterms = rd->typed.n_bterms; r->last_fterms = rd->typed.n_fterms; r->last_tcons = rd->typed.n_tcons;
        }
    }
    const uint32_t *bits = (const uint32_t *)((const char *)d_store + store_field_bytes(p, bstride));
    return launch_check(r, *rd, rd->typed, d_store, bits, B, bstride, d_first_bad, (cudaStream_t)stream);
}

extern "C" int cvmgpu_witness_import_dev(uint32_t n_wires, const void *d_wtns, uint64_t B, uint64_t bstride, void *d_store,
                                         void *stream) {
    if (!d_wtns || !d_store) return fail(CVMGPU_ERR_ARG, "null argument");
    if (B == 0 || n_wires == 0) return CVMGPU_OK;
    dim3 grid((unsigned)((B + 31) / 32), (n_wires + 31) / 32);
    kern::import_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const uint4 *)d_wtns, B, n_wires, (uint4 *)d_store, bstride);
    CUDA_TRY(cudaGetLastError());
    return CVMGPU_OK;
}

extern "C" int cvmgpu_r1cs_check(cvmgpu_r1cs *r, const uint8_t *witnesses, uint64_t B, uint32_t *first_bad) {
    if (!r || !witnesses || !first_bad) return fail(CVMGPU_ERR_ARG, "null argument");
    if (cvmgpu_device_count() <= 0) return fail(CVMGPU_ERR_CUDA, "no CUDA device available (there is no CPU fallback)");
    if (B == 0) return CVMGPU_OK;
    cvmgpu_r1cs::Dev *rd = nullptr;
    if (int rc = upload_r1cs(r, &rd)) return rc;
    const size_t row = (size_t)r->file.n_wires * 32;
    uint64_t chunk = std::min<uint64_t>(pick_chunk(B, 2 * row + 4), 1u << 22);
    if (chunk == 0) return fail(CVMGPU_ERR_CUDA, "cudaMemGetInfo failed");
    if (int rc = rd->d_store.ensure((size_t)r->file.n_wires * 32 * chunk)) return rc;
    if (int rc = rd->d_wtns.ensure(row * chunk)) return rc;
    if (int rc = rd->d_bad.ensure(4 * chunk)) return rc;
    cudaStream_t s = 0;
    for (uint64_t b0 = 0; b0 < B; b0 += chunk) {
        uint64_t n = std::min<uint64_t>(chunk, B - b0);
        CUDA_TRY(cudaMemcpyAsync(rd->d_wtns.p, witnesses + b0 * row, n * row, cudaMemcpyHostToDevice, s));
        if (int rc = cvmgpu_witness_import_dev(r->file.n_wires, rd->d_wtns.p, n, chunk, rd->d_store.p, s)) return rc;
        if (int rc = cvmgpu_r1cs_check_dev(r, rd->d_store.p, n, chunk, rd->d_bad.p, s)) return rc;
        CUDA_TRY(cudaMemcpyAsync(first_bad + b0, rd->d_bad.p, n * 4, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
    }
    return CVMGPU_OK;
}

// ------------------------------------------------------------------------------------------ field hooks
extern "C" int cvmgpu_fr_host_op(const char *op, const uint8_t *a, const uint8_t *b, uint8_t *out) {
    if (!op || !a || !out) return fail(CVMGPU_ERR_ARG, "null argument");
    hostfr::FfOp f = hostfr::op_from_name(op);
    if (f == hostfr::F_NONE) return fail(CVMGPU_ERR_ARG, std::string("unknown field op ") + op);
    fr::Fr x, y = fr::zero(), r = fr::zero();
    memcpy(x.v, a, 32);
    if (b) memcpy(y.v, b, 32);
    bool ok = hostfr::apply(f, x, y, r);
    memcpy(out, r.v, 32);
    return ok ? 0 : 1;
}

extern "C" int cvmgpu_fr_device_op(const char *op, const uint8_t *a, const uint8_t *b, uint8_t *out, uint64_t n) {
    if (!op || !a || !b || !out) return fail(CVMGPU_ERR_ARG, "null argument");
    hostfr::FfOp f = hostfr::op_from_name(op);
    if (f == hostfr::F_NONE) return fail(CVMGPU_ERR_ARG, std::string("unknown field op ") + op);
    if (cvmgpu_device_count() <= 0) return fail(CVMGPU_ERR_CUDA, "no CUDA device available (there is no CPU fallback)");
    if (n == 0) return CVMGPU_OK;
    DevBuf da, db, dout, dund;
    struct Guard {
        DevBuf &a, &b, &c, &d;
        ~Guard() { a.release(); b.release(); c.release(); d.release(); }
    } guard{da, db, dout, dund};
    if (int rc = da.ensure(n * 32)) return rc;
    if (int rc = db.ensure(n * 32)) return rc;
    if (int rc = dout.ensure(n * 32)) return rc;
    if (int rc = dund.ensure(n * 4)) return rc;
    CUDA_TRY(cudaMemcpy(da.p, a, n * 32, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(db.p, b, n * 32, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemset(dund.p, 0, n * 4));
    kern::fr_op_kernel<<<(unsigned)((n + 127) / 128), 128>>>((int)f, (const uint4 *)da.p, (const uint4 *)db.p, (uint4 *)dout.p, n,
                                                            (uint32_t *)dund.p);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy(out, dout.p, n * 32, cudaMemcpyDeviceToHost));
    return CVMGPU_OK;
}

extern "C" int cvmgpu_mul_peak(int variant, int ctas_per_sm, double *muls_per_second) {
    if (cvmgpu_device_count() <= 0) return fail(CVMGPU_ERR_CUDA, "no CUDA device available");
    uint4 *d = nullptr;
    CUDA_TRY(cudaMalloc(&d, 32));
    cudaDeviceProp prop;
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
    const uint32_t iters = 2048;
    unsigned blocks = (unsigned)prop.multiProcessorCount * (unsigned)std::max(1, ctas_per_sm);
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        CUDA_TRY(cudaEventRecord(e0));
        if (variant == 0) kern::mulbench_kernel<0><<<blocks, 128>>>(d, iters, 7u + rep);
        else if (variant == 2) kern::mulbench_kernel<2><<<blocks, 128>>>(d, iters, 7u + rep);
        else if (variant == 3) kern::mulbench_kernel<3><<<blocks, 128>>>(d, iters, 7u + rep);
        else if (variant == 4) kern::mulbench_kernel<4><<<blocks, 128>>>(d, iters, 7u + rep);
        else kern::mulbench_kernel<1><<<blocks, 128>>>(d, iters, 7u + rep);
        CUDA_TRY(cudaEventRecord(e1));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    CUDA_TRY(cudaGetLastError());
    if (muls_per_second) *muls_per_second = (double)blocks * 128.0 * iters * 2.0 / (best * 1e-3);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d);
    return CVMGPU_OK;
}

extern "C" int cvmgpu_imad_peak(int kind, double *macs_per_second, double *ms_out) {
    if (cvmgpu_device_count() <= 0) return fail(CVMGPU_ERR_CUDA, "no CUDA device available");
    uint32_t *d = nullptr;
    CUDA_TRY(cudaMalloc(&d, 4));
    cudaDeviceProp prop;
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
    const uint32_t iters = 4096;
    unsigned blocks = (unsigned)prop.multiProcessorCount * 8;
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        CUDA_TRY(cudaEventRecord(e0));
        switch (kind) {
            case 0: kern::imad_kernel<0><<<blocks, 256>>>(d, iters, 12345u + rep); break;
            case 1: kern::imad_kernel<1><<<blocks, 256>>>(d, iters, 12345u + rep); break;
            case 2: kern::imad_kernel<2><<<blocks, 256>>>(d, iters, 12345u + rep); break;
            case 3: kern::imad_kernel<3><<<blocks, 256>>>(d, iters, 12345u + rep); break;
            case 4: kern::imad_kernel<4><<<blocks, 256>>>(d, iters, 12345u + rep); break;
            case 5: kern::imad_kernel<5><<<blocks, 256>>>(d, iters, 12345u + rep); break;
            case 7: kern::imad_kernel<7><<<blocks, 256>>>(d, iters, 12345u + rep); break;
            case 8: kern::imad_kernel<8><<<blocks, 256>>>(d, iters, 12345u + rep); break;
            default: kern::imad_kernel<6><<<blocks, 256>>>(d, iters, 12345u + rep); break;
        }
        CUDA_TRY(cudaEventRecord(e1));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    CUDA_TRY(cudaGetLastError());
    double macs = (double)blocks * 256.0 * iters * (kind == 8 ? 16.0 : 8.0);
    if (macs_per_second) *macs_per_second = macs / (best * 1e-3);
    if (ms_out) *ms_out = best;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d);
    return CVMGPU_OK;
}
