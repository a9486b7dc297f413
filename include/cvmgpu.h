/* cvmgpu -- C ABI of the B200 batched witness-generation + R1CS-check engine.
 *
 * This is the drop-in boundary for the reference's one-witness-per-process calculator.  Each entry
 * point names the reference interface it replaces (paths relative to the reference tree,
 * MarioCalvarro/circom_cvm):
 *
 *   reference                                                        this library
 *   ---------------------------------------------------------------  ---------------------------------
 *   loadCircuit(<circuit>.dat) + generated <circuit>.cpp             cvmgpu_program_load(<circuit>.cvm)
 *     code_producers/src/c_elements/common/main.cpp:22-124             (the --cvm output of the same compile,
 *     compiler/src/circuit_design/circuit.rs:424-567                    circuit.rs:577-621)
 *   Circom_CalcWit::setInputSignal / tryRunCircuit -> run(ctx)       cvmgpu_witness_batch[_dev]
 *     common/calcwit.cpp:71-97, generated T_run bodies                 (B independent inputs at once)
 *   writeBinWitness(ctx, out.wtns)  common/main.cpp:286-332          cvmgpu_wtns_write / witness_batch output
 *   R1CSWriter (the file our checker consumes)                       cvmgpu_r1cs_load
 *     constraint_writers/src/r1cs_writer.rs:155-341
 *   `===` asserts compiled into the program (translate.rs:676-734)   cvmgpu_r1cs_check[_dev]  (the reference has no
 *     -- the reference's only satisfiability check                      stand-alone checker; semantics = file format:
 *                                                                      A*B - C = 0, constraints-json.md:17)
 *   extern "C" Fr_* (bn128/fr.hpp:28-70)                             device code (csrc/fr.cuh); cvmgpu_fr_host_op
 *                                                                      exposes the same limb routines on the host
 *
 * Conventions: every function returns 0 on success and a negative code on failure (never aborts --
 * the reference exits through assert()); cvmgpu_last_error() returns a thread-local message.
 * Field elements cross the boundary as 32-byte little-endian CANONICAL integers (the .wtns / .r1cs
 * encoding).  The caller owns every buffer; handles are opaque and released with *_free.
 * Host code can be Rust (`extern "C"`), C++, or Python ctypes; see INTEGRATION.md.
 */
#ifndef CVMGPU_H
#define CVMGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CVMGPU_OK 0
#define CVMGPU_ERR_IO (-1)          /* file missing / unreadable */
#define CVMGPU_ERR_PARSE (-2)       /* malformed .cvm / .r1cs */
#define CVMGPU_ERR_UNSUPPORTED (-3) /* program uses a feature the trace compiler rejects */
#define CVMGPU_ERR_CUDA (-4)        /* no device / CUDA failure: there is NO CPU fallback */
#define CVMGPU_ERR_ARG (-5)

/* per-witness status words (the reference aborts the process in these cases) */
#define CVMGPU_ST_OK 0
#define CVMGPU_ST_ASSERT 1   /* failed assert / `===`   (assert_bucket.rs:71-86) */
#define CVMGPU_ST_TOINT 2    /* Fr_toInt of a value outside [-2^31, 2^31): a data-dependent array index (generic/fr.cpp:1102-1170) */
#define CVMGPU_ST_DIVZERO 3  /* `\` or `%` by zero       (GMP division by zero in the reference) */
#define CVMGPU_ST_SPECULATION 6 /* only from a speculative program (cvmgpu_program_speculative): a main input is not 0 / 1 --
                                   recompute this witness with the program itself.  The host-buffer calls do that. */
#define CVMGPU_ST_LOOP 5     /* a data-dependent while loop needed more iterations than were traced (260) */

typedef struct cvmgpu_program cvmgpu_program;
typedef struct cvmgpu_r1cs cvmgpu_r1cs;

/* Info structs are versioned by size: set info.struct_size = sizeof(info) before the call; the library fills
 * min(struct_size, its own sizeof) bytes and writes back how many it filled.  New counters are only ever appended. */
typedef struct {
    uint32_t struct_size;        /* in: sizeof(cvmgpu_program_info) of the caller; out: bytes filled */
    uint32_t reserved0;
    uint64_t n_signals;          /* %%signals */
    uint32_t n_wires;            /* length of %%witness */
    uint32_t n_inputs;           /* main input signals (field elements) */
    uint32_t n_outputs;          /* main output signals */
    uint32_t n_slots;            /* on-chip slots per witness chosen for the tape */
    uint32_t n_rows;             /* rows of the device value store: n_frows + n_brows */
    uint64_t tape_len;           /* tape instructions executed per witness */
    uint64_t ref_mul;            /* N_mul: ff.mul (+1 per ff.div) the reference program executes per witness */
    uint64_t ref_field_ops;      /* all ff.* operations the reference program executes per witness */
    uint64_t cvm_instructions;   /* CVM instructions replayed on the host while tracing */
    uint64_t tape_mul, tape_div, tape_addsub, tape_other, tape_ld, tape_st, tape_spill_st;
    uint32_t n_consts;
    uint32_t dyn_branches;
    uint64_t ref_div;            /* ff.div the reference program executes per witness (Fr_div = mpz_invert + Fr_mul) */
    uint64_t tape_inv;           /* field inversions left on the tape after batching independent ones (Montgomery's trick) */
    uint64_t tape_sel;           /* selects on the tape (products with a 0/1 factor, if-converted branches, batch inversion) */
    uint64_t tape_dot;           /* fused dot products sum c_k*x_k (one Montgomery reduction each) */
    uint64_t tape_dot_terms;     /* their terms (64 multiply-accumulates each instead of 136) */
    uint64_t tape_macs;          /* 32x32->64 multiply-accumulates the tape kernel executes per witness */
    uint64_t tape_ld_streamed;   /* reloads served by the cp.async ring (requested 4 reloads ahead) */
    uint64_t unrolled_iterations; /* iterations of data-dependent while loops traced under predicates */
    uint64_t tape_lut;           /* boolean cones (XOR / Ch / Maj / ... written as field polynomials) run as one table look-up */
    uint64_t tape_ld_bool;       /* of tape_ld: reloads of values typed 0/1 (one word per warp from a bit row) */
    uint64_t tape_spill_st_bool; /* of tape_spill_st: spills of values typed 0/1 */
    uint64_t n_bool_wires;       /* witness wires typed 0/1: stored as bit rows */
    uint32_t n_bslots;           /* bit slots per warp (on-chip file of the values typed 0/1) */
    uint32_t n_frows;            /* field rows of the value store (field-typed wires + field spill rows), 32 B per witness */
    uint32_t n_brows;            /* bit rows (wires typed 0/1 + bit spill rows), one 32-bit word per 32 witnesses */
    uint32_t max_live_field;     /* most field-typed values live at once */
    uint32_t max_live_bool;      /* most 0/1-typed values live at once */
    uint32_t reserved1;
    uint64_t tape_int;           /* small-integer operations: bit-weighted sums kept as raw 64-bit integers instead of field values */
} cvmgpu_program_info;

typedef struct {
    uint32_t struct_size;        /* in: sizeof(cvmgpu_r1cs_info) of the caller; out: bytes filled */
    uint32_t n_wires, n_pub_out, n_pub_in, n_prv_in, n_constraints;
    uint64_t n_labels;
    uint64_t nnz;                /* non-zeros of A, B and C together */
    uint64_t nnz_pm1;            /* of which coefficient +1 or -1 */
    uint32_t n_coefs;            /* distinct coefficients (interned) */
    uint64_t nnz_small;          /* non-zeros evaluated on the small-coefficient path (|c| < 2^32, 8 MACs per term) */
    uint64_t macs;               /* 32x32->64 multiply-accumulates the check kernel executes per witness */
    uint64_t n_quadratic;        /* constraints with non-empty A and B (one Montgomery product each) */
    uint64_t nnz_const;          /* general coefficients on wire 0 (the constant 1): added, not multiplied */
    uint64_t n_squares;          /* quadratic constraints whose B repeats A: evaluated once and squared */
    /* the binding to the typed store of the last program this handle checked (0 before the first such check) */
    uint64_t bound_int_constraints; /* constraints over 0/1 wires with small coefficients: evaluated in 64-bit integers */
    uint64_t bound_bit_terms;       /* non-zeros on bit rows */
    uint64_t bound_field_terms;     /* non-zeros on field rows (the operand stream) */
    uint64_t bound_macs;            /* multiply-accumulates per witness on that layout (upper bound: 0 / +-1 factors skip the product) */
    uint64_t bound_bit_adds;        /* field additions of bit-row terms outside integer constraints */
    uint64_t bound_table_constraints; /* of bound_int_constraints: boolean predicates of <= 5 wires, evaluated bitwise for 32 witnesses at a time */
} cvmgpu_r1cs_info;

const char *cvmgpu_last_error(void);
int cvmgpu_device_count(void);
int cvmgpu_set_device(int device);
/* reserved (was: witnesses per thread of the tape kernel); accepted and ignored */
int cvmgpu_set_tape_mode(int mode);
/* frees the calling thread's pipeline buffers of the host-buffer entry points (also done by cvmgpu_program_free) */
void cvmgpu_release_buffers(void);

/* ---- program ---------------------------------------------------------------------------------- */
/* Parse + trace-compile a .cvm file.  n_slots = 0 picks the default.  Works without a GPU. */
int cvmgpu_program_load(const char *cvm_path, uint32_t n_slots, cvmgpu_program **out);
int cvmgpu_program_load_text(const char *cvm_text, size_t len, uint32_t n_slots, cvmgpu_program **out);
/* The fork's --cvm emitter prints nothing for component creation (create_component_bucket.rs:356-360).  Either compile
 * with patches/create_component_bucket.rs.diff (adds one `;;%%create_cmp` comment line per bucket), or pass the generated
 * <circuit>.cpp of the same compile here: the `<Sub>_create(...)` blocks of its `_run` bodies
 * (create_component_bucket.rs:206-354) supply the same information.  cpp_path / cpp_text may be NULL. */
int cvmgpu_program_load_with_cpp(const char *cvm_path, const char *cpp_path, uint32_t n_slots, cvmgpu_program **out);
int cvmgpu_program_load_text2(const char *cvm_text, size_t len, const char *cpp_text, size_t cpp_len, uint32_t n_slots,
                              cvmgpu_program **out);
/* Circuits with MIXED component arrays (an array whose positions hold different template instances, e.g. circomlib's
 * Poseidon) address the signals of those components through the io-map ("mapped" locations, location_rule.rs:86-171;
 * C++ twin templateInsId2IOSignalInfo, load_bucket.rs:262-318).  The .cvm file does not carry it (circuit.rs:577-621).
 * It is taken from the `;;%%io_map` comment lines of a patched emitter, or from the <circuit>.dat of the same compile
 * (c_code_generator.rs:617-674; reader main.cpp:59-92) -- whose section sizes are only in the generated C++
 * (`get_size_of_*`, circuit.rs:481-497), so dat requires cpp.  cpp / dat may be NULL. */
int cvmgpu_program_load_files(const char *cvm_path, const char *cpp_path, const char *dat_path, uint32_t n_slots,
                              cvmgpu_program **out);
int cvmgpu_program_load_text3(const char *cvm_text, size_t len, const char *cpp_text, size_t cpp_len, const void *dat,
                              size_t dat_len, uint32_t n_slots, cvmgpu_program **out);
/* Speculative typing.  Hash circuits take their message as unconstrained signals (nothing in Sha256(n) proves in[k] a
 * bit), so what is derived from the message before the first bit decomposition is field arithmetic on values that are
 * 0 / 1 in every sensible input.  Bit-heavy programs are traced a second time under the assumption that EVERY main input
 * is literally 0 or 1; that tape checks the assumption per witness and sets CVMGPU_ST_SPECULATION where it fails.  The
 * host-buffer entry points (cvmgpu_witness_batch*, _select, _multi) run it and recompute the flagged witnesses with the
 * general tape: results are identical for every input.  Device-API callers opt in: *spec (owned by p; NULL when p has
 * none) is a program handle with its own value-store layout for cvmgpu_witness_batch_dev / _export_dev /
 * cvmgpu_r1cs_check_store_dev / cvmgpu_store_bytes / cvmgpu_program_info_get; witnesses it flags must be redone with p.
 * CVMGPU_SPECULATE=0 in the environment: never build one. */
int cvmgpu_program_speculative(cvmgpu_program *p, cvmgpu_program **spec);
/* The same programs take their inputs as PACKED BITS: row = ceil(n_inputs / 8) bytes per witness, input k = bit (k & 7) of
 * byte (k >> 3).  One 32-byte field element per message bit is what the reference's input.json amounts to; for a batch
 * it is 256 times the bytes (Sha256(512) x 65 536: 1 GiB against 4 MiB, 19 ms of PCIe per call).  Outputs as
 * cvmgpu_witness_batch_select.  CVMGPU_ERR_UNSUPPORTED for programs without a bit-input tape. */
int cvmgpu_witness_batch_bits(cvmgpu_program *p, cvmgpu_r1cs *r, const uint8_t *input_bits, uint64_t B, uint32_t wire0,
                              uint32_t n_sel, uint8_t *wtns_out, uint32_t *status, uint32_t *first_bad);
/* PACKED witness rows: the whole witness in the types the trace compiler proved -- per witness the field wires as 32-byte
 * canonical values in wire order, then the 0/1 wires as bits in wire order (bit j & 31 of 32-bit word j >> 5, the words
 * following the field part, padded to a multiple of 16 bytes).  cvmgpu_packed_layout gives the row size, the two counts and
 * the wire -> typed-row map of the
 * same handle (entry & 0x80000000: a 0/1 wire).  Sha256(512): 8.6 KB per witness instead of 2.2 MB of 32-byte values.
 * bit_input_tape / inputs_are_bits != 0: the layout of the bit-input tape, inputs as packed bits (cvmgpu_witness_batch_bits);
 * 0: the layout of the program itself, inputs as 32-byte field elements. */
int cvmgpu_packed_layout(cvmgpu_program *p, int bit_input_tape, uint64_t *row_bytes, uint32_t *n_field, uint32_t *n_bits,
                         const uint32_t **wire_rows);
int cvmgpu_witness_batch_packed(cvmgpu_program *p, cvmgpu_r1cs *r, const uint8_t *inputs, int inputs_are_bits, uint64_t B,
                                uint8_t *packed_out, uint32_t *status, uint32_t *first_bad);
/* DEVICE: the packed rows of a value store written by program p (the handle whose layout applies) */
int cvmgpu_witness_export_packed_dev(cvmgpu_program *p, const void *d_store, uint64_t B, uint64_t bstride, void *d_out, void *stream);
int cvmgpu_program_info_get(const cvmgpu_program *p, cvmgpu_program_info *info);
void cvmgpu_program_free(cvmgpu_program *p);
/* Read-only view of the compiled tape (16-byte instructions, layout in csrc/tape.hpp) and of its constant table
 * (32-byte LE Montgomery values).  For inspection and for host-side tests of the trace compiler. */
int cvmgpu_program_tape(const cvmgpu_program *p, const void **ins, uint64_t *n_ins, const void **consts, uint32_t *n_consts);
/* constants of the tape's small-integer operations (raw 64-bit integers) */
int cvmgpu_program_iconsts(const cvmgpu_program *p, const uint64_t **iconsts, uint32_t *n);
/* The program's %%witness list: signal index of every witness wire (= the witness2SignalList of the .dat,
 * c_code_generator.rs:541-550; used to locate the input hash map of a .dat, circom_cvm_b200/inputs.py). */
int cvmgpu_program_witness(const cvmgpu_program *p, const uint64_t **signals, uint32_t *n);
/* The main component's input names, when the program text carries them: `;;%%main_input <name> <first signal> <size>` comment
 * lines (patches/main_input_directive.rs.diff; the fork's .cvm has no name table, the .dat hash map and the .sym file do).
 * text: one "name first_signal size" line per input, empty when there are none. */
int cvmgpu_program_main_inputs(const cvmgpu_program *p, const char **text, size_t *len);
/* Per witness wire: 1 when the trace compiler's value-range typing proves the wire 0/1 for every input (comparison
 * results, extracted bits, boolean combinations of those, constants 0 and 1).  Such wires are stored as bit rows. */
int cvmgpu_program_wire_types(const cvmgpu_program *p, const uint8_t **is_bool, uint32_t *n);
/* Per witness wire: its row in the typed value store -- a field row index, or 0x80000000 | bit row index. */
int cvmgpu_program_wire_rows(const cvmgpu_program *p, const uint32_t **wire_loc, uint32_t *n);

/* ---- witness generation ----------------------------------------------------------------------- */
/* HOST buffers.  inputs: B x n_inputs x 32 B (main inputs in signal order, canonical LE; values >= q are
 * reduced like Fr_str2element does).  wtns_out: B x n_wires x 32 B canonical LE -- row b is exactly the
 * data section of the .wtns the reference writes for input b.  status: B words.  Either output may be NULL. */
int cvmgpu_witness_batch(cvmgpu_program *p, const uint8_t *inputs, uint64_t B, uint8_t *wtns_out, uint32_t *status);
/* Same, plus the R1CS check of every witness on the device before it is exported: first_bad[b] as in
 * cvmgpu_r1cs_check.  r may be NULL (then first_bad is ignored).  Chunks of the batch are pipelined over two
 * streams; pass pinned host memory to let copies overlap the kernels. */
int cvmgpu_witness_batch_checked(cvmgpu_program *p, cvmgpu_r1cs *r, const uint8_t *inputs, uint64_t B,
                                 uint8_t *wtns_out, uint32_t *status, uint32_t *first_bad);

/* Same with an output selector: only wires [wire0, wire0 + n_sel) of every witness are exported and downloaded
 * (wtns_out: B x n_sel x 32 B).  Wires 0 .. n_outputs + n_pub_inputs are the public part of a witness (constant 1,
 * main outputs, public inputs): what a caller that keeps the witness on the device for a GPU prover needs back. */
int cvmgpu_witness_batch_select(cvmgpu_program *p, cvmgpu_r1cs *r, const uint8_t *inputs, uint64_t B, uint32_t wire0,
                                uint32_t n_sel, uint8_t *wtns_out, uint32_t *status, uint32_t *first_bad);

/* Single-process multi-device driver (SURVEY 8b `device_mask`): bit d of device_mask = use CUDA device d.  The batch is
 * cut into one contiguous slice per device; each slice runs on its device from its own host thread with its own
 * streams and buffers (program and constraint tables are replicated per device on first use).  Witnesses are
 * independent: there is no inter-device traffic.  Arguments otherwise as cvmgpu_witness_batch_select. */
int cvmgpu_witness_batch_multi(cvmgpu_program *p, cvmgpu_r1cs *r, const uint8_t *inputs, uint64_t B, uint32_t device_mask,
                               uint32_t wire0, uint32_t n_sel, uint8_t *wtns_out, uint32_t *status, uint32_t *first_bad);

/* DEVICE buffers on the current device.
 *   d_inputs  B x n_inputs x 32 B, as above
 *   d_store   typed value store of cvmgpu_store_bytes(p, bstride) bytes, B <= bstride < 2^27:
 *               field rows: row r, half h, witness w at ((r*2+h)*bstride + w)*16, Montgomery form;
 *               then bit rows (values proven 0/1): 32-bit word ((w >> 5) * n_brows + row), bit w & 31;
 *             cvmgpu_program_wire_rows maps witness wires to rows
 *   d_status  B words
 * stream: a cudaStream_t (0 = default stream). */
int cvmgpu_witness_batch_dev(cvmgpu_program *p, const void *d_inputs, uint64_t B, uint64_t bstride, void *d_store,
                             void *d_status, void *stream);
/* Witness generation AND R1CS check on device buffers (what cvmgpu_witness_batch_checked runs per chunk).  For programs
 * whose values are all field elements (no 0/1-typed values: Poseidon-like arithmetic circuits) the constraints are
 * scheduled into the tape -- each one is evaluated right after the last wire it mentions is produced, operands still
 * on chip -- and ONE kernel writes the store, the status and first_bad; other programs run the tape and then the check
 * kernels.  Same results either way (A.w * B.w = C.w on the stored values, first violated constraint or CVMGPU_NO_BAD).
 * d_store: cvmgpu_store_bytes_checked(p, r, bstride) bytes. */
int cvmgpu_witness_batch_checked_dev(cvmgpu_program *p, cvmgpu_r1cs *r, const void *d_inputs, uint64_t B, uint64_t bstride,
                                     void *d_store, void *d_status, void *d_first_bad, void *stream);
size_t cvmgpu_store_bytes_checked(cvmgpu_program *p, cvmgpu_r1cs *r, uint64_t bstride);
/* 0: never fuse; 1 (default): fuse field programs whose constraints are all evaluated in the field (constraints over 0/1
 * wires are cheaper in the table / integer check kernels); 2: fuse whenever the program allows it (tests, measurements).
 * Initial value: CVMGPU_FUSED in the environment. */
int cvmgpu_set_fused_mode(int mode);
/* counters / instructions of the fused tape of (p, r); CVMGPU_ERR_UNSUPPORTED when the pair runs separate kernels */
int cvmgpu_program_fused_info_get(cvmgpu_program *p, cvmgpu_r1cs *r, cvmgpu_program_info *info);
int cvmgpu_program_fused_tape(cvmgpu_program *p, cvmgpu_r1cs *r, const void **ins, uint64_t *n_ins, const void **consts,
                              uint32_t *n_consts);
/* value store -> .wtns row layout: d_wtns = B x n_wires x 32 B canonical LE */
int cvmgpu_witness_export_dev(cvmgpu_program *p, const void *d_store, uint64_t B, uint64_t bstride, void *d_wtns,
                              void *stream);
/* only wires [wire0, wire0 + n_sel): d_out = B x n_sel x 32 B */
int cvmgpu_witness_export_range_dev(cvmgpu_program *p, const void *d_store, uint64_t B, uint64_t bstride, uint32_t wire0,
                                    uint32_t n_sel, void *d_out, void *stream);
size_t cvmgpu_store_bytes(const cvmgpu_program *p, uint64_t bstride);

/* write one witness (n_wires x 32 B canonical) as a .wtns file, byte-identical to main.cpp:286-332 */
int cvmgpu_wtns_write(const char *path, const uint8_t *witness, uint32_t n_wires);

/* ---- R1CS ------------------------------------------------------------------------------------- */
int cvmgpu_r1cs_load(const char *r1cs_path, cvmgpu_r1cs **out);
int cvmgpu_r1cs_info_get(const cvmgpu_r1cs *r, cvmgpu_r1cs_info *info);
void cvmgpu_r1cs_free(cvmgpu_r1cs *r);
/* cvmgpu_r1cs_info_get with the bound_* counters computed for program p's value-store layout (host only, no GPU needed) */
int cvmgpu_r1cs_bind_info(const cvmgpu_r1cs *r, const cvmgpu_program *p, cvmgpu_r1cs_info *info);
/* HOST: witnesses B x n_wires x 32 B canonical; first_bad[b] = index of the first violated constraint or
 * 0xffffffff when witness b satisfies every constraint. */
int cvmgpu_r1cs_check(cvmgpu_r1cs *r, const uint8_t *witnesses, uint64_t B, uint32_t *first_bad);
/* DEVICE: checks the typed value store written by cvmgpu_witness_batch_dev for program p (the CSR is bound to p's
 * wire -> row map on first use; constraints over 0/1 wires with small coefficients are evaluated in 64-bit integers). */
int cvmgpu_r1cs_check_store_dev(cvmgpu_r1cs *r, cvmgpu_program *p, const void *d_store, uint64_t B, uint64_t bstride,
                                void *d_first_bad, void *stream);
/* DEVICE: checks a PLAIN store (row = wire, all field rows, Montgomery) as written by cvmgpu_witness_import_dev. */
int cvmgpu_r1cs_check_dev(cvmgpu_r1cs *r, const void *d_store, uint64_t B, uint64_t bstride, void *d_first_bad,
                          void *stream);
/* DEVICE: canonical AoS witnesses (B x n_wires x 32 B) -> plain value-store layout (n_wires x 2 x bstride x 16 B, Montgomery) */
int cvmgpu_witness_import_dev(uint32_t n_wires, const void *d_wtns, uint64_t B, uint64_t bstride, void *d_store,
                              void *stream);

/* ---- field arithmetic, host build of the device limb routines (test hook; no GPU needed) ------- */
/* op: add sub mul div idiv mod pow shl shr band bor bxor bnot lt leq gt geq eq neq land lor lnot neg inv square
 * a, b, out: 32-byte LE canonical.  Returns 1 where the reference aborts (`\` or `%` by zero: GMP raises); a field
 * division by zero yields 0, as the reference's Fr_div does (bn128/fr.cpp:146-163). */
int cvmgpu_fr_host_op(const char *op, const uint8_t *a, const uint8_t *b, uint8_t *out);
/* same operations executed by a CUDA kernel over n element pairs (device self-test of csrc/fr.cuh) */
int cvmgpu_fr_device_op(const char *op, const uint8_t *a, const uint8_t *b, uint8_t *out, uint64_t n);
/* dependency-free integer-pipe micro-benchmark.  The returned rate counts 8 (kind 8: 16) "units" per thread and iteration, where a
 * unit is: kind 0 one mad.lo+mad.hi pair, 1 one mad.wide.u32 (both = one 32x32->64 multiply-accumulate), 2 two mad.lo,
 * 3 two mad.hi, 4 two addc (carry chain), 5 two add, 6 two mad{c}.{lo,hi}.cc (carry chain on the multiply pipe),
 * 7 one mul.wide.u32 + one add, 8 one fused mad.lo.cc/madc.hi.cc pair = IMAD.WIDE.U32.X with carry in and out (the form
 * the field arithmetic uses). */
int cvmgpu_imad_peak(int kind, double *macs_per_second, double *ms);
/* register-resident Montgomery-multiplication throughput (no memory traffic): variant 0 = portable 64-bit CIOS,
 * 1 = mul.wide formulation with a second, warp-uniform chain (which ptxas runs on the uniform datapath), 2 = mul.wide
 * formulation, one dependent chain per thread; 3 = carry-chained IMAD.WIDE rows, one dependent chain per thread (the
 * multiplier the kernels use), 4 = the squaring (100 instead of 128 IMAD.WIDE);
 * ctas_per_sm x 128 threads per SM. */
int cvmgpu_mul_peak(int variant, int ctas_per_sm, double *muls_per_second);

#ifdef __cplusplus
}
#endif
#endif /* CVMGPU_H */
