"""Python host side over the C ABI (include/cvmgpu.h) -- the ctypes twin of the Rust `extern "C"`
binding the north star asks for (no Rust toolchain in this image).

The classes mirror the reference runtime's objects for this path:

  WitnessCalculator   <->  Circom_CalcWit + generated <circuit>.cpp   (common/calcwit.hpp:17-66)
      .calculate(inputs)        setInputSignal(...)/run(ctx)   for B inputs at once
      .write_wtns(path, row)    writeBinWitness                (common/main.cpp:286-332)
  R1cs                <->  the .r1cs written by constraint_writers/src/r1cs_writer.rs
      .check(witnesses)         (A.w)*(B.w) - C.w == 0 per constraint

There is no CPU fallback: if the CUDA library is missing or no GPU is visible, compute calls raise.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, byref, c_char_p, c_double, c_int, c_size_t, c_uint32, c_uint64, c_void_p

import numpy as np

_LIB = None
LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libcvmgpu.so")

ST_OK, ST_ASSERT, ST_TOINT, ST_DIVZERO, ST_INPUT, ST_LOOP, ST_SPECULATION = 0, 1, 2, 3, 4, 5, 6   # include/cvmgpu.h CVMGPU_ST_*
NO_BAD = 0xFFFFFFFF
ROW_BIT = 0x80000000       # cvmgpu_program_wire_rows: the wire is a bit row (csrc/tape.hpp ROW_BIT)

EXPORTS = [
    "cvmgpu_last_error", "cvmgpu_device_count", "cvmgpu_set_device",
    "cvmgpu_program_load", "cvmgpu_program_load_text", "cvmgpu_program_load_with_cpp", "cvmgpu_program_load_text2", "cvmgpu_program_load_files", "cvmgpu_program_load_text3", "cvmgpu_witness_batch_checked_dev", "cvmgpu_store_bytes_checked",
    "cvmgpu_program_fused_info_get", "cvmgpu_program_fused_tape", "cvmgpu_set_fused_mode", "cvmgpu_program_speculative", "cvmgpu_program_main_inputs", "cvmgpu_witness_batch_bits", "cvmgpu_packed_layout", "cvmgpu_witness_batch_packed",
    "cvmgpu_witness_export_packed_dev", "cvmgpu_program_info_get", "cvmgpu_program_free",
    "cvmgpu_program_tape", "cvmgpu_program_witness", "cvmgpu_program_wire_types", "cvmgpu_program_wire_rows", "cvmgpu_program_iconsts",
    "cvmgpu_witness_batch", "cvmgpu_witness_batch_checked", "cvmgpu_witness_batch_select", "cvmgpu_witness_batch_multi", "cvmgpu_witness_batch_dev",
    "cvmgpu_witness_export_dev", "cvmgpu_witness_export_range_dev", "cvmgpu_store_bytes", "cvmgpu_release_buffers",
    "cvmgpu_wtns_write",
    "cvmgpu_r1cs_load", "cvmgpu_r1cs_info_get", "cvmgpu_r1cs_free", "cvmgpu_r1cs_check", "cvmgpu_r1cs_check_dev",
    "cvmgpu_r1cs_check_store_dev", "cvmgpu_r1cs_bind_info", "cvmgpu_witness_import_dev",
    "cvmgpu_fr_host_op", "cvmgpu_fr_device_op", "cvmgpu_imad_peak", "cvmgpu_mul_peak", "cvmgpu_set_tape_mode",
]


class CvmGpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("cvmgpu error %d: %s" % (code, msg))
        self.code = code


class ProgramInfo(ctypes.Structure):
    _fields_ = [("struct_size", c_uint32), ("reserved0", c_uint32), ("n_signals", c_uint64), ("n_wires", c_uint32), ("n_inputs", c_uint32), ("n_outputs", c_uint32),
                ("n_slots", c_uint32), ("n_rows", c_uint32), ("tape_len", c_uint64), ("ref_mul", c_uint64),
                ("ref_field_ops", c_uint64), ("cvm_instructions", c_uint64), ("tape_mul", c_uint64),
                ("tape_div", c_uint64), ("tape_addsub", c_uint64), ("tape_other", c_uint64), ("tape_ld", c_uint64),
                ("tape_st", c_uint64), ("tape_spill_st", c_uint64), ("n_consts", c_uint32), ("dyn_branches", c_uint32),
                ("ref_div", c_uint64), ("tape_inv", c_uint64), ("tape_sel", c_uint64), ("tape_dot", c_uint64),
                ("tape_dot_terms", c_uint64), ("tape_macs", c_uint64), ("tape_ld_streamed", c_uint64), ("unrolled_iterations", c_uint64), ("tape_lut", c_uint64),
                ("tape_ld_bool", c_uint64), ("tape_spill_st_bool", c_uint64), ("n_bool_wires", c_uint64),
                ("n_bslots", c_uint32), ("n_frows", c_uint32), ("n_brows", c_uint32), ("max_live_field", c_uint32),
                ("max_live_bool", c_uint32), ("reserved1", c_uint32), ("tape_int", c_uint64)]

    def asdict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


class R1csInfo(ctypes.Structure):
    _fields_ = [("struct_size", c_uint32), ("n_wires", c_uint32), ("n_pub_out", c_uint32), ("n_pub_in", c_uint32), ("n_prv_in", c_uint32),
                ("n_constraints", c_uint32), ("n_labels", c_uint64), ("nnz", c_uint64), ("nnz_pm1", c_uint64),
                ("n_coefs", c_uint32), ("nnz_small", c_uint64), ("macs", c_uint64), ("n_quadratic", c_uint64),
                ("nnz_const", c_uint64), ("n_squares", c_uint64), ("bound_int_constraints", c_uint64),
                ("bound_bit_terms", c_uint64), ("bound_field_terms", c_uint64), ("bound_macs", c_uint64),
                ("bound_bit_adds", c_uint64), ("bound_table_constraints", c_uint64)]

    def asdict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


def lib():
    """Load csrc/libcvmgpu.so (built by circom_cvm_b200.build).  Raises if it is missing: the product
    path never substitutes a CPU implementation."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise CvmGpuError(-4, "CUDA library %s is not built; run `python -m circom_cvm_b200.build`" % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    L.cvmgpu_last_error.restype = c_char_p
    L.cvmgpu_program_load.argtypes = [c_char_p, c_uint32, POINTER(c_void_p)]
    L.cvmgpu_program_load_text.argtypes = [c_char_p, c_size_t, c_uint32, POINTER(c_void_p)]
    L.cvmgpu_program_load_with_cpp.argtypes = [c_char_p, c_char_p, c_uint32, POINTER(c_void_p)]
    L.cvmgpu_program_load_text2.argtypes = [c_char_p, c_size_t, c_char_p, c_size_t, c_uint32, POINTER(c_void_p)]
    L.cvmgpu_program_load_files.argtypes = [c_char_p, c_char_p, c_char_p, c_uint32, POINTER(c_void_p)]
    L.cvmgpu_program_load_text3.argtypes = [c_char_p, c_size_t, c_char_p, c_size_t, c_char_p, c_size_t, c_uint32, POINTER(c_void_p)]
    L.cvmgpu_program_info_get.argtypes = [c_void_p, POINTER(ProgramInfo)]
    L.cvmgpu_program_free.argtypes = [c_void_p]
    L.cvmgpu_program_free.restype = None
    L.cvmgpu_program_tape.argtypes = [c_void_p, POINTER(c_void_p), POINTER(c_uint64), POINTER(c_void_p), POINTER(c_uint32)]
    L.cvmgpu_program_witness.argtypes = [c_void_p, POINTER(c_void_p), POINTER(c_uint32)]
    L.cvmgpu_program_wire_types.argtypes = [c_void_p, POINTER(c_void_p), POINTER(c_uint32)]
    L.cvmgpu_program_wire_rows.argtypes = [c_void_p, POINTER(c_void_p), POINTER(c_uint32)]
    L.cvmgpu_program_iconsts.argtypes = [c_void_p, POINTER(c_void_p), POINTER(c_uint32)]
    L.cvmgpu_witness_batch_select.argtypes = [c_void_p, c_void_p, c_void_p, c_uint64, c_uint32, c_uint32, c_void_p, c_void_p, c_void_p]
    L.cvmgpu_packed_layout.argtypes = [c_void_p, c_int, POINTER(c_uint64), POINTER(c_uint32), POINTER(c_uint32), POINTER(c_void_p)]
    L.cvmgpu_witness_batch_packed.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_uint64, c_void_p, c_void_p, c_void_p]
    L.cvmgpu_witness_export_packed_dev.argtypes = [c_void_p, c_void_p, c_uint64, c_uint64, c_void_p, c_void_p]
    L.cvmgpu_witness_batch_bits.argtypes = [c_void_p, c_void_p, c_void_p, c_uint64, c_uint32, c_uint32, c_void_p, c_void_p, c_void_p]
    L.cvmgpu_witness_batch_multi.argtypes = [c_void_p, c_void_p, c_void_p, c_uint64, c_uint32, c_uint32, c_uint32, c_void_p, c_void_p, c_void_p]
    L.cvmgpu_witness_export_range_dev.argtypes = [c_void_p, c_void_p, c_uint64, c_uint64, c_uint32, c_uint32, c_void_p, c_void_p]
    L.cvmgpu_r1cs_bind_info.argtypes = [c_void_p, c_void_p, POINTER(R1csInfo)]
    L.cvmgpu_r1cs_check_store_dev.argtypes = [c_void_p, c_void_p, c_void_p, c_uint64, c_uint64, c_void_p, c_void_p]
    L.cvmgpu_release_buffers.argtypes = []
    L.cvmgpu_release_buffers.restype = None
    L.cvmgpu_witness_batch.argtypes = [c_void_p, c_void_p, c_uint64, c_void_p, c_void_p]
    L.cvmgpu_witness_batch_checked.argtypes = [c_void_p, c_void_p, c_void_p, c_uint64, c_void_p, c_void_p, c_void_p]
    L.cvmgpu_witness_batch_dev.argtypes = [c_void_p, c_void_p, c_uint64, c_uint64, c_void_p, c_void_p, c_void_p]
    L.cvmgpu_witness_export_dev.argtypes = [c_void_p, c_void_p, c_uint64, c_uint64, c_void_p, c_void_p]
    L.cvmgpu_store_bytes.argtypes = [c_void_p, c_uint64]
    L.cvmgpu_store_bytes.restype = c_size_t
    L.cvmgpu_store_bytes_checked.argtypes = [c_void_p, c_void_p, c_uint64]
    L.cvmgpu_store_bytes_checked.restype = c_size_t
    L.cvmgpu_witness_batch_checked_dev.argtypes = [c_void_p, c_void_p, c_void_p, c_uint64, c_uint64, c_void_p, c_void_p, c_void_p, c_void_p]
    L.cvmgpu_program_main_inputs.argtypes = [c_void_p, POINTER(c_char_p), POINTER(c_size_t)]
    L.cvmgpu_program_speculative.argtypes = [c_void_p, POINTER(c_void_p)]
    L.cvmgpu_program_fused_info_get.argtypes = [c_void_p, c_void_p, POINTER(ProgramInfo)]
    L.cvmgpu_program_fused_tape.argtypes = [c_void_p, c_void_p, POINTER(c_void_p), POINTER(c_uint64), POINTER(c_void_p), POINTER(c_uint32)]
    L.cvmgpu_wtns_write.argtypes = [c_char_p, c_void_p, c_uint32]
    L.cvmgpu_r1cs_load.argtypes = [c_char_p, POINTER(c_void_p)]
    L.cvmgpu_r1cs_info_get.argtypes = [c_void_p, POINTER(R1csInfo)]
    L.cvmgpu_r1cs_free.argtypes = [c_void_p]
    L.cvmgpu_r1cs_free.restype = None
    L.cvmgpu_r1cs_check.argtypes = [c_void_p, c_void_p, c_uint64, c_void_p]
    L.cvmgpu_r1cs_check_dev.argtypes = [c_void_p, c_void_p, c_uint64, c_uint64, c_void_p, c_void_p]
    L.cvmgpu_witness_import_dev.argtypes = [c_uint32, c_void_p, c_uint64, c_uint64, c_void_p, c_void_p]
    L.cvmgpu_fr_host_op.argtypes = [c_char_p, c_char_p, c_char_p, c_char_p]
    L.cvmgpu_fr_device_op.argtypes = [c_char_p, c_void_p, c_void_p, c_void_p, c_uint64]
    L.cvmgpu_imad_peak.argtypes = [c_int, POINTER(c_double), POINTER(c_double)]
    L.cvmgpu_mul_peak.argtypes = [c_int, c_int, POINTER(c_double)]
    L.cvmgpu_set_device.argtypes = [c_int]
    _LIB = L
    return L


def _check(rc):
    if rc != 0:
        raise CvmGpuError(rc, lib().cvmgpu_last_error().decode(errors="replace"))


def device_count():
    return int(lib().cvmgpu_device_count())


def set_device(i):
    _check(lib().cvmgpu_set_device(int(i)))


def _ptr(x):
    """host numpy array / bytes / torch tensor -> address"""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    raise TypeError(type(x))


def ints_to_le(rows, width):
    """rows of python ints -> uint8 array [len(rows), width, 32]"""
    out = np.zeros((len(rows), width, 32), dtype=np.uint8)
    for i, r in enumerate(rows):
        assert len(r) == width, "expected %d values per row, got %d" % (width, len(r))
        buf = b"".join(int(v).to_bytes(32, "little") for v in r)
        out[i] = np.frombuffer(buf, dtype=np.uint8).reshape(width, 32)
    return out


def le_to_ints(arr):
    """uint8 array [..., 32] -> nested lists of python ints (first two dims)"""
    a = np.ascontiguousarray(arr)
    flat = a.reshape(-1, 32)
    vals = [int.from_bytes(flat[i].tobytes(), "little") for i in range(flat.shape[0])]
    if a.ndim == 3:
        n = a.shape[1]
        return [vals[i * n:(i + 1) * n] for i in range(a.shape[0])]
    return vals


class WitnessCalculator:
    """Batched counterpart of the reference's generated witness calculator."""

    def __init__(self, cvm_path=None, cvm_text=None, n_slots=0, cpp_path=None, cpp_text=None, dat_path=None, dat_bytes=None):
        """cpp_path / cpp_text: the generated <circuit>.cpp of the same compile, from which component creation is
        recovered when the .cvm file does not carry `;;%%create_cmp` lines (the unpatched emitter prints nothing).
        dat_path / dat_bytes: the <circuit>.dat of the same compile, whose io-map circuits with mixed component arrays
        need (its section sizes are in the .cpp, so it is only read together with it)."""
        L = lib()
        h = c_void_p()
        if cvm_path is not None:
            _check(L.cvmgpu_program_load_files(os.fsencode(cvm_path), os.fsencode(cpp_path) if cpp_path else None,
                                               os.fsencode(dat_path) if dat_path else None, n_slots, byref(h)))
        else:
            data = cvm_text.encode() if isinstance(cvm_text, str) else cvm_text
            cpp = cpp_text.encode() if isinstance(cpp_text, str) else cpp_text
            _check(L.cvmgpu_program_load_text3(data, len(data), cpp, len(cpp) if cpp else 0, dat_bytes,
                                               len(dat_bytes) if dat_bytes else 0, n_slots, byref(h)))
        self._h = h
        self._owner = None
        self._finish_init()

    def _finish_init(self):
        L = lib()
        info = ProgramInfo()
        info.struct_size = ctypes.sizeof(ProgramInfo)
        _check(L.cvmgpu_program_info_get(self._h, byref(info)))
        self.info = info
        self.n_inputs = int(info.n_inputs)
        self.n_wires = int(info.n_wires)
        self.n_rows = int(info.n_rows)

    def speculative(self):
        """The same circuit traced under "every main input is 0 or 1" (bit-heavy programs only; None otherwise): a
        WitnessCalculator for the DEVICE API with its own value-store layout.  Witnesses it flags ST_SPECULATION must be
        recomputed with this calculator; the host-buffer calls (calculate*, calculate_checked ...) do all of that themselves.
        The handle belongs to this object."""
        h = c_void_p()
        _check(lib().cvmgpu_program_speculative(self._h, byref(h)))
        if not h.value:
            return None
        child = WitnessCalculator.__new__(WitnessCalculator)
        child._h, child._owner = h, self
        child._finish_init()
        return child

    def close(self):
        if getattr(self, "_h", None) and getattr(self, "_owner", None) is None:
            lib().cvmgpu_program_free(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- host-buffer API (what a user of the reference's calculator would call)
    def calculate(self, inputs, want_witness=True):
        """inputs: uint8 array [B, n_inputs, 32] (LE canonical) or list of rows of ints.
        Returns (witness uint8 [B, n_wires, 32] or None, status uint32 [B])."""
        if not isinstance(inputs, np.ndarray):
            inputs = ints_to_le(inputs, self.n_inputs)
        inputs = np.ascontiguousarray(inputs, dtype=np.uint8)
        B = inputs.shape[0]
        assert inputs.size == B * self.n_inputs * 32, "input shape mismatch"
        wt = np.empty((B, self.n_wires, 32), dtype=np.uint8) if want_witness else None
        st = np.empty((B,), dtype=np.uint32)
        _check(lib().cvmgpu_witness_batch(self._h, _ptr(inputs) if inputs.size else None, B, _ptr(wt), _ptr(st)))
        return wt, st

    def calculate_into(self, inputs, wtns_out, status_out, r1cs=None, first_bad_out=None):
        """Same, with caller-provided (e.g. pinned torch) host buffers; with `r1cs` every witness is also
        checked on the device and first_bad_out[b] receives the first violated constraint (NO_BAD = none)."""
        B = status_out.shape[0]
        _check(lib().cvmgpu_witness_batch_checked(self._h, r1cs._h if r1cs is not None else None, _ptr(inputs), B,
                                                  _ptr(wtns_out), _ptr(status_out), _ptr(first_bad_out)))

    def calculate_checked(self, inputs, r1cs):
        """-> (witness [B, n_wires, 32], status [B], first_bad [B])"""
        if not isinstance(inputs, np.ndarray):
            inputs = ints_to_le(inputs, self.n_inputs)
        inputs = np.ascontiguousarray(inputs, dtype=np.uint8)
        B = inputs.shape[0]
        wt = np.empty((B, self.n_wires, 32), dtype=np.uint8)
        st = np.empty((B,), dtype=np.uint32)
        bad = np.empty((B,), dtype=np.uint32)
        self.calculate_into(inputs, wt, st, r1cs, bad)
        return wt, st, bad

    # ---- device-buffer API (torch tensors on the current CUDA device)
    def store_bytes(self, bstride):
        return int(lib().cvmgpu_store_bytes(self._h, bstride))

    def run_dev(self, d_inputs, B, bstride, d_store, d_status, stream=0):
        _check(lib().cvmgpu_witness_batch_dev(self._h, _ptr(d_inputs), B, bstride, _ptr(d_store), _ptr(d_status), stream))

    def run_checked_dev(self, r1cs, d_inputs, B, bstride, d_store, d_status, d_first_bad, stream=0):
        """witness generation + R1CS check on device buffers: one kernel for field-only programs (the check is scheduled into
        the tape), tape + check kernels otherwise.  d_store: store_bytes_checked(r1cs, bstride) bytes."""
        _check(lib().cvmgpu_witness_batch_checked_dev(self._h, r1cs._h, _ptr(d_inputs), B, bstride, _ptr(d_store), _ptr(d_status),
                                                      _ptr(d_first_bad), stream))

    def store_bytes_checked(self, r1cs, bstride):
        return int(lib().cvmgpu_store_bytes_checked(self._h, r1cs._h, bstride))

    def fused_info(self, r1cs):
        """ProgramInfo of the tape with r1cs's check scheduled into it, or None when the pair runs separate kernels"""
        info = ProgramInfo()
        info.struct_size = ctypes.sizeof(ProgramInfo)
        if lib().cvmgpu_program_fused_info_get(self._h, r1cs._h, byref(info)) != 0:
            return None
        return info

    def fused_tape(self, r1cs):
        """-> (tape instructions, constants, layout) of the fused tape, for tests/tape_emulator.py; None if there is none"""
        info = self.fused_info(r1cs)
        if info is None:
            return None
        ins, n, cs, nc = c_void_p(), c_uint64(), c_void_p(), c_uint32()
        _check(lib().cvmgpu_program_fused_tape(self._h, r1cs._h, byref(ins), byref(n), byref(cs), byref(nc)))
        tape, consts = self._tape_arrays(ins, n, cs, nc)
        layout = dict(self.layout(), n_slots=int(info.n_slots), n_bslots=int(info.n_bslots), n_frows=int(info.n_frows),
                      n_brows=int(info.n_brows), iconsts=[])
        return tape, consts, layout

    def export_dev(self, d_store, B, bstride, d_wtns, stream=0):
        _check(lib().cvmgpu_witness_export_dev(self._h, _ptr(d_store), B, bstride, _ptr(d_wtns), stream))

    def export_range_dev(self, d_store, B, bstride, wire0, n_sel, d_out, stream=0):
        _check(lib().cvmgpu_witness_export_range_dev(self._h, _ptr(d_store), B, bstride, wire0, n_sel, _ptr(d_out), stream))

    def calculate_select_into(self, inputs, wire0, n_sel, wtns_out, status_out, r1cs=None, first_bad_out=None):
        """calculate_into with an output selector: only wires [wire0, wire0 + n_sel) come back (wtns_out [B, n_sel, 32])"""
        B = status_out.shape[0]
        _check(lib().cvmgpu_witness_batch_select(self._h, r1cs._h if r1cs is not None else None, _ptr(inputs), B, wire0, n_sel,
                                                 _ptr(wtns_out), _ptr(status_out), _ptr(first_bad_out)))

    def calculate_bits_into(self, input_bits, wire0, n_sel, wtns_out, status_out, r1cs=None, first_bad_out=None):
        """calculate_select_into with the inputs as packed bits: uint8 [B, ceil(n_inputs / 8)], input k = bit k & 7 of byte
        k >> 3 (np.packbits(..., bitorder="little")).  Only programs that have a bit-input tape (speculative())."""
        B = status_out.shape[0]
        _check(lib().cvmgpu_witness_batch_bits(self._h, r1cs._h if r1cs is not None else None, _ptr(input_bits), B, wire0, n_sel,
                                               _ptr(wtns_out), _ptr(status_out), _ptr(first_bad_out)))

    def packed_layout(self, bit_input_tape=False):
        """-> (row bytes, number of field wires, number of 0/1 wires, wire -> typed row map) of the packed witness rows"""
        rb, nf, nb, ptr = c_uint64(), c_uint32(), c_uint32(), c_void_p()
        _check(lib().cvmgpu_packed_layout(self._h, int(bit_input_tape), byref(rb), byref(nf), byref(nb), byref(ptr)))
        rows = np.ctypeslib.as_array(ctypes.cast(ptr, POINTER(c_uint32)), shape=(self.n_wires,)).copy()
        return int(rb.value), int(nf.value), int(nb.value), rows

    def calculate_packed_into(self, inputs, inputs_are_bits, packed_out, status_out, r1cs=None, first_bad_out=None):
        """the whole witness as packed rows (packed_layout(inputs_are_bits)): field wires 32 bytes each, 0/1 wires one bit each"""
        B = status_out.shape[0]
        _check(lib().cvmgpu_witness_batch_packed(self._h, r1cs._h if r1cs is not None else None, _ptr(inputs), int(inputs_are_bits), B,
                                                 _ptr(packed_out), _ptr(status_out), _ptr(first_bad_out)))

    @staticmethod
    def unpack_rows(packed, layout):
        """packed rows [B, row_bytes] + packed_layout() -> canonical ints per witness (test / inspection helper)"""
        row_bytes, n_f, n_b, rows = layout
        out = []
        for r in packed:
            raw = bytes(r)
            fvals = [int.from_bytes(raw[32 * j:32 * j + 32], "little") for j in range(n_f)]
            bits = np.unpackbits(np.frombuffer(raw[32 * n_f:], dtype=np.uint8), bitorder="little")
            w, fi, bi = [], 0, 0
            for loc in rows:
                if loc & ROW_BIT:
                    w.append(int(bits[bi]))
                    bi += 1
                else:
                    w.append(fvals[fi])
                    fi += 1
            out.append(w)
        return out

    def calculate_multi_into(self, inputs, device_mask, wire0, n_sel, wtns_out, status_out, r1cs=None, first_bad_out=None):
        """the same over several devices of this process (bit d of device_mask = CUDA device d)"""
        B = status_out.shape[0]
        _check(lib().cvmgpu_witness_batch_multi(self._h, r1cs._h if r1cs is not None else None, _ptr(inputs), B, device_mask, wire0,
                                                n_sel, _ptr(wtns_out), _ptr(status_out), _ptr(first_bad_out)))

    def wire_rows(self):
        """uint32 per witness wire: its row in the typed value store (ROW_BIT | bit row, or field row)"""
        ptr, n = c_void_p(), c_uint32()
        _check(lib().cvmgpu_program_wire_rows(self._h, byref(ptr), byref(n)))
        return np.ctypeslib.as_array(ctypes.cast(ptr, POINTER(c_uint32)), shape=(n.value,)).copy() if n.value else np.zeros(0, np.uint32)

    def layout(self):
        """what tests/tape_emulator.py needs to execute the tape: slot files, row counts, wire -> row map"""
        i = self.info
        ptr, n = c_void_p(), c_uint32()
        _check(lib().cvmgpu_program_iconsts(self._h, byref(ptr), byref(n)))
        iconsts = [int(x) for x in np.ctypeslib.as_array(ctypes.cast(ptr, POINTER(c_uint64)), shape=(n.value,))] if n.value else []
        return {"n_slots": int(i.n_slots), "n_bslots": int(i.n_bslots), "n_frows": int(i.n_frows), "n_brows": int(i.n_brows),
                "wire_loc": [int(x) for x in self.wire_rows()], "iconsts": iconsts}

    def write_wtns(self, path, witness_row):
        row = np.ascontiguousarray(witness_row, dtype=np.uint8)
        _check(lib().cvmgpu_wtns_write(os.fsencode(path), _ptr(row), self.n_wires))

    def main_inputs(self):
        """[(name, first signal, size)] from the program's `;;%%main_input` lines ([] when it has none)"""
        text, n = c_char_p(), c_size_t()
        _check(lib().cvmgpu_program_main_inputs(self._h, byref(text), byref(n)))
        out = []
        for line in (text.value or b"").decode().splitlines():
            name, start, size = line.rsplit(" ", 2)
            out.append((name, int(start), int(size)))
        return out

    def witness_signals(self):
        """the %%witness list: signal index of every witness wire"""
        sig, n = c_void_p(), c_uint32()
        _check(lib().cvmgpu_program_witness(self._h, byref(sig), byref(n)))
        return list(np.ctypeslib.as_array(ctypes.cast(sig, POINTER(c_uint64)), shape=(n.value,))) if n.value else []

    def wire_is_bool(self):
        """uint8 per witness wire: proven 0/1 by the trace compiler's typing"""
        ptr, n = c_void_p(), c_uint32()
        _check(lib().cvmgpu_program_wire_types(self._h, byref(ptr), byref(n)))
        return np.ctypeslib.as_array(ctypes.cast(ptr, POINTER(ctypes.c_uint8)), shape=(n.value,)).copy() if n.value else np.zeros(0, np.uint8)

    @property
    def n_outputs(self):
        return int(self.info.n_outputs)

    def tape(self):
        """-> (numpy structured array of tape instructions, constants as python ints in Montgomery form)"""
        ins, n, cs, nc = c_void_p(), c_uint64(), c_void_p(), c_uint32()
        _check(lib().cvmgpu_program_tape(self._h, byref(ins), byref(n), byref(cs), byref(nc)))
        return self._tape_arrays(ins, n, cs, nc)

    @staticmethod
    def _tape_arrays(ins, n, cs, nc):
        dt = np.dtype([("op", "u1"), ("flags", "u1"), ("dst", "<u2"), ("a", "<u4"), ("b", "<u4"), ("c", "<u4")])
        tape = np.frombuffer((ctypes.c_char * (n.value * 16)).from_address(ins.value), dtype=dt).copy() if n.value else np.zeros(0, dt)
        raw = bytes((ctypes.c_char * (nc.value * 32)).from_address(cs.value)) if nc.value else b""
        consts = [int.from_bytes(raw[i * 32:(i + 1) * 32], "little") for i in range(nc.value)]
        return tape, consts


class R1cs:
    def __init__(self, path):
        h = c_void_p()
        _check(lib().cvmgpu_r1cs_load(os.fsencode(path), byref(h)))
        self._h = h
        info = self.refresh_info()
        self.n_wires = int(info.n_wires)
        self.n_constraints = int(info.n_constraints)

    def close(self):
        if getattr(self, "_h", None):
            lib().cvmgpu_r1cs_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def refresh_info(self):
        """(the bound_* counters describe the binding to the last program layout checked)"""
        info = R1csInfo()
        info.struct_size = ctypes.sizeof(R1csInfo)
        _check(lib().cvmgpu_r1cs_info_get(self._h, byref(info)))
        self.info = info
        return info

    def bind_info(self, wc):
        """how this constraint system maps onto WitnessCalculator wc's typed store (host only)"""
        info = R1csInfo()
        info.struct_size = ctypes.sizeof(R1csInfo)
        _check(lib().cvmgpu_r1cs_bind_info(self._h, wc._h, byref(info)))
        return info

    def check(self, witnesses):
        """witnesses: uint8 [B, n_wires, 32] canonical -> uint32 [B] first violated constraint (NO_BAD = satisfied)"""
        w = np.ascontiguousarray(witnesses, dtype=np.uint8)
        B = w.shape[0]
        assert w.size == B * self.n_wires * 32
        bad = np.empty((B,), dtype=np.uint32)
        _check(lib().cvmgpu_r1cs_check(self._h, _ptr(w), B, _ptr(bad)))
        return bad

    def check_dev(self, d_store, B, bstride, d_first_bad, stream=0):
        """PLAIN store (row = wire, as written by cvmgpu_witness_import_dev)"""
        _check(lib().cvmgpu_r1cs_check_dev(self._h, _ptr(d_store), B, bstride, _ptr(d_first_bad), stream))

    def check_store_dev(self, wc, d_store, B, bstride, d_first_bad, stream=0):
        """typed store written by wc.run_dev (WitnessCalculator wc)"""
        _check(lib().cvmgpu_r1cs_check_store_dev(self._h, wc._h, _ptr(d_store), B, bstride, _ptr(d_first_bad), stream))


def set_fused_mode(mode):
    """0: tape and R1CS check always as separate kernels; 1 (default): the check is scheduled into the tape of field programs
    whose constraints are all evaluated in the field; 2: whenever the program allows it"""
    _check(lib().cvmgpu_set_fused_mode(int(mode)))


def set_tape_mode(mode):
    """0 = automatic, 1 / 2 = witnesses per thread of the tape kernel (experiments)"""
    lib().cvmgpu_set_tape_mode(int(mode))


def fr_host_op(op, a, b=0):
    out = ctypes.create_string_buffer(32)
    rc = lib().cvmgpu_fr_host_op(op.encode(), int(a).to_bytes(32, "little"), int(b).to_bytes(32, "little"), out)
    if rc < 0:
        _check(rc)
    return int.from_bytes(out.raw, "little"), rc


def fr_device_op(op, a_list, b_list):
    n = len(a_list)
    a = ints_to_le([a_list], n)[0]
    b = ints_to_le([b_list], n)[0]
    out = np.empty((n, 32), dtype=np.uint8)
    _check(lib().cvmgpu_fr_device_op(op.encode(), _ptr(a), _ptr(b), _ptr(out), n))
    return le_to_ints(out)


def mul_peak(variant, ctas_per_sm):
    v = c_double()
    _check(lib().cvmgpu_mul_peak(variant, ctas_per_sm, byref(v)))
    return v.value


def imad_peak(kind):
    v, ms = c_double(), c_double()
    _check(lib().cvmgpu_imad_peak(kind, byref(v), byref(ms)))
    return v.value, ms.value
