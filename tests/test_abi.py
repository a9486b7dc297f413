"""The C-ABI library loads and exports every symbol declared in include/cvmgpu.h; error behaviour
without touching a GPU."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "cvmgpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cvmgpu_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(cvmlib):
    from circom_cvm_b200 import engine as E
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(cvmlib, s), "missing export " + s
    assert sorted(E.EXPORTS) == syms


def test_load_errors_are_codes_not_aborts(cvmlib, tmp_path):
    from circom_cvm_b200 import engine as E
    with pytest.raises(E.CvmGpuError) as e:
        E.WitnessCalculator(cvm_path=str(tmp_path / "missing.cvm"))
    assert e.value.code == -1
    bad = tmp_path / "bad.cvm"
    bad.write_text("%%prime 7\n")
    with pytest.raises(E.CvmGpuError) as e:
        E.WitnessCalculator(cvm_path=str(bad))
    assert e.value.code == -2
    # a literal assigned to inside a copy loop is what the real emitter prints (store_bucket.rs:1026-1028) and loads
    # (tests/test_faithful_cvm.py); outside any loop it has no meaning
    with pytest.raises(E.CvmGpuError) as e:
        E.WitnessCalculator(cvm_text="%%prime 21888242871839275222246405745257275088548364400416034343698204186575808495617\n"
                                     "%%signals 2\n%%start T_0\n%%witness 0 1\n%%template T_0 [ ] [ ff 0 ] [1] [0]\n"
                                     "i64.3 = i64.add i64.3 i64.1\n")
    assert e.value.code == -2 and "outside a loop" in str(e.value)
    with pytest.raises(E.CvmGpuError) as e:
        E.R1cs(str(tmp_path / "missing.r1cs"))
    assert e.value.code == -1


def test_no_cpu_fallback(cvmlib):
    """Without a CUDA device every compute entry point must fail loudly."""
    from circom_cvm_b200 import engine as E
    from conftest import circuit
    if E.device_count() > 0:
        pytest.skip("a GPU is present")
    wc = E.WitnessCalculator(cvm_text=circuit("multiplier2").cvm)
    with pytest.raises(E.CvmGpuError) as e:
        wc.calculate([[3, 11]])
    assert e.value.code == -4
    with pytest.raises(E.CvmGpuError):
        E.fr_device_op("mul", [1], [2])


def test_plain_c_caller(cvmlib, tmp_path):
    """tests/abi_smoke.c: a C program built with gcc against include/cvmgpu.h and libcvmgpu.so -- struct sizes and offsets
    equal the ctypes twins, the info structs honour struct_size (a smaller struct of an older header is not overrun, an
    unset size is refused), and one batch runs (or fails with CVMGPU_ERR_CUDA where there is no GPU)."""
    import ctypes
    import os
    import subprocess

    from circom_cvm_b200 import engine as E
    from conftest import ROOT, circuit
    src = os.path.join(ROOT, "tests", "abi_smoke.c")
    exe = str(tmp_path / "abi_smoke")
    csrc = os.path.join(ROOT, "circom_cvm_b200", "csrc")
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), src, "-L", csrc, "-lcvmgpu",
                           "-Wl,-rpath," + csrc, "-o", exe])
    cvm = tmp_path / "m2.cvm"
    cvm.write_text(circuit("multiplier2").cvm)
    out = subprocess.run([exe, str(cvm), "3", "11"], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    kv = dict(l.split(" ", 1) for l in out.stdout.strip().split("\n"))
    assert int(kv["sizeof_program_info"]) == ctypes.sizeof(E.ProgramInfo)
    assert int(kv["sizeof_r1cs_info"]) == ctypes.sizeof(E.R1csInfo)
    assert int(kv["offsetof_n_wires"]) == E.ProgramInfo.n_wires.offset
    assert int(kv["offsetof_tape_len"]) == E.ProgramInfo.tape_len.offset
    assert int(kv["offsetof_tape_int"]) == E.ProgramInfo.tape_int.offset
    assert int(kv["offsetof_bound_table_constraints"]) == E.R1csInfo.bound_table_constraints.offset
    assert kv["n_wires"].startswith("4 n_inputs 2")
    if E.device_count() > 0:
        assert kv["batch"] == "ok status 0 product 33" and kv["multi"] == "rc 0"
    else:
        assert kv["batch"].startswith("rc -4") and kv["multi"] == "rc -4"


def test_integration_md_holds_the_generated_rust_binding():
    """INTEGRATION.md's `extern "C"` block is tools/gen_rust_ffi.py's output for the current header (round 1's hand-written
    ProgramInfo was three fields short of the C struct)."""
    import os
    from conftest import ROOT
    from tools.gen_rust_ffi import generate
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = text.split("<!-- BEGIN GENERATED: tools/gen_rust_ffi.py -->\n```rust\n")[1].split("```\n<!-- END GENERATED -->")[0]
    assert block == generate()
