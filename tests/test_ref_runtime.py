"""Program-level parity with the REFERENCE runtime.

oracle/_ref/<circuit> is the reference's own main.cpp + calcwit.cpp + generic/fr.cpp (compiled where they lie by
oracle/build_ref.py) linked with the circuit body that tools/circuitgen emits in the WriteC shapes.  Running
`./<circuit> input.json out.wtns` is therefore what a user of the reference does.  The CVM oracle must produce the
same witness, byte for byte -- this pins oracle/cvm_interp.py (and through it the GPU path) to the reference's
execution of the same programs.  The binaries are prebuilt (they travel to the GPU box); tests skip if absent.
"""
import json
import os
import random
import subprocess

import pytest

from circom_cvm_b200 import formats
from conftest import ROOT, circuit
from oracle import cvm_interp as I
from oracle import fr_model as M

REF = os.path.join(ROOT, "oracle", "_ref")


def ref_binary(name):
    p = os.path.join(REF, name)
    if not os.path.exists(p):
        pytest.skip("oracle/_ref/%s not built (needs /root/reference; run oracle/build_ref.py + build_ref_circuits.py)" % name)
    return p


def input_json(art, values):
    """main inputs in signal order -> the JSON object the reference's loadJson expects (main.cpp:241-284)"""
    doc, k = {}, 0
    for name, _start, size in art.main_inputs:
        vals = [str(v) for v in values[k:k + size]]
        k += size
        dims = [s for s in art.prog.main.tmpl.signals if s.name == name][0].dims
        doc[name] = vals[0] if dims == () else vals
    return doc


CASES = {
    "earlyret": [[5, 3], [M.Q - 5, 3], [3, 5], [1000, 7], [0, 0], [255, 1], [7, 1000], [M.Q - 1, M.Q - 2]],
    "nbits": [[0], [1], [255], [256], [1 << 253], [M.Q - 1]],
    "countdown": [[0], [1], [17], [200]],
    "babyadd4": [[995203441582195749578291179787384436505546430278305826713579947235728471134, 5472060717959818805561601436314318772137091100104008585924551046643952123905, 5299619240641551281634865583518297030282874472190772894086521144482721001553, 16950150798460657717958625567821834550301663161624707787222815936182638968203], [0, 1, 0, 1], [3, 5, 7, 11]],
    "multiplier2": [[3, 11], [M.Q - 7, 3]],
    "multiplier4": [[2, 3, 4, 5]],
    "num2bits8": [[0xA5], [255]],
    "iszero": [[0], [7]],
    "lessthan8": [[3, 200], [200, 3]],
    "sum3cmp": [[1, 0, 1, 1]],
    "dynindex": [[5, 10, 21, 32, 43, 54, 65, 76, 87], [0] + [0] * 8, [7] + [3] * 8, [2, 1, 2, 3, 4, 5, 6, 7, 8]],
    "mixedarr": [list(range(1, 10)) + list(range(11, 20)) + [3, 5, 7], [M.Q - 1] * 21],
    "opszoo": [[12345, 678, 3], [M.Q - 5, 17, 250], [1 << 200, (1 << 253) + 5, 254]],
    "poseidon2": [[1, 2], [M.Q - 1, 12345678901234567890]],
    "poseidon2m": [[1, 2], [M.Q - 1, 12345678901234567890]],
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_runtime_byte_for_byte(name, tmp_path):
    exe = ref_binary(name)
    art = circuit(name)
    prog = I.load(art.cvm)
    rng = random.Random(21)
    cases = list(CASES[name])
    if name in ("multiplier2", "opszoo", "poseidon2", "multiplier4", "babyadd4", "earlyret", "mixedarr", "poseidon2m"):
        cases += [[rng.randrange(M.Q) for _ in range(art.n_inputs)] for _ in range(3)]
    for k, values in enumerate(cases):
        jin, wout = tmp_path / ("in%d.json" % k), tmp_path / ("out%d.wtns" % k)
        jin.write_text(json.dumps(input_json(art, values)))
        subprocess.run([exe, str(jin), str(wout)], check=True, timeout=60)
        ours = formats.wtns_bytes(I.compute_witness(prog, values))
        assert wout.read_bytes() == ours, (name, values)


def test_eddsa_verifier_matches_reference_runtime(tmp_path):
    """BASELINE config 4: EdDSAPoseidonVerifier on signatures from the integer signer; byte-identical .wtns."""
    from tools.circuitgen.circuits import eddsa
    exe = ref_binary("eddsa")
    art = circuit("eddsa")
    prog = I.load(art.cvm)
    for k, (sk, nonce, msg) in enumerate([(123456789, 987654321, 42), (M.Q - 5, 7, 0)]):
        values = eddsa.sign(sk, nonce, msg)
        assert eddsa.verify(values)
        jin, wout = tmp_path / ("in%d.json" % k), tmp_path / ("out%d.wtns" % k)
        jin.write_text(json.dumps(input_json(art, values)))
        subprocess.run([exe, str(jin), str(wout)], check=True, timeout=120)
        assert wout.read_bytes() == formats.wtns_bytes(I.compute_witness(prog, values))
    # a forged message aborts the reference (assert) and raises ASSERT in the oracle
    values = eddsa.sign(11, 22, 33)
    values[6] = 34
    jin = tmp_path / "bad.json"
    jin.write_text(json.dumps(input_json(art, values)))
    r = subprocess.run([exe, str(jin), str(tmp_path / "bad.wtns")], capture_output=True, timeout=120)
    assert r.returncode != 0
    with pytest.raises(I.WitnessError) as e:
        I.compute_witness(prog, values)
    assert e.value.status == I.ST_ASSERT


@pytest.mark.parametrize("seed", [200, 201, 202, 203])
def test_random_circuits_against_reference_runtime(seed, tmp_path):
    """The oracle itself is pinned on random programs: a fuzz circuit (tests/fuzz_circuits.py: every operator, shifts,
    divisions incl. by zero, data-dependent branches) is emitted in the WriteC shapes, linked against the reference's
    own runtime + field arithmetic here, and run as `./circuit input.json out.wtns`; where the reference does not
    abort, the CVM oracle must produce the same bytes, and where it aborts the oracle must raise."""
    import shutil
    from fuzz_circuits import inputs_for, make_circuit
    from tools.circuitgen.build import compile_circuit, write_artifact
    need = [os.path.join(REF, f) for f in ("fr.o", "calcwit.o", "main.o", "ref_harness.o")]
    if not all(os.path.exists(p) for p in need) or shutil.which("g++") is None:
        pytest.skip("reference runtime objects not built (oracle/build_ref.py + build_ref_circuits.py)")
    art = compile_circuit(make_circuit(seed, n_stmts=30), (), name="fuzzref%d" % seed)
    paths = write_artifact(art, str(tmp_path), with_cpp=True)
    exe = str(tmp_path / art.name)
    flags = ["-std=c++11", "-O1", "-w", "-I", os.path.join(REF, "src"), "-I", os.path.join(ROOT, "oracle", "gmp_shim"),
             "-I", os.path.join(REF, "inc")]
    subprocess.run(["g++", *flags, paths["cpp"], *need[3:], *need[:3], "/usr/lib/x86_64-linux-gnu/libgmp.so.10", "-o", exe],
                   check=True, timeout=300)
    prog = I.load(art.cvm)
    agreed = aborted = 0
    for k, values in enumerate(inputs_for(seed, 10)):
        jin, wout = tmp_path / ("in%d.json" % k), tmp_path / ("out%d.wtns" % k)
        jin.write_text(json.dumps(input_json(art, values)))
        r = subprocess.run([exe, str(jin), str(wout)], capture_output=True, timeout=60)
        try:
            ours = formats.wtns_bytes(I.compute_witness(prog, values))
        except I.WitnessError:
            ours = None
        if r.returncode == 0:
            assert ours is not None and wout.read_bytes() == ours, (seed, values)
            agreed += 1
        else:
            assert ours is None, (seed, values, r.stderr[-200:])
            aborted += 1
    assert agreed + aborted == 10


def test_reference_runtime_aborts_where_we_flag(tmp_path):
    """A failing `===` aborts the reference process (assert); the oracle raises status ASSERT."""
    exe = ref_binary("num2bits8")
    art = circuit("num2bits8")
    jin = tmp_path / "in.json"
    jin.write_text(json.dumps(input_json(art, [256])))
    r = subprocess.run([exe, str(jin), str(tmp_path / "o.wtns")], capture_output=True, timeout=60)
    assert r.returncode != 0
    with pytest.raises(I.WitnessError) as e:
        I.compute_witness(I.load(art.cvm), [256])
    assert e.value.status == I.ST_ASSERT


def test_reference_bench_mode_runs():
    exe = ref_binary("poseidon2")
    out = subprocess.run([exe, "--bench", "0.3", "7"], capture_output=True, text=True, timeout=60, check=True).stdout
    d = json.loads(out.strip().splitlines()[-1])
    assert d["witnesses"] > 0 and d["witnesses_per_s"] > 0
