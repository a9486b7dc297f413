"""GPU parity tests (run with -m gpu on the B200 box).  Everything goes through the C ABI and is compared
bit-exactly with the oracle (oracle/fr_model.py, oracle/cvm_interp.py) and the committed golden vectors."""
import random

import numpy as np
import pytest

from conftest import circuit, load_kats
from oracle import cvm_interp as I
from oracle import fr_model as M

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def E(cvmlib):
    from circom_cvm_b200 import engine
    assert engine.device_count() > 0, "no CUDA device: the product path has no CPU fallback"
    return engine


def test_device_field_ops_match_reference_vectors(E):
    """csrc/fr.cuh on the GPU against KATs produced by the reference's generic/fr.cpp."""
    by_op = {}
    for op, _af, _bf, a, b, out in load_kats():
        if op in ("isTrue", "toInt", "rawMMul", "rawToMont", "rawFromMont", "copy"):
            continue
        by_op.setdefault(op, []).append((a, b or 0, int(out, 16)))
    assert len(by_op) >= 24
    for op, rows in by_op.items():
        got = E.fr_device_op(op, [r[0] for r in rows], [r[1] for r in rows])
        for (a, b, exp), g in zip(rows, got):
            assert g == exp, (op, hex(a), hex(b), hex(g), hex(exp))


def test_device_mul_random_bulk(E):
    rng = random.Random(11)
    a = [rng.randrange(M.Q) for _ in range(4096)]
    b = [rng.randrange(M.Q) for _ in range(4096)]
    for op in ("mul", "add", "sub", "div"):
        got = E.fr_device_op(op, a, b)
        f = M.BINOPS[op]
        for x, y, g in zip(a, b, got):
            assert g == f(x, y), op


def test_device_field_ops_extreme_pairs(E):
    """All pairs of carry-stressing operands (all-ones limbs, q - small, powers of two) through the multiplier, the
    adder chains and the inversion."""
    top = (M.Q >> 224) - 1
    pool = [M.Q - 1, M.Q - 2, (top << 224) | ((1 << 224) - 1), (1 << 253) - 1, (1 << 252) - 1, (1 << 224) - 1,
            (1 << 32) - 1, (1 << 64) - 1, 0xffffffff00000000ffffffff00000000ffffffff00000000ffffffff, 1, 0, 2,
            M.Q - (1 << 32), M.Q - (1 << 224), (M.Q - 1) // 2, (M.Q + 1) // 2, 1 << 253, (1 << 253) + (1 << 32) - 1]
    pool += [pow(2, 256, M.Q), pow(2, 512, M.Q), M.Q - pow(2, 256, M.Q), pow(pow(2, 256, M.Q), -1, M.Q)]
    rng = random.Random(19)
    sq = pool + [rng.randrange(M.Q) for _ in range(2000)] + [(1 << k) % M.Q for k in range(0, 300, 7)]
    for x, g in zip(sq, E.fr_device_op("square", sq, sq)):
        assert g == x * x % M.Q, hex(x)
    a = [x for x in pool for _ in pool]
    b = [y for _ in pool for y in pool]
    for op in ("mul", "add", "sub", "div"):
        got = E.fr_device_op(op, a, b)
        f = M.BINOPS[op]
        for x, y, g in zip(a, b, got):
            assert g == f(x, y), (op, hex(x), hex(y))


CASES = {
    "earlyret": [[5, 3], [M.Q - 5, 3], [3, 5], [1000, 7], [0, 0], [255, 1], [7, 1000], [M.Q - 1, M.Q - 2]],
    "nbits": [[0], [1], [255], [256], [1 << 253], [M.Q - 1]],
    "countdown": [[0], [1], [17], [200]],
    "babyadd4": [[995203441582195749578291179787384436505546430278305826713579947235728471134, 5472060717959818805561601436314318772137091100104008585924551046643952123905, 5299619240641551281634865583518297030282874472190772894086521144482721001553, 16950150798460657717958625567821834550301663161624707787222815936182638968203], [0, 1, 0, 1], [3, 5, 7, 11]],
    "multiplier2": [[3, 11], [M.Q - 1, 5], [0, 0]],
    "multiplier4": [[2, 3, 4, 5]],
    "num2bits8": [[0xA5], [0], [255], [256], [1 << 100]],
    "iszero": [[0], [7], [M.Q - 1]],
    "isequal": [[5, 5], [5, 6]],
    "lessthan8": [[3, 200], [200, 3], [7, 7]],
    "sum3cmp": [[1, 0, 1, 1], [0, 0, 0, 0]],
    "dynindex": [[5, 10, 21, 32, 43, 54, 65, 76, 87], [0] + [0] * 8, [7] + [3] * 8, [9, 1, 2, 3, 4, 5, 6, 7, 8],
                 [1 << 40] + [1] * 8, [M.Q - 1] + [1] * 8, [3] + [M.Q - 1] * 8],          # data-dependent array indices, ST_TOINT
    "mixedarr": [list(range(1, 10)) + list(range(11, 20)) + [3, 5, 7], [M.Q - 1] * 21, [0] * 21],   # mapped accesses (io-map)
    "opszoo": [[12345, 678, 3], [M.Q - 5, 17, 250], [0, 0, 0], [1 << 200, (1 << 253) + 5, 254]],
    "poseidon2": [[1, 2], [0, 0], [M.Q - 1, 12345678901234567890]],
    "poseidon2m": [[1, 2], [0, 0], [M.Q - 1, 12345678901234567890]],    # circomlib's shape: mapped accesses to ark[i]
    "widesums": [[M.Q - 1] * 40, [(1 << 253) - 1] * 40, list(range(40)),
                 [(M.Q - 1 - i) if i % 2 else ((1 << 224) - 1 + i) for i in range(40)]],
}


def oracle_batch(art, rows):
    prog = I.load(art.cvm)
    wit, st = [], []
    for r in rows:
        try:
            wit.append(I.compute_witness(prog, r))
            st.append(0)
        except I.WitnessError as e:
            wit.append(None)
            st.append(e.status)
    return wit, st


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("slots", [0, 5])
def test_witness_batch_matches_oracle(E, name, slots):
    art = circuit(name)
    rng = random.Random(3)
    rows = list(CASES[name])
    if name not in ("num2bits8", "sum3cmp", "lessthan8", "countdown"):
        rows += [[rng.randrange(M.Q) for _ in range(art.n_inputs)] for _ in range(200)]
    else:
        rows = rows * 50                                   # ragged batch size, not a multiple of 128
    wc = E.WitnessCalculator(cvm_text=art.cvm, n_slots=slots)
    wt, st = wc.calculate(rows)
    exp_w, exp_st = oracle_batch(art, rows)
    assert list(st) == exp_st
    got = E.le_to_ints(wt)
    for g, e, s in zip(got, exp_w, exp_st):
        if s == 0:
            assert g == e


@pytest.mark.parametrize("name", ["poseidon2", "opszoo", "babyadd4", "num2bits8", "lessthan8"])
def test_two_witnesses_per_thread_mode(E, name):
    """The large-batch variant of the tape kernel (two witnesses per thread) on ragged batches: same results."""
    art = circuit(name)
    rng = random.Random(31)
    rows = list(CASES[name])
    if name not in ("num2bits8", "lessthan8"):
        rows += [[rng.randrange(M.Q) for _ in range(art.n_inputs)] for _ in range(301)]
    else:
        rows = rows * 67
    wc = E.WitnessCalculator(cvm_text=art.cvm)
    E.set_tape_mode(2)
    try:
        wt, st = wc.calculate(rows)
    finally:
        E.set_tape_mode(0)
    exp_w, exp_st = oracle_batch(art, rows)
    assert list(st) == exp_st
    for g, e, s_ in zip(E.le_to_ints(wt), exp_w, exp_st):
        if s_ == 0:
            assert g == e


def test_empty_batch(E):
    wc = E.WitnessCalculator(cvm_text=circuit("multiplier2").cvm)
    wt, st = wc.calculate(np.zeros((0, 2, 32), dtype=np.uint8))
    assert wt.shape == (0, 4, 32) and st.shape == (0,)


def test_inputs_not_reduced_are_reduced_like_str2element(E):
    """Fr_str2element reduces mod q (bn128/fr.cpp:56-62)."""
    wc = E.WitnessCalculator(cvm_text=circuit("multiplier2").cvm)
    rows = [[M.Q + 3, 11], [(1 << 256) - 1, 2]]
    wt, st = wc.calculate(rows)
    got = E.le_to_ints(wt)
    assert got[0] == [1, 33, 3, 11]
    x = ((1 << 256) - 1) % M.Q
    assert got[1] == [1, (2 * x) % M.Q, x, 2]


def test_wtns_file_is_what_the_reference_writes(E, tmp_path):
    """Byte-identical .wtns (common/main.cpp:286-332) for the reference's own walk-through input
    (mkdocs/docs/getting-started/computing-the-witness.md:22-56: a=3, b=11)."""
    from circom_cvm_b200 import formats
    wc = E.WitnessCalculator(cvm_text=circuit("multiplier2").cvm)
    wt, st = wc.calculate([[3, 11]])
    p = tmp_path / "witness.wtns"
    wc.write_wtns(str(p), wt[0])
    assert p.read_bytes() == formats.wtns_bytes([1, 33, 3, 11])


def test_poseidon_published_vector(E):
    wc = E.WitnessCalculator(cvm_text=circuit("poseidon2").cvm)
    wt, st = wc.calculate([[1, 2]])
    assert st[0] == 0
    assert E.le_to_ints(wt)[0][1] == 7853200120776062878684798364095072458815029376092732009249414926327459813530


def _write_r1cs(art, path):
    from circom_cvm_b200 import formats
    formats.write_r1cs(str(path), art.constraints, art.n_wires, art.n_pub_out, art.n_pub_in, art.n_prv_in, art.witness,
                       n_labels=art.n_signals)


def test_long_dot_products_with_many_slots(E):
    """With 24 slots the tape fuses linear combinations of up to 16 terms into one lazy-reduction dot product (longer
    ones are split): carry-stressing inputs against the oracle."""
    art = circuit("widesums")
    rng = random.Random(41)
    rows = list(CASES["widesums"]) + [[rng.randrange(M.Q) for _ in range(40)] for _ in range(300)]
    wc = E.WitnessCalculator(cvm_text=art.cvm, n_slots=24)
    assert wc.info.tape_dot >= 6
    wt, st = wc.calculate(rows)
    assert not st.any()
    got = E.le_to_ints(wt)
    prog = I.load(art.cvm)
    for b in list(range(8)) + [100, 303]:
        assert got[b] == I.compute_witness(prog, rows[b]), b


@pytest.mark.parametrize("name", ["multiplier2", "lessthan8", "sum3cmp", "poseidon2", "num2bits8", "babyadd4", "widesums"])
def test_r1cs_check_accepts_valid_and_pinpoints_invalid(E, name, tmp_path):
    art = circuit(name)
    _write_r1cs(art, tmp_path / "c.r1cs")
    r = E.R1cs(str(tmp_path / "c.r1cs"))
    rng = random.Random(5)
    rows = [CASES[name][0]] * 3
    if name in ("multiplier2", "poseidon2", "babyadd4", "widesums"):
        rows += [[rng.randrange(M.Q) for _ in range(art.n_inputs)] for _ in range(130)]
    wc = E.WitnessCalculator(cvm_text=art.cvm)
    wt, st = wc.calculate(rows)
    assert not st.any()
    bad = r.check(wt)
    assert (bad == E.NO_BAD).all()
    # corrupt one wire of some witnesses; the CPU evaluation of the same .r1cs tells which constraint breaks first
    wt2 = wt.copy()
    victims = [0, len(rows) - 1]
    for v in victims:
        wire = 1 + (v % (art.n_wires - 1))
        wt2[v, wire, 0] ^= 1
    bad = r.check(wt2)
    vals = E.le_to_ints(wt2)
    for b in range(len(rows)):
        w = vals[b]
        first = E.NO_BAD
        for ci, (a, bb, c) in enumerate(art.constraints):
            ev = lambda lc: sum(v * w[k] for k, v in lc.items()) % M.Q
            if (ev(a) * ev(bb) - ev(c)) % M.Q:
                first = ci
                break
        assert bad[b] == first, (name, b)
    assert all(bad[v] != E.NO_BAD for v in victims)


def test_device_api_with_torch_buffers(E):
    """The device-pointer entry points on torch-allocated buffers (value store layout, export, check)."""
    import torch
    art = circuit("poseidon2")
    wc = E.WitnessCalculator(cvm_text=art.cvm)
    B = 1000
    rng = random.Random(9)
    rows = [[rng.randrange(M.Q) for _ in range(2)] for _ in range(B)]
    inp = torch.from_numpy(E.ints_to_le(rows, 2)).cuda()
    store = torch.empty(wc.store_bytes(B), dtype=torch.uint8, device="cuda")
    status = torch.empty(B, dtype=torch.int32, device="cuda")
    wtns = torch.empty((B, wc.n_wires, 32), dtype=torch.uint8, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    wc.run_dev(inp, B, B, store, status, s)
    wc.export_dev(store, B, B, wtns, s)
    torch.cuda.synchronize()
    assert int(status.abs().sum()) == 0
    from tools.circuitgen.circuits import poseidon
    got = E.le_to_ints(wtns.cpu().numpy())
    for b in (0, 1, 511, 999):
        assert got[b][1] == poseidon.poseidon_hash(rows[b])
        assert got[b] == I.compute_witness(I.load(art.cvm), rows[b])


def test_output_selector_and_typed_check_agree(E, tmp_path):
    """cvmgpu_witness_batch_select returns only the requested wire range (here: the public part and an inner slice),
    identical to the same columns of the full export; the typed check (bit rows, integer constraints) reports the same
    first violated constraint as the plain check of the exported rows for inputs that break constraints."""
    art = circuit("sha256_64")
    wc = E.WitnessCalculator(cvm_text=art.cvm)
    _write_r1cs(art, tmp_path / "s.r1cs")
    r = E.R1cs(str(tmp_path / "s.r1cs"))
    rng = random.Random(21)
    rows = [[rng.randrange(2) for _ in range(64)] for _ in range(70)]
    rows[3] = [M.Q - 1] + [0] * 63      # not a bit, and too large for the 35-bit sums: their decomposition constraint fails
    rows[40][0] = M.Q - 5
    rows[41][63] = 5                    # not a bit either, but every constraint still holds (the inputs are unconstrained)
    wt, st, bad = wc.calculate_checked(rows, r)
    assert (bad[[3, 40]] != E.NO_BAD).all() and (np.delete(bad, [3, 40]) == E.NO_BAD).all()
    assert (r.check(wt) == bad).all()
    vals = E.le_to_ints(wt[[3, 40]])
    for k, b in enumerate((3, 40)):
        w = vals[k]
        ev = lambda lc: sum(v * w[j] for j, v in lc.items()) % M.Q
        first = next(ci for ci, (a, bb, c) in enumerate(art.constraints) if (ev(a) * ev(bb) - ev(c)) % M.Q)
        assert bad[b] == first, b
    B = len(rows)
    inp = E.ints_to_le(rows, 64)
    for wire0, n_sel in ((0, 1 + 256 + 64), (1000, 777)):
        out = np.empty((B, n_sel, 32), dtype=np.uint8)
        st2 = np.empty(B, dtype=np.uint32)
        bad2 = np.empty(B, dtype=np.uint32)
        wc.calculate_select_into(inp, wire0, n_sel, out, st2, r, bad2)
        assert (out == wt[:, wire0:wire0 + n_sel]).all() and (st2 == st).all() and (bad2 == bad).all()


def test_multi_device_driver(E, tmp_path):
    """cvmgpu_witness_batch_multi (one host thread + streams per device of the mask): the same answers as the single-device
    call, on every visible device."""
    art = circuit("poseidon2")
    wc = E.WitnessCalculator(cvm_text=art.cvm)
    _write_r1cs(art, tmp_path / "p.r1cs")
    r = E.R1cs(str(tmp_path / "p.r1cs"))
    rng = random.Random(33)
    rows = [[rng.randrange(M.Q) for _ in range(2)] for _ in range(1001)]
    wt, st, bad = wc.calculate_checked(rows, r)
    inp = E.ints_to_le(rows, 2)
    n_dev = E.device_count()
    for mask in sorted({1, (1 << n_dev) - 1}):
        wt2 = np.zeros_like(wt)
        st2 = np.full_like(st, 7)
        bad2 = np.zeros_like(bad)
        wc.calculate_multi_into(inp, mask, 0, wc.n_wires, wt2, st2, r, bad2)
        assert (wt2 == wt).all() and (st2 == st).all() and (bad2 == bad).all(), mask
    with pytest.raises(E.CvmGpuError):
        wc.calculate_multi_into(inp, 1 << n_dev, 0, wc.n_wires, wt2, st2, r, bad2)    # a device that is not there


def test_plain_c_caller_on_the_gpu(tmp_path):
    from circom_cvm_b200 import build, engine
    from test_abi import test_plain_c_caller
    build.build()
    test_plain_c_caller(engine.lib(), tmp_path)


def test_eddsa_batch(E, tmp_path):
    """Config 4 shape: EdDSAPoseidonVerifier over signatures from the integer signer, 1 in 8 forged: per-witness
    flags, witness parity with the oracle, and the R1CS check."""
    from tools.circuitgen.circuits import eddsa
    art = circuit("eddsa")
    rng = random.Random(23)
    rows, forged = [], []
    for i in range(72):
        sig = eddsa.sign(rng.randrange(1, 1 << 250), rng.randrange(1, 1 << 250), rng.randrange(M.Q))
        if i % 8 == 5:
            sig[6] = (sig[6] + 1) % M.Q
            forged.append(i)
        rows.append(sig)
    wc = E.WitnessCalculator(cvm_text=art.cvm)
    _write_r1cs(art, tmp_path / "e.r1cs")
    r = E.R1cs(str(tmp_path / "e.r1cs"))
    wt, st, bad = wc.calculate_checked(rows, r)
    ok = [i for i in range(len(rows)) if i not in forged]
    assert (st[forged] == E.ST_ASSERT).all() and (bad[forged] != E.NO_BAD).all()
    assert not st[ok].any() and (bad[ok] == E.NO_BAD).all()
    prog = I.load(art.cvm)
    got = E.le_to_ints(wt[[0, 1, 70]])
    for g, i in zip(got, (0, 1, 70)):
        assert g == I.compute_witness(prog, rows[i])


@pytest.mark.parametrize("name,doc", [
    ("multiplier2", {"a": "3", "b": "11"}),
    ("poseidon2", {"inputs": ["1", "0x2"]}),
    ("babyadd4", {"q": ["5299619240641551281634865583518297030282874472190772894086521144482721001553",
                        "16950150798460657717958625567821834550301663161624707787222815936182638968203"],
                  "p": ["995203441582195749578291179787384436505546430278305826713579947235728471134",
                        "5472060717959818805561601436314318772137091100104008585924551046643952123905"]}),
])
def test_process_abi_drop_in(E, name, doc, tmp_path):
    """`python -m circom_cvm_b200 <circuit>.cvm input.json out.wtns` against the reference's own
    `./<circuit> input.json out.wtns` (oracle/_ref, when built): same input file, byte-identical output file."""
    import json
    import os
    import subprocess
    import sys
    from conftest import ROOT
    from tools.circuitgen.build import write_artifact
    art = circuit(name)
    paths = write_artifact(art, str(tmp_path), with_cpp=True)          # .cvm + .dat side by side
    jin = tmp_path / "input.json"
    jin.write_text(json.dumps(doc))
    out = tmp_path / "ours.wtns"
    env = dict(os.environ, PYTHONPATH=ROOT)
    subprocess.run([sys.executable, "-m", "circom_cvm_b200", paths["cvm"], str(jin), str(out)], check=True, env=env,
                   cwd=ROOT, timeout=300)
    # names resolved from the `circom --sym` file instead of the .dat, with the R1CS check on: same bytes
    out2 = tmp_path / "ours_sym.wtns"
    subprocess.run([sys.executable, "-m", "circom_cvm_b200", paths["cvm"], str(jin), str(out2), "--sym", paths["sym"],
                    "--r1cs", paths["r1cs"]], check=True, env=env, cwd=ROOT, timeout=300)
    assert out2.read_bytes() == out.read_bytes()
    # the native host program (csrc/calc_main.cpp) takes the same files
    from circom_cvm_b200 import build as cbuild
    cbuild.build()
    nout = tmp_path / "native.wtns"
    subprocess.run([cbuild.CALC, paths["cvm"], str(jin), str(nout)], check=True, timeout=300)
    assert nout.read_bytes() == out.read_bytes()
    # the program text alone (its ;;%%create_cmp / ;;%%main_input lines): no .dat, no .cpp, no .sym next to it
    lone = tmp_path / "lone"
    lone.mkdir()
    with open(paths["cvm"]) as fsrc:
        (lone / "c.cvm").write_text(fsrc.read())
    subprocess.run([sys.executable, "-m", "circom_cvm_b200", str(lone / "c.cvm"), str(jin), str(lone / "py.wtns")], check=True,
                   env=env, cwd=ROOT, timeout=300)
    subprocess.run([cbuild.CALC, str(lone / "c.cvm"), str(jin), str(lone / "native.wtns")], check=True, timeout=300)
    assert (lone / "py.wtns").read_bytes() == out.read_bytes() == (lone / "native.wtns").read_bytes()
    ref = os.path.join(ROOT, "oracle", "_ref", name)
    if os.path.exists(ref):
        rout = tmp_path / "ref.wtns"
        subprocess.run([ref, str(jin), str(rout)], check=True, timeout=60)
        assert out.read_bytes() == rout.read_bytes()
    else:
        from circom_cvm_b200 import formats
        from circom_cvm_b200.inputs import InputMap, row_from_json
        with open(paths["dat"], "rb") as f:
            row = row_from_json(InputMap(f.read(), art.witness, art.input_start, art.n_inputs), doc)
        assert out.read_bytes() == formats.wtns_bytes(I.compute_witness(I.load(art.cvm), row))


@pytest.mark.parametrize("name,doc", [
    ("poseidon2m", {"inputs": ["1", "2"]}),
    ("mixedarr", {"a": [str(v) for v in range(1, 10)], "b": [str(v) for v in range(11, 20)], "w": ["3", "5", "7"]}),
])
def test_process_drop_in_on_unpatched_compiler_outputs(E, name, doc, tmp_path):
    """The three files an UNPATCHED compile leaves (--cvm text as the fork prints it: no component creation, no io-map;
    the generated .cpp; the .dat) for circuits with a mixed component array: both entry-point programs pick the .cpp and
    .dat up next to the .cvm and write the bytes the reference's own binary writes."""
    import json
    import os
    import subprocess
    import sys
    from conftest import ROOT
    from tools.circuitgen.build import faithful_cvm, write_artifact
    art = circuit(name)
    paths = write_artifact(art, str(tmp_path), with_cpp=True)
    text = faithful_cvm(art)
    assert ";;%%" not in text and "get_template_id" in text
    with open(paths["cvm"], "w") as f:
        f.write(text)
    jin = tmp_path / "input.json"
    jin.write_text(json.dumps(doc))
    out = tmp_path / "ours.wtns"
    env = dict(os.environ, PYTHONPATH=ROOT)
    subprocess.run([sys.executable, "-m", "circom_cvm_b200", paths["cvm"], str(jin), str(out), "--r1cs", paths["r1cs"]],
                   check=True, env=env, cwd=ROOT, timeout=300)
    from circom_cvm_b200 import build as cbuild
    cbuild.build()
    nout = tmp_path / "native.wtns"
    subprocess.run([cbuild.CALC, paths["cvm"], str(jin), str(nout), "--r1cs", paths["r1cs"]], check=True, timeout=300)
    assert nout.read_bytes() == out.read_bytes()
    ref = os.path.join(ROOT, "oracle", "_ref", name)
    if os.path.exists(ref):
        rout = tmp_path / "ref.wtns"
        subprocess.run([ref, str(jin), str(rout)], check=True, timeout=60)
        assert out.read_bytes() == rout.read_bytes()
    if name == "poseidon2m":
        from circom_cvm_b200 import formats
        from tools.circuitgen.circuits import poseidon
        assert formats.read_wtns(out.read_bytes())["values"][1] == poseidon.poseidon_hash([1, 2])
    # without the .cpp/.dat next to it the same text cannot run, and says why
    lone = tmp_path / "lone"
    lone.mkdir()
    (lone / "c.cvm").write_text(text)
    r = subprocess.run([cbuild.CALC, str(lone / "c.cvm"), str(jin), str(lone / "o.wtns")], capture_output=True, timeout=300)
    assert r.returncode != 0 and b"never created" in r.stderr


def test_sha256_batch(E, tmp_path):
    """Config 3 shape (bit-decomposition heavy): Sha256 over 64-bit messages, checked against hashlib for every
    witness, against the CVM oracle for one, and through the R1CS check."""
    from tools.circuitgen.circuits import sha256
    art = circuit("sha256_64")
    rng = random.Random(17)
    rows = [[rng.randrange(2) for _ in range(64)] for _ in range(300)]
    rows[7] = [M.Q - 1] + [0] * 63                        # not a bit: asserts must flag it
    wc = E.WitnessCalculator(cvm_text=art.cvm)
    _write_r1cs(art, tmp_path / "s.r1cs")
    r = E.R1cs(str(tmp_path / "s.r1cs"))
    wt, st, bad = wc.calculate_checked(rows, r)
    assert st[7] == E.ST_ASSERT and bad[7] != E.NO_BAD
    ok = [i for i in range(len(rows)) if i != 7]
    assert not st[ok].any() and (bad[ok] == E.NO_BAD).all()
    vals = E.le_to_ints(wt[:, 1:257, :])
    for i in ok:
        assert vals[i] == sha256.sha256_bits(rows[i])
    full = E.le_to_ints(wt[3:4])[0]
    assert full == I.compute_witness(I.load(art.cvm), rows[3])
    # corrupted wires (small-coefficient and +-1 paths of the check): same first violated constraint as a CPU walk
    wt2 = wt[:3].copy()
    wt2[0, 300, 0] ^= 1
    wt2[1, art.n_wires - 5, 0] ^= 1
    bad2 = r.check(wt2)
    vals2 = E.le_to_ints(wt2)
    for b in range(3):
        w = vals2[b]
        first = E.NO_BAD
        for ci, (a, bb, c) in enumerate(art.constraints):
            ev = lambda lc: sum(v * w[k] for k, v in lc.items()) % M.Q
            if (ev(a) * ev(bb) - ev(c)) % M.Q:
                first = ci
                break
        assert bad2[b] == first, b
    assert bad2[2] == E.NO_BAD and bad2[0] != E.NO_BAD and bad2[1] != E.NO_BAD


# ---- BASELINE.json sizes: properties that do not need an oracle run per witness ----------------------------------
def _run_full(E, art, inputs_dev, B, tmp_path):
    import torch
    wc = E.WitnessCalculator(cvm_text=art.cvm)
    _write_r1cs(art, tmp_path / "full.r1cs")
    r = E.R1cs(str(tmp_path / "full.r1cs"))
    store = torch.empty(wc.store_bytes(B), dtype=torch.uint8, device="cuda")
    status = torch.empty(B, dtype=torch.int32, device="cuda")
    bad = torch.empty(B, dtype=torch.int32, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    wc.run_dev(inputs_dev, B, B, store, status, s)
    r.check_store_dev(wc, store, B, B, bad, s)
    torch.cuda.synchronize()
    return wc, r, store, status, bad


def test_poseidon_at_baseline_size(E, tmp_path):
    """Config 2 (Poseidon(2), 1 048 576 random inputs): no flag raised, every witness satisfies all 634 constraints (the
    check is an independent evaluation of the .r1cs), sampled outputs equal the plain-integer Poseidon, a second run
    writes the same store (determinism), and a corrupted store is caught by the check."""
    import torch
    from tools.circuitgen.circuits import poseidon
    art = circuit("poseidon2")
    B = 1 << 20
    g = torch.Generator(device="cuda")
    g.manual_seed(0xC1C00001)
    inp = torch.randint(0, 256, (B, 2, 32), dtype=torch.uint8, device="cuda", generator=g)
    inp[:, :, 31] &= 0x1F
    wc, r, store, status, bad = _run_full(E, art, inp, B, tmp_path)
    assert int((status != 0).sum()) == 0 and int((bad != -1).sum()) == 0
    sample = [0, 1, 31, 32, 4095]
    idx = torch.tensor(sample, device="cuda")
    sub_in = E.le_to_ints(inp[idx].cpu().numpy())
    # export works on a contiguous prefix of the batch: check the sampled witnesses of the first 4 096
    wt = torch.empty((1 << 12, wc.n_wires, 32), dtype=torch.uint8, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    wc.export_dev(store, 1 << 12, B, wt, s)
    torch.cuda.synchronize()
    head = E.le_to_ints(wt[[0, 1, 31, 32, 4095]].cpu().numpy())
    for k, b in enumerate([0, 1, 31, 32, 4095]):
        assert head[k][1] == poseidon.poseidon_hash(sub_in[k])
    checksum = store.view(torch.int64).sum().item()
    status2 = torch.empty_like(status)
    wc.run_dev(inp, B, B, store, status2, s)
    torch.cuda.synchronize()
    assert store.view(torch.int64).sum().item() == checksum
    # flip one byte of one wire of witnesses 5 and B-1: exactly those are reported
    view = store.view(torch.uint8)
    row3 = int(wc.wire_rows()[3])
    assert not row3 & E.ROW_BIT                # wire 3 of Poseidon(2) is a field row
    for b in (5, B - 1):
        off = ((row3 * 2) * B + b) * 16       # field row of wire 3, low half, witness b (layout in include/cvmgpu.h)
        view[off] ^= 1
    r.check_store_dev(wc, store, B, B, bad, s)
    torch.cuda.synchronize()
    hit = torch.nonzero(bad != -1).flatten().tolist()
    assert hit == [5, B - 1]


def test_sha256_at_baseline_size(E, tmp_path):
    """Config 3 (Sha256(512 bits), 65 536 random messages): no flag, all 68 640 constraints hold for every witness, and
    the digest wires of sampled witnesses equal hashlib's."""
    import hashlib
    import torch
    from tools.circuitgen.build import compile_circuit
    from tools.circuitgen.circuits import sha256
    art = compile_circuit(sha256.Sha256, (512,), name="sha256_512")
    B = 1 << 16
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    free, _total = torch.cuda.mem_get_info()
    if free < 12e9:
        pytest.skip("needs ~10 GB of free device memory (typed value store 5.6 GB + inputs)")
    g = torch.Generator(device="cuda")
    g.manual_seed(0xC1C00002)
    bits = torch.randint(0, 2, (B, 512), dtype=torch.uint8, device="cuda", generator=g)
    inp = torch.zeros((B, 512, 32), dtype=torch.uint8, device="cuda")
    inp[:, :, 0] = bits
    try:
        wc, r, store, status, bad = _run_full(E, art, inp, B, tmp_path)
    except torch.cuda.OutOfMemoryError:
        pytest.skip("not enough free device memory for the 64 K batch")
    assert int((status != 0).sum()) == 0 and int((bad != -1).sum()) == 0
    wt = torch.empty((64, wc.n_wires, 32), dtype=torch.uint8, device="cuda")
    wc.export_dev(store, 64, B, wt, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    digests = wt[:, 1:257, 0].cpu().numpy()            # 256 output bits, each a 0/1 field element
    assert int(wt[:, 1:257, 1:].sum()) == 0
    msgs = bits[:64].cpu().numpy()
    for k in range(64):
        m = bytes(int("".join(str(int(x)) for x in msgs[k, 8 * j:8 * j + 8]), 2) for j in range(64))
        want = "".join(format(byte, "08b") for byte in hashlib.sha256(m).digest())
        assert "".join(str(int(x)) for x in digests[k]) == want, k
    # typed store: 97.8 % of the wires are bit rows (one word per 32 witnesses); the exported rows are still full 32-byte
    # canonical values, and the host-side check of those rows (plain layout, field arithmetic only) agrees with the typed one
    info = wc.info
    assert info.n_bool_wires > 0.97 * wc.n_wires and info.n_brows >= info.n_bool_wires
    assert wc.store_bytes(B) < 8e9
    assert (r.check(wt.cpu().numpy()) == E.NO_BAD).all()
    ri = r.refresh_info()
    assert ri.bound_int_constraints > 0.8 * r.n_constraints
    # flip the bit of one bit-row wire for witnesses 7 and B-1: exactly those are reported
    rows = wc.wire_rows()
    wire = int(np.nonzero(rows & E.ROW_BIT)[0][5000])
    brow = int(rows[wire]) & ~E.ROW_BIT
    bits_base = info.n_frows * 32 * B
    words = store[bits_base:].view(torch.int32)
    for b in (7, B - 1):
        words[(b >> 5) * info.n_brows + brow] ^= (1 << (b & 31)) if (b & 31) < 31 else -(1 << 31)
    r.check_store_dev(wc, store, B, B, bad, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert torch.nonzero(bad != -1).flatten().tolist() == [7, B - 1]


def _extreme_values(rng):
    """Field elements that stress carry propagation: all-ones limbs, q - small, powers of two."""
    top = (M.Q >> 224) - 1
    pool = [M.Q - 1, M.Q - 2, (top << 224) | ((1 << 224) - 1), (1 << 253) - 1, (1 << 252) - 1, (1 << 224) - 1,
            (1 << 32) - 1, (1 << 64) - 1, 0xffffffff00000000ffffffff00000000ffffffff00000000ffffffff, 1, 0,
            M.Q - (1 << 32), M.Q - (1 << 224), (M.Q - 1) // 2, (M.Q + 1) // 2]
    return pool + [rng.randrange(M.Q) for _ in range(5)]


@pytest.mark.gpu
def test_r1cs_check_extreme_coefficients_and_values(E, tmp_path):
    """Synthetic constraint system whose linear combinations hit every coefficient class with carry-stressing values:
    16- and 17-term general dot products of (q-1)-like coefficients (the lazy accumulator's window-carry counters),
    long runs of 32-bit coefficients of both signs (the small-scalar accumulator), +-2^k, and general constants on
    wire 0.  The expected product wires come from Python integers."""
    from circom_cvm_b200 import formats
    rng = random.Random(77)
    ext = _extreme_values(rng)
    n_in = 48
    cons = []

    def lc_general(n, with_const):
        lc = {1 + rng.randrange(n_in): ext[rng.randrange(len(ext) - 5)] or 5 for _ in range(n * 3)}
        lc = dict(list(lc.items())[:n])
        for k in list(lc):
            if lc[k] in (0, 1, M.Q - 1) or lc[k] < (1 << 32) or M.Q - lc[k] < (1 << 32):
                lc[k] = M.Q - 1 - (1 << 40) - k
        if with_const:
            lc[0] = M.Q - 12345678901234567890
        return lc

    def lc_small(n, sign):
        out = {}
        while len(out) < n:
            c = rng.choice([(1 << 32) - 1, (1 << 32) - 2, 3, 5, 0x80000000, rng.randrange(9, 1 << 32)])
            out[1 + rng.randrange(n_in)] = c if sign > 0 else M.Q - c
        return out

    def lc_pow2(n):
        return {1 + rng.randrange(n_in): rng.choice([1, M.Q - 1, 2, M.Q - 2, 4, M.Q - 4, 8, M.Q - 8]) for _ in range(n)}

    shapes = [
        (lc_general(16, False), lc_general(17, True)),
        (lc_general(33, True), lc_general(1, False)),
        (lc_small(40, +1), lc_small(40, -1)),
        ({**lc_small(6, +1), **lc_general(5, True)}, {**lc_pow2(7), **lc_small(9, -1)}),
        ({**lc_pow2(5), **lc_general(16, False), **lc_small(4, 1)}, lc_pow2(1)),
        ({0: M.Q - 7}, lc_general(3, True)),              # constant-only combination times a short dot product
        (lc_general(2, True), {0: (1 << 200) + 9}),
    ]
    for g in (lc_general(5, True), {**lc_pow2(2), **lc_small(5, -1)}, {3: 1}, {0: M.Q - 3, 7: 1}):
        shapes.append((g, dict(g)))                       # B repeats A: evaluated once and squared
    for a, b in shapes:
        out_wire = 1 + n_in + len(cons)
        cons.append((a, b, {out_wire: 1}))
    # linear constraints: 0 = C with C = long mixed combination - its own value wire
    for k in range(3):
        out_wire = 1 + n_in + len(cons)
        c = {**lc_general(16 + k, k == 1), **lc_small(7, -1), **lc_pow2(3)}
        c[out_wire] = M.Q - 1
        cons.append(({}, {}, c))
    n_wires = 1 + n_in + len(cons)
    path = tmp_path / "x.r1cs"
    formats.write_r1cs(str(path), cons, n_wires, 0, 0, n_in, list(range(n_wires)))
    r = E.R1cs(str(path))
    info = r.info.asdict()
    assert info["nnz_const"] >= 5 and info["nnz_small"] >= 80 and info["n_squares"] == 4

    ev = lambda lc, w: sum(v * w[k] for k, v in lc.items()) % M.Q
    B = 96
    rows = []
    for b in range(B):
        w = [1] + [ext[(b + 3 * i) % len(ext)] if (b + i) % 4 else rng.randrange(M.Q) for i in range(n_in)]
        if b == 0:
            w = [1] + [M.Q - 1] * n_in
        w += [0] * len(cons)
        for ci, (a, bb, c) in enumerate(cons):
            ow = 1 + n_in + ci
            if a:
                w[ow] = ev(a, w) * ev(bb, w) % M.Q
            else:
                w[ow] = ev({k: v for k, v in c.items() if k != ow}, w)
        rows.append(w)
    wt = E.ints_to_le(rows, n_wires).reshape(B, n_wires, 32)
    bad = r.check(wt)
    assert (bad == E.NO_BAD).all(), bad
    # every constraint must also be *violated* when its value wire is off by one (the check is not vacuous)
    for ci in range(len(cons)):
        wt2 = wt.copy()
        wt2[:, 1 + n_in + ci, 0] ^= 1
        bad = r.check(wt2)
        assert (bad == ci).all(), (ci, bad[:8])
