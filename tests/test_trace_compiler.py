"""Host logic of the product (CVM parser, trace compiler, slot allocator) checked on CPU: the compiled
tape is executed by tests/tape_emulator.py and must reproduce the CVM oracle on every fixture circuit."""
import random

import pytest

from conftest import circuit
from oracle import cvm_interp as I
from oracle import fr_model as M
from tape_emulator import run_tape

CASES = {
    "earlyret": [[5, 3], [M.Q - 5, 3], [3, 5], [1000, 7], [0, 0], [255, 1], [7, 1000], [M.Q - 1, M.Q - 2]],
    "nbits": [[0], [1], [255], [256], [1 << 253], [M.Q - 1]],
    "countdown": [[0], [1], [17], [200]],
    "babyadd4": [[995203441582195749578291179787384436505546430278305826713579947235728471134, 5472060717959818805561601436314318772137091100104008585924551046643952123905, 5299619240641551281634865583518297030282874472190772894086521144482721001553, 16950150798460657717958625567821834550301663161624707787222815936182638968203], [0, 1, 0, 1], [3, 5, 7, 11]],
    "multiplier2": [[3, 11], [M.Q - 1, 5], [0, 0]],
    "multiplier4": [[2, 3, 4, 5], [M.Q - 1, M.Q - 2, 7, 0]],
    "num2bits8": [[0xA5], [0], [255], [256]],            # 256 does not fit: assert must fail
    "iszero": [[0], [7], [M.Q - 1]],
    "isequal": [[5, 5], [5, 6]],
    "lessthan8": [[3, 200], [200, 3], [7, 7]],
    "sum3cmp": [[1, 0, 1, 1], [0, 0, 0, 0]],
    # data-dependent array indices (Fr_toInt of a witness value): in range, out of range (reads / writes nothing that exists),
    # and values that do not fit an int (the reference asserts: ST_TOINT)
    "dynindex": [[5, 10, 21, 32, 43, 54, 65, 76, 87], [0] + [0] * 8, [7] + [3] * 8, [9, 1, 2, 3, 4, 5, 6, 7, 8],
                 [1 << 40] + [1] * 8, [M.Q - 1] + [1] * 8, [3] + [M.Q - 1] * 8],
    # a MIXED component array: every access to its components is a "mapped" location through the io-map
    "mixedarr": [list(range(1, 10)) + list(range(11, 20)) + [3, 5, 7], [M.Q - 1] * 21, [0] * 21],
    "opszoo": [[12345, 678, 3], [M.Q - 5, 17, 250], [0, 0, 0], [1 << 200, (1 << 253) + 5, 254]],
    "poseidon2": [[1, 2], [0, 0], [M.Q - 1, 12345678901234567890]],
    "poseidon2m": [[1, 2], [0, 0], [M.Q - 1, 12345678901234567890]],    # circomlib's shape: mapped accesses to ark[i]
    "widesums": [[M.Q - 1] * 40, [(1 << 253) - 1] * 40, list(range(40)),
                 [(M.Q - 1 - i) if i % 2 else ((1 << 224) - 1 + i) for i in range(40)]],
}


def oracle(prog, inp):
    try:
        return I.compute_witness(prog, inp), 0
    except I.WitnessError as e:
        return None, e.status


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("slots", [4, 7, 24])
def test_tape_matches_oracle(cvmlib, name, slots):
    from circom_cvm_b200 import engine as E
    art = circuit(name)
    prog = I.load(art.cvm)
    wc = E.WitnessCalculator(cvm_text=art.cvm, n_slots=slots)
    assert wc.n_wires == art.n_wires and wc.n_inputs == art.n_inputs
    tape, consts = wc.tape()
    rng = random.Random(7)
    cases = list(CASES[name])
    if name not in ("num2bits8", "sum3cmp", "lessthan8", "countdown"):
        cases += [[rng.randrange(M.Q) for _ in range(art.n_inputs)] for _ in range(3)]
    if name == "dynindex":
        cases += [[rng.randrange(8)] + [rng.randrange(M.Q) for _ in range(8)] for _ in range(6)]
    for inp in cases:
        w, st = oracle(prog, inp)
        rows, status = run_tape(tape, consts, wc.layout(), inp)
        assert status == st, (name, inp)
        if st == 0:
            assert rows == w, (name, inp)


def test_reference_op_counts(cvmlib):
    """N_mul reported by the tracer = dynamic ff.mul (+ff.div) count of the oracle interpreter."""
    from circom_cvm_b200 import engine as E
    for name in ("poseidon2", "opszoo", "lessthan8"):
        art = circuit(name)
        m = I.Machine(I.load(art.cvm))
        m.witness(CASES[name][0])
        wc = E.WitnessCalculator(cvm_text=art.cvm)
        if wc.info.dyn_branches == 0:
            assert wc.info.ref_mul == m.counters["mul"], name
            assert wc.info.cvm_instructions == m.counters["ops"], name
        else:   # both arms of a data-dependent branch are traced, so the tracer's count is an upper bound
            assert wc.info.ref_mul >= m.counters["mul"], name


def test_poseidon_known_answer():
    from tools.circuitgen.circuits import poseidon
    assert poseidon.poseidon_hash([1, 2]) == 7853200120776062878684798364095072458815029376092732009249414926327459813530
    art = circuit("poseidon2")
    w = I.compute_witness(I.load(art.cvm), [1, 2])
    assert w[1] == poseidon.poseidon_hash([1, 2])


def test_witness_satisfies_r1cs():
    for name, cases in CASES.items():
        art = circuit(name)
        prog = I.load(art.cvm)
        w, st = oracle(prog, cases[0])
        assert st == 0
        for (a, b, c) in art.constraints:
            ev = lambda lc: sum(v * w[k] for k, v in lc.items()) % M.Q
            assert (ev(a) * ev(b) - ev(c)) % M.Q == 0, name


def test_unsupported_programs_are_rejected_with_a_reason(cvmlib):
    from circom_cvm_b200 import engine as E
    q = "21888242871839275222246405745257275088548364400416034343698204186575808495617"
    head = "%%prime " + q + "\n%%signals 3\n%%start T_0\n%%witness 0 1 2\n%%template T_0 [ ff 0 ] [ ff 0 ] [2] [0]\n"
    # data-dependent `break` (a plain data-dependent `while` is unrolled under predicates instead, see the test above)
    text = head + "loop\nx_0 = get_signal i64.1\nif x_0\nbreak\nend\ncontinue\nend\n"
    with pytest.raises(E.CvmGpuError) as e:
        E.WitnessCalculator(cvm_text=text)
    assert e.value.code == -3 and "data-dependent" in str(e.value)
    # a signal STORE through a data-dependent address (loads and variable stores are traced: fixture `dynindex`)
    head1 = "%%prime " + q + "\n%%signals 4\n%%start T_0\n%%witness 0 1 2 3\n%%template T_0 [ ff 1 2 ] [ ff 0 ] [3] [0]\n"
    text = head1 + "x_0 = get_signal i64.1\nx_1 = ff.wrap_i64 x_0\nset_signal x_1 x_0\n"
    with pytest.raises(E.CvmGpuError) as e:
        E.WitnessCalculator(cvm_text=text)
    assert e.value.code == -3 and "data-dependent" in str(e.value)
    # the load form of the same program is accepted: out = in[in[0]] read through a computed index
    text = head1 + "x_0 = get_signal i64.1\nx_1 = ff.wrap_i64 x_0\nx_2 = get_signal x_1\nset_signal i64.0 x_2\n"
    wc = E.WitnessCalculator(cvm_text=text)
    tape, consts = wc.tape()
    for inp, out in (([1, 77], 1), ([2, 77], 77), ([0, 5], 0), ([3, 5], 0)):
        rows, status = run_tape(tape, consts, wc.layout(), inp)
        # (index 0 names the output itself, still unset: 0; index 3 is outside the component: nothing)
        assert status == 0 and rows[1] == out, (inp, rows)
    rows, status = run_tape(tape, consts, wc.layout(), [1 << 35, 1])
    assert status == 2          # CVMGPU_ST_TOINT: the reference's Fr_toInt asserts


def test_data_dependent_while_is_unrolled_under_predicates(cvmlib):
    """A `while` whose trip count depends on a signal is traced iteration by iteration inside the if-converted arm
    (tracer.hpp OP_CONTINUE), up to 260 iterations; a witness that needs more gets CVMGPU_ST_LOOP instead of a wrong
    value (the reference would simply keep looping)."""
    from circom_cvm_b200 import engine as E
    art = circuit("countdown")
    prog = I.load(art.cvm)
    wc = E.WitnessCalculator(cvm_text=art.cvm)
    assert wc.info.unrolled_iterations == 260
    tape, consts = wc.tape()
    for n in (0, 3, 259, 260):
        rows, status = run_tape(tape, consts, wc.layout(), [n])
        assert status == 0 and rows == I.compute_witness(prog, [n]), n
        assert rows[1] == sum(k * k for k in range(n + 1))
    for n in (261, 5000, M.Q - 1):
        rows, status = run_tape(tape, consts, wc.layout(), [n])
        assert status == 5, n


def test_eddsa_verifier_tape(cvmlib):
    """BASELINE config 4 (EdDSAPoseidonVerifier: Baby Jubjub scalar multiplications, 1320 `<--` divisions).  The tape
    (batched inversions, selects for 0/1 factors, fused dot products) reproduces the oracle's witness; forged
    signatures and out-of-range S raise the assert flag; the witness satisfies the R1CS."""
    from circom_cvm_b200 import engine as E
    from tools.circuitgen.circuits import babyjub, eddsa
    art = circuit("eddsa")
    prog = I.load(art.cvm)
    wc = E.WitnessCalculator(cvm_text=art.cvm)
    info = wc.info
    assert info.ref_div > 1000 and info.tape_inv * 4 < info.ref_div      # Montgomery's trick across independent divisions
    tape, consts = wc.tape()
    good = eddsa.sign(123456789, 987654321, 42)
    assert eddsa.verify(good)
    w = I.compute_witness(prog, good)
    rows, status = run_tape(tape, consts, wc.layout(), good)
    assert status == 0 and rows == w
    for (a, b, c) in art.constraints:
        ev = lambda lc: sum(v * w[k] for k, v in lc.items()) % M.Q
        assert (ev(a) * ev(b) - ev(c)) % M.Q == 0
    forged = list(good)
    forged[6] = 43
    big_s = list(good)
    big_s[3] = good[3] + babyjub.SUBORDER            # same point, but S >= l must be rejected by CompConstant
    disabled = list(forged)
    disabled[0] = 0                                   # enabled = 0 switches every check off
    for inp, want in ((forged, 1), (big_s, 1), (disabled, 0)):
        _w, st = oracle(prog, inp)
        rows, status = run_tape(tape, consts, wc.layout(), inp)
        assert status == st == want, inp
        if want == 0:
            assert rows == _w


def test_sha256_circuit_matches_hashlib(cvmlib):
    """Sha256(64 bits): one compression block, ~34K constraints.  The compiled tape must hash like hashlib
    (FIPS 180-4) and its witness must satisfy every R1CS constraint."""
    from circom_cvm_b200 import engine as E
    from tools.circuitgen.circuits import sha256
    art = circuit("sha256_64")
    assert len(art.constraints) > 30000
    wc = E.WitnessCalculator(cvm_text=art.cvm)
    tape, consts = wc.tape()
    rng = random.Random(5)
    for _ in range(2):
        bits = [rng.randrange(2) for _ in range(64)]
        rows, status = run_tape(tape, consts, wc.layout(), bits)
        assert status == 0
        w = rows
        assert w[1:257] == sha256.sha256_bits(bits)
        for (a, b, c) in art.constraints:
            ev = lambda lc: sum(v * w[k] for k, v in lc.items()) % M.Q
            assert (ev(a) * ev(b) - ev(c)) % M.Q == 0
    # a non-bit input breaks the bit decompositions downstream (BinSum: lin === lout)
    bad = [M.Q - 1] + [0] * 63
    rows, status = run_tape(tape, consts, wc.layout(), bad)
    assert status == 1


@pytest.mark.parametrize("name", ["num2bits8", "lessthan8", "opszoo", "sum3cmp", "earlyret", "poseidon2"])
def test_wire_typing_is_sound(cvmlib, name):
    """Wires the trace compiler types as 0/1 (cvmgpu_program_wire_types) really are 0 or 1 in the oracle's witness, for
    ordinary and extreme inputs."""
    from circom_cvm_b200 import engine as E
    art = circuit(name)
    prog = I.load(art.cvm)
    wc = E.WitnessCalculator(cvm_text=art.cvm)
    typed = wc.wire_is_bool()
    assert len(typed) == art.n_wires and typed[0] == 1          # wire 0 is the constant 1
    rng = random.Random(13)
    cases = list(CASES[name]) + [[rng.randrange(M.Q) for _ in range(art.n_inputs)] for _ in range(20)]
    for inp in cases:
        w, st = oracle(prog, inp)
        if st:
            continue
        for k, v in enumerate(w):
            assert not typed[k] or v in (0, 1), (name, inp, k, v)


def test_sha256_wires_are_almost_all_bits(cvmlib):
    from circom_cvm_b200 import engine as E
    art = circuit("sha256_64")
    typed = E.WitnessCalculator(cvm_text=art.cvm).wire_is_bool()
    assert typed.sum() > 0.99 * len(typed)
