"""Random circuits over the whole operator set: compiled tape (tests/tape_emulator.py) vs the CVM oracle."""
import pytest

from fuzz_circuits import inputs_for, make_circuit
from oracle import cvm_interp as I
from tape_emulator import run_tape


@pytest.mark.parametrize("seed", range(80))
def test_random_circuit_tape_matches_oracle(cvmlib, seed):
    from circom_cvm_b200 import engine as E
    from tools.circuitgen.build import compile_circuit
    art = compile_circuit(make_circuit(seed), (), name="fuzz%d" % seed)
    prog = I.load(art.cvm)
    for slots in (4, 9):
        wc = E.WitnessCalculator(cvm_text=art.cvm, n_slots=slots)
        tape, consts = wc.tape()
        for inp in inputs_for(seed, 6):
            try:
                w, st = I.compute_witness(prog, inp), 0
            except I.WitnessError as e:
                w, st = None, e.status
            rows, status = run_tape(tape, consts, wc.layout(), inp)
            if st == 0:
                assert status == 0, (seed, inp)
                assert rows == w, (seed, inp)
            else:
                # the oracle stops at the first failure; the tape runs on and reports the first one it meets in ITS order
                assert status != 0, (seed, inp)


@pytest.mark.parametrize("seed", range(300, 340))
def test_random_bit_circuit_tape_matches_oracle(cvmlib, seed):
    """The generator's bit-heavy mode (sums of bits taken apart again, boolean polynomials): the typed paths of the tape
    compiler -- bit-slot file, small integers, fused sums, warp-cooperative groups -- against the oracle, with the automatic
    slot choice and with tiny slot files (spills of every type)."""
    from circom_cvm_b200 import engine as E
    from tools.circuitgen.build import compile_circuit
    art = compile_circuit(make_circuit(seed, n_stmts=50, bits=True), (), name="fuzzb%d" % seed)
    prog = I.load(art.cvm)
    for slots in (0, 4):
        wc = E.WitnessCalculator(cvm_text=art.cvm, n_slots=slots)
        tape, consts = wc.tape()
        for inp in inputs_for(seed, 5):
            try:
                w, st = I.compute_witness(prog, inp), 0
            except I.WitnessError as e:
                w, st = None, e.status
            rows, status = run_tape(tape, consts, wc.layout(), inp)
            if st == 0:
                assert status == 0 and rows == w, (seed, inp)
            else:
                assert status != 0, (seed, inp)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(400, 424))
def test_random_bit_circuit_gpu_matches_oracle_and_checks_agree(cvmlib, seed, tmp_path):
    """Bit-heavy random circuits on the device: witnesses and status words against the oracle; the typed check (bit rows,
    integer and truth-table constraints) must report the same first violated constraint as the plain check of the
    exported rows, and as a walk of the constraints in Python for the witnesses that violate something."""
    from circom_cvm_b200 import engine as E
    from circom_cvm_b200 import formats
    from oracle import fr_model as M
    from tools.circuitgen.build import compile_circuit
    art = compile_circuit(make_circuit(seed, n_stmts=60, bits=True), (), name="fuzzb%d" % seed, constraint_assert_disabled=True)
    prog = I.load(art.cvm)
    rows = inputs_for(seed, 70)
    p = str(tmp_path / "f.r1cs")
    formats.write_r1cs(p, art.constraints, art.n_wires, art.n_pub_out, art.n_pub_in, art.n_prv_in, art.witness, n_labels=art.n_signals)
    r = E.R1cs(p)
    wc = E.WitnessCalculator(cvm_text=art.cvm, n_slots=0 if seed % 2 else 5)
    wt, st, bad = wc.calculate_checked(rows, r)
    assert (r.check(wt) == bad).all(), seed
    got = E.le_to_ints(wt)
    n_bad = 0
    for b, inp in enumerate(rows):
        try:
            w, ost = I.compute_witness(prog, inp), 0
        except I.WitnessError as e:
            w, ost = None, e.status
        if ost == 0:
            assert st[b] == 0 and got[b] == w, (seed, inp)
            if n_bad < 6 or bad[b] == E.NO_BAD:
                ev = lambda lc: sum(v * w[k] for k, v in lc.items()) % M.Q
                first = next((ci for ci, (a, bb, c) in enumerate(art.constraints) if (ev(a) * ev(bb) - ev(c)) % M.Q), E.NO_BAD)
                assert bad[b] == first, (seed, inp)
                n_bad += bad[b] != E.NO_BAD
        else:
            assert st[b] != 0, (seed, inp)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(100, 116))
def test_random_circuit_gpu_matches_oracle(cvmlib, seed):
    """Same generator through the C ABI on the device: status words and witnesses against the oracle."""
    from circom_cvm_b200 import engine as E
    from tools.circuitgen.build import compile_circuit
    art = compile_circuit(make_circuit(seed, n_stmts=60), (), name="fuzz%d" % seed)
    prog = I.load(art.cvm)
    rows = inputs_for(seed, 70)
    wc = E.WitnessCalculator(cvm_text=art.cvm, n_slots=0 if seed % 2 else 5)
    wt, st = wc.calculate(rows)
    got = E.le_to_ints(wt)
    for b, inp in enumerate(rows):
        try:
            w, ost = I.compute_witness(prog, inp), 0
        except I.WitnessError as e:
            w, ost = None, e.status
        if ost == 0:
            assert st[b] == 0 and got[b] == w, (seed, inp)
        else:
            assert st[b] != 0, (seed, inp)


@pytest.mark.gpu
def test_random_constraint_systems_gpu_check(cvmlib):
    """Random R1CS over every coefficient class, empty / repeated combinations and carry-stressing values: the first
    violated constraint per witness against Python integers (tools/fuzz_r1cs_gpu.py is the long-running version)."""
    from tools.fuzz_r1cs_gpu import main
    n, bad = main(40)
    assert n == 40 * 64 and bad == 0
