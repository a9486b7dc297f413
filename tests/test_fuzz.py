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
