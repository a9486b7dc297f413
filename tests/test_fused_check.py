"""The R1CS check scheduled into the tape (csrc/fused.hpp): for field-only programs cvmgpu_witness_batch_checked runs ONE
kernel whose tape evaluates every constraint right after its last wire is produced.  Its verdict -- first violated
constraint per witness -- must be what a walk of the constraints over the witness gives, and what the stand-alone check
kernel (r1cs_kernel, on the exported rows) gives.  CPU tests run the fused tape in tests/tape_emulator.py; the GPU tests
run the product path."""
import random

import pytest

from conftest import circuit
from oracle import cvm_interp as I
from oracle import fr_model as M
from tape_emulator import run_tape

NO_BAD = 0xFFFFFFFF


def make_arith_circuit(seed, n_stmts=24):
    """Random arithmetic circuit (signals, products, linear combinations with +-1 / small / general coefficients, long
    sums, squares, constants) whose constraints are deliberately NOT all implied by the assignments: some signals are
    assigned with `<--` one expression and constrained with `===` to another that agrees only for some inputs."""
    rng = random.Random(seed)

    def tmpl(T):
        n_in = rng.randint(2, 5)
        ins = [T.input("i%d" % k) for k in range(n_in)]
        sigs = list(ins)
        n_out = rng.randint(1, 3)
        outs = [T.output("o%d" % k) for k in range(n_out)]

        def coef():
            r = rng.random()
            if r < 0.3:
                return rng.choice([1, M.Q - 1])
            if r < 0.5:
                return rng.choice([2, 3, 5, 1 << 20, M.Q - 2, M.Q - 12345])
            return rng.randrange(M.Q)

        def lin(n_terms):
            e = None
            for _ in range(n_terms):
                t = rng.choice(sigs)
                c = coef()
                term = t if c == 1 else t * c
                e = term if e is None else e + term
            if rng.random() < 0.4:
                e = e + rng.randrange(M.Q)
            return e

        for k in range(n_stmts):
            s = T.signal("s%d" % k)
            r = rng.random()
            if r < 0.35:                      # product of two linear combinations
                T.bind(s, lin(rng.randint(1, 3)) * lin(rng.randint(1, 3)))
            elif r < 0.45:                    # square
                x = rng.choice(sigs)
                T.bind(s, x * x)
            elif r < 0.65:                    # linear, sometimes long (more terms than a slot file holds)
                T.bind(s, lin(rng.choice([1, 2, 3, 4, 9, 20])))
            elif r < 0.75:                    # constant
                T.bind(s, rng.randrange(M.Q))
            elif r < 0.85:                    # inverse with the usual constraint (holds unless the operand is 0)
                x = rng.choice(sigs)
                T.assign(s, 1 / x)
                T.constrain(s * x, 1)
            else:                             # assigned one thing, constrained to another: violated for most inputs
                T.assign(s, lin(2))
                if rng.random() < 0.5:
                    T.constrain(s, lin(2))
                else:
                    T.constrain(s * rng.choice(sigs), lin(1))
            sigs.append(s)
        for o in outs:
            T.bind(o, lin(3) * rng.choice(sigs) if rng.random() < 0.5 else lin(4))
    tmpl.__name__ = "Arith%d" % seed
    return tmpl


def arith_inputs(seed, n_in, n):
    rng = random.Random(seed * 7919 + 1)
    rows = [[0] * n_in, [1] * n_in, [M.Q - 1] * n_in]
    while len(rows) < n:
        rows.append([rng.choice([0, 1, 2, M.Q - 1, rng.randrange(M.Q), rng.randrange(1 << 32)]) for _ in range(n_in)])
    return rows[:n]


def walk(art, w):
    """first violated constraint of witness w (canonical ints), NO_BAD if none"""
    def ev(lc):
        return sum(c * w[i] for i, c in lc.items()) % M.Q
    for ci, (a, b, c) in enumerate(art.constraints):
        if (ev(a) * ev(b) - ev(c)) % M.Q:
            return ci
    return NO_BAD


def compile_arith(seed, tmp_path, **kw):
    from circom_cvm_b200 import formats
    from tools.circuitgen.build import compile_circuit
    art = compile_circuit(make_arith_circuit(seed), (), name="arith%d" % seed, constraint_assert_disabled=True, **kw)
    p = str(tmp_path / ("arith%d.r1cs" % seed))
    formats.write_r1cs(p, art.constraints, art.n_wires, art.n_pub_out, art.n_pub_in, art.n_prv_in, art.witness, n_labels=art.n_signals)
    return art, p


@pytest.mark.parametrize("seed", range(40))
def test_fused_tape_gives_the_first_violated_constraint(cvmlib, seed, tmp_path):
    from circom_cvm_b200 import engine as E
    art, p = compile_arith(seed, tmp_path, o1=seed % 3 != 0)
    prog = I.load(art.cvm)
    r = E.R1cs(p)
    n_viol = 0
    for slots in (0, 4, 7):
        wc = E.WitnessCalculator(cvm_text=art.cvm, n_slots=slots)
        E.set_fused_mode(2)           # (a wire that happens to be typed 0/1 makes the default mode keep the separate kernels)
        try:
            fused = wc.fused_tape(r)
        finally:
            E.set_fused_mode(1)
        assert fused is not None, "an arithmetic circuit must be fusable"
        tape, consts, layout = fused
        base_tape, _c = wc.tape()
        assert len(tape) > len(base_tape)
        for inp in arith_inputs(seed, art.n_inputs, 8):
            w = I.compute_witness(prog, inp)
            rows, status, first_bad = run_tape(tape, consts, layout, inp, want_first_bad=True)
            assert status == 0 and rows == w, (seed, slots, inp)
            assert first_bad == walk(art, w), (seed, slots, inp)
            n_viol += first_bad != NO_BAD
    VIOLATIONS.append(n_viol)


VIOLATIONS = []


def test_the_generator_produces_violated_constraints():
    """(so that the agreement above is not only about witnesses that satisfy everything)"""
    if not VIOLATIONS:
        pytest.skip("runs after the parametrized test above")
    assert sum(1 for v in VIOLATIONS if v) >= len(VIOLATIONS) // 2


@pytest.mark.parametrize("name", ["poseidon2", "poseidon2m", "multiplier2", "multiplier4", "mixedarr"])
def test_fixture_circuits_fuse_and_hold(cvmlib, name, tmp_path):
    from circom_cvm_b200 import engine as E
    from test_trace_compiler import CASES
    from tools.circuitgen.build import write_artifact
    art = circuit(name)
    paths = write_artifact(art, str(tmp_path))
    wc, r = E.WitnessCalculator(cvm_text=art.cvm), E.R1cs(paths["r1cs"])
    info = wc.fused_info(r)
    assert info is not None
    # the check's multiply-accumulates are counted in the fused tape: generation + check
    assert info.tape_macs > wc.info.tape_macs
    tape, consts, layout = wc.fused_tape(r)
    prog = I.load(art.cvm)
    for inp in CASES[name]:
        w = I.compute_witness(prog, inp)
        rows, status, first_bad = run_tape(tape, consts, layout, inp, want_first_bad=True)
        assert (rows, status, first_bad) == (w, 0, NO_BAD), (name, inp)


def test_bit_heavy_programs_keep_the_separate_check(cvmlib, tmp_path):
    """Integer-typed values, groups or many 0/1 values: constraints over bits go through the table / integer kernels, no
    fused tape (a few 0/1 values -- the zero tests of batched inversions -- do not prevent fusing)."""
    from circom_cvm_b200 import engine as E
    from tools.circuitgen.build import write_artifact
    for name in ("num2bits8", "lessthan8", "opszoo", "sha256_64"):
        art = circuit(name)
        paths = write_artifact(art, str(tmp_path))
        wc, r = E.WitnessCalculator(cvm_text=art.cvm), E.R1cs(paths["r1cs"])
        assert wc.fused_info(r) is None
        assert wc.store_bytes_checked(r, 1024) == wc.store_bytes(1024)


@pytest.mark.parametrize("name", ["iszero", "isequal", "babyadd4", "widesums"])
def test_programs_with_a_few_typed_values_can_fuse(cvmlib, name, tmp_path):
    """0/1-typed values among the operands of constraints (the bit file, bit rows): converted on fetch, compared by T_RNE.
    Not what the default mode picks for them; forced here (cvmgpu_set_fused_mode(2)) so that the path stays correct."""
    from circom_cvm_b200 import engine as E
    from test_trace_compiler import CASES
    from tools.circuitgen.build import write_artifact
    art = circuit(name)
    paths = write_artifact(art, str(tmp_path))
    wc, r = E.WitnessCalculator(cvm_text=art.cvm), E.R1cs(paths["r1cs"])
    E.set_fused_mode(2)
    try:
        tape, consts, layout = wc.fused_tape(r)
    finally:
        E.set_fused_mode(1)
    prog = I.load(art.cvm)
    for inp in CASES[name]:
        w = I.compute_witness(prog, inp)
        rows, status, first_bad = run_tape(tape, consts, layout, inp, want_first_bad=True)
        assert (rows, status, first_bad) == (w, 0, walk(art, w)), (name, inp)


def test_eddsa_verifier_can_fuse(cvmlib, tmp_path):
    """BASELINE config 4: a valid signature satisfies everything; a forged message raises the assert AND violates the
    constraint the assert stands for -- the same one a walk of the constraints over the stored witness finds."""
    from circom_cvm_b200 import engine as E
    from tools.circuitgen.build import write_artifact
    from tools.circuitgen.circuits import eddsa
    art = circuit("eddsa")
    paths = write_artifact(art, str(tmp_path))
    wc, r = E.WitnessCalculator(cvm_text=art.cvm), E.R1cs(paths["r1cs"])
    assert wc.fused_info(r) is None          # measured: 58.3 ms fused against 36.6 + 17.9 ms (untyped tape + stand-alone check)
    assert wc.info.n_bslots == 0             # 3 % of its values are 0/1: compiled untyped, field-only kernel
    E.set_fused_mode(2)
    try:
        tape, consts, layout = wc.fused_tape(r)
    finally:
        E.set_fused_mode(1)
    inp = eddsa.sign(123456789, 987654321, 42)
    rows, status, first_bad = run_tape(tape, consts, layout, inp, want_first_bad=True)
    assert status == 0 and first_bad == NO_BAD and rows == I.compute_witness(I.load(art.cvm), inp)
    inp[6] = 43
    rows, status, first_bad = run_tape(tape, consts, layout, inp, want_first_bad=True)
    assert status != 0 and first_bad == walk(art, [x or 0 for x in rows]) != NO_BAD


def test_a_foreign_constraint_system_is_caught_by_the_fused_check(cvmlib, tmp_path):
    """The fused check evaluates the .r1cs it is given, not what the program computed: a system with one coefficient changed
    reports that constraint."""
    from circom_cvm_b200 import engine as E
    from circom_cvm_b200 import formats
    art = circuit("poseidon2")
    cons = [tuple(dict(lc) for lc in abc) for abc in art.constraints]
    victim = 317
    wire, coef = next(iter(cons[victim][2].items()))
    cons[victim][2][wire] = (coef + 1) % M.Q
    p = str(tmp_path / "tampered.r1cs")
    formats.write_r1cs(p, cons, art.n_wires, art.n_pub_out, art.n_pub_in, art.n_prv_in, art.witness, n_labels=art.n_signals)
    wc, r = E.WitnessCalculator(cvm_text=art.cvm), E.R1cs(p)
    tape, consts, layout = wc.fused_tape(r)
    rows, status, first_bad = run_tape(tape, consts, layout, [1, 2], want_first_bad=True)
    assert status == 0 and first_bad == victim
    assert rows == I.compute_witness(I.load(art.cvm), [1, 2])


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(100, 124))
def test_fused_kernel_agrees_with_the_check_kernel(cvmlib, seed, tmp_path):
    """Product path on the device: cvmgpu_witness_batch_checked (one fused kernel) against the stand-alone check kernel run
    on the exported rows, against a walk of the constraints, and the witnesses against the oracle."""
    import numpy as np
    from circom_cvm_b200 import engine as E
    art, p = compile_arith(seed, tmp_path, o1=seed % 2 == 0)
    prog = I.load(art.cvm)
    r = E.R1cs(p)
    rows = arith_inputs(seed, art.n_inputs, 150)
    wc = E.WitnessCalculator(cvm_text=art.cvm, n_slots=[0, 4, 6, 9][seed % 4])
    E.set_fused_mode(2)
    try:
        assert wc.fused_info(r) is not None
        wt, st, bad = wc.calculate_checked(rows, r)
    finally:
        E.set_fused_mode(1)
    assert not st.any()
    assert (r.check(wt) == bad).all(), seed
    wt2, st2 = wc.calculate(rows)            # the plain tape writes the same rows
    assert np.array_equal(wt, wt2) and np.array_equal(st, st2)
    got = E.le_to_ints(wt)
    for b in range(0, len(rows), 7):
        w = I.compute_witness(prog, rows[b])
        assert got[b] == w, (seed, b)
        assert int(bad[b]) == walk(art, w), (seed, b)


@pytest.mark.gpu
def test_fused_kernel_on_device_buffers_and_tampered_system(cvmlib, tmp_path):
    """cvmgpu_witness_batch_checked_dev with torch buffers: Poseidon(2) x 5 000, every witness valid; against a tampered
    constraint system every witness reports the tampered constraint; the store it leaves exports the same rows."""
    import torch
    from circom_cvm_b200 import engine as E
    from circom_cvm_b200 import formats
    from tools.circuitgen.build import write_artifact
    art = circuit("poseidon2")
    paths = write_artifact(art, str(tmp_path))
    cons = [tuple(dict(lc) for lc in abc) for abc in art.constraints]
    victim = max(ci for ci in range(len(cons)) if cons[ci][0])          # the last quadratic constraint
    wire, coef = next(iter(cons[victim][0].items()))
    cons[victim][0][wire] = (coef + 5) % M.Q
    p = str(tmp_path / "tampered.r1cs")
    formats.write_r1cs(p, cons, art.n_wires, art.n_pub_out, art.n_pub_in, art.n_prv_in, art.witness, n_labels=art.n_signals)
    wc, r, rt = E.WitnessCalculator(cvm_text=art.cvm), E.R1cs(paths["r1cs"]), E.R1cs(p)
    B = 5000
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(9)
    inputs = torch.randint(0, 256, (B, 2, 32), dtype=torch.uint8, device=dev, generator=g)
    inputs[:, :, 31] &= 0x1F
    stream = torch.cuda.current_stream().cuda_stream
    out = {}
    for tag, rr in (("good", r), ("tampered", rt)):
        store = torch.zeros(wc.store_bytes_checked(rr, B), dtype=torch.uint8, device=dev)
        st = torch.full((B,), 7, dtype=torch.int32, device=dev)
        bad = torch.full((B,), 7, dtype=torch.int32, device=dev)
        wc.run_checked_dev(rr, inputs, B, B, store, st, bad, stream)
        wt = torch.empty((B, wc.n_wires, 32), dtype=torch.uint8, device=dev)
        wc.export_dev(store, B, B, wt, stream)
        torch.cuda.synchronize()
        assert not st.any()
        out[tag] = (bad.cpu(), wt.cpu())
    assert (out["good"][0] == -1).all()
    assert (out["tampered"][0] == victim).all()
    assert torch.equal(out["good"][1], out["tampered"][1])
    # the same rows as the host-buffer API without a check
    wt_ref, _st = wc.calculate(inputs[:64].cpu().numpy())
    assert (out["good"][1][:64].numpy() == wt_ref).all()
