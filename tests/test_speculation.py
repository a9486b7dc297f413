"""Speculative typing of main inputs (csrc/cvmgpu.cu build_program, tracer.hpp assume_bit_inputs).

Sha256(n) takes its message as unconstrained signals, so nothing proves in[k] a bit and what is derived from the message
before the first bit decomposition is field arithmetic.  Bit-heavy programs are traced a second time under "every main
input is 0 or 1"; that tape checks the assumption per witness (ST_SPECULATION) and the host-buffer API recomputes flagged
witnesses with the general tape -- so the results must be the oracle's for EVERY input, bits or not."""
import random

import pytest

from conftest import circuit
from oracle import cvm_interp as I
from oracle import fr_model as M
from tape_emulator import ST_SPECULATION, run_tape


def test_speculative_tape_of_sha256(cvmlib):
    from circom_cvm_b200 import engine as E
    art = circuit("sha256_64")
    wc = E.WitnessCalculator(cvm_text=art.cvm)
    sp = wc.speculative()
    assert sp is not None
    assert sp.n_wires == wc.n_wires and sp.n_inputs == wc.n_inputs
    # the field rows are gone (wire 0 keeps one), and so are the Montgomery products
    assert wc.info.n_frows > 100 and sp.info.n_frows == 1
    assert wc.info.tape_mul > 0 and sp.info.tape_mul == 0
    assert sp.info.tape_len < wc.info.tape_len
    tape, consts = sp.tape()
    prog = I.load(art.cvm)
    rng = random.Random(3)
    for _ in range(3):
        inp = [rng.randrange(2) for _ in range(art.n_inputs)]
        rows, st = run_tape(tape, consts, sp.layout(), inp)
        assert st == 0 and rows == I.compute_witness(prog, inp)
    for bad in (2, 7, M.Q - 1, 1 << 200):
        inp = [rng.randrange(2) for _ in range(art.n_inputs)]
        inp[rng.randrange(art.n_inputs)] = bad
        _rows, st = run_tape(tape, consts, sp.layout(), inp)
        assert st == ST_SPECULATION
    sp.close()
    assert wc.speculative() is not None          # the handle belongs to the program, closing the view does not free it


@pytest.mark.parametrize("name", ["poseidon2", "eddsa", "num2bits8", "lessthan8", "opszoo", "multiplier2", "babyadd4"])
def test_programs_that_do_not_speculate(cvmlib, name):
    """field programs have nothing to gain; a handful of inputs feeding bit decompositions are numbers, not bits"""
    from circom_cvm_b200 import engine as E
    assert E.WitnessCalculator(cvm_text=circuit(name).cvm).speculative() is None


@pytest.mark.gpu
def test_host_api_is_exact_for_bits_and_for_anything_else(cvmlib, tmp_path):
    """Sha256(64): messages of bits, and rows where some `bit` is 2, q-1, a random field element.  The host-buffer call runs the
    speculative tape and redoes the flagged rows with the general one: every row equals the oracle's witness, the R1CS verdict
    is that of the stand-alone check on the returned rows, and the speculative program alone flags exactly the non-bit rows."""
    import numpy as np
    from circom_cvm_b200 import engine as E
    from tools.circuitgen.build import write_artifact
    art = circuit("sha256_64")
    paths = write_artifact(art, str(tmp_path))
    wc, r = E.WitnessCalculator(cvm_text=art.cvm), E.R1cs(paths["r1cs"])
    rng = random.Random(17)
    rows = [[rng.randrange(2) for _ in range(art.n_inputs)] for _ in range(97)]
    odd = {5: 2, 31: M.Q - 1, 32: rng.randrange(M.Q), 96: 1 << 64}
    for b, v in odd.items():
        rows[b][rng.randrange(art.n_inputs)] = v
    wt, st, bad = wc.calculate_checked(rows, r)
    keep = [b for b in range(97) if b not in odd]
    assert not st[keep].any() and (st != E.ST_SPECULATION).all()
    assert (r.check(wt) == bad).all()
    assert (bad[keep] == E.NO_BAD).all()
    got = E.le_to_ints(wt)
    prog = I.load(art.cvm)
    for b in list(odd) + [0, 1, 50]:
        try:                                  # (a message `bit` that is not one trips the circuit's own asserts)
            w, ost = I.compute_witness(prog, rows[b]), 0
        except I.WitnessError as e:
            w, ost = None, e.status
        assert int(st[b]) == ost, b
        if ost == 0:
            assert got[b] == w, b
    # what the general tape alone gives for the same batch (device API, no speculation): the same bytes and flags
    import torch
    dev = torch.device("cuda", 0)
    B = len(rows)
    d_in = torch.from_numpy(E.ints_to_le(rows, art.n_inputs)).to(dev)
    g_store = torch.zeros(wc.store_bytes(B), dtype=torch.uint8, device=dev)
    g_st = torch.zeros(B, dtype=torch.int32, device=dev)
    g_wt = torch.empty((B, wc.n_wires, 32), dtype=torch.uint8, device=dev)
    wc.run_dev(d_in, B, B, g_store, g_st, torch.cuda.current_stream().cuda_stream)
    wc.export_dev(g_store, B, B, g_wt, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(g_st.cpu().numpy().astype(np.uint32), st)
    assert np.array_equal(g_wt.cpu().numpy(), wt)
    # the same batch without the check
    wt2, st2 = wc.calculate(rows)
    assert np.array_equal(wt2, wt) and np.array_equal(st2, st)
    # device API, opted in: the speculative program flags exactly the rows that are not bits
    sp = wc.speculative()
    store = torch.zeros(sp.store_bytes(B), dtype=torch.uint8, device=dev)
    d_st = torch.zeros(B, dtype=torch.int32, device=dev)
    sp.run_dev(d_in, B, B, store, d_st, torch.cuda.current_stream().cuda_stream)
    d_wt = torch.empty((B, sp.n_wires, 32), dtype=torch.uint8, device=dev)
    sp.export_dev(store, B, B, d_wt, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    flagged = set(int(b) for b in torch.nonzero(d_st == E.ST_SPECULATION).flatten().cpu())
    assert flagged == set(odd) and int((d_st != 0).sum()) == len(odd)
    assert np.array_equal(d_wt.cpu().numpy()[keep], wt[keep])


@pytest.mark.gpu
def test_speculation_switches_itself_off_when_inputs_are_numbers(cvmlib, tmp_path):
    """A bit-heavy circuit whose many inputs are NUMBERS (each one range-checked by a Num2Bits): the first batch shows it, the
    program stops speculating, results are exact throughout."""
    import numpy as np
    from circom_cvm_b200 import engine as E
    from tools.circuitgen.build import compile_circuit
    from tools.circuitgen.circuits import basic

    def RangeChecks(T):
        x = T.input("x", (12,))
        out = T.output("out", (12,))
        n2b = T.component("n2b", (12,))
        i = T.var("i")
        with T.for_(i, 0, i < 12):
            T.new(n2b[i], basic.Num2Bits, 16)
        with T.for_(i, 0, i < 12):
            T.bind(n2b[i].pin("in"), x[i])
            T.bind(out[i], n2b[i].pin("out")[3])
    art = compile_circuit(RangeChecks, (), name="rangechecks")
    wc = E.WitnessCalculator(cvm_text=art.cvm)
    assert wc.speculative() is not None
    rng = random.Random(4)
    rows = [[rng.randrange(1 << 16) for _ in range(12)] for _ in range(64)]
    prog = I.load(art.cvm)
    for _round in range(2):
        wt, st = wc.calculate(rows)
        assert not st.any()
        got = E.le_to_ints(wt)
        for b in (0, 7, 63):
            assert got[b] == I.compute_witness(prog, rows[b])


def test_packed_layout(cvmlib):
    """row layout of the packed witness: field wires x 32 B, then one bit per 0/1 wire (32-bit words)"""
    from circom_cvm_b200 import engine as E
    art = circuit("sha256_64")
    wc = E.WitnessCalculator(cvm_text=art.cvm)
    rb, nf, nb, rows = wc.packed_layout(bit_input_tape=True)
    assert nf == 1 and nf + nb == wc.n_wires and rb == 32 + 16 * ((nb + 127) // 128)
    assert not rows[0] & E.ROW_BIT and all(r & E.ROW_BIT for r in rows[1:])
    rb0, nf0, nb0, _rows0 = wc.packed_layout()
    assert nf0 > nf and nf0 + nb0 == wc.n_wires and rb0 == 32 * nf0 + 16 * ((nb0 + 127) // 128)
    pos = E.WitnessCalculator(cvm_text=circuit("poseidon2").cvm)
    assert pos.packed_layout()[:3] == (32 * pos.n_wires, pos.n_wires, 0)
    with pytest.raises(E.CvmGpuError) as e:
        pos.packed_layout(bit_input_tape=True)
    assert e.value.code == -3


@pytest.mark.gpu
def test_packed_bit_inputs_and_packed_witness_rows(cvmlib, tmp_path):
    """Sha256(64): messages as packed bits in (8 bytes per witness instead of 2 KB), the whole witness as packed rows out
    (field wires 32 B, 0/1 wires one bit): the same witnesses, flags and R1CS verdicts as the 32-byte interfaces."""
    import numpy as np
    from circom_cvm_b200 import engine as E
    from tools.circuitgen.build import write_artifact
    art = circuit("sha256_64")
    paths = write_artifact(art, str(tmp_path))
    wc, r = E.WitnessCalculator(cvm_text=art.cvm), E.R1cs(paths["r1cs"])
    rng = random.Random(23)
    B = 133
    rows = [[rng.randrange(2) for _ in range(art.n_inputs)] for _ in range(B)]
    wt, st, bad = wc.calculate_checked(rows, r)
    assert not st.any() and (bad == E.NO_BAD).all()
    bits = np.packbits(np.array(rows, dtype=np.uint8), axis=1, bitorder="little")
    assert bits.shape == (B, (art.n_inputs + 7) // 8)
    # packed bits in, a wire range out
    n_pub = 1 + art.n_pub_out
    wt_b = np.zeros((B, n_pub, 32), dtype=np.uint8)
    st_b, bad_b = np.full(B, 9, dtype=np.uint32), np.zeros(B, dtype=np.uint32)
    wc.calculate_bits_into(bits, 0, n_pub, wt_b, st_b, r, bad_b)
    assert not st_b.any() and (bad_b == E.NO_BAD).all() and np.array_equal(wt_b, wt[:, :n_pub])
    # packed bits in, packed rows out
    for as_bits in (True, False):
        layout = wc.packed_layout(bit_input_tape=as_bits)
        out = np.zeros((B, layout[0]), dtype=np.uint8)
        st_p, bad_p = np.full(B, 9, dtype=np.uint32), np.zeros(B, dtype=np.uint32)
        wc.calculate_packed_into(bits if as_bits else E.ints_to_le(rows, art.n_inputs), as_bits, out, st_p, r, bad_p)
        assert not st_p.any() and (bad_p == E.NO_BAD).all()
        got = E.WitnessCalculator.unpack_rows(out[[0, 1, 64, B - 1]], layout)
        want = E.le_to_ints(wt[[0, 1, 64, B - 1]])
        assert got == want, as_bits
    # a field program: its packed rows are its .wtns rows
    pos_art = circuit("poseidon2")
    pos = E.WitnessCalculator(cvm_text=pos_art.cvm)
    prow = [[rng.randrange(M.Q), rng.randrange(M.Q)] for _ in range(40)]
    pw, pst = pos.calculate(prow)
    pout = np.zeros((40, pos.packed_layout()[0]), dtype=np.uint8)
    pst2 = np.zeros(40, dtype=np.uint32)
    pos.calculate_packed_into(E.ints_to_le(prow, 2), False, pout, pst2)
    assert np.array_equal(pout.reshape(40, pos.n_wires, 32), pw) and not pst2.any()
    with pytest.raises(E.CvmGpuError):
        pos.calculate_bits_into(np.zeros((4, 1), dtype=np.uint8), 0, 1, np.zeros((4, 1, 32), dtype=np.uint8), np.zeros(4, dtype=np.uint32))


@pytest.mark.gpu
def test_typed_check_on_corrupted_bit_rows_agrees_with_the_plain_check(cvmlib, tmp_path):
    """The three check kernels of a bit-heavy layout (truth tables, shift-sums as bit-matrix transposes, per-witness integers)
    against the plain check of the exported rows, on a store whose bit rows were corrupted at random: the first violated
    constraint of every witness must be the same."""
    import numpy as np
    import torch
    from circom_cvm_b200 import engine as E
    from tools.circuitgen.build import write_artifact
    art = circuit("sha256_64")
    paths = write_artifact(art, str(tmp_path))
    wc, r = E.WitnessCalculator(cvm_text=art.cvm), E.R1cs(paths["r1cs"])
    rng = random.Random(29)
    B = 320
    rows = [[rng.randrange(2) for _ in range(art.n_inputs)] for _ in range(B)]
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream().cuda_stream
    d_in = torch.from_numpy(E.ints_to_le(rows, art.n_inputs)).to(dev)
    for prog in (wc.speculative(), wc):
        info = prog.info
        store = torch.zeros(prog.store_bytes(B), dtype=torch.uint8, device=dev)
        st = torch.zeros(B, dtype=torch.int32, device=dev)
        prog.run_dev(d_in, B, B, store, st, stream)
        torch.cuda.synchronize()
        assert not st.any()
        # the bit region follows the field rows: n_brows words per group of 32 witnesses
        field_bytes = int(info.n_frows) * 32 * B
        n_words = (B // 32) * int(info.n_brows)
        words = store[field_bytes:field_bytes + 4 * n_words].view(torch.int32)
        assert words.numel() == n_words
        idx = torch.tensor([rng.randrange(n_words) for _ in range(400)], device=dev)
        words[idx] ^= torch.tensor([1 << rng.randrange(31) for _ in range(400)], dtype=torch.int32, device=dev)
        bad = torch.empty(B, dtype=torch.int32, device=dev)
        r.check_store_dev(prog, store, B, B, bad, stream)
        wt = torch.empty((B, prog.n_wires, 32), dtype=torch.uint8, device=dev)
        prog.export_dev(store, B, B, wt, stream)
        torch.cuda.synchronize()
        plain = r.check(wt.cpu().numpy())
        got = bad.cpu().numpy().astype(np.uint32)
        assert np.array_equal(got, plain)
        assert (got != E.NO_BAD).sum() > 50
