"""The program text the fork's --cvm emitters REALLY print -- defects included -- must load and mean the
same as the cleaned-up dialect the other fixtures use:

  * no component creation (create_component_bucket.rs:356-360): recovered from the generated <circuit>.cpp
    (`impl WriteC for CreateCmpBucket`, create_component_bucket.rs:206-354);
  * copy loops that assign to literal addresses (`i64.5 = i64.add i64.5 i64.1`, store_bucket.rs:1026-1028, call_bucket.rs:985-987);
  * the array-equality loop (compute_bucket.rs:538-586);
  * the loaded first element as operand of a multi-element return (return_bucket.rs:131).

tools/circuitgen/emit_cvm.py faithful=True reproduces those emitters line by line; both the product's parser
(csrc/cvm_parse.hpp) and the oracle (oracle/cvm_interp.py, written independently) must honour them."""
import random

import pytest

from conftest import circuit
from oracle import cvm_interp as I
from oracle import fr_model as M
from tape_emulator import run_tape
from test_trace_compiler import CASES, oracle


def _texts(name):
    from tools.circuitgen.build import faithful_cvm
    from tools.circuitgen.emit_cpp import emit_cpp
    art = circuit(name)
    return art, faithful_cvm(art), emit_cpp(art)


def _dat(art):
    """<circuit>.dat of the same compile (tools/circuitgen/build.py write_artifact): carries the io-map of mixed arrays"""
    from circom_cvm_b200 import formats
    consts = sorted(art.compiled.constants, key=art.compiled.constants.get)
    return formats.dat_bytes(art.main_inputs, art.witness, consts, io_map=art.compiled.io_map)


@pytest.mark.parametrize("name", sorted(CASES) + ["eddsa"])
def test_faithful_text_means_the_same(cvmlib, name):
    from circom_cvm_b200 import engine as E
    art, text, cpp = _texts(name)
    assert ";;%%create_cmp" not in text and ";;%%io_map" not in text
    ref = I.load(art.cvm)
    dat = _dat(art)
    prog = I.load(text, cpp_text=cpp, dat=dat)
    wc = E.WitnessCalculator(cvm_text=text, cpp_text=cpp, dat_bytes=dat)
    base = E.WitnessCalculator(cvm_text=art.cvm)
    assert wc.n_wires == art.n_wires and wc.n_inputs == art.n_inputs
    assert wc.info.tape_len == base.info.tape_len          # the same trace comes out of both dialects
    tape, consts = wc.tape()
    rng = random.Random(11)
    if name == "eddsa":
        from tools.circuitgen.circuits import eddsa
        cases = [eddsa.sign(123456789, 987654321, 42)]
    else:
        cases = list(CASES[name])[:4]
        if name not in ("num2bits8", "sum3cmp", "lessthan8", "countdown"):
            cases += [[rng.randrange(M.Q) for _ in range(art.n_inputs)] for _ in range(2)]
    for inp in cases:
        w, st = oracle(ref, inp)
        w2, st2 = oracle(prog, inp)
        assert (w2, st2) == (w, st), (name, inp)
        rows, status = run_tape(tape, consts, wc.layout(), inp)
        assert status == st, (name, inp)
        if st == 0:
            assert rows == w, (name, inp)


def test_the_defective_shapes_are_really_there():
    """(so that the test above cannot pass by the emitter silently falling back to the clean dialect)"""
    _art, text, _cpp = _texts("sum3cmp")
    lines = text.split("\n")
    assert any(l.startswith("i64.") and " = i64.add i64." in l for l in lines)           # literal assigned to
    k = next(i for i, l in enumerate(lines) if l == ";; OP(EQ)")
    body = [l for l in lines[k + 1:k + 14]]
    assert body[1] == "loop" and body[3].split()[1:3] == ["=", "ff.eq"] and body[-1] == "end" and body[-2] == "break"
    _art, text, _cpp = _texts("earlyret")
    assert "ff.call $" in text and "= spr" in text


def test_missing_component_creation_is_reported(cvmlib):
    """Without the extension lines and without the .cpp the program cannot run: a clear error, not a crash."""
    from circom_cvm_b200 import engine as E
    _art, text, _cpp = _texts("sum3cmp")
    with pytest.raises(E.CvmGpuError) as e:
        E.WitnessCalculator(cvm_text=text)
    assert e.value.code == -3 and "never created" in str(e.value)


def test_mapped_accesses_need_the_io_map(cvmlib, tmp_path):
    """A circuit with a mixed component array prints `get_template_id` / `get_template_signal_position` ... (location_rule.rs:
    86-171).  As the fork prints it, the .cvm carries no io-map: it loads with the .cpp AND .dat of the same compile, from
    memory or from files, and says what is missing otherwise."""
    from circom_cvm_b200 import engine as E
    from tools.circuitgen.build import write_artifact
    art, text, cpp = _texts("mixedarr")
    assert "get_template_id" in text and "get_template_signal_dimension" in text and "get_template_signal_size" in text
    with pytest.raises(E.CvmGpuError) as e:
        E.WitnessCalculator(cvm_text=text, cpp_text=cpp)
    assert "io-map" in str(e.value)
    with pytest.raises(NotImplementedError):
        I.compute_witness(I.load(text, cpp_text=cpp), CASES["mixedarr"][0])
    with pytest.raises(E.CvmGpuError) as e:        # the .dat alone cannot be read: its section sizes are in the .cpp
        E.WitnessCalculator(cvm_text=art.cvm, dat_bytes=_dat(art))
    assert e.value.code == -5 and "section sizes" in str(e.value)          # CVMGPU_ERR_ARG
    # from files, as a compile leaves them
    paths = write_artifact(art, str(tmp_path), with_cpp=True)
    with open(paths["cvm"], "w") as f:
        f.write(text)
    with open(paths["dat"], "rb") as f:
        assert f.read() == _dat(art)
    wc = E.WitnessCalculator(cvm_path=paths["cvm"], cpp_path=paths["cpp"], dat_path=paths["dat"])
    base = E.WitnessCalculator(cvm_text=art.cvm)          # the default dialect carries ;;%%io_map lines
    assert ";;%%io_map" in art.cvm
    assert wc.info.tape_len == base.info.tape_len and wc.n_wires == art.n_wires
    with pytest.raises(E.CvmGpuError):
        E.WitnessCalculator(cvm_path=paths["cvm"], cpp_path=paths["cpp"])


def test_dat_io_map_layout():
    """The io-map section of the .dat as c_code_generator.rs:617-674 writes it (ids, then per id: n, n x {offset, len,
    lengths[1..], size, busId}), read back by the oracle's reader with the sizes of the generated C++."""
    import struct
    art, _text, cpp = _texts("mixedarr")
    dat = _dat(art)
    io = art.compiled.io_map
    assert sorted(io) == [0, 1, 2] and "uint get_size_of_io_map() {return 3;}" in cpp
    n_words = 3 + sum(1 + sum(4 + max(len(d) - 1, 0) for _o, d, _s in defs) for defs in io.values())
    words = struct.unpack("<%dI" % n_words, dat[-4 * n_words:])
    assert words[:3] == (0, 1, 2) and words[3] == 4                       # partial, total, m, w
    assert words[4:8] == (0, 0, 1, 0)                                     # partial[2]: offset 0, no further dimension
    assert words[12:17] == (3, 1, 2, 1, 0)                                # m[2][2]: offset 3, lengths[1] = 2
    prog = I.Program(art.cvm.replace(";;%%io_map", ";;-"))
    assert prog.io_map == {}
    I.read_dat_io_map(prog, cpp, dat)
    assert prog.io_map == I.load(art.cvm).io_map
    assert prog.io_map[2][2] == (5, [4], 1, 0)


def test_literal_register_scope(cvmlib):
    """A literal that a copy loop assigns to is a register only inside that loop: the same token used afterwards is the
    literal again (what the C++ twin of the same bucket means, store_bucket.rs:612-648)."""
    from circom_cvm_b200 import engine as E
    q = "21888242871839275222246405745257275088548364400416034343698204186575808495617"
    text = ("%%prime " + q + "\n%%signals 8\n%%start T_0\n%%witness 0 1 2 3 4 5 6 7\n"
            "%%template T_0 [ ff 1 3 ] [ ff 1 4 ] [7] [0]\n"
            # out[0..2] = in[0..2] through the emitter's copy loop over literal addresses 4 -> 0
            "x_1 = i64.3\nloop\nif x_1 \nx_0 = get_signal i64.4\nset_signal i64.0 x_0\nx_1 = i64.sub x_1 i64.1\n"
            "i64.4 = i64.add i64.4 i64.1\ni64.0 = i64.add i64.0 i64.1\ncontinue\nend\nbreak\nend\n"
            # afterwards i64.4 is signal 4 again: out[3] = in[0] * in[0]
            "x_2 = get_signal i64.4\nx_3 = ff.mul x_2 x_2\nset_signal i64.3 x_3\n")
    for prog_w in (I.compute_witness(I.load(text), [5, 6, 7]),):
        assert prog_w == [1, 5, 6, 7, 25, 5, 6, 7]
    wc = E.WitnessCalculator(cvm_text=text)
    tape, consts = wc.tape()
    rows, status = run_tape(tape, consts, wc.layout(), [5, 6, 7])
    assert status == 0 and rows == [1, 5, 6, 7, 25, 5, 6, 7]
