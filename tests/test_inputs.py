"""Named inputs (circom_cvm_b200/inputs.py): the .dat input hash map + input.json handling of the reference's
generated calculator (common/main.cpp:144-284, calcwit.cpp:51-97), checked on CPU against the fixture artifacts."""
import json

import pytest

from circom_cvm_b200 import formats
from circom_cvm_b200.inputs import InputError, InputMap, SymInputMap, qualify, row_from_json, rows_from_json_text
from conftest import circuit


def imap_of(art):
    consts = sorted(art.compiled.constants, key=art.compiled.constants.get)
    dat = formats.dat_bytes(art.main_inputs, art.witness, consts)
    return InputMap(dat, art.witness, art.input_start, art.n_inputs)


def test_scalar_and_array_inputs_in_any_key_order():
    m = imap_of(circuit("multiplier2"))
    assert row_from_json(m, {"b": "11", "a": 3}) == [3, 11]
    m = imap_of(circuit("babyadd4"))
    assert row_from_json(m, {"q": ["0x10", "0b11"], "p": ["7", "0o17"]}) == [7, 15, 16, 3]
    m = imap_of(circuit("poseidon2"))
    assert row_from_json(m, {"inputs": [1, 2]}) == [1, 2]
    with pytest.raises(InputError, match="Types are not the same"):       # check_type, main.cpp:192-207
        row_from_json(m, {"inputs": [1, "2"]})


def test_number_forms_follow_json2FrElements():
    m = imap_of(circuit("multiplier2"))
    q = 21888242871839275222246405745257275088548364400416034343698204186575808495617
    assert row_from_json(m, {"a": str(q + 5), "b": "0X1f"}) == [5, 31]            # Fr_str2element: mpz_fdiv_r by q
    assert row_from_json(m, {"a": 2 ** 60, "b": 1.0}) == [int(format(float(2 ** 60), ".0f")), 1]   # numbers pass through a double
    # a negative JSON number survives the double and is reduced with a floor modulus (bn128/fr.cpp:56-62): q - 5;
    # values of any size are reduced, never refused
    assert row_from_json(m, {"a": -5, "b": -2.0}) == [q - 5, q - 2]
    assert row_from_json(m, {"a": str(2 ** 300 + 7), "b": 1e40}) == [(2 ** 300 + 7) % q, int(format(1e40, ".0f")) % q]
    # an empty digit string passes check_valid_number (main.cpp:126-142) and mpz_init_set_str leaves 0
    assert row_from_json(m, {"a": "", "b": "0x"}) == [0, 0]
    for bad in ("12a", "0b102", "1e5", "-5", "+5", "0x-1"):           # a sign is not a digit of any base
        with pytest.raises(InputError, match="Invalid number"):
            row_from_json(m, {"a": bad, "b": "1"})
    with pytest.raises(InputError, match="Invalid JSON type"):
        row_from_json(m, {"a": None, "b": "1"})


def test_errors_match_the_reference_messages():
    m = imap_of(circuit("babyadd4"))
    with pytest.raises(InputError, match="Not enough values"):
        row_from_json(m, {"p": ["1"], "q": ["1", "2"]})
    with pytest.raises(InputError, match="Too many values"):
        row_from_json(m, {"p": ["1", "2", "3"], "q": ["1", "2"]})
    with pytest.raises(InputError, match="Signal not found"):
        row_from_json(m, {"p": ["1", "2"], "q": ["1", "2"], "zz": "1"})
    with pytest.raises(InputError, match="Not all inputs have been set"):
        row_from_json(m, {"p": ["1", "2"]})


def test_nested_names_are_qualified_like_the_reference():
    out = {}
    qualify("", {"a": {"b": [{"c": 1}, {"c": [2, 3]}]}, "d": [[1, 2], [3, 4]]}, out)
    assert out == {"a.b[0].c": 1, "a.b[1].c": [2, 3], "d": [[1, 2], [3, 4]]}


def test_batches_and_foreign_dat():
    art = circuit("multiplier2")
    m = imap_of(art)
    assert rows_from_json_text(m, json.dumps([{"a": "1", "b": "2"}, {"a": "3", "b": "4"}])) == [[1, 2], [3, 4]]
    with pytest.raises(InputError, match="does not belong"):
        InputMap(b"\x07" * 9000, art.witness, art.input_start, art.n_inputs)


def test_sym_file_resolves_the_same_inputs(tmp_path):
    """`circom --sym` lines (sym_writer.rs:4-14, formats/sym.md) as the name source instead of the .dat hash map: same
    rows, same error messages; the fixture generator's .sym lists every signal once with its witness position."""
    from tools.circuitgen.build import sym_entries
    for name, doc in (("multiplier2", {"b": "11", "a": 3}), ("babyadd4", {"q": ["0x10", "0b11"], "p": ["7", "0o17"]}),
                      ("poseidon2", {"inputs": [1, 2]}), ("multiplier4", {"in": [2, 3, 4, 5]})):
        art = circuit(name)
        ent = sym_entries(art)
        p = tmp_path / (name + ".sym")
        formats.write_sym(str(p), ent)
        assert formats.read_sym(str(p)) == ent
        # every signal once, in label order; witness positions 1..n_wires-1 each exactly once (sym.md)
        assert [e[0] for e in ent] == list(range(1, art.n_signals))
        assert sorted(e[1] for e in ent if e[1] >= 0) == list(range(1, art.n_wires))
        assert all(art.witness[e[1]] == e[0] for e in ent if e[1] >= 0)
        ms = SymInputMap(ent, art.input_start, art.n_inputs)
        assert row_from_json(ms, doc) == row_from_json(imap_of(art), doc)
    art = circuit("multiplier2")
    ms = SymInputMap(sym_entries(art), art.input_start, art.n_inputs)
    with pytest.raises(InputError, match="Signal not found"):
        row_from_json(ms, {"a": 1, "c": 2})
    with pytest.raises(InputError, match="Not all inputs have been set"):
        row_from_json(ms, {"a": 1})
    with pytest.raises(InputError, match="does not belong"):
        SymInputMap(sym_entries(art), art.input_start, art.n_inputs + 3)     # fewer main signals than the program has inputs
    # the doc example's shape (sym.md:43-50): main.c.* lines are not inputs of main
    doc_sym = [(1, 1, 1, "main.out"), (2, 2, 1, "main.in[0]"), (3, 3, 1, "main.in[1]"), (4, -1, 0, "main.c.out"),
               (5, -1, 0, "main.c.in[0]"), (6, 4, 0, "main.c.in[1]")]
    assert row_from_json(SymInputMap(doc_sym, 2, 2), {"in": ["5", "6"]}) == [5, 6]


def test_program_witness_list_through_the_abi(cvmlib):
    from circom_cvm_b200 import engine as E
    art = circuit("babyadd4")
    wc = E.WitnessCalculator(cvm_text=art.cvm)
    assert [int(x) for x in wc.witness_signals()] == list(art.witness)
    assert wc.n_outputs == art.n_outputs


# ---- the native host program (csrc/calc_main.cpp): same input handling, in the reference's own language
def _calc():
    import os
    from circom_cvm_b200 import build
    build.build()
    assert os.path.exists(build.CALC)
    return build.CALC


def test_native_calculator_input_errors_and_no_cpu_fallback(cvmlib, tmp_path):
    import subprocess
    from tools.circuitgen.build import write_artifact
    exe = _calc()
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1 and "Usage" in r.stderr
    art = circuit("babyadd4")
    paths = write_artifact(art, str(tmp_path), with_cpp=True)

    def run(doc_text):
        j = tmp_path / "in.json"
        j.write_text(doc_text)
        return subprocess.run([exe, paths["cvm"], str(j), str(tmp_path / "o.wtns")], capture_output=True, text=True, timeout=120)

    for text, msg in (('{"p": ["1"], "q": ["1", "2"]}', "Not enough values"),
                      ('{"p": ["1", "2", "3"], "q": ["1", "2"]}', "Too many values"),
                      ('{"p": ["1", "2"], "q": ["1", "2"], "zz": "1"}', "Signal not found"),
                      ('{"p": ["1", "2"]}', "Not all inputs have been set"),
                      ('{"p": ["1", 2], "q": ["1", "2"]}', "Types are not the same"),
                      ('{"p": ["0x1g", "2"], "q": ["1", "2"]}', "Invalid number"),
                      ('{"p": ["-1", "2"], "q": ["1", "2"]}', "Invalid number"),
                      ('{"p": ["1", "2"], "q": ["1", "2"]', "invalid JSON")):
        r = run(text)
        assert r.returncode == 1 and msg in r.stderr, (text, r.stderr)
    # well-formed input: without a CUDA device the library refuses (no CPU fallback); with one the file is written
    r = run('{"q": ["0x10", "0b11"], "p": ["7", "0o17"]}')
    from circom_cvm_b200 import engine as E
    if E.device_count() == 0:
        assert r.returncode == 1 and "no CPU fallback" in r.stderr
    else:
        assert r.returncode == 0, r.stderr


def test_native_calculator_reads_sym(cvmlib, tmp_path):
    """cvmgpu_calc --sym: the same input errors through the .sym-built table; a foreign .sym is refused."""
    import subprocess
    from tools.circuitgen.build import write_artifact
    exe = _calc()
    art = circuit("babyadd4")
    paths = write_artifact(art, str(tmp_path), with_cpp=False)          # no .dat: only the .sym names the inputs

    def run(doc_text, sym):
        j = tmp_path / "in.json"
        j.write_text(doc_text)
        return subprocess.run([exe, paths["cvm"], str(j), str(tmp_path / "o.wtns"), "--sym", sym], capture_output=True, text=True,
                              timeout=120)

    for text, msg in (('{"p": ["1"], "q": ["1", "2"]}', "Not enough values"),
                      ('{"p": ["1", "2"], "q": ["1", "2"], "zz": "1"}', "Signal not found"),
                      ('{"p": ["1", "2"]}', "Not all inputs have been set")):
        r = run(text, paths["sym"])
        assert r.returncode == 1 and msg in r.stderr, (text, r.stderr)
    other = write_artifact(circuit("multiplier2"), str(tmp_path / "m2"))
    r = run('{"q": ["1", "2"], "p": ["7", "8"]}', other["sym"])
    assert r.returncode == 1 and "does not belong" in r.stderr
    r = run('{"q": ["0x10", "0b11"], "p": ["7", "0o17"]}', paths["sym"])
    from circom_cvm_b200 import engine as E
    if E.device_count() == 0:
        assert r.returncode == 1 and "no CPU fallback" in r.stderr
    else:
        assert r.returncode == 0, r.stderr


def test_input_names_from_the_program_text(cvmlib):
    """`;;%%main_input <name> <first signal> <size>` lines (patches/main_input_directive.rs.diff): the same rows as the .dat
    hash map gives, with neither the .dat nor the .sym; a program without them says so."""
    from circom_cvm_b200 import engine as E
    from circom_cvm_b200 import formats
    from circom_cvm_b200.inputs import InputError, InputMap, NamedInputMap, row_from_json
    from conftest import circuit
    from tools.circuitgen.build import faithful_cvm
    from tools.circuitgen.emit_cpp import emit_cpp
    art = circuit("mixedarr")
    wc = E.WitnessCalculator(cvm_text=art.cvm)
    assert wc.main_inputs() == art.main_inputs
    doc = {"a": [str(v) for v in range(1, 10)], "b": list(range(11, 20)), "w": ["0x3", "5", "7"]}
    consts = sorted(art.compiled.constants, key=art.compiled.constants.get)
    dat = formats.dat_bytes(art.main_inputs, art.witness, consts, io_map=art.compiled.io_map)
    want = row_from_json(InputMap(dat, art.witness, art.input_start, art.n_inputs), doc)
    assert row_from_json(NamedInputMap.from_program(wc), doc) == want == list(range(1, 10)) + list(range(11, 20)) + [3, 5, 7]
    with pytest.raises(InputError):
        row_from_json(NamedInputMap.from_program(wc), {"a": [1] * 9, "b": [1] * 9})
    bare = E.WitnessCalculator(cvm_text=faithful_cvm(art), cpp_text=emit_cpp(art), dat_bytes=dat)
    assert bare.main_inputs() == []
    with pytest.raises(InputError):
        NamedInputMap.from_program(bare)
