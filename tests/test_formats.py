"""File formats: .r1cs / .wtns round trips against the layouts in SURVEY.md App. A, and the C++ loaders."""
import os
import struct

from circom_cvm_b200 import formats
from conftest import circuit

Q = formats.BN254_R


def test_wtns_layout_matches_reference_writer():
    # common/main.cpp:286-332; Multiplier2 probe of SURVEY App. C: 204 bytes = 12 + 12 + 40 + 12 + 128
    data = formats.wtns_bytes([1, 33, 3, 11])
    assert len(data) == 204
    assert data[:4] == b"wtns" and struct.unpack_from("<II", data, 4) == (2, 2)
    assert struct.unpack_from("<IQ", data, 12) == (1, 40)
    assert struct.unpack_from("<I", data, 24)[0] == 32
    assert int.from_bytes(data[28:60], "little") == Q
    assert struct.unpack_from("<I", data, 60)[0] == 4
    assert struct.unpack_from("<IQ", data, 64) == (2, 128)
    back = formats.read_wtns(data)
    assert back["values"] == [1, 33, 3, 11] and back["prime"] == Q


def test_cabi_wtns_writer_is_byte_identical(cvmlib, tmp_path):
    from circom_cvm_b200 import engine as E
    wc = E.WitnessCalculator(cvm_text=circuit("multiplier2").cvm)
    row = E.ints_to_le([[1, 33, 3, 11]], 4)[0]
    p = tmp_path / "w.wtns"
    wc.write_wtns(str(p), row)
    assert p.read_bytes() == formats.wtns_bytes([1, 33, 3, 11])


def test_r1cs_roundtrip_and_section_order(tmp_path):
    art = circuit("lessthan8")
    p = tmp_path / "c.r1cs"
    formats.write_r1cs(str(p), art.constraints, art.n_wires, art.n_pub_out, art.n_pub_in, art.n_prv_in, art.witness,
                       n_labels=art.n_signals)
    raw = p.read_bytes()
    # on-disk order: constraints (2), header (1), wire2label (3)   (r1cs_porting.rs:19-53)
    assert struct.unpack_from("<I", raw, 12)[0] == 2
    back = formats.read_r1cs(str(p))
    assert back["n_wires"] == art.n_wires and back["n_constraints"] == len(art.constraints)
    assert back["constraints"] == [tuple(dict(lc) for lc in c) for c in art.constraints]
    assert back["wire2label"] == art.witness
    assert back["n_pub_out"] == 1 and back["n_prv_in"] == 2


def test_cpp_r1cs_loader_agrees(cvmlib, tmp_path):
    from circom_cvm_b200 import engine as E
    art = circuit("poseidon2")
    p = tmp_path / "p.r1cs"
    formats.write_r1cs(str(p), art.constraints, art.n_wires, art.n_pub_out, art.n_pub_in, art.n_prv_in, art.witness,
                       n_labels=art.n_signals)
    r = E.R1cs(str(p))
    info = r.info.asdict()
    assert info["n_wires"] == art.n_wires and info["n_constraints"] == len(art.constraints)
    assert info["nnz"] == sum(len(lc) for c in art.constraints for lc in c)
    pm1 = sum(1 for c in art.constraints for lc in c for v in lc.values() if v in (1, Q - 1))
    assert info["nnz_pm1"] == pm1
    assert info["n_labels"] == art.n_signals
    # coefficient classes of the check kernel: full-size coefficients on wire 0 (the constant 1) are added, not multiplied
    small = lambda v: v < (1 << 32) or Q - v < (1 << 32)
    assert info["nnz_const"] == sum(1 for c in art.constraints for lc in c for w, v in lc.items() if w == 0 and not small(v))
    assert info["n_quadratic"] == sum(1 for a, b, _ in art.constraints if a and b)
    assert info["n_squares"] == sum(1 for a, b, _ in art.constraints if a and a == b) == 162


def test_doc_example_constraint_convention():
    """mkdocs/docs/circom-language/formats/constraints-json.md:17,26 -- A*B - C = 0, signal 0 is the constant 1,
    -1 is written as q-1."""
    art = circuit("multiplier2")
    (a, b, c), = art.constraints
    assert a == {2: 1} and b == {3: 1} and c == {1: 1}
    art = circuit("iszero")           # out <== -in*inv + 1  ->  (-in)*inv - (out - 1) = 0
    a, b, c = art.constraints[0]
    assert a == {2: Q - 1} and b == {3: 1} and c == {1: 1, 0: Q - 1}


def test_dat_hash_and_constants():
    # FNV-1a 64 (calcwit.cpp:17-24) and the 40-byte constant records (c_code_generator.rs:552-615)
    assert formats.fnv1a("a") == 0xAF63DC4C8601EC8C
    d = formats.dat_bytes([("a", 2, 1), ("b", 3, 1)], [0, 1, 2, 3], [5, Q - 1, 1 << 40])
    assert len(d) == 256 * 24 + 4 * 8 + 3 * 40
    rec = d[256 * 24 + 32:]
    assert struct.unpack_from("<iI", rec, 0) == (5, 0x40000000)
    assert int.from_bytes(rec[8:40], "little") == (5 << 256) % Q
    assert struct.unpack_from("<iI", rec, 40) == (-1, 0x40000000)
    assert struct.unpack_from("<iI", rec, 80) == (0, 0xC0000000)
