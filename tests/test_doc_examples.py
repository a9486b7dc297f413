"""The reference tree's only golden constraint systems are its documentation examples:
mkdocs/docs/circom-language/formats/constraints-json.md:26-100 gives `basic.circom` and its constraints under --O1 (the
default) and --O0, with explicit BN254 coefficients.  They pin (1) the stand-in compiler's constraint generation and O1
simplification (tools/circuitgen/execute.py) against the reference compiler's documented output, (2) the `A*B - C = 0`
convention with signal 0 = the constant 1 and q-1 for -1 through our .r1cs writer, the C++ loader and (on a GPU) the check."""
import pytest

from oracle import cvm_interp as I

Q = 21888242871839275222246405745257275088548364400416034343698204186575808495617
M1 = Q - 1
# constraints-json.md:55-62 (--O1) and :75-84 (--O0), keys as integers
DOC_O1 = [({2: M1}, {4: 1}, {1: M1}),
          ({}, {}, {0: 1, 2: 2, 3: 1, 4: M1})]
DOC_O0 = [({}, {}, {2: 1, 5: M1}),
          ({}, {}, {0: 1, 2: 2, 3: 1, 6: M1}),
          ({}, {}, {1: M1, 4: 1}),
          ({5: M1}, {6: 1}, {4: M1})]


def Internal(T):
    inp = T.input("in", (2,))
    out = T.output("out")
    T.bind(out, inp[0] * inp[1])


def Main(T):
    inp = T.input("in", (2,))
    out = T.output("out")
    c = T.component("c")
    T.new(c, Internal)
    T.bind(c.pin("in")[0], inp[0])
    T.bind(c.pin("in")[1], inp[1] + 2 * inp[0] + 1)
    T.bind(out, c.pin("out"))


def _norm(cons):
    return sorted((tuple(sorted(a.items())), tuple(sorted(b.items())), tuple(sorted(c.items()))) for a, b, c in cons)


def _same_system(ours, doc):
    """equal as sets of constraints, up to the sign of a constraint (c*(A*B - C) = 0 for c = -1: the compiler and the stand-in
    may normalise a linear constraint with either sign)"""
    def canon(con):
        a, b, c = con
        neg = lambda lc: {k: (Q - v) % Q for k, v in lc.items()}
        alts = [(a, b, c), (neg(a), b, neg(c)), (a, neg(b), neg(c))]
        return min(_norm([x])[0] for x in alts)
    return sorted(canon(x) for x in ours) == sorted(canon(x) for x in doc)


def test_basic_circom_constraints_match_the_documentation():
    from tools.circuitgen.build import compile_circuit
    art = compile_circuit(Main, (), name="basic")
    assert art.n_wires == 5 and _same_system(art.constraints, DOC_O1), art.constraints
    art0 = compile_circuit(Main, (), name="basic", o1=False)
    assert art0.n_wires == 7 and _same_system(art0.constraints, DOC_O0), art0.constraints
    # the witness of the compiled program satisfies the DOCUMENTED systems (wire numbering: sym.md / constraints-json.md:26)
    for a, docs in ((art, DOC_O1), (art0, DOC_O0)):
        w = I.compute_witness(I.load(a.cvm), [3, 11])
        assert w[1] == 3 * (11 + 2 * 3 + 1) and w[2:4] == [3, 11]
        for (la, lb, lc) in docs:
            ev = lambda lc_: sum(v * w[k] for k, v in lc_.items()) % Q
            assert (ev(la) * ev(lb) - ev(lc)) % Q == 0


def test_documented_system_through_the_r1cs_loader(cvmlib, tmp_path):
    from circom_cvm_b200 import engine as E
    from circom_cvm_b200 import formats
    p = str(tmp_path / "basic.r1cs")
    formats.write_r1cs(p, DOC_O1, 5, 1, 0, 2, [0, 1, 2, 3, 4], n_labels=7)
    r = E.R1cs(p)
    i = r.info
    assert (i.n_wires, i.n_constraints, i.n_pub_out, i.n_prv_in) == (5, 2, 1, 2)
    assert i.nnz == 7 and i.nnz_pm1 == 6 and i.n_quadratic == 1           # -1 appears as q-1 (constraints-json.md:57-60)
    if E.device_count() == 0:
        pytest.skip("the check itself needs a GPU (tests -m gpu run it)")
    _check_on_gpu(E, r)


def _check_on_gpu(E, r):
    good = [1, 3 * 18, 3, 11, 18]
    bad_lin = [1, 3 * 18, 3, 11, 19]          # breaks the linear constraint (index 1) and the product (index 0)
    bad_out = [1, 55, 3, 11, 18]
    w = E.ints_to_le([good, bad_lin, bad_out], 5)
    assert list(r.check(w)) == [E.NO_BAD, 0, 0]
    w = E.ints_to_le([[1, 0, 0, 5, 6]], 5)    # in[0] = 0: product constraint holds (0 = 0), linear one 1 + 0 + 5 - 6 = 0 holds
    assert list(r.check(w)) == [E.NO_BAD]
    w = E.ints_to_le([[1, 0, 0, 5, 7]], 5)
    assert list(r.check(w)) == [1]


@pytest.mark.gpu
def test_documented_system_checked_on_the_gpu(tmp_path):
    from circom_cvm_b200 import build, engine as E, formats
    build.build()
    p = str(tmp_path / "basic.r1cs")
    formats.write_r1cs(p, DOC_O1, 5, 1, 0, 2, [0, 1, 2, 3, 4], n_labels=7)
    _check_on_gpu(E, E.R1cs(p))


def test_symbols_circom_sym_file_matches_the_documentation():
    """mkdocs/docs/circom-language/formats/sym.md:23-78: signal numbering (outputs, inputs, then the sub-component), which
    signals --O1 eliminates (witness position -1) and where the survivors land.  (The component column of the doc counts
    components in the compiler's DAG order -- main = 1, Internal = 0 -- which the stand-in numbers in creation order; the
    runtime never reads that column, calcwit.cpp / main.cpp, so only #s, #w and the name are compared.)"""
    from tools.circuitgen.build import compile_circuit, sym_entries
    doc_o1 = [(1, 1, "main.out"), (2, 2, "main.in[0]"), (3, 3, "main.in[1]"), (4, -1, "main.c.out"), (5, -1, "main.c.in[0]"),
              (6, 4, "main.c.in[1]")]
    doc_o0 = [(s, s, n) for s, _w, n in doc_o1]
    for o1, doc in ((True, doc_o1), (False, doc_o0)):
        art = compile_circuit(Main, (), name="symbols", o1=o1)
        assert [(s, w, n) for s, w, _c, n in sym_entries(art)] == doc
