import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def cvmlib():
    """The C-ABI library, built in-tree on first use."""
    from circom_cvm_b200 import build, engine
    build.build()
    return engine.lib()


def load_kats():
    out = []
    with open(os.path.join(ROOT, "tests", "golden", "fr_kat.txt")) as f:
        for line in f:
            if line.startswith("#") or not line.strip():
                continue
            op, af, bf, a, b, o = line.split()
            out.append((op, int(af), int(bf), int(a, 16), None if b == "-" else int(b, 16), o))
    return out


CIRCUITS = {}


def circuit(name):
    """Compile a fixture circuit once per session: -> tools.circuitgen Artifact"""
    if name in CIRCUITS:
        return CIRCUITS[name]
    from tools.circuitgen.build import compile_circuit
    from tools.circuitgen.circuits import babyjub, basic, poseidon
    table = {
        "babyadd4": (babyjub.BabyAddChain, (4,)),
        "nbits": (basic.NBits, ()),
        "earlyret": (basic.EarlyReturns, ()),
        "countdown": (basic.CountDown, ()),
        "multiplier2": (basic.Multiplier2, ()),
        "multiplier4": (basic.MultiplierN, (4,)),
        "num2bits8": (basic.Num2Bits, (8,)),
        "iszero": (basic.IsZero, ()),
        "isequal": (basic.IsEqual, ()),
        "lessthan8": (basic.LessThan, (8,)),
        "sum3cmp": (basic.Sum3Cmp, ()),
        "mixedarr": (basic.MixedArr, ()),
        "dynindex": (basic.DynIndex, ()),
        "opszoo": (basic.OpsZoo, ()),
        "poseidon2": (poseidon.Poseidon, (2,)),
        "poseidon2m": (poseidon.PoseidonMixed, (2,)),     # circomlib 0.5.x shape: `ark` is a mixed component array
        "widesums": (basic.WideSums, ()),
    }
    if name == "eddsa":
        from tools.circuitgen.circuits import eddsa
        art = compile_circuit(eddsa.EdDSAPoseidonVerifier, (), name=name)
    elif name == "sha256_64":
        from tools.circuitgen.circuits import sha256
        art = compile_circuit(sha256.Sha256, (64,), name=name)
    else:
        fn, args = table[name]
        art = compile_circuit(fn, args, name=name)
    CIRCUITS[name] = art
    return art
