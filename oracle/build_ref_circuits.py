#!/usr/bin/env python3
"""Build circuit binaries on top of the REFERENCE runtime: oracle/_ref/<name> (+ <name>.dat).

TEST / BASELINE INFRASTRUCTURE ONLY.  The circuit bodies are emitted by tools/circuitgen in the shapes of the
reference's WriteC emitters (the Rust compiler itself cannot run here); everything else that executes --
loadCircuit, loadJson, Circom_CalcWit, Fr_*, writeBinWitness -- is the reference's own code (oracle/build_ref.py).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
OUT = os.path.join(HERE, "_ref")
LIBGMP = "/usr/lib/x86_64-linux-gnu/libgmp.so.10"


def circuits():
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
        "lessthan8": (basic.LessThan, (8,)),
        "sum3cmp": (basic.Sum3Cmp, ()),
        "mixedarr": (basic.MixedArr, ()),
        "dynindex": (basic.DynIndex, ()),
        "opszoo": (basic.OpsZoo, ()),
        "poseidon2": (poseidon.Poseidon, (2,)),
        "poseidon2m": (poseidon.PoseidonMixed, (2,)),     # circomlib 0.5.x shape: `ark` is a mixed component array
    }
    try:
        from tools.circuitgen.circuits import eddsa
        table["eddsa"] = (eddsa.EdDSAPoseidonVerifier, ())   # BASELINE config 4
    except ImportError:
        pass
    try:
        from tools.circuitgen.circuits import sha256
        table["sha256_64"] = (sha256.Sha256, (64,))
        table["sha256_512"] = (sha256.Sha256, (512,))      # BASELINE config 3 (CPU baseline of bench.py --workload sha256_512)
    except ImportError:
        pass
    return table


def build(names=None):
    from tools.circuitgen.build import compile_circuit, write_artifact
    need = [os.path.join(OUT, f) for f in ("fr.o", "calcwit.o", "main.o")]
    if not all(os.path.exists(p) for p in need):
        print("oracle/_ref runtime objects missing; run oracle/build_ref.py first")
        return False
    src = os.path.join(OUT, "src")
    cdir = os.path.join(OUT, "circuits")
    flags = ["-std=c++11", "-O3", "-w", "-I", src, "-I", os.path.join(HERE, "gmp_shim"), "-I", os.path.join(OUT, "inc")]
    harness = os.path.join(OUT, "ref_harness.o")
    hsrc = os.path.join(HERE, "ref_harness.cpp")
    if not os.path.exists(harness) or os.path.getmtime(harness) < max(os.path.getmtime(hsrc), *(os.path.getmtime(p) for p in need)):
        subprocess.check_call(["g++", *flags, "-c", hsrc, "-o", harness])
    # a binary is up to date when it is newer than everything it is made from (the generator, the harness, the runtime)
    gen_dir = os.path.join(ROOT, "tools", "circuitgen")
    deps = [harness, *need, os.path.join(ROOT, "circom_cvm_b200", "formats.py")]
    for d, _sub, files in os.walk(gen_dir):
        deps += [os.path.join(d, f) for f in files if f.endswith(".py")]
    newest = max(os.path.getmtime(p) for p in deps)
    for name, (fn, args) in circuits().items():
        if names and name not in names:
            continue
        exe = os.path.join(OUT, name)
        if not names and os.path.exists(exe) and os.path.exists(exe + ".dat") and os.path.getmtime(exe) > newest:
            continue
        art = compile_circuit(fn, args, name=name)
        paths = write_artifact(art, cdir, with_cpp=True)
        exe = os.path.join(OUT, name)
        cmd = ["g++", *flags, paths["cpp"], harness, *need, LIBGMP, "-o", exe]
        print("+", " ".join(cmd), flush=True)
        subprocess.check_call(cmd)
        with open(paths["dat"], "rb") as fsrc, open(exe + ".dat", "wb") as fdst:
            fdst.write(fsrc.read())
    # the bench binary name bench.py looks for
    p2 = os.path.join(OUT, "poseidon2")
    if os.path.exists(p2):
        for ext in ("", ".dat"):
            with open(p2 + ext, "rb") as fsrc, open(os.path.join(OUT, "poseidon2_bench") + ext, "wb") as fdst:
                fdst.write(fsrc.read())
        os.chmod(os.path.join(OUT, "poseidon2_bench"), 0o755)
    return True


if __name__ == "__main__":
    sys.exit(0 if build(sys.argv[1:] or None) else 1)
