"""Extended GPU fuzz run of the fused R1CS check (not part of the test-suite): random arithmetic circuits with deliberately
violated constraints (tests/test_fused_check.py make_arith_circuit) -- the fused kernel's first violated constraint against the
stand-alone check kernel on the returned rows and against a walk of the constraints; witnesses against the oracle.
python tools/fuzz_fused_gpu.py [first_seed] [n_seeds]"""
import os
import sys
import tempfile
import pathlib

sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
import numpy as np

from test_fused_check import NO_BAD, arith_inputs, compile_arith, walk
from oracle import cvm_interp as I
from circom_cvm_b200 import engine as E

first = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
count = int(sys.argv[2]) if len(sys.argv) > 2 else 300
tmp = pathlib.Path(tempfile.mkdtemp())
n = bad = viol = fused_default = 0
for seed in range(first, first + count):
    art, p = compile_arith(seed, tmp, o1=seed % 2 == 0)
    prog = I.load(art.cvm)
    r = E.R1cs(p)
    rows = arith_inputs(seed, art.n_inputs, 96)
    wc = E.WitnessCalculator(cvm_text=art.cvm, n_slots=[0, 4, 6, 9][seed % 4])
    fused_default += wc.fused_info(r) is not None
    E.set_fused_mode(2 if seed % 3 else 1)
    try:
        wt, st, fb = wc.calculate_checked(rows, r)
    finally:
        E.set_fused_mode(1)
    if st.any() or not (r.check(wt) == fb).all():
        bad += 1
        print("MISMATCH (fused vs stand-alone check / status)", seed)
    got = E.le_to_ints(wt)
    for b in range(0, len(rows), 5):
        w = I.compute_witness(prog, rows[b])
        n += 1
        if got[b] != w or int(fb[b]) != walk(art, w):
            bad += 1
            print("MISMATCH", seed, b)
    viol += int((fb != NO_BAD).sum())
    wc.close()
    r.close()
print("checked", count, "circuits (", fused_default, "fused by default),", n, "witnesses against the oracle,", viol, "violations reported; bad", bad)
