"""Extended GPU fuzz run (not part of the test-suite): random circuits over the whole operator set, device results
against the CVM oracle.  python tools/fuzz_gpu.py"""
import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
from fuzz_circuits import inputs_for, make_circuit
from oracle import cvm_interp as I
from circom_cvm_b200 import engine as E
from tools.circuitgen.build import compile_circuit
bad=0; n=0
for seed in range(5000, 5600):
    try:
        art = compile_circuit(make_circuit(seed, n_stmts=70), (), name="fz%d"%seed)
    except ZeroDivisionError:
        continue          # the generator drew a constant division by zero: a compile-time error in circom as well
    prog = I.load(art.cvm)
    rows = inputs_for(seed, 96)
    wc = E.WitnessCalculator(cvm_text=art.cvm, n_slots=[0,4,6,9][seed%4])
    if seed % 3 == 0: E.set_tape_mode(2)
    wt, st = wc.calculate(rows)
    E.set_tape_mode(0)
    got = E.le_to_ints(wt)
    for b, inp in enumerate(rows):
        try: w, ost = I.compute_witness(prog, inp), 0
        except I.WitnessError as e: w, ost = None, e.status
        n+=1
        if ost == 0:
            if st[b] != 0 or got[b] != w: bad+=1; print('MISMATCH', seed, inp, st[b])
        elif st[b] == 0: bad+=1; print('MISSED FAILURE', seed, inp, ost)
print('checked', n, 'bad', bad)
