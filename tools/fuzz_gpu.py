"""Extended GPU fuzz run (not part of the test-suite): random circuits over the whole operator set -- every other seed in
the generator's bit-heavy mode (typed paths: bit-slot file, small integers, fused sums, warp-cooperative groups) -- device
witnesses and status words against the CVM oracle, and the typed R1CS check against the plain check of the exported rows.
python tools/fuzz_gpu.py [first_seed] [n_seeds]"""
import os
import sys
import tempfile

sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
from fuzz_circuits import inputs_for, make_circuit
from oracle import cvm_interp as I
from circom_cvm_b200 import engine as E, formats
from tools.circuitgen.build import compile_circuit

first = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
count = int(sys.argv[2]) if len(sys.argv) > 2 else 600
bad = n = n_viol = 0
tmp = tempfile.mkdtemp()
for seed in range(first, first + count):
    try:
        art = compile_circuit(make_circuit(seed, n_stmts=70, bits=seed % 2 == 0), (), name="fz%d" % seed,
                              constraint_assert_disabled=seed % 4 == 0)
    except ZeroDivisionError:
        continue          # the generator drew a constant division by zero: a compile-time error in circom as well
    prog = I.load(art.cvm)
    rows = inputs_for(seed, 96)
    wc = E.WitnessCalculator(cvm_text=art.cvm, n_slots=[0, 4, 6, 9][seed % 4])
    p = os.path.join(tmp, "f.r1cs")
    formats.write_r1cs(p, art.constraints, art.n_wires, art.n_pub_out, art.n_pub_in, art.n_prv_in, art.witness, n_labels=art.n_signals)
    r = E.R1cs(p)
    wt, st, fb = wc.calculate_checked(rows, r)
    if not (r.check(wt) == fb).all():
        bad += 1
        print('CHECK MISMATCH (typed vs plain)', seed)
    n_viol += int((fb != E.NO_BAD).sum())
    got = E.le_to_ints(wt)
    for b, inp in enumerate(rows):
        try:
            w, ost = I.compute_witness(prog, inp), 0
        except I.WitnessError as e:
            w, ost = None, e.status
        n += 1
        if ost == 0:
            if st[b] != 0 or got[b] != w:
                bad += 1
                print('MISMATCH', seed, inp, st[b])
        elif st[b] == 0:
            bad += 1
            print('MISSED FAILURE', seed, inp, ost)
    wc.close()
    r.close()
print('checked', n, 'witnesses,', n_viol, 'with a violated constraint; bad', bad)
