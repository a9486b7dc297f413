"""GPU experiment: where the time of the host-buffer pipeline goes (flags-only call, Poseidon(2), 1M witnesses)."""
import os, sys, time, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from circom_cvm_b200 import engine as E, formats
from tools.circuitgen.build import compile_circuit
from tools.circuitgen.circuits import poseidon
art = compile_circuit(poseidon.Poseidon, (2,))
d = tempfile.mkdtemp()
p = os.path.join(d, "p.r1cs")
formats.write_r1cs(p, art.constraints, art.n_wires, art.n_pub_out, art.n_pub_in, art.n_prv_in, art.witness, n_labels=art.n_signals)
wc = E.WitnessCalculator(cvm_text=art.cvm); r = E.R1cs(p)
B = 1 << 20
h_in = torch.randint(0, 256, (B, 2, 32), dtype=torch.uint8); h_in[:, :, 31] &= 0x1f; h_in = h_in.pin_memory()
h_st = torch.empty(B, dtype=torch.int32).pin_memory(); h_bad = torch.empty(B, dtype=torch.int32).pin_memory()
def t(fn, n=12):
    for _ in range(4): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]
print("flags-only, check: %.1f ms" % t(lambda: wc.calculate_select_into(h_in, 0, 0, None, h_st, r, h_bad)))
print("flags-only, no check: %.1f ms" % t(lambda: wc.calculate_select_into(h_in, 0, 0, None, h_st, None, None)))
h_pub = torch.empty((B, 4, 32), dtype=torch.uint8).pin_memory()
print("public outputs, check: %.1f ms" % t(lambda: wc.calculate_select_into(h_in, 0, 4, h_pub, h_st, r, h_bad)))
for sub in (1 << 18, 1 << 17):
    print("flags-only B=%d: %.2f ms" % (sub, t(lambda: wc.calculate_select_into(h_in[:sub], 0, 0, None, h_st[:sub], r, h_bad[:sub]))))

h_wt = torch.empty((1 << 17, wc.n_wires, 32), dtype=torch.uint8).pin_memory()
print("full rows B=131072: %.1f ms" % t(lambda: wc.calculate_select_into(h_in[:1 << 17], 0, wc.n_wires, h_wt, h_st[:1 << 17], r, h_bad[:1 << 17]), 6))

for mode in (0, 1, 0, 1):
    E.set_fused_mode(mode)
    print("fused mode %d: full rows B=131072: %.1f ms | flags-only 1M: %.1f ms" % (
        mode, t(lambda: wc.calculate_select_into(h_in[:1 << 17], 0, wc.n_wires, h_wt, h_st[:1 << 17], r, h_bad[:1 << 17]), 6),
        t(lambda: wc.calculate_select_into(h_in, 0, 0, None, h_st, r, h_bad))))
d_wt = torch.empty(h_wt.shape, dtype=torch.uint8, device="cuda")
print("plain D2H of the same bytes: %.1f ms" % t(lambda: h_wt.copy_(d_wt, non_blocking=True), 6))
