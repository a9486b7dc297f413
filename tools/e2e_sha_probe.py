"""GPU experiment: the flags-only host-buffer call on Sha256(512) x 65 536 (1 GiB of inputs): where the time goes."""
import os, sys, time, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from circom_cvm_b200 import engine as E

bench.select_workload("sha256_512")
art, cvm_path, r1cs_path = bench.build_workload(tempfile.mkdtemp())
wc, r = E.WitnessCalculator(cvm_path=cvm_path), E.R1cs(r1cs_path)
B = 65536
h_in = torch.zeros((B, wc.n_inputs, 32), dtype=torch.uint8)
h_in[:, :, 0] = torch.randint(0, 2, (B, wc.n_inputs), dtype=torch.uint8)
h_in = h_in.pin_memory()
h_st = torch.empty(B, dtype=torch.int32).pin_memory(); h_bad = torch.empty(B, dtype=torch.int32).pin_memory()
def t(fn, n=8):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    ts.sort()
    return "%.1f ms (min %.1f, max %.1f)" % (ts[len(ts) // 2], ts[0], ts[-1])
d_in = torch.empty(h_in.shape, dtype=torch.uint8, device="cuda")
print("plain H2D of the inputs:", t(lambda: d_in.copy_(h_in, non_blocking=True)))
print("flags-only, check:", t(lambda: wc.calculate_select_into(h_in, 0, 0, None, h_st, r, h_bad)))
print("flags-only, no check:", t(lambda: wc.calculate_select_into(h_in, 0, 0, None, h_st, None, None)))
for sub in (32768, 16384, 8192):
    print("flags-only B=%d:" % sub, t(lambda: wc.calculate_select_into(h_in[:sub], 0, 0, None, h_st[:sub], r, h_bad[:sub])))
sp = wc.speculative()
store = torch.empty(sp.store_bytes(B), dtype=torch.uint8, device="cuda")
d_st = torch.empty(B, dtype=torch.int32, device="cuda"); d_bad = torch.empty(B, dtype=torch.int32, device="cuda")
s = torch.cuda.current_stream().cuda_stream
for sub in (65536, 32768, 16384, 8192):
    print("device only B=%d: tape" % sub, t(lambda: sp.run_dev(d_in, sub, B, store, d_st, s)), "check", t(lambda: r.check_store_dev(sp, store, sub, B, d_bad, s)))
