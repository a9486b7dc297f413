"""GPU experiment: SM clock while the full-row host-buffer call runs (copy-dominated: kernels ~10 % of the time)."""
import os, sys, time, tempfile, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, pynvml
from circom_cvm_b200 import engine as E, formats
from tools.circuitgen.build import compile_circuit
from tools.circuitgen.circuits import poseidon
art = compile_circuit(poseidon.Poseidon, (2,))
d = tempfile.mkdtemp()
p = os.path.join(d, "p.r1cs")
formats.write_r1cs(p, art.constraints, art.n_wires, art.n_pub_out, art.n_pub_in, art.n_prv_in, art.witness, n_labels=art.n_signals)
wc = E.WitnessCalculator(cvm_text=art.cvm); r = E.R1cs(p)
B = 1 << 17
h_in = torch.randint(0, 256, (B, 2, 32), dtype=torch.uint8); h_in[:, :, 31] &= 0x1f; h_in = h_in.pin_memory()
h_st = torch.empty(B, dtype=torch.int32).pin_memory(); h_bad = torch.empty(B, dtype=torch.int32).pin_memory()
h_wt = torch.empty((B, wc.n_wires, 32), dtype=torch.uint8).pin_memory()
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
samples = []; stop = False
def poll():
    while not stop:
        samples.append((time.perf_counter(), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_MEM)))
        time.sleep(0.002)
th = threading.Thread(target=poll, daemon=True); th.start()
for _ in range(3): wc.calculate_select_into(h_in, 0, wc.n_wires, h_wt, h_st, r, h_bad)
for k in range(12):
    t0 = time.perf_counter(); wc.calculate_select_into(h_in, 0, wc.n_wires, h_wt, h_st, r, h_bad); t1 = time.perf_counter()
    sm = [s[1] for s in samples if t0 <= s[0] <= t1]
    print("call %2d: %.1f ms  SM clock min %s median %s max %s MHz (%d samples)" % (k, (t1 - t0) * 1e3, min(sm), sorted(sm)[len(sm)//2], max(sm), len(sm)))
stop = True
