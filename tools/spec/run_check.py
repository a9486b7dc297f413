"""GPU experiment: generic check kernel vs the generated one (tools/spec/gen_check.py) on Poseidon(2) x 1M."""
import ctypes, os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from circom_cvm_b200 import engine as E, formats
from tools.circuitgen.build import compile_circuit
from tools.circuitgen.circuits import poseidon
art = compile_circuit(poseidon.Poseidon, (2,))
d = tempfile.mkdtemp(); p = os.path.join(d, "p.r1cs")
formats.write_r1cs(p, art.constraints, art.n_wires, art.n_pub_out, art.n_pub_in, art.n_prv_in, art.witness, n_labels=art.n_signals)
wc = E.WitnessCalculator(cvm_text=art.cvm); r = E.R1cs(p)
B = 1 << 20
dev = torch.device("cuda")
inp = torch.randint(0, 256, (B, 2, 32), dtype=torch.uint8, device=dev); inp[:, :, 31] &= 0x1f
store = torch.empty(wc.store_bytes(B), dtype=torch.uint8, device=dev)
status = torch.empty(B, dtype=torch.int32, device=dev)
bad = torch.empty(B, dtype=torch.int32, device=dev); bad2 = torch.empty(B, dtype=torch.int32, device=dev)
s = torch.cuda.current_stream().cuda_stream
wc.run_dev(inp, B, B, store, status, s)
# corrupt a few witnesses
view = store.view(torch.uint8)
rows = wc.wire_rows()
for b, wire in ((5, 3), (77777, 400), (B - 1, 600)):
    assert not int(rows[wire]) & 0x80000000
    view[((int(rows[wire]) * 2) * B + b) * 16] ^= 1
L = ctypes.CDLL(os.path.join(ROOT, "tools", "spec", "_libspec.so"))
L.spec_check_launch.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_void_p]
def t(fn, n=5):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
print("generic check: %.2f ms" % t(lambda: r.check_store_dev(wc, store, B, B, bad, s)))
print("generated check: %.2f ms" % t(lambda: L.spec_check_launch(store.data_ptr(), B, B, bad2.data_ptr(), s)))
print("same answers:", bool((bad == bad2).all()), "violations:", int((bad != -1).sum()), torch.nonzero(bad != -1).flatten().tolist(), bad[[5, 77777, B - 1]].tolist())
