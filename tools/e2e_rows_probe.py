"""GPU experiment: the full-row host-buffer call on Poseidon(2) x 131 072 against a plain D2H of the same bytes, repeated,
with the pipeline's chunk size varied (CVMGPU_CHUNK)."""
import os, sys, time, tempfile, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
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
    def t(fn, n=6):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(n):
            t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
        ts.sort()
        return "%.1f ms (min %.1f max %.1f)" % (ts[len(ts) // 2], ts[0], ts[-1])
    d_wt = torch.empty(h_wt.shape, dtype=torch.uint8, device="cuda")
    print("chunk", os.environ.get("CVMGPU_CHUNK", "default"), "| full rows:", t(lambda: wc.calculate_select_into(h_in, 0, wc.n_wires, h_wt, h_st, r, h_bad)),
          "| plain D2H:", t(lambda: h_wt.copy_(d_wt, non_blocking=True)), "| again full rows:", t(lambda: wc.calculate_select_into(h_in, 0, wc.n_wires, h_wt, h_st, r, h_bad)))
else:
    for chunk, nbuf in ((None, None), (None, "1"), ("65536", None), ("65536", "1"), (None, None), (None, "1"), ("32768", "1"), ("32768", None)):
        env = dict(os.environ)
        if chunk: env["CVMGPU_CHUNK"] = chunk
        if nbuf: env["CVMGPU_NBUF"] = nbuf
        print("nbuf", nbuf or "2", end=" ", flush=True)
        subprocess.run([sys.executable, __file__, "child"], env=env)
