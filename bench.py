#!/usr/bin/env python3
"""Benchmark of the hot path: batched witness generation + R1CS check on B200.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference ...                      (the reference's CPU calculator on the host cores)

Workload (BASELINE.json configs[1]): Poseidon(2), 1,048,576 random inputs per GPU, witness generation
followed by the R1CS check of every witness.  A "step" is one pass of both kernels over the batch.
One JSON line is printed by rank 0 (contract in the task description).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

UNIT = "witnesses/s"
MACS_PER_MUL = 136      # 8x8 + 8x8 + 8 32x32->64 multiply-accumulates per BN254 Montgomery product

# BASELINE.json configs: [1] Poseidon(2) x 1M (the configuration the metric is quoted on: default), [2] Sha256(512) x 64K
WORKLOADS = {
    "poseidon2": {"label": "Poseidon(2)", "module": "poseidon", "fn": "Poseidon", "args": (2,), "batch": 1 << 20,
                  "chunk": 1 << 20, "e2e_batch": 1 << 17, "bits": False, "ref": "poseidon2_bench"},
    "sha256_512": {"label": "Sha256(512 bits)", "module": "sha256", "fn": "Sha256", "args": (512,), "batch": 1 << 16,
                   "chunk": 1 << 16, "e2e_batch": 1 << 10, "bits": True, "ref": "sha256_512"},
    # stand-in for the SCALE of config 5 (circom-ecdsa, ~1.5 M constraints, batch 1 K; the circuit itself is not built,
    # DESIGN.md section 8): Sha256 over 44 blocks = 1.51 M constraints / 1.5 M wires, 1 024 inputs -- the large sparse check
    # at a batch far too small to fill the GPU with one thread per witness.  Takes minutes to compile; never the default.
    "sha256_44blocks": {"label": "Sha256(22 000 bits, 44 blocks)", "module": "sha256", "fn": "Sha256", "args": (22000,),
                        "batch": 1 << 10, "chunk": 1 << 10, "e2e_batch": 1 << 5, "bits": True, "ref": "sha256_44blocks"},
    # config 4: valid signatures from the integer signer (tools/circuitgen/circuits/eddsa.py); a pool of distinct
    # signatures is tiled over the batch (signing in Python is slow; the instruction stream does not depend on the data)
    "eddsa": {"label": "EdDSAPoseidonVerifier", "module": "eddsa", "fn": "EdDSAPoseidonVerifier", "args": (),
              "batch": 1 << 16, "chunk": 1 << 16, "e2e_batch": 1 << 13, "bits": False, "ref": "eddsa", "pool": 1024},
}
WL = WORKLOADS["poseidon2"]
METRIC = "witnesses/sec (Poseidon(2) witness generation + R1CS check)"


def select_workload(name):
    global WL, METRIC
    WL = WORKLOADS[name]
    METRIC = "witnesses/sec (%s witness generation + R1CS check)" % WL["label"]


def build_workload(tmpdir):
    import importlib

    from circom_cvm_b200 import formats
    from tools.circuitgen.build import compile_circuit
    mod = importlib.import_module("tools.circuitgen.circuits." + WL["module"])
    name = WL["ref"].replace("_bench", "")
    art = compile_circuit(getattr(mod, WL["fn"]), WL["args"], name=name)
    cvm_path = os.path.join(tmpdir, name + ".cvm")
    r1cs_path = os.path.join(tmpdir, name + ".r1cs")
    with open(cvm_path, "w") as f:
        f.write(art.cvm)
    formats.write_r1cs(r1cs_path, art.constraints, art.n_wires, art.n_pub_out, art.n_pub_in, art.n_prv_in,
                       art.witness, n_labels=art.n_signals)
    return art, cvm_path, r1cs_path


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.stop_flag = False
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 7:
                    continue
                try:
                    self.samples.append((float(parts[0]), float(parts[1]), float(parts[2])))
                except ValueError:
                    continue
                for n, v in zip(names, parts[3:7]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
        except FileNotFoundError:
            pass

    def stop(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(s[0] for s in self.samples)
        hi = [s for s in sm if s >= 0.5 * max(sm)] or sm
        return {"sm_mhz": hi[len(hi) // 2], "sm_max_mhz": max(s[1] for s in self.samples),
                "power_w_max": max(s[2] for s in self.samples), "reasons": sorted(self.reasons)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


def cpu_baseline_port(art, seconds=12.0):
    """Oracle port (pure Python CVM interpreter) on one core: a bounded sample of the same workload."""
    import random

    from oracle import cvm_interp as I
    prog = I.load(art.cvm)
    rng = random.Random(0xC1C00001)
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        I.compute_witness(prog, [rng.randrange(2 if WL["bits"] else I.M.Q) for _ in range(art.n_inputs)])
        n += 1
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "%d %s witnesses, pure-Python CVM oracle (oracle/cvm_interp.py), witness generation only"
                      % (n, WL["label"])}


_POOL = {}


def input_pool(art):
    """Workloads whose inputs must be valid (EdDSA): -> uint8 array [pool, n_inputs, 32] of LE canonical values."""
    import random

    import numpy as np
    if "rows" not in _POOL:
        from tools.circuitgen.circuits import eddsa
        rng = random.Random(0xC1C00003)
        rows = [eddsa.sign(rng.randrange(1, 1 << 250), rng.randrange(1, 1 << 250), rng.randrange(1 << 253))
                for _ in range(WL["pool"])]
        buf = b"".join(int(v).to_bytes(32, "little") for r in rows for v in r)
        _POOL["rows"] = np.frombuffer(buf, dtype=np.uint8).reshape(len(rows), art.n_inputs, 32).copy()
    return _POOL["rows"]


def reference_binary():
    p = os.path.join(ROOT, "oracle", "_ref", WL["ref"])
    return p if os.path.exists(p) else None


def cpu_baseline_reference(n_threads, seconds=10.0, art=None):
    """The reference's own C++ runtime + field arithmetic (oracle/_ref) running the Poseidon(2) program emitted
    in the WriteC shapes; run(ctx) only is timed inside the binary; one process per thread."""
    exe = reference_binary()
    if exe is None:
        return None
    mode = ["1" if WL["bits"] else "0"]
    if WL.get("pool"):
        pool_path = os.path.join(tempfile.gettempdir(), "cvmbench_pool_%d.bin" % os.getpid())
        input_pool(art).tofile(pool_path)
        mode = ["2", pool_path]
    procs = [subprocess.Popen([exe, "--bench", str(seconds), str(1234 + i)] + mode,
                              stdout=subprocess.PIPE, text=True, cwd=os.path.dirname(exe))
             for i in range(n_threads)]
    total, count = 0.0, 0
    for p in procs:
        out, _ = p.communicate()
        try:
            d = json.loads(out.strip().splitlines()[-1])
            total += d["witnesses_per_s"]
            count += d["witnesses"]
        except Exception:
            return None
    return {"value": total, "unit": UNIT, "cores": n_threads, "kind": "reference", "value_per_core": total / max(1, n_threads),
            "sample": "%d %s witnesses over %d processes x %.0f s: reference common/calcwit.cpp + generic/fr.cpp "
                      "(--no_asm arithmetic; bn128/fr.asm cannot be assembled here) running the circuit body emitted "
                      "by tools/circuitgen in the WriteC shapes; run(ctx) only" % (count, WL["label"], n_threads, seconds)}


def run_reference_arm(args, art):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    res = None
    per = max(2.0, min(20.0, 6.0))
    vals = []
    for _ in range(args.warmup):
        cpu_baseline_reference(cores, 1.0, art) if reference_binary() else None
    for _ in range(args.steps):
        res = cpu_baseline_reference(cores, per, art)
        if res is None:
            break
        vals.append(res["value"])
    if res is None:
        res = cpu_baseline_port(art, seconds=10.0)
        vals = [res["value"]]
    value = sum(vals) / len(vals)
    res["value"] = value
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 * (time.perf_counter() - t0) / max(1, args.steps),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32x8 (BN254 Fr)",
            "data": "synthetic",
            "config": {"workload": "%s batch of %d random inputs per GPU: witness generation + R1CS check" % (WL["label"], WL["batch"]),
                       "reference_arm": "witness generation (run(ctx)) on the host cores, bounded sample per step; the reference "
                                        "has no R1CS checker (SURVEY F4)"},
            "cpu_baseline": res,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default="poseidon2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="witnesses per GPU per step (default: the BASELINE config's)")
    ap.add_argument("--chunk", type=int, default=0, help="witnesses per kernel launch (the value store is sized for it)")
    ap.add_argument("--e2e-batch", type=int, default=0)
    ap.add_argument("--slots", type=int, default=0)
    ap.add_argument("--tape-mode", type=int, default=0, help="witnesses per thread of the tape kernel: 0 auto, 1, 2")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs only: leave the host-buffer leg out")
    args = ap.parse_args()
    select_workload(args.workload)
    args.batch = args.batch or WL["batch"]
    args.chunk = min(args.batch, args.chunk or WL["chunk"])
    args.e2e_batch = args.e2e_batch or WL["e2e_batch"]

    tmpdir = tempfile.mkdtemp(prefix="cvmbench_")
    art, cvm_path, r1cs_path = build_workload(tmpdir)
    if args.impl == "reference":
        run_reference_arm(args, art)
        return

    import torch
    import torch.distributed as dist

    from circom_cvm_b200 import build as cbuild
    from circom_cvm_b200 import engine as E

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if rank == 0:
        cbuild.build()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()
    E.set_device(local)
    E.set_tape_mode(args.tape_mode)
    dev = torch.device("cuda", local)

    wc = E.WitnessCalculator(cvm_path=cvm_path, n_slots=args.slots)
    r1 = E.R1cs(r1cs_path)
    info, rinfo = wc.info.asdict(), r1.info.asdict()
    B, CH = args.batch, args.chunk
    n_chunks = (B + CH - 1) // CH
    g = torch.Generator(device=dev)
    g.manual_seed(0xC1C00001 + rank)
    if WL.get("pool"):
        pool = torch.from_numpy(input_pool(art)).to(dev)
        inputs = pool.repeat((B + pool.shape[0] - 1) // pool.shape[0], 1, 1)[:B].contiguous()
    elif WL["bits"]:
        inputs = torch.zeros((B, wc.n_inputs, 32), dtype=torch.uint8, device=dev)
        inputs[:, :, 0] = torch.randint(0, 2, (B, wc.n_inputs), dtype=torch.uint8, device=dev, generator=g)
    else:
        inputs = torch.randint(0, 256, (B, wc.n_inputs, 32), dtype=torch.uint8, device=dev, generator=g)
        inputs[:, :, 31] &= 0x1F          # < 2^253 < q: canonical field elements
    store = torch.empty(wc.store_bytes(CH), dtype=torch.uint8, device=dev)
    status = torch.empty(B, dtype=torch.int32, device=dev)
    bad = torch.empty(B, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def gen(k):
        n = min(CH, B - k * CH)
        wc.run_dev(inputs[k * CH:], n, CH, store, status[k * CH:], stream)

    def check(k):
        n = min(CH, B - k * CH)
        r1.check_dev(store, n, CH, bad[k * CH:], stream)

    def step():
        for k in range(n_chunks):
            gen(k)
            check(k)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n_launch = args.steps * n_chunks
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3 * n_launch)]
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t_start.record()
    for k in range(n_launch):
        ev[3 * k].record()
        gen(k % n_chunks)
        ev[3 * k + 1].record()
        check(k % n_chunks)
        ev[3 * k + 2].record()
    t_end.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    if rank == 0:
        sampler.stop()
    ms_total = t_start.elapsed_time(t_end)
    # per step (= n_chunks launches of each kernel)
    ms_tape = sum(ev[3 * k].elapsed_time(ev[3 * k + 1]) for k in range(n_launch)) / args.steps
    ms_check = sum(ev[3 * k + 1].elapsed_time(ev[3 * k + 2]) for k in range(n_launch)) / args.steps
    n_fail = int((status != 0).sum()) + int((bad != -1).sum())
    # the path's only exchange (north_star): the final gather of the per-witness flags to rank 0, outside the timed region
    gather_ms = 0.0
    if world > 1:
        from circom_cvm_b200.sharding import gather_flags
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        gather_flags(status, world * B)                  # first call sets the communicator up
        g0.record()
        all_status = gather_flags(status, world * B)
        all_bad = gather_flags(bad, world * B)
        g1.record()
        torch.cuda.synchronize()
        gather_ms = g0.elapsed_time(g1)
        if rank == 0:
            assert all_status.numel() == world * B and all_bad.numel() == world * B
            n_fail = int((all_status != 0).sum()) + int((all_bad != -1).sum())
    t = torch.tensor([ms_total, ms_tape, ms_check, float(n_fail)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # the only exchange of the path: gather of per-witness flags (here: their count) at the end
    ms_total, ms_tape, ms_check, n_fail = [float(x) for x in t.tolist()]
    ms_step = ms_total / args.steps
    value = world * B / (ms_step * 1e-3)

    # ---- end-to-end through the public host-buffer API (pinned host memory, copies inside the timed region)
    # ---- the HBM-bound kernel of the path, timed on its own (not part of `value`): value store -> .wtns rows
    n_exp = min(CH, max(1, (8 << 30) // (wc.n_wires * 32)))       # at most 8 GiB of .wtns rows
    wtns_dev = torch.empty((n_exp, wc.n_wires, 32), dtype=torch.uint8, device=dev)
    ex0, ex1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wc.export_dev(store, n_exp, CH, wtns_dev, stream)
    ex0.record()
    for _ in range(3):
        wc.export_dev(store, n_exp, CH, wtns_dev, stream)
    ex1.record()
    torch.cuda.synchronize()
    ms_export = ex0.elapsed_time(ex1) / 3
    export_bytes = 2 * n_exp * wc.n_wires * 32
    del wtns_dev

    Be = min(args.e2e_batch, B)
    h_in = torch.empty((Be, wc.n_inputs, 32), dtype=torch.uint8).pin_memory()
    h_in.copy_(inputs[:Be].cpu())
    del store                          # the host-buffer API brings its own device buffers
    torch.cuda.empty_cache()
    h_wt = torch.empty((Be, wc.n_wires, 32), dtype=torch.uint8).pin_memory()
    h_st = torch.empty(Be, dtype=torch.int32).pin_memory()
    h_bad = torch.empty(Be, dtype=torch.int32).pin_memory()
    for _ in range(0 if args.skip_e2e else 2):
        wc.calculate_into(h_in, h_wt, h_st, r1, h_bad)
    if world > 1:
        dist.barrier()
    e2e_steps = 0 if args.skip_e2e else max(2, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        wc.calculate_into(h_in, h_wt, h_st, r1, h_bad)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / max(1, e2e_steps)
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    e2e_ok = int((h_st != 0).sum()) == 0 and int((h_bad != -1).sum()) == 0
    # same call without the witness download (wtns_out = NULL): generate + check, only the per-witness flags come back.
    # Reported next to `e2e`, not instead of it: it shows what the PCIe transfer of the 20 KB witness rows costs.
    e2e_flags = None
    if not args.skip_e2e:
        Bf = B
        f_in = torch.empty((Bf, wc.n_inputs, 32), dtype=torch.uint8).pin_memory()
        f_in.copy_(inputs[:Bf].cpu())
        f_st = torch.empty(Bf, dtype=torch.int32).pin_memory()
        f_bad = torch.empty(Bf, dtype=torch.int32).pin_memory()
        wc.calculate_into(f_in, None, f_st, r1, f_bad)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            wc.calculate_into(f_in, None, f_st, r1, f_bad)
        torch.cuda.synchronize()
        tf = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tf, op=dist.ReduceOp.MAX)
        e2e_flags = {"value": world * Bf / float(tf.item()), "unit": UNIT, "h2d_bytes_per_step": Bf * wc.n_inputs * 32,
                     "d2h_bytes_per_step": Bf * 8, "batch_per_gpu": Bf, "ms_per_step": 1000 * float(tf.item()),
                     "all_witnesses_valid": int((f_st != 0).sum()) == 0 and int((f_bad != -1).sum()) == 0,
                     "api": "same call with wtns_out = NULL: witness generation + R1CS check, flags only"}
    e2e = None if args.skip_e2e else {"value": world * Be / e2e_s, "unit": UNIT, "h2d_bytes_per_step": Be * wc.n_inputs * 32,
           "d2h_bytes_per_step": Be * (wc.n_wires * 32 + 8), "batch_per_gpu": Be, "ms_per_step": 1000 * e2e_s,
           "all_witnesses_valid": bool(e2e_ok),
           "api": "WitnessCalculator.calculate_into -> cvmgpu_witness_batch_checked (full .wtns rows returned)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- rooflines.  Field arithmetic is bound by the integer pipes (north_star): the unit is one 32x32->64
    # multiply-accumulate, the peak is measured in this run by dependency-free chains of mad.wide.u32 (IMAD.WIDE) and
    # of mad.lo+mad.hi pairs (the multiplicand depends on the chain, or ptxas hoists the product).  "algorithmic" counts the reference program's multiplications (136 x N_mul);
    # "executed" counts what the kernels really issue after lazy reduction / fusion (program.tape_macs, r1cs.macs).
    macs0, _ = E.imad_peak(0)
    macs1 = max(E.imad_peak(1)[0], E.imad_peak(7)[0])     # mad.wide with a 64-bit addend / mul.wide (zero addend)
    macs8 = E.imad_peak(8)[0]                             # carry-chained IMAD.WIDE.U32.X (the form fr.cuh uses)
    peak_macs = max(macs0, macs1, macs8)
    peaks, peak_kind = measured_peaks()
    hbm_peak = float(peaks["hbm_gbs"])
    peak_src = ("in-run micro-benchmark cvmgpu_imad_peak: max(mad.lo+mad.hi pairs %.2f, mad.wide %.2f, carry-chained "
                "IMAD.WIDE.X %.2f) Tmac/s" % (macs0 / 1e12, macs1 / 1e12, macs8 / 1e12))
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(args.workload, {})

    def dram(kernel):
        t = traffic.get(kernel)
        return None if t is None else t["dram_bytes_per_witness"] * CH      # per launch, like `achieved`

    t_tape, t_check = ms_tape * 1e-3, ms_check * 1e-3
    tape_bytes = B * (info["tape_st"] + info["tape_ld"]) * 32 + B * info["n_inputs"] * 32
    check_bytes_alg = B * rinfo["n_wires"] * 32 + B * 4 + n_chunks * (rinfo["nnz"] * 8 + 3 * (rinfo["n_constraints"] + 1) * 4)
    kernels = {
        "tape_kernel": {
            "ms": ms_tape, "launches_per_step": n_chunks,
            "imad": {"bound": "imad", "unit": "Tmac/s", "peak": peak_macs / 1e12,
                     "achieved": B * info["ref_mul"] * MACS_PER_MUL / t_tape / 1e12,
                     "frac": B * info["ref_mul"] * MACS_PER_MUL / t_tape / peak_macs,
                     "achieved_executed": B * info["tape_macs"] / t_tape / 1e12,
                     "frac_executed": B * info["tape_macs"] / t_tape / peak_macs,
                     "algorithmic_unit": "%d macs per field multiplication x N_mul=%d (reference program) per witness; "
                                         "executed: %d macs per witness" % (MACS_PER_MUL, info["ref_mul"], info["tape_macs"])},
            "hbm": {"bound": "hbm", "unit": "GB/s", "peak": hbm_peak, "peak_kind": peak_kind,
                    "achieved": tape_bytes / t_tape / 1e9, "frac": tape_bytes / t_tape / 1e9 / hbm_peak,
                    "bytes": "B*(wire stores + spill stores + reloads)*32 + inputs", "traffic": dram("tape_kernel")},
        },
        "r1cs_kernel": {
            "ms": ms_check, "launches_per_step": n_chunks,
            "constraints_per_s": world * B * rinfo["n_constraints"] / t_check,
            "pm1_fraction": rinfo["nnz_pm1"] / max(1, rinfo["nnz"]),
            "imad": {"bound": "imad", "unit": "Tmac/s", "peak": peak_macs / 1e12,
                     "achieved": B * rinfo["macs"] / t_check / 1e12, "frac": B * rinfo["macs"] / t_check / peak_macs,
                     "algorithmic_unit": "%d macs per witness: 64 per general-coefficient term + 72 per LC reduction + 8 per "
                                         "small-coefficient term + 136 per quadratic constraint (%d of %d)"
                                         % (rinfo["macs"], rinfo["n_quadratic"], rinfo["n_constraints"])},
            "hbm": {"bound": "hbm", "unit": "GB/s", "peak": hbm_peak, "peak_kind": peak_kind,
                    "achieved": check_bytes_alg / t_check / 1e9, "frac": check_bytes_alg / t_check / 1e9 / hbm_peak,
                    "bytes": "B*nWires*32 + B*4 + nnz*8 + row pointers (compulsory, SURVEY 8d)", "traffic": dram("r1cs_kernel")},
        },
    }
    kernels["export_kernel"] = {
        "ms": ms_export, "witnesses": n_exp,
        "hbm": {"bound": "hbm", "unit": "GB/s", "peak": hbm_peak, "peak_kind": peak_kind,
                "achieved": export_bytes / (ms_export * 1e-3) / 1e9, "frac": export_bytes / (ms_export * 1e-3) / 1e9 / hbm_peak,
                "bytes": "2 * B * nWires * 32 (Montgomery SoA read, canonical AoS written); timed alone, outside `value`"}}
    dom = "tape_kernel" if ms_tape >= ms_check else "r1cs_kernel"
    # the resource that actually binds the dominant kernel: the one with the larger fraction of its peak
    kd = kernels[dom]
    view = "imad" if kd["imad"]["frac"] >= kd["hbm"]["frac"] else "hbm"
    roofline = dict(kd[view])
    roofline.update({"kernel": dom, "ms": kd["ms"], "peak_source": peak_src if view == "imad" else "MEASURED_PEAKS.json hbm_gbs (%s)" % peak_kind,
                     "traffic": kd["hbm"]["traffic"],
                     "other_view": {k: kd["hbm" if view == "imad" else "imad"][k] for k in ("bound", "achieved", "peak", "unit", "frac")},
                     "note": "integer-multiply bound (north_star: field arithmetic): every multiply-accumulate is one IMAD.WIDE.U32.X, "
                             "one warp-instruction per 4 cycles per scheduler; ncu on Poseidon(2): FMA-heavy pipe 82% busy in the tape "
                             "kernel, 67% in the check, DRAM 27-33% (profiles/r01_summary.md)"})
    roofline_hbm = {k: v["hbm"] for k, v in kernels.items()}
    if args.skip_cpu:
        cpu = None
    else:
        cpu = cpu_baseline_reference(os.cpu_count() or 1, art=art) or cpu_baseline_port(art)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32x8 (BN254 Fr, Montgomery)", "data": "synthetic",
        "config": {"workload": "%s batch of %d random inputs per GPU: witness generation + R1CS check" % (WL["label"], B),
                   "circuit": "tools/circuitgen %s (circomlib structure), %d signals, %d wires, %d constraints"
                              % (WL["label"], info["n_signals"], info["n_wires"], rinfo["n_constraints"]),
                   "batch_per_gpu": B, "witnesses_per_launch": CH,
                   "l2": "working set %.1f GB per launch >> 126 MB L2" % (wc.store_bytes(CH) / 1e9),
                   "parallelism": "batch sharded over %d GPU(s), no data-path collective" % world,
                   "n_slots": info["n_slots"], "tape_len": info["tape_len"], "failures": n_fail,
                   "flags_gather_ms": gather_ms},
        "kernels_ms": {"tape_kernel": ms_tape, "r1cs_kernel": ms_check},
        "witnesses_per_s_gen_only": world * B / (ms_tape * 1e-3),
        "constraints_per_s_check_only": world * B * rinfo["n_constraints"] / (ms_check * 1e-3),
        "roofline": roofline, "roofline_hbm": roofline_hbm, "kernels": kernels,
        "cpu_baseline": cpu, "e2e": e2e, "e2e_flags_only": e2e_flags, "gpu_launches": 2 * n_launch,
        "clocks": sampler.summary(),
        "program": info, "r1cs": rinfo,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
