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
    """SM clock / throttle reasons during the timed region (B200_PROFILING.md): NVML polled every 10 ms from this process
    (a timed region of ~125 ms is over before an `nvidia-smi -lms` child has printed its first line); `nvidia-smi` as fallback."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.stop_flag = False
        self.proc = None
        self.source = None

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:
            import torch
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            return pynvml, pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode() if not uuid.startswith("GPU-") else uuid.encode())
        except Exception:
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.index)

    def run(self):
        try:
            nv, h = self._nvml_handle()
            sm_max = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            self.source = "nvml"
            while not self.stop_flag:
                sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                except Exception:
                    pw = 0.0
                self.samples.append((sm, sm_max, pw))
                mask = int(get_reasons(h))
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
                time.sleep(0.01)
            return
        except Exception:
            pass
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.source = "nvidia-smi"
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
        self.join(timeout=1.0)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock samples (NVML and nvidia-smi unavailable)"]}
        sm = sorted(s[0] for s in self.samples)
        hi = [s for s in sm if s >= 0.5 * max(sm)] or sm
        return {"sm_mhz": hi[len(hi) // 2], "sm_max_mhz": max(s[1] for s in self.samples),
                "power_w_max": max(s[2] for s in self.samples), "reasons": sorted(self.reasons),
                "samples": len(self.samples), "source": self.source}


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
    emit(line)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default="poseidon2", choices=sorted(WORKLOADS))
    ap.add_argument("--secondary", default=None,
                    help="second workload reported under `secondary` (default: sha256_512 when the primary is poseidon2; 'none')")
    ap.add_argument("--batch", type=int, default=0, help="witnesses per GPU per step (default: the BASELINE config's)")
    ap.add_argument("--chunk", type=int, default=0, help="witnesses per kernel launch (the value store is sized for it)")
    ap.add_argument("--e2e-batch", type=int, default=0)
    ap.add_argument("--slots", type=int, default=0)
    ap.add_argument("--tape-mode", type=int, default=0, help="reserved")
    ap.add_argument("--no-speculation", action="store_true",
                    help="device-resident legs run the general tape even when the program has a speculative (bit-input) one")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs only: leave the host-buffer leg out")
    return ap.parse_args()


def load_traffic(workload):
    """DRAM bytes per witness of each kernel from the committed ncu --set full captures (profiles/r02_traffic.json, else
    round 1's): `roofline.traffic`."""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", name)
        if os.path.exists(tpath):
            with open(tpath) as f:
                t = json.load(f).get(workload)
            if t:
                return t, name
    return {}, None


def bench_workload(args, name, dist_ctx, peak_ctx, batch=0, chunk=0, e2e_batch=0):
    """One workload through the device-resident API (`value`), the host-buffer API (`e2e`), its rooflines and its CPU
    baseline.  -> the JSON line's dict (rank 0) or None."""
    import torch
    import torch.distributed as dist

    from circom_cvm_b200 import engine as E
    rank, world, local, dev = dist_ctx
    select_workload(name)
    B = batch or WL["batch"]
    CH = min(B, chunk or WL["chunk"])
    e2e_b = e2e_batch or WL["e2e_batch"]
    tmpdir = tempfile.mkdtemp(prefix="cvmbench_")
    art, cvm_path, r1cs_path = build_workload(tmpdir)

    wc_host = E.WitnessCalculator(cvm_path=cvm_path, n_slots=args.slots)
    r1 = E.R1cs(r1cs_path)
    # Bit-heavy programs come with a second tape traced under "every main input is 0 or 1" that checks the assumption per
    # witness (speculative typing, cvmgpu_program_speculative).  The host-buffer API (the e2e legs) uses it on its own and
    # recomputes flagged witnesses with the general tape; the device-resident legs opt in here -- the synthetic messages are
    # bits, `speculation_fallbacks` counts the witnesses that would have to be redone (0).
    wc_spec = None if args.no_speculation else wc_host.speculative()
    wc = wc_spec or wc_host
    info = wc.info.asdict()
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
    # field programs run ONE kernel: the tape with the R1CS check scheduled into it (csrc/fused.hpp); bit-heavy ones run the
    # tape and then the check kernels.  Either way this is what cvmgpu_witness_batch_checked_dev launches.
    finfo = wc.fused_info(r1)
    fused = finfo is not None
    if fused:
        finfo = finfo.asdict()
    store_bytes = wc.store_bytes_checked(r1, CH)
    store = torch.empty(store_bytes, dtype=torch.uint8, device=dev)
    status = torch.empty(B, dtype=torch.int32, device=dev)
    bad = torch.empty(B, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    # the store of a bit-heavy circuit can be smaller than the 126 MB L2: flush it between timed iterations then
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if store_bytes < (512 << 20) else None

    def gen(k):
        n = min(CH, B - k * CH)
        wc.run_dev(inputs[k * CH:], n, CH, store, status[k * CH:], stream)

    def check(k):
        n = min(CH, B - k * CH)
        r1.check_store_dev(wc, store, n, CH, bad[k * CH:], stream)

    def gen_checked(k):
        n = min(CH, B - k * CH)
        wc.run_checked_dev(r1, inputs[k * CH:], n, CH, store, status[k * CH:], bad[k * CH:], stream)

    for _ in range(args.warmup):
        for k in range(n_chunks):
            if fused:
                gen_checked(k)
            else:
                gen(k)
                check(k)
    torch.cuda.synchronize()
    rinfo = r1.refresh_info().asdict()
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
        if flush is not None:
            flush.zero_()
        ev[3 * k].record()
        if fused:
            gen_checked(k % n_chunks)
            ev[3 * k + 1].record()
        else:
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
    # per step (= n_chunks launches of each kernel)
    ms_tape = sum(ev[3 * k].elapsed_time(ev[3 * k + 1]) for k in range(n_launch)) / args.steps
    ms_check = sum(ev[3 * k + 1].elapsed_time(ev[3 * k + 2]) for k in range(n_launch)) / args.steps
    # the L2 flush (when there is one) sits between the iterations, outside the step
    ms_total = (ms_tape + ms_check) * args.steps if flush is not None else t_start.elapsed_time(t_end)
    n_fail = int((status != 0).sum()) + int((bad != -1).sum())
    n_spec_fail = int((status == E.ST_SPECULATION).sum())
    sep_tape = sep_check = None
    if fused:
        # the two stand-alone kernels on the same data, outside the step (what CVMGPU_FUSED=0 would run): reported next to it
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        gen(0)
        check(0)
        e[0].record()
        for _ in range(3):
            gen(0)
        e[1].record()
        for _ in range(3):
            check(0)
        e[2].record()
        torch.cuda.synchronize()
        sep_tape, sep_check = e[0].elapsed_time(e[1]) / 3 * n_chunks, e[1].elapsed_time(e[2]) / 3 * n_chunks
        n_fail += int((bad[:min(CH, B)] != -1).sum())
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
    ms_total, ms_tape, ms_check, n_fail = [float(x) for x in t.tolist()]
    ms_step = ms_total / args.steps
    value = world * B / (ms_step * 1e-3)

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
    stored_per_witness = info["n_frows"] * 32 + info["n_brows"] / 8.0
    wire_bytes_per_witness = (wc.n_wires - info["n_bool_wires"]) * 32 + info["n_bool_wires"] / 8.0
    export_bytes = n_exp * (wire_bytes_per_witness + wc.n_wires * 32)
    del wtns_dev

    # ---- end-to-end through the public host-buffer API (pinned host memory, copies inside the timed region)
    del store, flush                   # the host-buffer API brings its own device buffers
    torch.cuda.empty_cache()

    def e2e_leg(Be, wire0, n_sel, api):
        h_in = torch.empty((Be, wc.n_inputs, 32), dtype=torch.uint8).pin_memory()
        h_in.copy_(inputs[:Be].cpu())
        h_wt = torch.empty((Be, n_sel, 32), dtype=torch.uint8).pin_memory() if n_sel else None
        h_st = torch.empty(Be, dtype=torch.int32).pin_memory()
        h_bad = torch.empty(Be, dtype=torch.int32).pin_memory()
        for _ in range(3):
            wc_host.calculate_select_into(h_in, wire0, n_sel, h_wt, h_st, r1, h_bad)
        if world > 1:
            dist.barrier()
        steps = max(3, min(args.steps, 8))
        per_step = []
        t0 = time.perf_counter()
        for _ in range(steps):
            t1 = time.perf_counter()
            wc_host.calculate_select_into(h_in, wire0, n_sel, h_wt, h_st, r1, h_bad)     # synchronous: returns with the results
            per_step.append(time.perf_counter() - t1)
        torch.cuda.synchronize()
        te = torch.tensor([(time.perf_counter() - t0) / steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        sec = float(te.item())
        ok = int((h_st != 0).sum()) == 0 and int((h_bad != -1).sum()) == 0
        out = {"value": world * Be / sec, "unit": UNIT, "h2d_bytes_per_step": Be * wc.n_inputs * 32,
               "d2h_bytes_per_step": Be * (n_sel * 32 + 8), "batch_per_gpu": Be, "ms_per_step": 1000 * sec,
               "steps": steps, "ms_per_step_median": 1000 * sorted(per_step)[len(per_step) // 2], "ms_per_step_max": 1000 * max(per_step),
               "all_witnesses_valid": bool(ok), "api": api}
        if h_wt is not None and h_wt.numel() >= (64 << 20):
            # what the host lets through: the same bytes copied device -> the same pinned buffer by plain cudaMemcpy, every
            # rank at once.  The full-row call cannot beat this; with N processes on one host it is the shared PCIe / memory fabric.
            d_wt = torch.empty(h_wt.shape, dtype=torch.uint8, device=dev)
            h_wt.copy_(d_wt, non_blocking=True)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            for _ in range(3):
                h_wt.copy_(d_wt, non_blocking=True)
            torch.cuda.synchronize()
            tc = torch.tensor([(time.perf_counter() - t0) / 3], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tc, op=dist.ReduceOp.MAX)
            out["d2h_ceiling"] = {"gb_per_s_all_gpus": world * h_wt.numel() / float(tc.item()) / 1e9,
                                  "witnesses_per_s": world * Be / float(tc.item()),
                                  "what": "plain D2H cudaMemcpy of the same %d bytes per GPU into the same pinned buffers, all ranks "
                                          "concurrently: the host-side ceiling of the full-row call" % h_wt.numel()}
            del d_wt
        return out

    def compact_leg(Be):
        """The compact wire formats of a program whose inputs are bits: the messages go up as packed bits (n_inputs / 8 bytes per
        witness) and the WHOLE witness comes back as packed rows (field wires 32 B, 0/1 wires one bit each)."""
        import numpy as np
        layout = wc_host.packed_layout(bit_input_tape=True)
        h_bits = torch.from_numpy(np.packbits(inputs[:Be, :, 0].cpu().numpy(), axis=1, bitorder="little")).pin_memory()
        h_out = torch.empty((Be, layout[0]), dtype=torch.uint8).pin_memory()
        h_st = torch.empty(Be, dtype=torch.int32).pin_memory()
        h_bad = torch.empty(Be, dtype=torch.int32).pin_memory()
        for _ in range(3):
            wc_host.calculate_packed_into(h_bits, True, h_out, h_st, r1, h_bad)
        if world > 1:
            dist.barrier()
        steps = max(3, min(args.steps, 8))
        per_step = []
        t0 = time.perf_counter()
        for _ in range(steps):
            t1 = time.perf_counter()
            wc_host.calculate_packed_into(h_bits, True, h_out, h_st, r1, h_bad)
            per_step.append(time.perf_counter() - t1)
        torch.cuda.synchronize()
        te = torch.tensor([(time.perf_counter() - t0) / steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        sec = float(te.item())
        ok = int((h_st != 0).sum()) == 0 and int((h_bad != -1).sum()) == 0
        return {"value": world * Be / sec, "unit": UNIT, "h2d_bytes_per_step": int(h_bits.numel()), "d2h_bytes_per_step": int(h_out.numel()) + 8 * Be,
                "batch_per_gpu": Be, "ms_per_step": 1000 * sec, "steps": steps,
                "ms_per_step_median": 1000 * sorted(per_step)[len(per_step) // 2], "ms_per_step_max": 1000 * max(per_step),
                "all_witnesses_valid": bool(ok), "packed_row_bytes": layout[0], "field_wires": layout[1], "bit_wires": layout[2],
                "api": "cvmgpu_witness_batch_packed(inputs_are_bits = 1): packed message bits in, witness generation + R1CS check, the WHOLE "
                       "witness back as packed rows (every wire, in the type the program proved: 32 bytes or 1 bit)"}

    e2e = e2e_public = e2e_flags = e2e_compact = None
    if not args.skip_e2e:
        Be = min(e2e_b, B)
        e2e = e2e_leg(Be, 0, wc.n_wires, "WitnessCalculator.calculate_into -> cvmgpu_witness_batch_checked "
                                         "(full .wtns rows returned, %d B per witness)" % (wc.n_wires * 32))
        n_pub = 1 + art.n_pub_out + art.n_pub_in
        e2e_public = e2e_leg(B, 0, n_pub, "cvmgpu_witness_batch_select: witness generation + R1CS check, only the public part "
                                          "of every witness (constant 1, %d outputs, %d public inputs) comes back"
                             % (art.n_pub_out, art.n_pub_in))
        e2e_flags = e2e_leg(B, 0, 0, "same call with wtns_out = NULL: witness generation + R1CS check, flags only")
        if wc_spec is not None and WL["bits"]:
            e2e_compact = compact_leg(B)
    E.lib().cvmgpu_release_buffers()

    if rank != 0:
        wc_host.close()
        r1.close()
        return None

    # ---- rooflines.  Witness generation is integer-multiply bound: unit = one 32x32->64 multiply-accumulate,
    # algorithmic work = 136 x N_mul of the reference program, executed = what the kernel issues after lazy reduction and
    # fusion; peak = in-run dependency-free IMAD.WIDE chains.  The R1CS check is HBM bound by SURVEY 8d's definition:
    # bytes = B * nWires * 32 + B * 4 + matrix; the IMAD view of the same kernel is reported next to it.
    peak_macs, peak_src, hbm_peak, peak_kind = peak_ctx
    traffic, traffic_src = load_traffic(name)

    def dram(kernel):
        t = traffic.get(kernel)
        if t is None and kernel == "r1cs_kernel":
            # the check of a bit-heavy layout is r1cs_table_kernel + r1cs_shift_kernel (+ r1cs_kernel when anything is left for it)
            parts = [v for k, v in traffic.items() if k.startswith("r1cs_")]
            if parts:
                return sum(v["dram_bytes_per_witness"] for v in parts) * CH
        return None if t is None else t["dram_bytes_per_witness"] * CH      # per launch, like `achieved`

    ms_fused = ms_tape if fused else None
    if fused:           # the stand-alone kernels were timed outside the step (rank 0's numbers)
        ms_tape, ms_check = sep_tape, sep_check
    t_tape, t_check = ms_tape * 1e-3, ms_check * 1e-3
    tape_stored = B * ((info["tape_st"] - info["tape_spill_st_bool"]) * 32 + (info["tape_ld"] - info["tape_ld_bool"]) * 32
                       + info["n_inputs"] * 32)
    matrix_bytes = n_chunks * (rinfo["nnz"] * 8 + 3 * (rinfo["n_constraints"] + 1) * 4)
    check_bytes_alg = B * rinfo["n_wires"] * 32 + B * 4 + matrix_bytes
    check_bytes_stored = B * wire_bytes_per_witness + B * 4 + matrix_bytes
    check_macs = rinfo["bound_macs"]
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
                    "achieved": tape_stored / t_tape / 1e9, "frac": tape_stored / t_tape / 1e9 / hbm_peak,
                    "bytes": "B*(field-row stores + field-row reloads)*32 + inputs (bit rows: 4 B per warp, not counted)",
                    "traffic": dram("tape_kernel")},
            "tape_instructions_per_s": B * info["tape_len"] / t_tape,
        },
        "r1cs_kernel": {
            "ms": ms_check, "launches_per_step": n_chunks,
            "constraints_per_s": world * B * rinfo["n_constraints"] / t_check,
            "pm1_fraction": rinfo["nnz_pm1"] / max(1, rinfo["nnz"]),
            "integer_constraints": rinfo["bound_int_constraints"],
            "hbm": {"bound": "hbm", "unit": "GB/s", "peak": hbm_peak, "peak_kind": peak_kind,
                    "achieved": check_bytes_alg / t_check / 1e9, "frac": check_bytes_alg / t_check / 1e9 / hbm_peak,
                    "bytes": "B*nWires*32 + B*4 + nnz*8 + row pointers (SURVEY 8d's compulsory bytes: every wire as a 32-byte value)",
                    "achieved_stored": check_bytes_stored / t_check / 1e9,
                    "frac_stored": check_bytes_stored / t_check / 1e9 / hbm_peak,
                    "bytes_stored": "the same with wires as stored: %d field rows x 32 B + %d bit rows x 1/8 B per witness"
                                    % (wc.n_wires - info["n_bool_wires"], info["n_bool_wires"]),
                    "traffic": dram("r1cs_kernel")},
            "imad": {"bound": "imad", "unit": "Tmac/s", "peak": peak_macs / 1e12,
                     "achieved": B * check_macs / t_check / 1e12, "frac": B * check_macs / t_check / peak_macs,
                     "algorithmic_unit": "%d macs per witness issued at most (64 per general-coefficient term + 72 per LC "
                                         "reduction + 8 per small-coefficient term + 136 per quadratic constraint outside the "
                                         "%d integer-evaluated constraints; products with a 0 / +-1 factor are skipped at run "
                                         "time, so this is an upper bound)" % (check_macs, rinfo["bound_int_constraints"])},
        },
        "export_kernel": {
            "ms": ms_export, "witnesses": n_exp,
            "hbm": {"bound": "hbm", "unit": "GB/s", "peak": hbm_peak, "peak_kind": peak_kind,
                    "achieved": export_bytes / (ms_export * 1e-3) / 1e9, "frac": export_bytes / (ms_export * 1e-3) / 1e9 / hbm_peak,
                    "bytes": "B * (stored wires read + nWires * 32 canonical AoS written); timed alone, outside `value`"}},
    }
    if fused:
        # ONE kernel does both: integer-multiply bound.  Algorithmic work = the reference program's multiplications plus the
        # multiply-accumulates of the constraint system as the stand-alone check counts them; executed = the fused tape's
        # own count.  Bytes = what it has to write (the value store) + inputs: it reads no wire back.
        t_f = ms_fused * 1e-3
        alg_macs = info["ref_mul"] * MACS_PER_MUL + rinfo["macs"]
        fused_bytes = B * ((finfo["tape_st"] - finfo["tape_spill_st_bool"]) * 32 + (finfo["tape_ld"] - finfo["tape_ld_bool"]) * 32
                           + info["n_inputs"] * 32 + 8)
        for kk in ("tape_kernel", "r1cs_kernel"):
            kernels[kk]["note"] = "stand-alone kernel timed outside the step, for comparison (the step runs tape_check_kernel)"
        kernels["tape_check_kernel"] = {
            "ms": ms_fused, "launches_per_step": n_chunks, "what": "tape_kernel running the tape with the R1CS check scheduled into it "
            "(csrc/fused.hpp): %d instructions (tape alone: %d), %d slots" % (finfo["tape_len"], info["tape_len"], finfo["n_slots"]),
            "imad": {"bound": "imad", "unit": "Tmac/s", "peak": peak_macs / 1e12,
                     "achieved": B * alg_macs / t_f / 1e12, "frac": B * alg_macs / t_f / peak_macs,
                     "achieved_executed": B * finfo["tape_macs"] / t_f / 1e12, "frac_executed": B * finfo["tape_macs"] / t_f / peak_macs,
                     "algorithmic_unit": "%d macs per field multiplication x N_mul=%d (reference program) + %d macs of the constraint "
                                         "system (64 per general-coefficient term, 72 per reduction, 136 per product) per witness; "
                                         "executed: %d macs per witness" % (MACS_PER_MUL, info["ref_mul"], rinfo["macs"], finfo["tape_macs"])},
            "hbm": {"bound": "hbm", "unit": "GB/s", "peak": hbm_peak, "peak_kind": peak_kind,
                    "achieved": fused_bytes / t_f / 1e9, "frac": fused_bytes / t_f / 1e9 / hbm_peak,
                    "bytes": "B*(field-row stores + reloads)*32 + inputs + flags: no wire is read back for the check",
                    "traffic": dram("tape_check_kernel")},
            "constraints_per_s": world * B * rinfo["n_constraints"] / t_f,
        }
    dom = "tape_check_kernel" if fused else ("tape_kernel" if ms_tape >= ms_check else "r1cs_kernel")
    kd = kernels[dom]
    # the view SURVEY 8d prescribes for the dominant kernel: witness generation = IMAD, check = HBM
    view, other = ("hbm", "imad") if dom == "r1cs_kernel" else ("imad", "hbm")
    roofline = dict(kd[view])
    roofline.update({"kernel": dom, "ms": kd["ms"],
                     "peak_source": peak_src if view == "imad" else "MEASURED_PEAKS.json hbm_gbs (%s)" % peak_kind,
                     "traffic": kd["hbm"]["traffic"], "traffic_source": traffic_src,
                     "other_view": {k: kd[other][k] for k in ("bound", "achieved", "peak", "unit", "frac")}})
    if roofline.get("frac", 0) > 1.0:
        roofline["note"] = ("frac > 1 is the algorithmic count SURVEY 8(d) prescribes (the reference program's field multiplications x 136) "
                            "over the multiplier peak: the trace compiler removed that arithmetic (fused dot products; for 0/1-typed values "
                            "the reference multiplies field elements to AND two bits, the typed tape does not multiply there).  "
                            "frac_executed is what the kernel issues.")
    if args.skip_cpu:
        cpu = None
    else:
        cpu = cpu_baseline_reference(os.cpu_count() or 1, art=art) or cpu_baseline_port(art)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32x8 (BN254 Fr, Montgomery); wires proven 0/1: 1 bit", "data": "synthetic",
        "config": {"workload": "%s batch of %d random inputs per GPU: witness generation + R1CS check" % (WL["label"], B),
                   "circuit": "tools/circuitgen %s (circomlib structure), %d signals, %d wires, %d constraints"
                              % (WL["label"], info["n_signals"], info["n_wires"], rinfo["n_constraints"]),
                   "batch_per_gpu": B, "witnesses_per_launch": CH,
                   "l2": ("working set %.2f GB per launch >> 126 MB L2" % (store_bytes / 1e9)) if store_bytes >= (512 << 20)
                         else ("value store %.0f MB: a 256 MB buffer is rewritten between timed iterations (L2 flush, outside "
                               "the per-kernel events that make up the step)" % (store_bytes / 1e6)),
                   "parallelism": "batch sharded over %d GPU(s), no data-path collective" % world,
                   "n_slots": (finfo if fused else info)["n_slots"], "n_bit_slots": (finfo if fused else info)["n_bslots"],
                   "tape_len": (finfo if fused else info)["tape_len"],
                   "speculative_bit_inputs": wc_spec is not None, "speculation_fallbacks": n_spec_fail,
                   "check": ("scheduled into the tape: one kernel per launch (csrc/fused.hpp)" if fused
                             else "separate kernels on the value store (r1cs_table_kernel + r1cs_shift_kernel + r1cs_kernel for what is left)"),
                   "stored_bytes_per_witness": stored_per_witness, "failures": n_fail, "flags_gather_ms": gather_ms},
        "kernels_ms": {"tape_check_kernel": ms_fused} if fused else {"tape_kernel": ms_tape, "r1cs_kernel": ms_check},
        "separate_kernels_ms": {"tape_kernel": ms_tape, "r1cs_kernel": ms_check, "note": "timed outside the step"} if fused else None,
        "witnesses_per_s_gen_only": world * B / (ms_tape * 1e-3),
        "constraints_per_s_check_only": world * B * rinfo["n_constraints"] / (ms_check * 1e-3),
        "roofline": roofline, "kernels": kernels,
        "cpu_baseline": cpu, "e2e": e2e, "e2e_public_outputs": e2e_public, "e2e_flags_only": e2e_flags, "e2e_compact": e2e_compact,
        "gpu_launches": n_launch if fused else 2 * n_launch,
        "clocks": sampler.summary(),
        "program": info, "fused_program": finfo if fused else None, "r1cs": rinfo,
    }
    wc_host.close()
    r1.close()
    return line


_RESULT_FD = None


def protect_stdout():
    """stdout carries exactly one JSON line: keep a private copy of fd 1 for it and point fd 1 at stderr, so banners that
    libraries print from C (NCCL's version line under torchrun) cannot land in front of the result."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    fd = 1 if _RESULT_FD is None else _RESULT_FD
    while data:
        data = data[os.write(fd, data):]


def main():
    args = parse_args()
    protect_stdout()
    select_workload(args.workload)
    if args.impl == "reference":
        tmpdir = tempfile.mkdtemp(prefix="cvmbench_")
        art, _cvm_path, _r1cs_path = build_workload(tmpdir)
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
    dev = torch.device("cuda", local)
    # integer-multiply peak, measured in this run by dependency-free chains (the multiplicand depends on the chain, or
    # ptxas hoists the product): mad.lo+mad.hi pairs, mad.wide / mul.wide, and the carry-chained IMAD.WIDE.U32.X form fr.cuh uses
    macs0, _ = E.imad_peak(0)
    macs1 = max(E.imad_peak(1)[0], E.imad_peak(7)[0])
    macs8 = E.imad_peak(8)[0]
    peak_macs = max(macs0, macs1, macs8)
    peaks, peak_kind = measured_peaks()
    peak_src = ("in-run micro-benchmark cvmgpu_imad_peak: max(mad.lo+mad.hi pairs %.2f, mad.wide %.2f, carry-chained "
                "IMAD.WIDE.X %.2f) Tmac/s" % (macs0 / 1e12, macs1 / 1e12, macs8 / 1e12))
    peak_ctx = (peak_macs, peak_src, float(peaks["hbm_gbs"]), peak_kind)
    dist_ctx = (rank, world, local, dev)

    line = bench_workload(args, args.workload, dist_ctx, peak_ctx, args.batch, args.chunk, args.e2e_batch)
    sec_name = args.secondary or ("sha256_512" if args.workload == "poseidon2" else "none")
    if sec_name != "none" and sec_name != args.workload:
        # BASELINE.json's metric names two circuits: the second one (config 3) rides in the same line under `secondary`
        sec = bench_workload(args, sec_name, dist_ctx, peak_ctx)
        if rank == 0:
            keep = ("metric", "value", "unit", "ms_per_step", "config", "kernels_ms", "separate_kernels_ms", "witnesses_per_s_gen_only",
                    "constraints_per_s_check_only", "roofline", "kernels", "cpu_baseline", "e2e", "e2e_public_outputs",
                    "e2e_flags_only", "e2e_compact", "gpu_launches", "program", "r1cs")
            line["secondary"] = {k: sec[k] for k in keep}
            line["gpu_launches"] += sec["gpu_launches"]
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
