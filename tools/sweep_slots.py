"""GPU experiment: tape / check kernel times of one workload for several (field slots, bit slots) choices.
usage: python tools/sweep_slots.py <workload> [slots:bslots ...]   (0:0 = the library's own choice)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
wl = sys.argv[1]
for spec in sys.argv[2:] or ["0:0"]:
    ns, nb = spec.split(":")
    env = dict(os.environ)
    if int(nb):
        env["CVMGPU_BSLOTS"] = nb
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", wl, "--secondary", "none", "--steps", "3",
                          "--warmup", "2", "--skip-cpu", "--skip-e2e", "--slots", ns], env=env, capture_output=True, text=True)
    try:
        d = json.loads(out.stdout.strip().splitlines()[-1])
        pr = d["program"]
        print(spec, "tape %.2f ms check %.2f ms | slots %d bslots %d tape_len %d ld %d (bit %d) spill_st %d" % (
            d["kernels_ms"]["tape_kernel"], d["kernels_ms"]["r1cs_kernel"], pr["n_slots"], pr["n_bslots"], pr["tape_len"],
            pr["tape_ld"], pr["tape_ld_bool"], pr["tape_spill_st"]), flush=True)
    except Exception:
        print(spec, "FAILED", out.stdout[-300:], out.stderr[-600:], flush=True)
