"""GPU probe: tape + check as two kernels against the tape with the check scheduled into it (csrc/fused.hpp).
usage: python tools/fused_probe.py [workload] [batch]"""
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from circom_cvm_b200 import engine as E


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "poseidon2"
    bench.select_workload(name)
    B = int(sys.argv[2]) if len(sys.argv) > 2 else bench.WL["batch"]
    art, cvm_path, r1cs_path = bench.build_workload(tempfile.mkdtemp(prefix="fusedprobe_"))
    wc, r1 = E.WitnessCalculator(cvm_path=cvm_path), E.R1cs(r1cs_path)
    fi = wc.fused_info(r1)
    print("fused:", None if fi is None else dict(tape_len=fi.tape_len, n_slots=fi.n_slots, macs=fi.tape_macs, ld=fi.tape_ld,
                                                    spill=fi.tape_spill_st), "base:", dict(tape_len=wc.info.tape_len, n_slots=wc.info.n_slots,
                                                                                           macs=wc.info.tape_macs))
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(5)
    inputs = torch.randint(0, 256, (B, wc.n_inputs, 32), dtype=torch.uint8, device=dev, generator=g)
    inputs[:, :, 31] &= 0x1F
    inputs[7, 0, :] = 0xFF          # an input >= q
    nbytes = wc.store_bytes_checked(r1, B)
    s1, s2 = torch.zeros(nbytes, dtype=torch.uint8, device=dev), torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    st1, st2 = torch.empty(B, dtype=torch.int32, device=dev), torch.empty(B, dtype=torch.int32, device=dev)
    b1, b2 = torch.empty(B, dtype=torch.int32, device=dev), torch.empty(B, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def separate():
        wc.run_dev(inputs, B, B, s1, st1, stream)
        r1.check_store_dev(wc, s1, B, B, b1, stream)

    def fused():
        wc.run_checked_dev(r1, inputs, B, B, s2, st2, b2, stream)

    for fn in (separate, fused):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print("%-9s %.3f ms  %.2f M witnesses/s" % (fn.__name__, ms, B / ms / 1e3))
    print("status equal:", bool((st1 == st2).all()), "first_bad equal:", bool((b1 == b2).all()), "violations:", int((b1 != -1).sum()))
    w1 = torch.empty((min(B, 4096), wc.n_wires, 32), dtype=torch.uint8, device=dev)
    w2 = torch.empty_like(w1)
    wc.export_dev(s1, w1.shape[0], B, w1, stream)
    wc.export_dev(s2, w2.shape[0], B, w2, stream)
    torch.cuda.synchronize()
    print("witness rows equal:", bool((w1 == w2).all()))


if __name__ == "__main__":
    main()
