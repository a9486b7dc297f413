"""Experiment: per-circuit specialised R1CS check kernel (straight-line CUDA source generated from the constraints).
usage: python tools/spec/gen_check.py <n_chunks> -> tools/spec/_spec_check.cu + build of tools/spec/_libspec.so"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from circom_cvm_b200 import engine as E  # noqa: E402
from tools.circuitgen.build import compile_circuit  # noqa: E402
from tools.circuitgen.circuits import poseidon  # noqa: E402

Q = 21888242871839275222246405745257275088548364400416034343698204186575808495617
R = 1 << 256


def fr_lit(v):
    m = v * R % Q
    return "Fr{{" + ",".join("0x%08xu" % ((m >> (32 * i)) & 0xffffffff) for i in range(8)) + "}}"


def gen(n_chunks):
    art = compile_circuit(poseidon.Poseidon, (2,))
    wc = E.WitnessCalculator(cvm_text=art.cvm)
    rows = [int(x) for x in wc.wire_rows()]
    cons = art.constraints
    per = (len(cons) + n_chunks - 1) // n_chunks
    out = ['#include <stdint.h>', '#include <cuda_runtime.h>', '#include "%s/circom_cvm_b200/csrc/fr.cuh"' % ROOT, 'using fr::Fr;',
           '__device__ __forceinline__ Fr ldw(const uint4 *wb, uint32_t row, uint64_t bs) { const uint4 lo = __ldg(wb + (uint64_t)(2 * row) * bs), '
           'hi = __ldg(wb + (uint64_t)(2 * row + 1) * bs); Fr r; r.v[0]=lo.x; r.v[1]=lo.y; r.v[2]=lo.z; r.v[3]=lo.w; r.v[4]=hi.x; r.v[5]=hi.y; '
           'r.v[6]=hi.z; r.v[7]=hi.w; return r; }']

    def lc_code(lc, name):
        code = ["Fr %s = fr::zero();" % name]
        gen_terms = []
        for wire, coef in lc.items():
            loc = rows[wire]
            if loc & 0x80000000:
                continue                      # (a wire bound to the constant 0)
            c = coef % Q
            neg = Q - c
            if wire == 0:
                code.append("%s = fr::add(%s, %s);" % (name, name, fr_lit(c)))
            elif c in (1, 2, 4, 8):
                k = {1: 0, 2: 1, 4: 2, 8: 3}[c]
                code.append("{ Fr v = ldw(wb, %d, bs); %s %s = fr::add(%s, v); }" % (loc, "v = fr::add(v, v); " * k, name, name))
            elif neg in (1, 2, 4, 8):
                k = {1: 0, 2: 1, 4: 2, 8: 3}[neg]
                code.append("{ Fr v = ldw(wb, %d, bs); %s %s = fr::sub(%s, v); }" % (loc, "v = fr::add(v, v); " * k, name, name))
            else:
                gen_terms.append((loc, c))
        for i in range(0, len(gen_terms), 16):
            grp = gen_terms[i:i + 16]
            code.append("{ fr::Wide T; fr::wide_zero(T);")
            for loc, c in grp:
                code.append("  fr::wide_mac(T, %s, ldw(wb, %d, bs));" % (fr_lit(c), loc))
            code.append("  %s = fr::add(%s, fr::wide_reduce(T, %d)); }" % (name, name, len(grp)))
        return code

    for ch in range(n_chunks):
        out.append('extern "C" __global__ void __launch_bounds__(128, 4) spec_check_%d(const uint4 *store, uint64_t bs, uint64_t B, '
                   'uint32_t *first_bad) {' % ch)
        out.append("  uint64_t w = (uint64_t)blockIdx.x * 128 + threadIdx.x; const bool active = w < B; if (!active) w = B - 1;")
        out.append("  const uint4 *wb = store + w; uint32_t bad = 0xffffffffu;")
        for ci in range(ch * per, min(len(cons), (ch + 1) * per)):
            a, b, c = cons[ci]
            out.append("  { // constraint %d" % ci)
            if a and b:
                out += ["    " + l for l in lc_code(a, "sa")]
                if a == b:
                    out.append("    Fr prod = fr::mont_sqr(sa);")
                else:
                    out += ["    " + l for l in lc_code(b, "sb")]
                    out.append("    Fr prod = fr::mont_mul(sa, sb);")
            else:
                out.append("    Fr prod = fr::zero();")
            out += ["    " + l for l in lc_code(c, "sc")]
            out.append("    if (bad == 0xffffffffu && !fr::equal(prod, sc)) bad = %du;" % ci)
            out.append("  }")
        out.append("  if (active && bad != 0xffffffffu) atomicMin(first_bad + w, bad);")
        out.append("}")
    out.append('extern "C" int spec_check_launch(const void *store, uint64_t bs, uint64_t B, void *first_bad, void *stream) {')
    out.append("  cudaMemsetAsync(first_bad, 0xff, B * 4, (cudaStream_t)stream);")
    for ch in range(n_chunks):
        out.append("  spec_check_%d<<<(unsigned)((B + 127) / 128), 128, 0, (cudaStream_t)stream>>>((const uint4 *)store, bs, B, "
                   "(uint32_t *)first_bad);" % ch)
    out.append("  return (int)cudaGetLastError(); }")
    src = os.path.join(ROOT, "tools", "spec", "_spec_check.cu")
    open(src, "w").write("\n".join(out) + "\n")
    return src


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    src = gen(n)
    lib = os.path.join(ROOT, "tools", "spec", "_libspec.so")
    import time
    t0 = time.time()
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
                           "-cudart", "static", "-Xptxas", "-v", "-o", lib, src])
    print("built in %.1f s, %d bytes" % (time.time() - t0, os.path.getsize(lib)))
