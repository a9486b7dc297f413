"""Build the C-ABI shared library (csrc/libcvmgpu.so) with nvcc for sm_100a, in-tree."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libcvmgpu.so")
SOURCES = ["cvmgpu.cu"]
HEADERS = ["fr.cuh", "kernels.cuh", "cvm_parse.hpp", "tracer.hpp", "tape.hpp", "r1cs.hpp", "host_fr.hpp",
           os.path.join("..", "..", "include", "cvmgpu.h")]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    for f in SOURCES + HEADERS + [os.path.join("..", "build.py")]:
        p = os.path.join(CSRC, f)
        if os.path.exists(p) and os.path.getmtime(p) > t:
            return True
    return False


def build(force=False, verbose=False, extra=()):
    if not force and not needs_build():
        return LIB
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "-cudart", "static",
           "-Xptxas", "-v" if verbose else "-O3", *extra,
           "-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    print("+", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
