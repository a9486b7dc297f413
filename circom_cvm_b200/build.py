"""Build the C-ABI shared library (csrc/libcvmgpu.so) with nvcc for sm_100a, in-tree."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libcvmgpu.so")
CALC = os.path.join(CSRC, "cvmgpu_calc")        # native host program with the reference calculator's process interface
SOURCES = ["cvmgpu.cu"]
HEADERS = ["fr.cuh", "kernels.cuh", "cvm_parse.hpp", "tracer.hpp", "tape.hpp", "fused.hpp", "r1cs.hpp", "host_fr.hpp",
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


def build_calc(force=False):
    """g++ the native host program against the library (rpath = its own directory)."""
    src = os.path.join(CSRC, "calc_main.cpp")
    hdr = os.path.join(CSRC, "..", "..", "include", "cvmgpu.h")
    if not force and os.path.exists(CALC) and os.path.getmtime(CALC) > max(os.path.getmtime(src), os.path.getmtime(hdr), os.path.getmtime(LIB)):
        return CALC
    cmd = ["g++", "-O2", "-std=c++17", "-Wall", src, "-L", CSRC, "-lcvmgpu", "-Wl,-rpath,$ORIGIN", "-o", CALC]
    print("+", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return CALC


def build(force=False, verbose=False, extra=()):
    if not force and not needs_build():
        build_calc()
        return LIB
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "-cudart", "static",
           "-Xptxas", "-v" if verbose else "-O3", *extra,
           "-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    print("+", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    build_calc(force=True)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
