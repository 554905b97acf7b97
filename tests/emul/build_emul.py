"""Compile the kernel sources for the HOST against tests/emul/cuda_emul.h (test-only).

The result, tests/emul/_build/libmacjd_emul.so, exports the same C ABI as the product
library but runs every kernel on host memory with fibers; the CPU test-suite uses it to
check kernel logic without a GPU.  The product package never loads it.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "ma-cjd-cooperative-jamming-decision-making-via-marl_b200", "csrc")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libmacjd_emul.so")


def _newest():
    m = os.path.getmtime(os.path.join(HERE, "cuda_emul.h"))
    for root in (CSRC, os.path.join(ROOT, "include")):
        for f in os.listdir(root):
            if f.endswith((".cu", ".cuh", ".h")):
                m = max(m, os.path.getmtime(os.path.join(root, f)))
    return m


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _newest():
        return LIB
    cmd = ["g++", "-std=c++17", "-O2", "-g", "-fPIC", "-shared", "-x", "c++",
           "-DMACJD_TEST_HOST_EMULATION", "-Wall", "-Wno-unused-function", "-Wno-unknown-pragmas",
           "-Wno-unused-variable", "-Wno-sign-compare",
           "-I", HERE, "-I", CSRC, os.path.join(CSRC, "macjd_api.cu"), "-o", LIB]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("host-emulation build failed:\n" + res.stdout + res.stderr)
    if res.stderr.strip():
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
