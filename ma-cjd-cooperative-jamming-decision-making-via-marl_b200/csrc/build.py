"""Build libmacjd_b200.so (sm_100a) in-tree with nvcc.  Used by __graft_entry__.build()."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libmacjd_b200.so")
SOURCES = ["macjd_api.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def newest_source_mtime():
    m = 0.0
    for root in (HERE, os.path.join(HERE, "..", "..", "include")):
        for f in os.listdir(root):
            if f.endswith((".cu", ".cuh", ".h")):
                m = max(m, os.path.getmtime(os.path.join(root, f)))
    return m


def build(force=False, verbose=False, out=None):
    global LIB
    if out:
        LIB = out
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= newest_source_mtime():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("NVCC_EXTRA", "").split()
    cmd = [nvcc] + NVCC_FLAGS + extra + [os.path.join(HERE, s) for s in SOURCES] + ["-o", LIB]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + log)
    if verbose:
        print(log)
    return LIB


if __name__ == "__main__":
    outs = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose=True, out=outs[0] if outs else None))
