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


def source_hash(extra=()):
    """sha256 over every source the library is compiled from (+ the flags): what a reuse is checked against."""
    import hashlib
    h = hashlib.sha256()
    for root in (HERE, os.path.join(HERE, "..", "..", "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cu", ".cuh", ".h")):
                h.update(f.encode())
                with open(os.path.join(root, f), "rb") as fh:
                    h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS + list(extra)).encode())
    return h.hexdigest()


def build(force=False, verbose=False, out=None):
    """force=True: compile.  force=False: reuse the in-tree library only if build_info.json says it was
    compiled from exactly these sources and flags; otherwise compile."""
    import json
    import time
    global LIB
    if out:
        LIB = out
    extra = os.environ.get("NVCC_EXTRA", "").split()
    digest = source_hash(extra)
    info_path = LIB + ".build_info.json"
    if not force and os.path.exists(LIB):
        try:
            with open(info_path) as f:
                if json.load(f).get("source_sha256") == digest:
                    return LIB
        except (OSError, ValueError):
            pass
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not os.path.exists(nvcc):
        raise RuntimeError(f"{LIB} is missing or stale and nvcc ({nvcc}) is not available to rebuild it")
    cmd = [nvcc] + NVCC_FLAGS + extra + [os.path.join(HERE, s) for s in SOURCES] + ["-o", LIB]
    t0 = time.time()
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + log)
    ver = subprocess.run([nvcc, "--version"], capture_output=True, text=True).stdout.strip().splitlines()[-1:]
    with open(info_path, "w") as f:
        json.dump({"source_sha256": digest, "compiled": True, "seconds": round(time.time() - t0, 1), "nvcc": ver,
                   "flags": NVCC_FLAGS + extra, "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())}, f, indent=1)
    if verbose:
        print(log)
    return LIB


if __name__ == "__main__":
    outs = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose=True, out=outs[0] if outs else None))
