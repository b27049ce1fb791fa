"""Build the C-ABI CUDA library (fcd_b200/csrc/*.cu -> fcd_b200/libfcd_b200.so) for sm_100a with nvcc.

The library has no torch dependency: plain `extern "C"` entry points declared in include/fcd_b200.h.
"""
from __future__ import annotations

import glob
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfcd_b200.so")
STAMP = os.path.join(HERE, "csrc", ".build_stamp")

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
              "-I", os.path.join(os.path.dirname(HERE), "include")]


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _digest():
    h = hashlib.sha256()
    for f in _sources() + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + \
            sorted(glob.glob(os.path.join(os.path.dirname(HERE), "include", "*.h"))):
        h.update(os.path.basename(f).encode())      # names, not absolute paths: the tree is copied to the GPU box
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(a for a in NVCC_FLAGS if not os.path.isabs(a)).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    # one builder at a time: the ranks of a multi-process job may all find a stale library at import
    import fcntl
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    with open(os.path.join(HERE, "build", ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            return _build_locked(force, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(force: bool, verbose: bool) -> str:
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == dig:
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for src in _sources():
        obj = os.path.join(HERE, "build", os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [nvcc, *NVCC_FLAGS, "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- nvcc {os.path.basename(src)} (rc={p.returncode})\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    tmp = LIB + ".tmp"
    subprocess.check_call([nvcc, "-shared", "-o", tmp, *objs, "-gencode", "arch=compute_100a,code=sm_100a"])
    os.replace(tmp, LIB)        # atomic: a concurrent loader sees the old or the new file, never half of one
    with open(STAMP, "w") as f:
        f.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
