"""In-tree build of libpmp_edgecheck.so (sm_100a only) with nvcc.

The shared object is written next to the sources (plant_motion_planning_b200/_lib/) so it travels with
the repository snapshot; nothing is JIT-compiled at import time.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_DIR = os.path.join(PKG_DIR, "_lib")
LIB_PATH = os.path.join(LIB_DIR, "libpmp_edgecheck.so")
STAMP = os.path.join(LIB_DIR, "libpmp_edgecheck.stamp")
SOURCES = ["pmp_edgecheck.cu"]
HEADERS = ["pmp_kernels.cuh", "pmp_tables.h", os.path.join("..", "..", "include", "pmp_edgecheck.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--shared", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def source_digest() -> str:
    h = hashlib.sha256()
    for name in SOURCES + HEADERS:
        with open(os.path.join(CSRC, name), "rb") as fp:
            h.update(fp.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current() -> bool:
    if not (os.path.exists(LIB_PATH) and os.path.exists(STAMP)):
        return False
    with open(STAMP) as fp:
        return fp.read().strip() == source_digest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA library if the sources changed; returns the .so path."""
    if not force and is_current():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [_nvcc()] + NVCC_FLAGS + ["-o", LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    log = os.path.join(LIB_DIR, "build.log")
    with open(log, "w") as fp:
        fp.write(" ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
        raise RuntimeError(f"nvcc failed (see {log})")
    if verbose:
        sys.stderr.write(proc.stderr)
    with open(STAMP, "w") as fp:
        fp.write(source_digest())
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
