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
SOURCES = ["pmp_edgecheck.cu", "pmp_inst_a12.cu", "pmp_inst_a16.cu", "pmp_inst_a18.cu", "pmp_inst_a20.cu", "pmp_inst_a32.cu",
           "pmp_inst_a12f.cu", "pmp_inst_a16f.cu", "pmp_inst_a18f.cu", "pmp_inst_a20f.cu", "pmp_inst_a32f.cu"]
HEADERS = ["pmp_kernels.cuh", "pmp_tables.h", "pmp_launch.h", "pmp_launch_impl.cuh", "pmp_tree.cuh", "pmp_hull.cuh", "pmp_sampler.cuh",
           os.path.join("..", "..", "include", "pmp_edgecheck.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


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
    obj_dir = os.path.join(LIB_DIR, "obj")
    os.makedirs(obj_dir, exist_ok=True)
    nvcc = _nvcc()
    log = os.path.join(LIB_DIR, "build.log")

    def compile_one(src: str):
        obj = os.path.join(obj_dir, os.path.splitext(src)[0] + ".o")
        cmd = [nvcc] + NVCC_FLAGS + ["-c", "-o", obj, os.path.join(CSRC, src)]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        return src, obj, cmd, proc

    # one translation unit per sphere-count padding: the template instantiations compile in parallel
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        results = list(pool.map(compile_one, SOURCES))
    text = ""
    for src, obj, cmd, proc in results:
        text += " ".join(cmd) + "\n" + proc.stdout + proc.stderr + "\n"
    link = [nvcc, "--shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + [r[1] for r in results]
    failed = [r for r in results if r[3].returncode != 0]
    lproc = None
    if not failed:
        lproc = subprocess.run(link, capture_output=True, text=True)
        text += " ".join(link) + "\n" + lproc.stdout + lproc.stderr
    with open(log, "w") as fp:
        fp.write(text)
    if failed or lproc.returncode != 0:
        sys.stderr.write(text[-4000:])
        raise RuntimeError(f"nvcc failed (see {log})")
    if verbose:
        sys.stderr.write(text)
    shutil.rmtree(obj_dir, ignore_errors=True)      # 18 MB of objects that need not travel with the snapshot
    with open(STAMP, "w") as fp:
        fp.write(source_digest())
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
