"""Build libcrisper.so (sm_100a only) in-tree with nvcc.  Used by __graft_entry__.build() and by developers.

    python -m crisperwhisper_b200.build [--force]

The shared library is a plain C-ABI (include/crisper.h): it links cudart statically and nothing from torch, so it
is loaded with ctypes and travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "libcrisper.so")
SOURCES = ["api.cu", "logmel.cu", "align.cu", "gemm.cu", "encoder.cu", "decoder.cu", "resample.cu", "postproc.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "--use_fast_math=false"]
FLAGS = [f for f in FLAGS if f != "--use_fast_math=false"]  # precise math everywhere (parity), flag kept for clarity


def _newer(a: str, bs) -> bool:
    if not os.path.exists(a):
        return False
    ta = os.path.getmtime(a)
    return all(os.path.getmtime(b) <= ta for b in bs)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "crisper.h"))
    jobs = []
    objs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, s.replace(".cu", ".o"))
        objs.append(obj)
        if force or not _newer(obj, [src] + headers):
            jobs.append([NVCC] + FLAGS + ["-c", src, "-o", obj])

    def run(cmd):
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return r

    if jobs:
        with ThreadPoolExecutor(max_workers=min(7, len(jobs))) as ex:
            list(ex.map(run, jobs))
    if force or jobs or not _newer(LIB, objs):
        run([NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
