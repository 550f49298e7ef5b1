"""Builds liblsk.so (hand-written sm_100a CUDA behind a C ABI) in-tree with nvcc.

    python -m liestationarykernels_b200.build        # or __graft_entry__.build()

No JIT cache, no torch extension machinery: the shared object sits next to this file so that it travels to the
GPU box with the repository snapshot and shows up as an in-tree .so in the loaded-library record.
"""
from __future__ import annotations

import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "liblsk.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "lsk.h"))
    headers = [h for h in headers if os.path.exists(h)]
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    jobs = []
    for src in _sources():
        obj = os.path.join(objdir, src[:-3] + ".o")
        if force or _stale(obj, [os.path.join(CSRC, src)] + headers):
            cmd = [NVCC] + ARCH + FLAGS + ["-I", os.path.join(os.path.dirname(HERE), "include"), "-c", os.path.join(CSRC, src), "-o", obj]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append(cmd)
    with concurrent.futures.ThreadPoolExecutor(max_workers=8) as pool:
        for res in pool.map(lambda c: subprocess.run(c, capture_output=True, text=True), jobs):
            if verbose and res.stderr:
                print(res.stderr, file=sys.stderr)
            if res.returncode != 0:
                raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(res.args), res.stderr))
    objs = [os.path.join(objdir, s[:-3] + ".o") for s in _sources()]
    if force or jobs or _stale(LIB, objs):
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("link failed:\n%s" % res.stderr)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
