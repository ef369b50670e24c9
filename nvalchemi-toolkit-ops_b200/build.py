"""Build libnvalchemi_nl_b200.so for sm_100a with nvcc (in-tree, next to the sources).

Usage:  python build.py [--force]
The library is plain CUDA C++ with a C ABI (include/nvalchemi_nl_b200.h); it does not link torch.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libnvalchemi_nl_b200.so")
SOURCES = ["nvnl_api.cu"]
HEADERS = ["nvnl_common.cuh", "nvnl_build.cuh", "nvnl_sweep.cuh", "nvnl_fast.cuh", "nvnl_rows.cuh", "nvnl_pair.cuh", "nvnl_cache.cuh", os.path.join("..", "..", "include", "nvalchemi_nl_b200.h")]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=...)")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    cmd = [
        _nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
        "-Xcompiler", "-fPIC", "-shared", "-o", LIB,
    ] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.check_call(cmd, cwd=CSRC)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
