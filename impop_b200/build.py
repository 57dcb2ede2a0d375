"""Build libimpop_b200.so in-tree with nvcc for sm_100a (the only target)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libimpop_b200.so")
SOURCES = ("api.cu", "window_kernels.cu", "aux_kernels.cu", "ingest.cpp")
HEADERS = ("common.cuh", "stats_math.cuh", os.path.join("..", "..", "include", "impop_b200.h"))
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-O3,-pthread", "-shared"]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libimpop_b200.so cannot be built")


def stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB_PATH
    extra = os.environ.get("IMPOP_NVCC_DEFS", "").split()      # experiment switches, e.g. -DIMPOP_EPI_I2F=1
    cmd = [_nvcc(), *NVCC_FLAGS, *extra, "-o", LIB_PATH, *[os.path.join(CSRC, s) for s in SOURCES]]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    subprocess.check_call(cmd)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
