"""Build libtealeaf_b200.so in-tree with nvcc for sm_100a (and nothing else)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(CSRC, "libtealeaf_b200.so")
SOURCES = ["tl_api.cu"]
DEPS = ["tl_api.cu", "tl_device.cuh", "tl_kernels_basic.cuh", "tl_kernels_fused.cuh", "tl_kernels_ring.cuh", "tl_kernels_persist.cuh", "tl_kernels_tma.cuh", "tl_multi.inl", "tl_eigen.h",
        os.path.join("..", "..", "include", "tealeaf_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    # no FMA contraction: per-cell arithmetic is expression-for-expression the oracle's
    # (and Julia's, which never contracts implicitly), so fields match bit for bit
    "-fmad=false",
    "-Xcompiler", "-fPIC", "-shared", "-DTL_WITH_NCCL",
]


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, *NVCC_FLAGS, "-o", OUT, *SOURCES, "-ldl"]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
    r = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libtealeaf_b200.so")
    if verbose:
        sys.stderr.write(r.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
