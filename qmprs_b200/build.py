"""Build libqmprs_b200.so in-tree with nvcc for sm_100a (no JIT cache: the built .so
travels with the repo snapshot to the GPU box)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = [os.path.join(HERE, "csrc", f) for f in ("api.cu", "gemm.cu", "svd.cu", "qr.cu", "mps_ops.cu", "dense.cu", "dense_small.cu", "dense_persist.cu", "small_svd.cu", "small_mps.cu")]
OUT = os.path.join(HERE, "libqmprs_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"),
    "-I", os.path.join(HERE, "csrc"), "-cudart", "static",
]


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = SRC + [os.path.join(ROOT, "include", "qmprs_b200.h"), os.path.join(HERE, "csrc", "common.cuh"),
            os.path.join(HERE, "csrc", "polar.cuh"), os.path.join(HERE, "csrc", "small_linalg.cuh"),
            os.path.join(HERE, "csrc", "chi2_select.cuh")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + SRC + ["-o", OUT]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libqmprs_b200.so")
    if verbose:
        sys.stderr.write(res.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
