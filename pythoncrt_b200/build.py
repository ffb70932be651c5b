"""Build `pythoncrt_b200/libcrt_b200.so` in-tree with nvcc for sm_100a.

    python -m pythoncrt_b200.build [--force]

nvcc cross-compiles without a GPU.  -fmad=false: no implicit multiply-add
contraction anywhere (fused multiply-adds are written explicitly where OpenCV
uses them, see csrc/crt_math.cuh); -lineinfo so ncu's source page maps to the
.cu/.cuh files.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libcrt_b200.so")
SOURCES = ["crt_abi.cu"]
DEPS = ["crt_abi.cu", "crt_math.cuh", "crt_stages.cuh", "crt_kernels.cuh", "crt_fused.cuh", "crt_fused_gauss.cuh", "crt_fused_ps2.cuh", "crt_fused_gauss_ps2.cuh", "crt_gather_tile.cuh", "crt_tma.cuh", "crt_pow_tables.h", "crt_derive.h",
        os.path.join("..", "..", "include", "crt_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-fmad=false",
              "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v"]


def nvcc_path() -> str:
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def up_to_date() -> bool:
    if not os.path.isfile(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(os.path.join(CSRC, d)) <= t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return OUT
    cmd = [nvcc_path(), *NVCC_FLAGS, "-o", OUT, *[os.path.join(CSRC, s) for s in SOURCES]]
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    log = os.path.join(HERE, "build_ptxas.log")
    with open(log, "w") as f:
        f.write(res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
