"""Build `pythoncrt_b200/libcrt_b200.so` in-tree with nvcc for sm_100a.

    python -m pythoncrt_b200.build [--force]

nvcc cross-compiles without a GPU.  The library is several translation units (one per kernel
family, csrc/crt_tu_*.cu, plus the C-ABI / host logic in csrc/crt_abi.cu) compiled in parallel and
linked into one shared object.  -fmad=false: no implicit multiply-add contraction anywhere (fused
multiply-adds are written explicitly where OpenCV uses them, see csrc/crt_math.cuh); -lineinfo so
ncu's source page maps to the .cu/.cuh files.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
OUT = os.path.join(HERE, "libcrt_b200.so")
SOURCES = ["crt_abi.cu", "crt_tu_ps2.cu", "crt_tu_gauss_ps2.cu", "crt_tu_gauss.cu", "crt_tu_fused.cu", "crt_tu_staged.cu", "crt_tu_warp_ps2.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-fmad=false",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def nvcc_path() -> str:
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _deps():
    return (glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) +
            [os.path.join(HERE, "..", "include", "crt_b200.h")])


def _sources():
    return [s for s in SOURCES if os.path.isfile(os.path.join(CSRC, s))]


def up_to_date() -> bool:
    if not os.path.isfile(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(os.path.getmtime(d) <= t for d in _deps())


def _compile(src: str, force: bool):
    """One translation unit -> object file; returns (src, ptxas log).  A unit is rebuilt when any header is newer
    (coarse, but a header change almost always touches every unit)."""
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    log = os.path.join(OBJ, src.replace(".cu", ".log"))
    if not force and os.path.isfile(obj) and os.path.isfile(log):
        t = os.path.getmtime(obj)
        own = os.path.join(CSRC, src)
        headers = [d for d in _deps() if not d.endswith(".cu")]
        if os.path.getmtime(own) <= t and all(os.path.getmtime(h) <= t for h in headers):
            return src, open(log).read()
    cmd = [nvcc_path(), *NVCC_FLAGS, "-c", "-o", obj, os.path.join(CSRC, src)]
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    with open(log, "w") as f:
        f.write(res.stdout + res.stderr)
    return src, res.stdout + res.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return OUT
    os.makedirs(OBJ, exist_ok=True)
    srcs = _sources()
    with ThreadPoolExecutor(max_workers=min(len(srcs), os.cpu_count() or 4)) as pool:
        logs = list(pool.map(lambda s: _compile(s, force), srcs))
    objs = [os.path.join(OBJ, s.replace(".cu", ".o")) for s in srcs]
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC", "-o", OUT, *objs]
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    with open(os.path.join(HERE, "build_ptxas.log"), "w") as f:
        for s, text in logs:
            f.write(f"==== {s} ====\n{text}\n")
    if verbose:
        for s, text in logs:
            print(f"==== {s} ====\n{text}")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
