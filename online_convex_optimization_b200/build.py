"""Build liboco_b200.so (hand-written sm_100a CUDA + the C ABI) in-tree with nvcc.

    python -m online_convex_optimization_b200.build [--force]

Each .cu is compiled to an object in parallel, then linked into
online_convex_optimization_b200/csrc/liboco_b200.so.  Flags that matter:
  -gencode arch=compute_100a,code=sm_100a   Blackwell B200 only, no other arch
  -fmad=false                               the reference's numba kernels emit no FMA; fusing
                                            mul+add would break bit-exact parity
  -lineinfo                                 ncu source view
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(CSRC, "liboco_b200.so")
SOURCES = ["abi.cu", "kernels_seeded.cu", "kernels_sweep_gpath.cu", "kernels_sweep_iid.cu", "kernels_gsweep_lo.cu", "kernels_gsweep_hi.cu", "kernels_gsweep_wide.cu", "kernels_stream.cu", "kernels_warp.cu", "kernels_wide_tma.cu", "kernels_exact.cu", "kernels_wide_exact.cu", "kernels_fill_warp.cu", "kernels_wide_smart.cu", "kernels_wide_ipm.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build liboco_b200.so")


def _digest() -> str:
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for name in sorted(os.listdir(CSRC)):
        if name.endswith((".cu", ".cuh", ".h")):
            h.update(name.encode())
            h.update(open(os.path.join(CSRC, name), "rb").read())
    h.update(open(os.path.join(HERE, "..", "include", "oco_b200.h"), "rb").read())
    return h.hexdigest()


def _compile(nvcc: str, src: str) -> str:
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    with open(obj + ".log", "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile if sources changed since the last build; returns the .so path."""
    stamp = os.path.join(OBJ, "digest.txt")
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == digest:
        return LIB
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    with cf.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(lambda s: _compile(nvcc, s), SOURCES))
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    open(stamp, "w").write(digest)
    if verbose:
        print("built", LIB)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
