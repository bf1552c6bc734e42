"""Builds librf_b200.so (the product) in-tree with nvcc for sm_100a.

    python -m retrofire_b200.build [--force]

Flags are part of the numerics contract: no FMA contraction, IEEE division and
square root, denormals preserved, never --use_fast_math.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "librf_b200.so")
SOURCES = ["rf_api.cu"]
HEADERS = ["rf_device.cuh", "rf_geometry.cuh", "rf_raster.cuh", "rf_order.cuh", "rf_peer.cuh", os.path.join("..", "..", "include", "retrofire_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-shared", "-Xcompiler", "-fPIC", "-diag-suppress", "177",
]


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if force or stale():
        cmd = [nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + [os.path.join(CSRC, s) for s in SOURCES]
        subprocess.run(cmd, check=True, cwd=CSRC)
    return OUT


def build_variant(name: str, defines: dict) -> str:
    """Tuning aid: the same library with -D overrides of the #ifndef-guarded constants, as retrofire_b200/_variants/<name>.so.
    `RF_B200_LIB=<path>` makes _ffi load it (A/B timing of two builds inside one GPU session, see scratch/ab.sh)."""
    vdir = os.path.join(HERE, "_variants")
    os.makedirs(vdir, exist_ok=True)
    out = os.path.join(vdir, f"{name}.so")
    cmd = [nvcc()] + NVCC_FLAGS + [f"-D{k}={v}" for k, v in defines.items()] + ["-o", out] + [os.path.join(CSRC, s) for s in SOURCES]
    subprocess.run(cmd, check=True, cwd=CSRC)
    return out


def ptx(path: str) -> str:
    """Emit PTX (for the no-FMA check in tests/test_build.py)."""
    cmd = [nvcc(), "-gencode", "arch=compute_100a,code=compute_100a", "-O3", "-std=c++17", "-fmad=false", "-prec-div=true",
           "-prec-sqrt=true", "-ftz=false", "-diag-suppress", "177", "-ptx", "-o", path, os.path.join(CSRC, SOURCES[0])]
    subprocess.run(cmd, check=True, cwd=CSRC)
    return path


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
