#!/usr/bin/env python3
"""Build gauss-process_b200/lib/libb200id.so (hand-written sm_100a CUDA behind the C ABI).

    python gauss-process_b200/build.py [--force] [--verbose]

nvcc cross-compiles without a GPU.  The library is built IN-TREE so that it travels to the GPU
box with the repo snapshot; it is git-ignored (*.so).  No --use_fast_math: the path is FP64 parity
work.
"""
from __future__ import annotations

import argparse
import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB_DIR = PKG / "lib"
LIB = LIB_DIR / "libb200id.so"
STAMP = LIB_DIR / "libb200id.stamp"
SOURCES = ["capi.cu", "frf_project.cu", "gp_batch.cu", "gp_batch_tiled.cu", "gp_pinv.cu", "gp_large.cu", "chol_blocked.cu", "fir_conv.cu", "fir_fft.cu", "kr_fir.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def source_digest() -> str:
    h = hashlib.sha256()
    files = sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "b200id.h", Path(__file__)]
    for f in files:
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(os.environ.get("B200ID_NVCC_FLAGS", "").encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    LIB_DIR.mkdir(parents=True, exist_ok=True)
    digest = source_digest()
    if not force and LIB.exists() and STAMP.exists() and STAMP.read_text().strip() == digest:
        return LIB
    srcs = [str(CSRC / s) for s in SOURCES if (CSRC / s).exists()]
    extra = os.environ.get("B200ID_NVCC_FLAGS", "").split()      # e.g. -DGT_PROFILE for the phase-timing build
    cmd = [nvcc_path(), "-O3", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-lineinfo", *ARCH,
           "-Xcompiler", "-ffp-contract=off", *extra, "-o", str(LIB), *srcs]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libb200id.so")
    STAMP.write_text(digest)
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(a.force, a.verbose))
