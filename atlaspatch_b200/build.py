"""Builds libatlaspatch_b200.so (in-tree) with nvcc for sm_100a.  No torch dependency in the library."""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
LIB = PKG / "libatlaspatch_b200.so"
STAMP = PKG / ".build_stamp"

SOURCES = ["ctx.cu", "gemm_tcgen05.cu", "encoder_kernels.cu", "preprocess_resize.cu", "attention_tcgen05.cu", "attention_units.cu", "encoder.cu", "coords.cu", "patch_filter.cu", "slide_kernels.cu", "sam2_kernels.cu", "sam2.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--use_fast_math",
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function", "-shared", "-cudart", "static",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libatlaspatch_b200.so cannot be built")


def _digest() -> str:
    h = hashlib.sha256()
    for f in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [ROOT / "include" / "atlaspatch_b200.h"]):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    digest = _digest()
    if not force and LIB.exists() and STAMP.exists() and STAMP.read_text().strip() == digest:
        return LIB
    cmd = [_nvcc(), *NVCC_FLAGS, f"-I{ROOT / 'include'}", f"-I{CSRC}", "-o", str(LIB)]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [str(CSRC / s) for s in SOURCES]
    cmd += ["-lpthread"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed ({res.returncode}) building {LIB.name}")
    STAMP.write_text(digest + "\n")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
