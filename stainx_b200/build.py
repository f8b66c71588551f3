"""Build ``libstainx_b200.so`` in-tree with nvcc for sm_100a.

    python -m stainx_b200.build [--force]

One translation unit per ``csrc/*.cu``, linked into ``stainx_b200/_lib/libstainx_b200.so``
(git-ignored; it travels to the GPU box with the working tree).  nvcc cross-compiles without a
GPU.  No ``--use_fast_math``: fast intrinsics are chosen per call site in the kernels, and the
LUT arithmetic of histogram matching must stay IEEE-exact.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB_DIR = PKG / "_lib"
LIB = LIB_DIR / "libstainx_b200.so"
SOURCES = ["lib.cu", "hm.cu", "reinhard.cu", "macenko.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-fvisibility=default",
]


def _nvcc() -> str:
    cand = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(cand).exists():
        raise RuntimeError("nvcc not found; set NVCC or install the CUDA toolkit")
    return cand


def _stale(target: Path, deps: list[Path]) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    LIB_DIR.mkdir(exist_ok=True)
    headers = [CSRC / "common.cuh", PKG.parent / "include" / "stainx_b200.h"]
    objects = []
    nvcc = _nvcc()
    for src in SOURCES:
        obj = LIB_DIR / (src + ".o")
        if force or _stale(obj, [CSRC / src, *headers]):
            cmd = [nvcc, *NVCC_FLAGS, *os.environ.get("SX_EXTRA_NVCC_FLAGS", "").split(), "-c", str(CSRC / src), "-o", str(obj)]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd))
            subprocess.run(cmd, check=True)
        objects.append(obj)
    if force or _stale(LIB, objects):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB), *map(str, objects)]
        if verbose:
            print(" ".join(cmd))
        subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv or "--verbose" in sys.argv)
    print(path)
