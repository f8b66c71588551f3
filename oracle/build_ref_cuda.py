"""Builds the REFERENCE's own CUDA extension (rendeirolab/stainx v0.1.4, `stainx_cuda_torch`) as a SECONDARY timing
comparator: the reference's GPU path, run next to this repo's on the same B200.

    python oracle/build_ref_cuda.py          # build container only: needs /root/reference

TEST / BENCH INFRASTRUCTURE ONLY.  Nothing in `stainx_b200/` loads the result.  The four sources are compiled where
they lie under /root/reference (never copied into this repo) with nvcc directly -- not through the reference's
setup.py, which refuses to build without a visible GPU (setup.py:L138-139) -- using the reference's own flags
(setup.py:L103: --use_fast_math -O3, one -gencode for the device's compute capability; sm_100 here).  Output:
`oracle/_ref/stainx_cuda_torch.so`, git-ignored, travels to the GPU box with the working tree.  It links against
the torch of this image (same image on the GPU box).

The module exposes the reference's four native entry points (bindings.cpp:L19-35):
histogram_matching / reinhard / macenko / macenko_fast (images, params...) -> Tensor.
"""
from __future__ import annotations

import subprocess
import sys
import sysconfig
from pathlib import Path

REF = Path("/root/reference")
OUT_DIR = Path(__file__).resolve().parent / "_ref"
OUT = OUT_DIR / "stainx_cuda_torch.so"
SOURCES = ["bindings.cpp", "histogram_matching.cu", "reinhard.cu", "macenko.cu"]


def build(force: bool = False) -> Path | None:
    src_dir = REF / "src" / "stainx_cuda_torch" / "csrc"
    if not src_dir.exists():
        return OUT if OUT.exists() else None  # GPU box: use the prebuilt file
    srcs = [src_dir / s for s in SOURCES]
    if OUT.exists() and not force and all(OUT.stat().st_mtime > s.stat().st_mtime for s in srcs):
        return OUT
    import torch
    from torch.utils import cpp_extension as ext

    OUT_DIR.mkdir(exist_ok=True)
    inc = [f"-I{p}" for p in ext.include_paths(device_type="cuda")] + [f"-I{sysconfig.get_paths()['include']}", f"-I{src_dir}", f"-I{REF}"]
    lib_dir = Path(torch.__file__).parent / "lib"
    flags = ["--expt-relaxed-constexpr", "--use_fast_math", "-std=c++17", "-O3", "-DNDEBUG", "-Xcompiler", "-funroll-loops", "-Xcompiler", "-ffast-math",
             "-Xcompiler", "-finline-functions", "-gencode", "arch=compute_100,code=sm_100", "-DTARGET_CUDA_ARCH=100",
             "-DTORCH_EXTENSION_NAME=stainx_cuda_torch", "-DTORCH_API_INCLUDE_EXTENSION_H", f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}",
             "-Xcompiler", "-fPIC", "-x", "cu"]
    objs = []
    for s in srcs:
        o = OUT_DIR / (s.name + ".o")
        subprocess.run(["nvcc", *flags, *inc, "-c", str(s), "-o", str(o)], check=True)
        objs.append(str(o))
    subprocess.run(["nvcc", "-shared", "-o", str(OUT), *objs, f"-L{lib_dir}", "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python", "-lcudart",
                    f"-Xlinker=-rpath={lib_dir}"], check=True)
    for o in objs:
        Path(o).unlink()
    return OUT


def load():
    """The extension module, or None when it was not built."""
    if not OUT.exists():
        return None
    import importlib.util

    import torch  # noqa: F401 - libtorch must be loaded first

    spec = importlib.util.spec_from_file_location("stainx_cuda_torch", OUT)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
