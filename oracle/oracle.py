"""ctypes front-end of the CPU oracle ``oracle/stainx_oracle.c``.

TEST INFRASTRUCTURE ONLY: imported by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.  The product package
``stainx_b200`` never imports this module.

All functions take and return numpy arrays (NCHW, C=3, uint8 or float32) and mirror the
reference's torch CPU backend entry points (``src/stainx/backends/torch_backend.py``):

=============================  ==========================================================
``hm_fit / hm_transform``      ``HistogramMatchingTorch.compute_reference_histograms_torch``
                               (L143-179) / ``.transform`` (L194-301)
``reinhard_fit / _transform``  ``ReinhardTorch.compute_reference_mean_std_torch`` (L308-323)
                               / ``.transform`` (L325-355)
``macenko_fit / _transform``   ``MacenkoTorch.compute_reference_stain_matrix_torch``
                               (L463-519) / ``.transform`` (L521-560)
=============================  ==========================================================
"""
from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_build" / "libstainx_oracle.so"
_lib = None

_c_i64 = ctypes.c_int64
_c_int = ctypes.c_int
_c_vp = ctypes.c_void_p


def build(force: bool = False) -> Path:
    """Compile the oracle with gcc (``make -C oracle``)."""
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < (_HERE / "stainx_oracle.c").stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-B" if force else "-s"], check=True, capture_output=True)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(str(_LIB_PATH))
        _lib.ox_num_threads.restype = _c_int
        _lib.ox_sum_f32.restype = ctypes.c_float
        _lib.ox_sum_f32.argtypes = [_c_vp, _c_i64]
        _lib.ox_macenko_fit.restype = _c_i64
    return _lib


def num_threads() -> int:
    return int(lib().ox_num_threads())


def set_num_threads(n: int | None = None) -> int:
    """Set the OpenMP thread count explicitly (default: every host core).  torchrun exports
    OMP_NUM_THREADS=1 to its workers, which would silently make a CPU baseline single-threaded."""
    import os

    lib().ox_set_num_threads(_c_int(int(n) if n else (os.cpu_count() or 1)))
    return num_threads()


def _prep(img: np.ndarray) -> tuple[np.ndarray, int, int, int]:
    if img.ndim != 4 or img.shape[1] != 3:
        raise ValueError(f"oracle expects NCHW with C=3, got {img.shape}")
    if img.dtype == np.uint8:
        dt = 0
    elif img.dtype == np.float32:
        dt = 1
    else:
        raise TypeError(f"oracle handles uint8/float32, got {img.dtype}")
    img = np.ascontiguousarray(img)
    return img, dt, img.shape[0], img.shape[2] * img.shape[3]


def _p(a: np.ndarray) -> ctypes.c_void_p:
    return ctypes.c_void_p(a.ctypes.data)


def sum_f32(a: np.ndarray) -> np.float32:
    a = np.ascontiguousarray(a, dtype=np.float32)
    return np.float32(lib().ox_sum_f32(_p(a), _c_i64(a.size)))


# ---------------------------------------------------------------- histogram matching
def hm_counts(img: np.ndarray) -> np.ndarray:
    img, dt, n, hw = _prep(img)
    counts = np.zeros((3, 256), dtype=np.int64)
    lib().ox_hm_counts(_p(img), _c_int(dt), _c_i64(n), _c_i64(hw), _p(counts))
    return counts


def hm_fit(img: np.ndarray) -> np.ndarray:
    """-> float32 (3, 256): the reference's ``_ref_histograms_256`` stacked."""
    img, dt, n, hw = _prep(img)
    ref = np.zeros((3, 256), dtype=np.float32)
    lib().ox_hm_fit(_p(img), _c_int(dt), _c_i64(n), _c_i64(hw), _p(ref))
    return ref


def hm_ref_cdf(ref_hist: np.ndarray) -> np.ndarray:
    ref_hist = np.ascontiguousarray(ref_hist, dtype=np.float32)
    out = np.zeros((3, 256), dtype=np.float32)
    lib().ox_hm_ref_cdf(_p(ref_hist), _p(out))
    return out


def hm_lut(counts: np.ndarray, npix: int, ref_cdf: np.ndarray) -> np.ndarray:
    counts = np.ascontiguousarray(counts, dtype=np.int64)
    ref_cdf = np.ascontiguousarray(ref_cdf, dtype=np.float32)
    lut = np.zeros((3, 256), dtype=np.float32)
    lib().ox_hm_lut(_p(counts), _c_i64(npix), _p(ref_cdf), _p(lut))
    return lut


def hm_transform(img: np.ndarray, ref_hist: np.ndarray) -> np.ndarray:
    img, dt, n, hw = _prep(img)
    ref_hist = np.ascontiguousarray(ref_hist, dtype=np.float32)
    out = np.empty_like(img)
    lib().ox_hm_transform(_p(img), _c_int(dt), _c_i64(n), _c_i64(hw), _p(ref_hist), _p(out))
    return out


# ---------------------------------------------------------------- reinhard
def reinhard_fit(img: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    img, dt, n, hw = _prep(img)
    mean = np.zeros(3, dtype=np.float32)
    std = np.zeros(3, dtype=np.float32)
    lib().ox_reinhard_fit(_p(img), _c_int(dt), _c_i64(n), _c_i64(hw), _p(mean), _p(std))
    return mean, std


def reinhard_transform(img: np.ndarray, ref_mean: np.ndarray, ref_std: np.ndarray) -> np.ndarray:
    img, dt, n, hw = _prep(img)
    ref_mean = np.ascontiguousarray(ref_mean, dtype=np.float32)
    ref_std = np.ascontiguousarray(ref_std, dtype=np.float32)
    out = np.empty_like(img)
    lib().ox_reinhard_transform(_p(img), _c_int(dt), _c_i64(n), _c_i64(hw), _p(ref_mean), _p(ref_std), _p(out))
    return out


def rgb_to_lab(img: np.ndarray) -> np.ndarray:
    img, dt, n, hw = _prep(img)
    lab = np.empty(img.shape, dtype=np.float32)
    lib().ox_rgb_to_lab_image(_p(img), _c_int(dt), _c_i64(n), _c_i64(hw), _p(lab))
    return lab


def lab_to_rgb(lab: np.ndarray) -> np.ndarray:
    lab = np.ascontiguousarray(lab, dtype=np.float32)
    rgb = np.empty_like(lab)
    lib().ox_lab_to_rgb_image(_p(lab), _c_i64(lab.shape[0]), _c_i64(lab.shape[2] * lab.shape[3]), _p(rgb))
    return rgb


# ---------------------------------------------------------------- macenko
def macenko_fit(img: np.ndarray, mid_sign: int = 1) -> tuple[np.ndarray, np.ndarray]:
    """-> HE float32 (3, 2), maxC float32 (2,). ``mid_sign`` flips the middle eigenvector."""
    img, dt, n, hw = _prep(img)
    he = np.zeros((3, 2), dtype=np.float32)
    maxc = np.zeros(2, dtype=np.float32)
    kept = lib().ox_macenko_fit(_p(img), _c_int(dt), _c_i64(n), _c_i64(hw), _c_int(mid_sign), _p(he), _p(maxc))
    if kept <= 0:
        raise RuntimeError("macenko_fit: no pixel passes the OD mask")
    return he, maxc


def macenko_transform(img: np.ndarray, he_ref: np.ndarray, maxc_ref: np.ndarray, mid_signs=None, return_fit: bool = False):
    """-> output in the input dtype (float32 stays in [0, 255]); optionally per-image HE/maxC."""
    img, dt, n, hw = _prep(img)
    he_ref = np.ascontiguousarray(he_ref, dtype=np.float32).reshape(3, 2)
    maxc_ref = np.ascontiguousarray(maxc_ref, dtype=np.float32).reshape(2)
    out = np.empty_like(img)
    he_out = np.zeros((n, 3, 2), dtype=np.float32)
    maxc_out = np.zeros((n, 2), dtype=np.float32)
    signs_p = None
    if mid_signs is not None:
        signs = np.ascontiguousarray(mid_signs, dtype=np.int32).reshape(n)
        signs_p = _p(signs)
    lib().ox_macenko_transform(_p(img), _c_int(dt), _c_i64(n), _c_i64(hw), _p(he_ref), _p(maxc_ref), signs_p, _p(out), _p(he_out), _p(maxc_out))
    if return_fit:
        return out, he_out, maxc_out
    return out
