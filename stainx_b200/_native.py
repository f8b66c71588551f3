"""ctypes binding of ``libstainx_b200.so`` (the C ABI declared in ``include/stainx_b200.h``).

Plays the role of the reference's extension package ``stainx_cuda_torch``
(``src/stainx_cuda_torch/__init__.py:L30-51``): it exposes ``FUNCTIONS_AVAILABLE`` and the native
entry points.  There is no fallback: if the shared library is missing, every call raises.
"""
from __future__ import annotations

import ctypes
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("STAINX_B200_LIB", _PKG / "_lib" / "libstainx_b200.so"))

_i64 = ctypes.c_int64
_int = ctypes.c_int
_vp = ctypes.c_void_p
_f32 = ctypes.c_float

# name -> argtypes; every function returns int status unless listed in _RESTYPES.
PROTOTYPES: dict[str, list] = {
    "sx_abi_version": [],
    "sx_last_error": [],
    "sx_device_info": [ctypes.POINTER(_int), ctypes.POINTER(_int), ctypes.POINTER(_int), ctypes.POINTER(_i64)],
    "sx_kernel_launches": [],
    "sx_peer_status": [ctypes.POINTER(_int), ctypes.POINTER(_int), ctypes.POINTER(ctypes.c_uint32)],
    "sx_peer_status_clear": [],
    # histogram matching
    "sx_hm_hist": [_vp, _int, _int, _i64, _i64, _i64, _vp, _vp],
    "sx_hm_ref_hist": [_vp, _vp, _vp],
    "sx_hm_ref_cdf": [_vp, _vp, _vp],
    "sx_hm_build_lut": [_vp, _i64, _vp, _vp, _vp],
    "sx_hm_apply": [_vp, _int, _int, _i64, _i64, _i64, _vp, _vp, _vp],
    "sx_hm_peer_buffer_bytes": [],
    "sx_hm_build_lut_peers": [_vp, _int, _int, ctypes.c_uint32, _vp, _vp, _vp, _vp],
    "sx_hm_transform_peers": [_vp, _int, _int, _i64, _i64, _i64, _vp, _vp, _int, _int, ctypes.c_uint32, _vp, _vp, _vp, _i64, _vp],
    "sx_hm_workspace_bytes": [],
    "sx_hm_transform": [_vp, _int, _int, _i64, _i64, _i64, _vp, _vp, _vp, _i64, _vp],
    "sx_hm_fit": [_vp, _int, _int, _i64, _i64, _i64, _vp, _vp, _i64, _vp],
    # reinhard
    "sx_reinhard_stats": [_vp, _int, _i64, _i64, _i64, _vp, _vp],
    "sx_reinhard_finalize": [_vp, _vp, _vp, _vp],
    "sx_reinhard_apply": [_vp, _int, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp],
    "sx_reinhard_peer_buffer_bytes": [],
    "sx_reinhard_finalize_peers": [_vp, _int, _int, ctypes.c_uint32, _vp, _vp, _vp],
    "sx_reinhard_workspace_bytes": [],
    "sx_reinhard_transform": [_vp, _int, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _i64, _vp],
    "sx_reinhard_fit": [_vp, _int, _i64, _i64, _i64, _vp, _vp, _vp, _i64, _vp],
    # macenko
    "sx_macenko_workspace_bytes": [_i64],
    "sx_macenko_region": [_i64, _int, ctypes.POINTER(_i64), ctypes.POINTER(_i64)],
    "sx_macenko_begin": [_vp, _i64, _vp],
    "sx_macenko_peer_buffer_bytes": [],
    "sx_macenko_peer_scratch_bytes": [],
    "sx_macenko_peer_combine": [_vp, _int, _int, ctypes.c_uint32, _int, _vp, _vp],
    "sx_macenko_fit_peers": [_vp, _int, _i64, _i64, _i64, _vp, _vp, _int, _int, ctypes.c_uint32, _int, _vp, _vp, _vp, _vp],
    "sx_macenko_moments": [_vp, _int, _i64, _i64, _i64, _int, _i64, _vp, _i64, _vp],
    "sx_macenko_basis": [_vp, _i64, _i64, _i64, _int, _vp],
    "sx_macenko_moments_fallback": [_vp, _int, _i64, _i64, _i64, _i64, _vp, _i64, _vp],
    "sx_macenko_hist": [_vp, _int, _i64, _i64, _i64, _int, _i64, _int, _int, _vp, _i64, _vp],
    "sx_macenko_select": [_vp, _i64, _i64, _i64, _int, _int, _vp],
    "sx_macenko_apply": [_vp, _int, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _int, _f32, _vp, _i64, _vp],
    "sx_macenko_transform": [_vp, _int, _i64, _i64, _i64, _vp, _vp, _vp, _int, _f32, _vp, _i64, _vp],
    "sx_macenko_fit": [_vp, _int, _i64, _i64, _i64, _vp, _vp, _vp, _i64, _vp],
    "sx_macenko_fit_transform": [_vp, _int, _i64, _i64, _i64, _vp, _vp, _vp, _int, ctypes.c_float, _vp, _i64, _vp],
    "sx_macenko_fit_transform_peers": [_vp, _int, _i64, _i64, _i64, _vp, _vp, _int, _int, ctypes.c_uint32, _int, _vp, _vp, _vp, _vp, _int, ctypes.c_float, _vp, _i64, _vp],
    # tuning hooks (bench / profiling only)
    "sx_hm_set_tuning": [_int, _int, _int],
    "sx_reinhard_set_tuning": [_int],
    "sx_macenko_set_tuning": [_int, _i64],
    "sx_macenko_trace": [_int, ctypes.c_char_p, _i64],
}
_RESTYPES = {
    "sx_last_error": ctypes.c_char_p,
    "sx_kernel_launches": _i64,
    "sx_hm_workspace_bytes": _i64,
    "sx_hm_peer_buffer_bytes": _i64,
    "sx_reinhard_workspace_bytes": _i64,
    "sx_reinhard_peer_buffer_bytes": _i64,
    "sx_macenko_workspace_bytes": _i64,
    "sx_macenko_peer_buffer_bytes": _i64,
    "sx_macenko_peer_scratch_bytes": _i64,
}

ABI_VERSION = 2  # include/stainx_b200.h: SX_ABI_VERSION
SX_U8, SX_F32, SX_F16, SX_BF16 = 0, 1, 2, 3
SX_NCHW, SX_NHWC = 0, 1
SX_STAGE_ANGLE, SX_STAGE_CONC = 0, 1
REGIONS = {"moments": 0, "odrange": 1, "hist1": 2, "hist2": 3, "vmin": 4, "vmax": 5, "fit": 6, "counters": 7, "status": 8}

_lib: ctypes.CDLL | None = None
_load_error: str | None = None


def _load() -> ctypes.CDLL:
    lib = ctypes.CDLL(str(LIB_PATH))
    for name, argtypes in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the .so is stale
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, _int)
    if lib.sx_abi_version() != ABI_VERSION:
        raise ImportError(f"libstainx_b200 ABI {lib.sx_abi_version()} != {ABI_VERSION}; rebuild with `python -m stainx_b200.build`")
    return lib


def lib() -> ctypes.CDLL:
    """The loaded library; raises ``ImportError`` when it has not been built."""
    global _lib, _load_error
    if _lib is None:
        try:
            _lib = _load()
            _load_error = None
        except (OSError, AttributeError) as exc:
            _load_error = f"{type(exc).__name__}: {exc}"
            raise ImportError(f"libstainx_b200.so is not available at {LIB_PATH} ({_load_error}). Build it with `python -m stainx_b200.build`; there is no fallback path.") from exc
    return _lib


def available() -> bool:
    try:
        lib()
        return True
    except ImportError:
        return False


FUNCTIONS_AVAILABLE = available()


class StainxNativeError(RuntimeError):
    """Raised when a native call reports a non-zero status (reference: TORCH_CHECK -> RuntimeError)."""


def check(status: int, what: str) -> None:
    if status != 0:
        msg = lib().sx_last_error()
        raise StainxNativeError(f"{what} failed (status {status}): {msg.decode() if msg else 'unknown error'}")


def peer_status() -> tuple[bool, int, int]:
    """(timed_out, rank waited for, epoch) of the current device's peer-exchange status record."""
    t, r, e = _int(0), _int(0), ctypes.c_uint32(0)
    check(lib().sx_peer_status(ctypes.byref(t), ctypes.byref(r), ctypes.byref(e)), "sx_peer_status")
    return bool(t.value), int(r.value), int(e.value)


def kernel_launches() -> int:
    return int(lib().sx_kernel_launches())
