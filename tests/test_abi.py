"""CPU: the C-ABI library loads and exports every symbol include/stainx_b200.h declares, and the
host-only entry points behave (no compute calls without a GPU)."""
from __future__ import annotations

import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
HEADER = ROOT / "include" / "stainx_b200.h"


def declared_functions() -> list[str]:
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(sx_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    names = declared_functions()
    for must in ("sx_hm_hist", "sx_hm_build_lut", "sx_hm_apply", "sx_reinhard_stats", "sx_reinhard_apply", "sx_macenko_moments", "sx_macenko_hist", "sx_macenko_select", "sx_macenko_apply", "sx_last_error"):
        assert must in names
    assert len(names) >= 29


def test_library_exports_every_declared_symbol():
    from stainx_b200 import _native

    lib = ctypes.CDLL(str(_native.LIB_PATH))
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, f"declared in the header but not exported: {missing}"


def test_bindings_cover_the_header():
    from stainx_b200 import _native

    assert set(declared_functions()) <= set(_native.PROTOTYPES), set(declared_functions()) - set(_native.PROTOTYPES)
    assert _native.FUNCTIONS_AVAILABLE
    assert _native.lib().sx_abi_version() == _native.ABI_VERSION == 2


def test_host_only_entry_points():
    from stainx_b200 import _native

    lib = _native.lib()
    assert lib.sx_hm_workspace_bytes() >= 768 * 16
    assert lib.sx_reinhard_workspace_bytes() >= 64
    one, many = lib.sx_macenko_workspace_bytes(1), lib.sx_macenko_workspace_bytes(64)
    assert 0 < one < many and many % 256 == 0
    off, size = ctypes.c_int64(), ctypes.c_int64()
    seen = []
    for name, rid in _native.REGIONS.items():
        assert lib.sx_macenko_region(64, rid, ctypes.byref(off), ctypes.byref(size)) == 0, name
        assert 0 <= off.value and off.value + size.value <= many
        seen.append((off.value, off.value + size.value))
    seen.sort()
    assert all(a[1] <= b[0] for a, b in zip(seen, seen[1:])), "regions overlap"
    # errors: status code + message, no exception at the C level
    assert lib.sx_macenko_region(64, 99, ctypes.byref(off), ctypes.byref(size)) != 0
    assert b"unknown region" in lib.sx_last_error()
    assert lib.sx_hm_hist(None, 0, 0, 1, 1, 1, None, None) != 0
    assert b"NULL" in lib.sx_last_error()
    assert lib.sx_kernel_launches() >= 0


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    """No fallback: with the .so absent the product path raises ImportError."""
    from stainx_b200 import _native

    monkeypatch.setattr(_native, "LIB_PATH", tmp_path / "nope.so")
    monkeypatch.setattr(_native, "_lib", None)
    with pytest.raises(ImportError, match="no fallback"):
        _native.lib()
    from stainx_b200.backends.torch_cuda_backend import ReinhardCUDA

    with pytest.raises(ImportError):
        ReinhardCUDA("cuda")


def test_product_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may touch oracle/."""
    pkg = ROOT / "stainx_b200"
    offenders = [str(p) for p in pkg.rglob("*") if p.suffix in {".py", ".cu", ".cuh", ".h"} and re.search(r"(from|import)\s+oracle|oracle[/.]|stainx_oracle|\box_[a-z]", p.read_text())]
    assert not offenders, offenders
