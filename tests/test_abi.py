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


def _integration_stub_source() -> str:
    """The ``stainx_cuda_torch/__init__.py`` replacement printed in INTEGRATION.md section 2, pointed at the in-tree library."""
    from stainx_b200 import _native

    text = (ROOT / "INTEGRATION.md").read_text()
    m = re.search(r"```python\n(# src/stainx_cuda_torch/__init__\.py\n.*?)```", text, flags=re.S)
    assert m, "INTEGRATION.md no longer holds the stub"
    src = m.group(1)
    assert 'ctypes.CDLL("libstainx_b200.so")' in src
    return src.replace('ctypes.CDLL("libstainx_b200.so")', f'ctypes.CDLL({str(_native.LIB_PATH)!r})')


def test_integration_stub_is_valid_and_matches_the_bindings(tmp_path):
    """The stub a reference maintainer would add must at least import against the built library, expose
    the reference's four names, and declare the same argument lists as this package's own bindings."""
    import importlib.util

    import torch

    from stainx_b200 import _native

    pkg = tmp_path / "stainx_cuda_torch"
    pkg.mkdir()
    (pkg / "__init__.py").write_text(_integration_stub_source())
    spec = importlib.util.spec_from_file_location("_sx_stub_under_test", pkg / "__init__.py")
    stub = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(stub)
    assert stub.FUNCTIONS_AVAILABLE is True
    for name in ("histogram_matching", "reinhard", "macenko", "macenko_fast"):  # csrc/bindings.cpp:L27-35
        assert callable(getattr(stub, name))
    for fn in ("sx_hm_transform", "sx_reinhard_transform", "sx_macenko_transform"):
        ours = _native.PROTOTYPES[fn]
        theirs = getattr(stub._lib, fn).argtypes
        assert len(theirs) == len(ours), fn
        assert [ctypes.sizeof(a) for a in theirs] == [ctypes.sizeof(a) for a in ours], fn
    with pytest.raises(AssertionError):  # host tensors are refused before any native call (the reference: TORCH_CHECK is_cuda)
        stub.reinhard(torch.zeros(1, 3, 8, 8), torch.zeros(3), torch.ones(3))


@pytest.mark.skipif(not Path("/root/reference/src/stainx/__init__.py").exists(), reason="the reference is only present in the build container")
def test_integration_stub_turns_on_the_reference_cuda_backend(tmp_path):
    """With the stub on the path in front of the reference's own (unbuilt) extension package, the UNMODIFIED reference
    sees CUDA_AVAILABLE and builds its torch_cuda backend classes (a fresh interpreter: the reference caches the probe)."""
    import subprocess
    import sys

    pkg = tmp_path / "stainx_cuda_torch"
    pkg.mkdir()
    (pkg / "__init__.py").write_text(_integration_stub_source())
    code = (
        "import sys; sys.path[:0] = [%r, '/root/reference/src']\n"
        "import stainx_cuda_torch, stainx\n"
        "from stainx.backends import torch_cuda_backend as b\n"
        "assert stainx_cuda_torch.__file__.startswith(%r), stainx_cuda_torch.__file__\n"
        "assert b.CUDA_AVAILABLE is True\n"
        "for cls in (b.ReinhardCUDA, b.MacenkoCUDA, b.HistogramMatchingCUDA): cls('cuda')\n"
        "n = stainx.Reinhard(device='cuda', backend='torch_cuda')\n"
        "assert n.backend == 'torch_cuda'\n"
        "print('stub ok')\n"
    ) % (str(tmp_path), str(tmp_path))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "stub ok" in r.stdout, r.stderr[-2000:]
