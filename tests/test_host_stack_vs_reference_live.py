"""CPU, build container only: the HOST stack of this package against the reference's own API, live.

The normalizer front-ends and backend classes of ``stainx_b200`` (dtype gate, cast-back, ``normalize_to_0_1``,
``channel_axis``, ``fit_transform``) are ordinary Python around the kernel layer.  Here the kernel layer is the CPU
stand-in of ``tests/cpu_ops.py`` (numpy + the oracle), so the very same front-end / backend code that drives the GPU runs
on this box, and every call is mirrored on the reference (``/root/reference/src``, ``backend="torch"``,
``device="cpu"``): same output dtype, same shape, same values within the parity bars, same fitted attributes.
Skipped where the reference is absent (the GPU box).
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import pytest
import torch

from tests.helpers import he_batch, he_tile, noise_f32, noise_u8

REFERENCE_SRC = Path("/root/reference/src")
if not (REFERENCE_SRC / "stainx" / "__init__.py").exists():
    pytest.skip("the reference is not present on this machine", allow_module_level=True)
if str(REFERENCE_SRC) not in sys.path:
    sys.path.append(str(REFERENCE_SRC))
try:
    import stainx as ref_pkg
except Exception as exc:  # noqa: BLE001
    pytest.skip(f"the reference does not import here: {exc}", allow_module_level=True)

import stainx_b200 as our_pkg  # noqa: E402
from stainx_b200.backends import torch_cuda_backend as backends  # noqa: E402
from tests import cpu_ops  # noqa: E402


def _ours(name: str, **kwargs):
    """A product normalizer whose backend instance runs the CPU stand-in kernel layer."""
    n = getattr(our_pkg, name)(device="cpu", **kwargs)
    cls = {"Reinhard": backends.ReinhardCUDA, "Macenko": backends.MacenkoCUDA, "HistogramMatching": backends.HistogramMatchingCUDA}[name]
    n._backend_impl = cpu_ops.cpu_backend(cls, "cpu", **n._get_backend_kwargs())
    return n


def _theirs(name: str, **kwargs):
    return getattr(ref_pkg, name)(device="cpu", backend="torch", **kwargs)


def _cast(x: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    if dtype == torch.uint8:
        return x if x.dtype == torch.uint8 else (x * 255).round().to(torch.uint8)
    return (x.float() / 255.0 if x.dtype == torch.uint8 else x).to(dtype)


def _close(got: torch.Tensor, want: torch.Tensor, float_bar: float, grey_levels_over_255: bool = False) -> None:
    assert got.dtype == want.dtype, (got.dtype, want.dtype)
    assert got.shape == want.shape
    d = (got.double() - want.double()).abs()
    if grey_levels_over_255:  # uint8 input with normalize_to_0_1: float32 values k / 255 of a TRUNCATED uint8 result
        d = (d * 255).round()
    if want.dtype == torch.uint8 or grey_levels_over_255:
        assert float(d.max()) <= 1 and float((d > 0).double().mean()) < 1e-2  # truncation knife edge: the bar of the GPU parity tests
    else:
        assert float(d.max()) <= float_bar, float(d.max())


@pytest.mark.parametrize("dtype", [torch.uint8, torch.float32, torch.float64])
@pytest.mark.parametrize("channel_axis", [1, -1])
def test_histogram_matching_front_end(dtype, channel_axis):
    ref, src = _cast(noise_u8((1, 3, 33, 47), 1, 2.0), dtype), _cast(noise_u8((3, 3, 29, 31), 2, 0.6), dtype)
    if channel_axis == -1:
        ref, src = ref.permute(0, 2, 3, 1).contiguous(), src.permute(0, 2, 3, 1).contiguous()
    a, b = _ours("HistogramMatching", channel_axis=channel_axis).fit(ref), _theirs("HistogramMatching", channel_axis=channel_axis).fit(ref)
    assert torch.equal(torch.stack(a._ref_histograms_256), torch.stack(b._ref_histograms_256))
    got, want = a.transform(src), b.transform(src)
    assert got.dtype == want.dtype and got.shape == want.shape
    assert float((got.double() - want.double()).abs().max()) <= (0 if dtype != torch.float64 else 1e-7)  # bit-exact; float64: the cast-back of the same float32
    got, want = _ours("HistogramMatching", channel_axis=channel_axis).fit_transform(src), _theirs("HistogramMatching", channel_axis=channel_axis).fit_transform(src)
    assert got.dtype == want.dtype and float((got.double() - want.double()).abs().max()) <= (0 if dtype != torch.float64 else 1e-7)
    # the reference keeps (values, cdf) per channel; they are derived lazily here
    assert torch.allclose(a._reference_histogram.float(), b._reference_histogram.float(), atol=1e-6)


@pytest.mark.parametrize("dtype", [torch.uint8, torch.float32, torch.float64])
def test_reinhard_front_end(dtype):
    ref, src = _cast(noise_f32((2, 3, 31, 45), 3, 1.4), dtype), _cast(noise_f32((3, 3, 27, 33), 4, 0.7), dtype)
    a, b = _ours("Reinhard").fit(ref), _theirs("Reinhard").fit(ref)
    assert torch.allclose(a._reference_mean.flatten(), b._reference_mean.flatten(), atol=1e-4)
    assert torch.allclose(a._reference_std.flatten(), b._reference_std.flatten(), atol=1e-4)
    _close(a.transform(src), b.transform(src), 1e-4)
    _close(_ours("Reinhard").fit_transform(src), _theirs("Reinhard").fit_transform(src), 1e-4)


@pytest.mark.parametrize("dtype", [torch.uint8, torch.float32, torch.float64])
@pytest.mark.parametrize("unit", [False, True])
def test_macenko_front_end(dtype, unit):
    ref, src = _cast(he_tile(72, 88, 42), dtype), _cast(he_batch(3, 64, 80), dtype)
    a, b = _ours("Macenko", normalize_to_0_1=unit).fit(ref), _theirs("Macenko", normalize_to_0_1=unit).fit(ref)
    assert float((a._stain_matrix - b._stain_matrix).abs().max()) <= 1e-4
    assert float((a._target_max_conc.flatten() / b._target_max_conc.flatten() - 1).abs().max()) <= 1e-3
    levels = unit and dtype == torch.uint8
    got, want = a.transform(src), b.transform(src)
    _close(got, want, 1e-3 if unit else 1e-3 * 255, levels)
    got, want = _ours("Macenko", normalize_to_0_1=unit).fit_transform(src), _theirs("Macenko", normalize_to_0_1=unit).fit_transform(src)
    _close(got, want, 1e-3 if unit else 1e-3 * 255, levels)


def test_errors_match_the_reference():
    x = noise_u8((1, 3, 16, 16), 5)
    for name in ("Reinhard", "Macenko", "HistogramMatching"):
        for make in (_ours, _theirs):
            with pytest.raises(ValueError, match=r"Must call fit\(\) before transform\(\)"):
                make(name).transform(x)
    for pkg in (our_pkg, ref_pkg):
        with pytest.raises(ValueError, match="precision must be"):
            pkg.Macenko(device="cpu", precision="ultra")
        with pytest.raises(ValueError, match="precision='fast' requires backend='torch_cuda'"):
            pkg.Macenko(device="cpu", backend="torch", precision="fast")
    assert set(our_pkg.__all__) == set(ref_pkg.__all__)
    assert np.all([hasattr(our_pkg, n) for n in ref_pkg.__all__])


# ---------------------------------------------------------------- StainNormalizerTransform (transforms.py)
@pytest.fixture
def cpu_module_stack(monkeypatch):
    """The nn.Module layer on this box: every normalizer it builds gets the CPU stand-in kernel layer, and the
    'batches must be CUDA tensors' rule of the product (there is no CPU compute path) is lifted for the test."""

    def stand_in(cls):
        class _CPU(cls):  # noqa: N801
            @staticmethod
            def _kernel_layer():
                return cpu_ops

            def _check_device(self) -> None:
                pass

        return _CPU

    for norm, cls in ((our_pkg.Reinhard, backends.ReinhardCUDA), (our_pkg.Macenko, backends.MacenkoCUDA), (our_pkg.HistogramMatching, backends.HistogramMatchingCUDA)):
        sub = stand_in(cls)
        monkeypatch.setattr(norm, "_get_torch_cuda_class", lambda self, sub=sub: sub)
    monkeypatch.setattr(our_pkg.StainNormalizerTransform, "_follow_device", lambda self, device: None)
    return our_pkg.StainNormalizerTransform


def _module_inputs(method: str, dtype: torch.dtype):
    if method == "macenko":
        return _cast(he_tile(56, 72, 42), dtype), _cast(he_batch(3, 56, 72), dtype)
    return _cast(noise_u8((1, 3, 40, 52), 11, 1.7), dtype), _cast(noise_u8((3, 3, 40, 52), 12, 0.8), dtype)


@pytest.mark.parametrize("method", ["macenko", "reinhard", "histogram_matching"])
@pytest.mark.parametrize("dtype", [torch.uint8, torch.float32])
def test_module_reference_and_batch_mode(cpu_module_stack, method, dtype):
    ref, batch = _module_inputs(method, dtype)
    ours = cpu_module_stack(method, mode="reference", reference=ref)
    theirs = ref_pkg.StainNormalizerTransform(method, mode="reference", reference=ref, device="cpu", backend="torch")
    levels = method == "macenko" and dtype == torch.uint8  # Macenko built by the module defaults to normalize_to_0_1=True
    bar = 1e-3 if method == "macenko" else 1e-4
    _close(ours(batch), theirs(batch), bar, levels)
    _close(ours(batch[1]), theirs(batch[1]), bar, levels)  # CHW in, CHW out
    assert ours(batch[1]).dim() == 3
    assert len(ours.state_dict()) == len(theirs.state_dict()) == 0  # fitted parameters are plain attributes
    # batch mode: re-fit on image `batch_ref_index` of every call
    ours = cpu_module_stack(method, mode="batch", batch_ref_index=2)
    theirs = ref_pkg.StainNormalizerTransform(method, mode="batch", batch_ref_index=2, device="cpu", backend="torch")
    _close(ours(batch), theirs(batch), bar, levels)
    with pytest.raises(IndexError):
        cpu_module_stack(method, mode="batch", batch_ref_index=7)(batch)


def test_module_channels_last_and_prebuilt_normalizer(cpu_module_stack):
    ref, batch = _module_inputs("histogram_matching", torch.uint8)
    ref_l, batch_l = ref.permute(0, 2, 3, 1).contiguous(), batch.permute(0, 2, 3, 1).contiguous()
    ours = cpu_module_stack("histogram_matching", reference=ref_l, channel_axis=-1)
    theirs = ref_pkg.StainNormalizerTransform("histogram_matching", reference=ref_l, channel_axis=-1, device="cpu", backend="torch")
    assert torch.equal(ours(batch_l), theirs(batch_l))
    # a pre-fitted normalizer handed to the module; normalize_to_0_1 forwarded to Macenko
    mref, mbatch = _module_inputs("macenko", torch.float32)
    ours = cpu_module_stack(normalizer=_ours("Macenko").fit(mref), normalize_to_0_1=False)
    theirs = ref_pkg.StainNormalizerTransform(normalizer=_theirs("Macenko").fit(mref), normalize_to_0_1=False, device="cpu", backend="torch")
    _close(ours(mbatch), theirs(mbatch), 1e-3 * 255)
    for make in (cpu_module_stack, lambda *a, **k: ref_pkg.StainNormalizerTransform(*a, device="cpu", backend="torch", **k)):
        with pytest.raises(ValueError, match="only applies to Macenko"):
            make("reinhard", reference=ref, normalize_to_0_1=True)
        with pytest.raises(ValueError, match="only supported for histogram_matching"):
            make("reinhard", reference=ref, channel_axis=-1)
        with pytest.raises(ValueError, match="Expected NCHW"):
            make("reinhard", reference=ref)(batch_l)
        with pytest.raises(ValueError):
            make("macenko", mode="reference")  # no reference, nothing fitted
