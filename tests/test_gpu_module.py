"""GPU tests of the module layer: ``StainNormalizerTransform`` on CUDA (row T1 of SURVEY.md section 8a).

Mirrors the reference's ``tests/torch_interface/test_stain_normalizer_transform.py`` (L62-84 device
following, L149-157 fit on a CPU reference / forward on CUDA, L159-173 ``normalize_to_0_1`` on the
torch_cuda backend, L175-179 batch-mode refit) and adds what the reference cannot assert offline:
every result is compared with the CPU oracle.  BASELINE config 5's shape class (uint8 2048x2048 ->
float32 [0, 1]) is checked on one noise tile (both-signs protocol) and one Beer-Lambert tile.
"""
from __future__ import annotations

import numpy as np
import pytest
import torch

from tests.helpers import best_sign_diff, he_batch, he_tile, noise_f32, noise_u8

pytestmark = pytest.mark.gpu


def _np(t: torch.Tensor) -> np.ndarray:
    return t.detach().cpu().numpy()


@pytest.fixture
def reference():
    return noise_u8((1, 3, 64, 64), 0)


def test_default_device_follows_cuda_input(cuda, ox, reference):
    from stainx_b200 import StainNormalizerTransform

    src = noise_u8((2, 3, 64, 64), 8)
    t = StainNormalizerTransform(method="reinhard", mode="reference", reference=reference.to(cuda))
    out = t(src.to(cuda))
    assert out.device.type == "cuda" and out.shape == src.shape and out.dtype == torch.uint8
    assert torch.device(t.normalizer.device).type == "cuda"
    mean, std = ox.reinhard_fit(reference.numpy())
    want = ox.reinhard_transform(src.numpy(), mean, std)
    d = np.abs(_np(out).astype(np.int32) - want.astype(np.int32))
    assert d.max() <= 1 and (d > 0).mean() < 0.01


def test_explicit_backend_and_device_none(cuda, reference):
    from stainx_b200 import StainNormalizerTransform

    t = StainNormalizerTransform(method="reinhard", mode="reference", reference=reference.to(cuda), backend="torch_cuda")
    out = t(noise_u8((1, 3, 64, 64), 3).to(cuda))
    assert out.device.type == "cuda"
    with pytest.raises(ValueError, match="requires CUDA tensors"):
        t(noise_u8((1, 3, 64, 64), 3))  # device=None follows the batch, and there is no CPU path


def test_device_none_fit_cpu_reference_forward_cuda(cuda, ox, reference):
    """Reference L149-157: a device=None transform is fitted from a HOST reference and then follows CUDA batches."""
    from stainx_b200 import StainNormalizerTransform

    t = StainNormalizerTransform(method="reinhard", mode="reference", reference=reference)  # CPU tensor
    assert t.normalizer._is_fitted
    src = noise_u8((2, 3, 64, 64), 4)
    out = t(src.to(cuda))
    assert out.device.type == "cuda" and torch.device(t.normalizer.device).type == "cuda"
    mean, std = ox.reinhard_fit(reference.numpy())
    assert np.abs(_np(t.normalizer._reference_mean) - mean).max() <= 1e-3
    want = ox.reinhard_transform(src.numpy(), mean, std)
    assert np.abs(_np(out).astype(np.int32) - want.astype(np.int32)).max() <= 1


def test_explicit_device_moves_host_batches(cuda, reference):
    from stainx_b200 import StainNormalizerTransform

    t = StainNormalizerTransform(method="histogram_matching", mode="reference", reference=reference, device="cuda")
    src = noise_u8((2, 3, 64, 64), 5)
    out = t(src)  # host batch, explicit device: moved to the GPU
    assert out.device.type == "cuda"
    assert torch.equal(out, t(src.to(cuda)))


def test_single_image_chw_roundtrip(cuda, reference):
    from stainx_b200 import StainNormalizerTransform

    img = noise_u8((3, 64, 64), 2)
    t = StainNormalizerTransform(method="reinhard", mode="reference", reference=reference.to(cuda))
    out = t(img.to(cuda))
    assert out.shape == img.shape
    assert torch.equal(out, t(img.unsqueeze(0).to(cuda))[0])


def test_macenko_normalize_to_0_1_torch_cuda(cuda, ox):
    """Reference L159-173, plus the oracle comparison (both-signs protocol on noise)."""
    from stainx_b200 import Macenko, StainNormalizerTransform

    ref, src = noise_f32((1, 3, 64, 64), 10), noise_f32((2, 3, 64, 64), 11)
    t = StainNormalizerTransform(method="macenko", mode="reference", reference=ref.to(cuda), backend="torch_cuda", device="cuda")
    assert t.normalizer.normalize_to_0_1 is True
    out = t(src.to(cuda))
    assert out.device.type == "cuda" and out.dtype == torch.float32
    assert float(out.amax()) <= 1.0 + 1e-4 and float(out.amin()) >= -1e-5 and float(out.mean()) > 0.05
    he, maxc = _np(t.normalizer._stain_matrix), _np(t.normalizer._target_max_conc)
    cand = [ox.macenko_transform(src.numpy(), he, maxc, mid_signs=[s, s]) / 255.0 for s in (1, -1)]
    assert best_sign_diff(_np(out), cand[0], cand[1]).max() <= 1e-3
    # matches an explicitly built Macenko(normalize_to_0_1=True) (reference L133-141)
    n = Macenko(device=cuda, backend="torch_cuda", normalize_to_0_1=True).fit(ref.to(cuda))
    assert torch.allclose(out, n.transform(src.to(cuda)), rtol=0, atol=1e-4)
    # without the flag a unit-range float input stays in [0, 255] (reference L111-118)
    t255 = StainNormalizerTransform(method="macenko", mode="reference", reference=ref.to(cuda), device="cuda", normalize_to_0_1=False)
    assert float(t255(src.to(cuda)).amax()) > 1.0


def test_float_above_one_is_not_rescaled(cuda):
    """Reference L120-131: ColorJitter can push floats above 1; the dtype gate must not silently divide by 255."""
    from stainx_b200 import StainNormalizerTransform

    ref = noise_f32((1, 3, 64, 64), 9)
    src = (noise_f32((2, 3, 64, 64), 12) * 1.3).clamp(0.0, 1.5)
    t = StainNormalizerTransform(method="macenko", mode="reference", reference=ref.to(cuda), device="cuda")
    out = t(src.to(cuda))
    assert float(out.mean()) > 0.05 and float(out.amax()) <= 1.0 + 1e-4


@pytest.mark.parametrize("method", ["reinhard", "macenko", "histogram_matching"])
def test_batch_mode_refits_on_every_call(cuda, ox, method):
    """Reference L175-179 + T1: mode='batch' fits on batch[batch_ref_index] of EVERY incoming batch."""
    from stainx_b200 import StainNormalizerTransform

    t = StainNormalizerTransform(method=method, mode="batch", device="cuda", batch_ref_index=1)
    batches = [he_batch(3, 96, 96, seed0=50), he_batch(3, 96, 96, seed0=80)]
    for b in batches:
        out = _np(t(b.to(cuda)))
        assert t.normalizer._is_fitted and out.shape[0] == 3
        r = b[1:2].numpy()
        if method == "reinhard":
            mean, std = ox.reinhard_fit(r)
            want = ox.reinhard_transform(b.numpy(), mean, std)
        elif method == "histogram_matching":
            want = ox.hm_transform(b.numpy(), ox.hm_fit(r))
            assert np.array_equal(out, want)
            continue
        else:
            he, maxc = ox.macenko_fit(r)
            assert np.abs(_np(t.normalizer._stain_matrix) - he).max() <= 1e-4
            want = ox.macenko_transform(b.numpy(), _np(t.normalizer._stain_matrix), _np(t.normalizer._target_max_conc)).astype(np.float32) / 255.0
            d = np.abs(out - want)
            assert d.max() <= 1.0 / 255.0 + 1e-6 and (d > 1e-6).mean() < 0.01  # uint8 in: one grey level at the truncation knife edge
            continue
        d = np.abs(out.astype(np.int32) - want.astype(np.int32))
        assert d.max() <= 1 and (d > 0).mean() < 0.01
    with pytest.raises(IndexError):
        StainNormalizerTransform(method="reinhard", mode="batch", device="cuda", batch_ref_index=7)(batches[0].to(cuda))


def test_hm_channels_last_module(cuda, ox):
    from stainx_b200 import HistogramMatching, StainNormalizerTransform

    ref, src = noise_u8((1, 3, 32, 40), 6, 1.5), noise_u8((2, 3, 32, 40), 7, 0.7)
    t = StainNormalizerTransform(method="histogram_matching", mode="reference", reference=ref.permute(0, 2, 3, 1).contiguous().to(cuda), device="cuda", channel_axis=-1)
    out = t(src.permute(0, 2, 3, 1).contiguous().to(cuda))
    want = ox.hm_transform(src.numpy(), ox.hm_fit(ref.numpy())).transpose(0, 2, 3, 1)
    assert out.shape == (2, 32, 40, 3) and np.array_equal(_np(out), want)
    n = HistogramMatching(device=cuda, backend="torch_cuda", channel_axis=-1).fit(ref.permute(0, 2, 3, 1).contiguous().to(cuda))
    t2 = StainNormalizerTransform(mode="reference", normalizer=n, device="cuda")
    assert t2.channel_axis == -1
    assert torch.equal(t2(src.permute(0, 2, 3, 1).contiguous().to(cuda)), out)


def test_c5_uint8_2048_tiles_to_float32(cuda, ox):
    """BASELINE config 5's shape class: StainNormalizerTransform('macenko') on uint8 2048x2048 tiles -> float32
    in [0, 1].  One noise tile (the bench distribution; both-signs protocol) and one Beer-Lambert tile against
    the oracle: the reference truncates to a grey level and divides by 255, so the bar is one grey level on
    < 1 % of the pixels and exact k/255 values everywhere."""
    from stainx_b200 import StainNormalizerTransform

    ref = noise_u8((1, 3, 2048, 2048), 42)
    tiles = torch.cat([noise_u8((1, 3, 2048, 2048), 43), he_tile(2048, 2048, 77, 1.05)])
    t = StainNormalizerTransform(method="macenko", mode="reference", reference=ref.to(cuda), device="cuda", backend="torch_cuda")
    out = t(tiles.to(cuda))
    assert out.dtype == torch.float32 and tuple(out.shape) == (2, 3, 2048, 2048)
    levels = out * 255.0
    assert float((levels - levels.round()).abs().max()) <= 2e-5  # k / 255 exactly
    he, maxc = _np(t.normalizer._stain_matrix), _np(t.normalizer._target_max_conc)
    o = _np(levels.round())
    cand = [ox.macenko_transform(tiles.numpy(), he, maxc, mid_signs=[s, s]).astype(np.float64) for s in (1, -1)]
    for i in range(2):  # per image: the better of the two middle-eigenvector signs
        d = min((np.abs(o[i] - c[i]) for c in cand), key=lambda x: x.max() * 1e6 + (x > 0).mean())
        assert d.max() <= 1 and (d > 0).mean() < 0.01


def test_host_stream_orders_behind_fit_without_host_sync(cuda, ox):
    """HostStream runs the transform on its own non-blocking stream; a fit() enqueued on the caller's
    stream right before submit() must be complete before the transform reads the fitted tensors.  The fit is
    held back behind a long spin kernel so that an unordered transform would certainly overtake it."""
    from stainx_b200 import HistogramMatching
    from stainx_b200.ingest import HostStream

    ref, src = noise_u8((1, 3, 128, 128), 21, 2.0), noise_u8((2, 3, 128, 128), 22, 0.5).pin_memory()
    ref_dev = ref.to(cuda)
    hm = HistogramMatching(device=cuda, backend="torch_cuda")
    pipe = HostStream(hm, device=cuda, depth=2)
    hm._is_fitted = True  # HostStream binds transform at construction; fit happens below, on the caller's stream
    torch.cuda.synchronize()
    torch.cuda._sleep(200_000_000)  # ~0.1 s of GPU time in front of the fit kernels
    hm.fit(ref_dev)
    ticket = pipe.submit(src)  # no host synchronisation between fit and submit
    out = ticket.wait()
    assert np.array_equal(out.numpy(), ox.hm_transform(src.numpy(), ox.hm_fit(ref.numpy())))


def test_device_stream_pool_matches_direct_transform(cuda):
    """ingest.DeviceStream issues independent device-resident batches round-robin on a pool of streams (their phases
    overlap); every result must equal the plain transform of the same batch, for all three methods."""
    from stainx_b200 import HistogramMatching, Macenko, Reinhard
    from stainx_b200.ingest import DeviceStream

    g = torch.Generator(device=cuda).manual_seed(17)
    ref = (torch.rand((1, 3, 256, 256), device=cuda, generator=g).pow(1.5) * 255).round().to(torch.uint8)
    batches = [(torch.rand((6, 3, 512, 512), device=cuda, generator=g).pow(p) * 255).round().to(torch.uint8) for p in (0.6, 1.0, 1.7, 0.8, 1.3)]
    for cls in (HistogramMatching, Reinhard, Macenko):
        n = cls(device=cuda, backend="torch_cuda").fit(ref)
        want = [n.transform(b) for b in batches]
        torch.cuda.synchronize()
        pool = DeviceStream(n, streams=3)
        for _ in range(3):  # repeated use: recycled outputs / workspaces across the streams
            got = pool.map(batches)
            checks = [(a != b).sum() for a, b in zip(got, want)]
            assert all(int(c) == 0 for c in checks), cls.__name__
