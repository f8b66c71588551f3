"""float16 / bfloat16 image tensors read and written by the kernels themselves.

The reference widens such tensors to float32, computes, and casts the result back to the input dtype
(torch_backend.py:L103-131).  The kernels here load the 16-bit values directly (8 pixels per 128-bit vector),
compute in float32 registers and store with round-to-nearest-even.  Bars, per dtype:

* histogram matching: the float32 result is bit-exact, so the 16-bit output must EQUAL the reference's;
* Reinhard / Macenko: |out - reference| <= (float32 bar) + one unit in the last place of the output dtype
  (the float32 values differ by up to the float32 bar, so the rounding can land one step apart).

References: ``tests/golden/half_io.npz`` (the reference's own outputs, ``make_golden_half.py``) and the CPU oracle
on the widened input, cast back with torch.
"""
from __future__ import annotations

import numpy as np
import pytest
import torch

from tests.conftest import golden
from tests.helpers import he_batch, he_tile, noise_f32

pytestmark = pytest.mark.gpu

DTYPES = {"f16": torch.float16, "bf16": torch.bfloat16}
# one unit in the last place at the top of the output range: [0.5, 1) for unit outputs, [128, 256) for [0, 255] outputs
ULP_UNIT = {"f16": 2.0**-11, "bf16": 2.0**-8}
ULP_255 = {"f16": 2.0**-3, "bf16": 1.0}


def _from_bits(a: np.ndarray, dt: torch.dtype) -> torch.Tensor:
    return torch.from_numpy(a.copy()).view(dt)


@pytest.mark.parametrize("name", ["f16", "bf16"])
def test_half_goldens_all_methods(cuda, name):
    from stainx_b200 import HistogramMatching, Macenko, Reinhard

    g, dt = golden("half_io"), DTYPES[name]
    ref, src = _from_bits(g[f"hm_{name}_ref"], dt).to(cuda), _from_bits(g[f"hm_{name}_src"], dt).to(cuda)
    hm = HistogramMatching(device=cuda, backend="torch_cuda").fit(ref)
    assert np.array_equal(torch.stack(hm._ref_histograms_256).cpu().numpy(), g[f"hm_{name}_ref_hist"])
    out = hm.transform(src)
    assert out.dtype == dt
    assert torch.equal(out.cpu(), _from_bits(g[f"hm_{name}_out"], dt))

    ref, src = _from_bits(g[f"rh_{name}_ref"], dt).to(cuda), _from_bits(g[f"rh_{name}_src"], dt).to(cuda)
    rh = Reinhard(device=cuda, backend="torch_cuda").fit(ref)
    assert np.abs(rh._reference_mean.cpu().numpy() - g[f"rh_{name}_mean"]).max() <= 1e-3
    rh._reference_mean, rh._reference_std = torch.from_numpy(g[f"rh_{name}_mean"]).to(cuda), torch.from_numpy(g[f"rh_{name}_std"]).to(cuda)
    out = rh.transform(src)
    assert out.dtype == dt
    assert float((out.float().cpu() - _from_bits(g[f"rh_{name}_out"], dt).float()).abs().max()) <= 1e-3 + ULP_UNIT[name]

    mk = Macenko(device=cuda, backend="torch_cuda").fit(ref)
    assert np.abs(mk._stain_matrix.cpu().numpy() - g[f"mk_{name}_he"]).max() <= 1e-4
    assert np.abs(mk._target_max_conc.cpu().numpy() / g[f"mk_{name}_maxc"] - 1).max() <= 1e-3
    mk._stain_matrix, mk._target_max_conc = torch.from_numpy(g[f"mk_{name}_he"]).to(cuda), torch.from_numpy(g[f"mk_{name}_maxc"]).to(cuda)
    out = mk.transform(src)
    assert out.dtype == dt
    assert float((out.float().cpu() - _from_bits(g[f"mk_{name}_out"], dt).float()).abs().max()) <= 1e-3 * 255.0 + ULP_255[name]
    mk.normalize_to_0_1 = True
    out01 = mk.transform(src)
    assert out01.dtype == dt
    assert float((out01.float().cpu() - _from_bits(g[f"mk_{name}_out01"], dt).float()).abs().max()) <= 1e-3 + ULP_UNIT[name] * (1 + 1)  # two roundings


@pytest.mark.parametrize("name", ["f16", "bf16"])
@pytest.mark.parametrize("shape", [(1, 3, 1, 3), (2, 3, 33, 35), (3, 3, 256, 256), (1, 3, 321, 199), (4, 3, 512, 512)])
def test_half_hm_vs_oracle_equal(cuda, ox, name, shape):
    """Vector path (H*W a multiple of 8, aligned), scalar heads / tails, NCHW and NHWC: outputs EQUAL the oracle's
    float32 result cast to the dtype."""
    from stainx_b200 import HistogramMatching

    dt = DTYPES[name]
    ref, src = noise_f32((1, 3, shape[2], shape[3]), 42, 1.7).to(dt), noise_f32(shape, 43, 0.6).to(dt)
    ref_hist = ox.hm_fit(ref.float().numpy())
    want = torch.from_numpy(ox.hm_transform(src.float().numpy(), ref_hist)).to(dt)
    n = HistogramMatching(device=cuda, backend="torch_cuda").fit(ref.to(cuda))
    assert np.array_equal(torch.stack(n._ref_histograms_256).cpu().numpy(), ref_hist)
    out = n.transform(src.to(cuda))
    assert out.dtype == dt and torch.equal(out.cpu(), want)
    nl = HistogramMatching(device=cuda, backend="torch_cuda", channel_axis=-1).fit(ref.permute(0, 2, 3, 1).contiguous().to(cuda))
    out = nl.transform(src.permute(0, 2, 3, 1).contiguous().to(cuda))
    assert torch.equal(out.cpu().permute(0, 3, 1, 2), want)


@pytest.mark.parametrize("name", ["f16", "bf16"])
@pytest.mark.parametrize("shape", [(2, 3, 9, 7), (3, 3, 128, 256), (1, 3, 321, 199), (2, 3, 1024, 1024)])
def test_half_reinhard_and_macenko_vs_oracle(cuda, ox, name, shape):
    from stainx_b200 import Macenko, Reinhard

    dt = DTYPES[name]
    h, w = shape[2], shape[3]
    ref = (he_tile(max(h, 16), max(w, 16), 42)[:, :, :h, :w].float() / 255.0).to(dt)
    src = (he_batch(shape[0], max(h, 16), max(w, 16))[:, :, :h, :w].float() / 255.0).to(dt)
    rh = Reinhard(device=cuda, backend="torch_cuda").fit(ref.to(cuda))
    mean, std = ox.reinhard_fit(ref.float().numpy())
    assert np.abs(rh._reference_mean.cpu().numpy() - mean).max() <= 1e-3 and np.abs(rh._reference_std.cpu().numpy() - std).max() <= 1e-3
    want = ox.reinhard_transform(src.float().numpy(), rh._reference_mean.cpu().numpy(), rh._reference_std.cpu().numpy())
    out = rh.transform(src.to(cuda))
    assert out.dtype == dt
    assert float((out.float().cpu() - torch.from_numpy(want).to(dt).float()).abs().max()) <= 1e-3 + ULP_UNIT[name]
    if h * w < 64:
        return  # Macenko needs a stain plane
    mk = Macenko(device=cuda, backend="torch_cuda").fit(ref.to(cuda))
    he, maxc = ox.macenko_fit(ref.float().numpy())
    assert np.abs(mk._stain_matrix.cpu().numpy() - he).max() <= 1e-4
    want = torch.from_numpy(ox.macenko_transform(src.float().numpy(), mk._stain_matrix.cpu().numpy(), mk._target_max_conc.cpu().numpy())).to(dt)
    out = mk.transform(src.to(cuda))
    assert out.dtype == dt
    assert float((out.float().cpu() - want.float()).abs().max()) <= 1e-3 * 255.0 + ULP_255[name]
    mk.normalize_to_0_1 = True
    out01 = mk.transform(src.to(cuda))
    assert float((out01.float().cpu() - (want / 255.0).float()).abs().max()) <= 1e-3 + 2 * ULP_UNIT[name]


def test_half_matches_widened_float32_path(cuda):
    """The native 16-bit path == widening with torch, running the float32 kernels, casting back (what round 1 did)."""
    from stainx_b200 import ops

    g = torch.Generator(device=cuda).manual_seed(3)
    x32 = torch.rand((3, 3, 256, 384), device=cuda, generator=g)
    for dt in (torch.float16, torch.bfloat16):
        x = x32.to(dt)
        ref_hist = ops.hm_fit(x[:1].contiguous())
        assert torch.equal(ref_hist, ops.hm_fit(x[:1].float().contiguous()))
        assert torch.equal(ops.hm_transform(x, ref_hist), ops.hm_transform(x.float(), ref_hist).to(dt))
        mean, std = ops.reinhard_fit(x[:1].contiguous())
        m32, s32 = ops.reinhard_fit(x[:1].float().contiguous())
        assert float((mean - m32).abs().max()) <= 1e-5 and float((std - s32).abs().max()) <= 1e-5
        assert torch.equal(ops.reinhard_transform(x, m32, s32), ops.reinhard_transform(x.float(), m32, s32).to(dt))
        he, maxc = ops.macenko_fit(x[:1].float().contiguous())
        a, b = ops.macenko_transform(x, he, maxc), ops.macenko_transform(x.float(), he, maxc).to(dt)
        assert float((a.float() - b.float()).abs().max()) <= (0.125 if dt == torch.float16 else 1.0)  # one step: moment sums are accumulated in a different order
