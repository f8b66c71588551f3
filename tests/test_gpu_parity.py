"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and against golden
vectors produced by the reference itself.

Bars: histogram matching bit-exact (uint8 and float32); Reinhard / Macenko
float32 max-abs <= 1e-3 on [0, 1]; uint8 outputs within 1 grey level (truncation knife edge, the
reference's own PARITY_ATOL, tests/torch_cuda_interface/test_cuda_backend_parity_against_torch.py
L27-28); Macenko HE <= 1e-4, maxC rel <= 1e-3 on Beer-Lambert tiles; on near-isotropic noise the
oracle is evaluated with both middle-eigenvector signs (SURVEY.md section 7 H-a).
"""
from __future__ import annotations

import numpy as np
import pytest
import torch

from tests.conftest import golden
from tests.helpers import best_sign_diff, he_batch, he_tile, noise_f32, noise_u8

pytestmark = pytest.mark.gpu

F32_TOL = 1e-3  # on [0, 1] outputs (BASELINE.json north_star)


def _np(t: torch.Tensor) -> np.ndarray:
    return t.detach().cpu().numpy()


# ======================================================================== histogram matching
@pytest.mark.parametrize("name", ["hm_u8", "hm_f32", "hm_u8_uniform", "hm_u8_self", "hm_u8_sparse"])
def test_hm_golden_bit_exact(cuda, name):
    from stainx_b200 import HistogramMatching

    g = golden(name)
    n = HistogramMatching(device=cuda, backend="torch_cuda", channel_axis=1).fit(torch.from_numpy(g["ref"]).to(cuda))
    assert np.array_equal(_np(torch.stack(n._ref_histograms_256)), g["ref_hist"])
    out = n.transform(torch.from_numpy(g["src"]).to(cuda))
    assert out.dtype == torch.from_numpy(g["out"]).dtype
    assert np.array_equal(_np(out), g["out"])


def test_hm_golden_nhwc(cuda):
    from stainx_b200 import HistogramMatching

    g = golden("hm_u8_nhwc")
    n = HistogramMatching(device=cuda, backend="torch_cuda", channel_axis=-1).fit(torch.from_numpy(g["ref"]).to(cuda))
    assert np.array_equal(_np(torch.stack(n._ref_histograms_256)), g["ref_hist"])
    out = n.transform(torch.from_numpy(g["src"]).to(cuda))
    assert tuple(out.shape) == g["out"].shape
    assert np.array_equal(_np(out), g["out"])


@pytest.mark.parametrize("dtype", ["u8", "f32"])
@pytest.mark.parametrize("shape", [(1, 3, 1, 1), (2, 3, 7, 5), (3, 3, 64, 64), (1, 3, 321, 199), (2, 3, 256, 512), (5, 3, 130, 131), (1, 3, 1024, 1024)])
def test_hm_vs_oracle(cuda, ox, dtype, shape):
    from stainx_b200 import HistogramMatching

    make = noise_u8 if dtype == "u8" else noise_f32
    ref = make((1, 3, shape[2], shape[3]), 42, 1.7)
    src = make(shape, 43, 0.6)
    n = HistogramMatching(device=cuda, backend="torch_cuda").fit(ref.to(cuda))
    ref_hist = ox.hm_fit(ref.numpy())
    assert np.array_equal(_np(torch.stack(n._ref_histograms_256)), ref_hist)
    out = n.transform(src.to(cuda))
    assert np.array_equal(_np(out), ox.hm_transform(src.numpy(), ref_hist))


@pytest.mark.parametrize("dtype", ["u8", "f32"])
def test_hm_nhwc_vs_oracle(cuda, ox, dtype):
    from stainx_b200 import HistogramMatching

    make = noise_u8 if dtype == "u8" else noise_f32
    ref = make((1, 3, 50, 70), 1, 2.0)
    src = make((3, 3, 61, 47), 2, 0.8)
    n = HistogramMatching(device=cuda, backend="torch_cuda", channel_axis=-1).fit(ref.permute(0, 2, 3, 1).contiguous().to(cuda))
    out = n.transform(src.permute(0, 2, 3, 1).contiguous().to(cuda))
    want = ox.hm_transform(src.numpy(), ox.hm_fit(ref.numpy())).transpose(0, 2, 3, 1)
    assert np.array_equal(_np(out), want)


@pytest.mark.parametrize("hw", [(700, 500), (701, 499), (1024, 1024)])
def test_hm_nhwc_large_vs_planar(cuda, hw):
    """Interleaved uint8 batches above 1 MB take the coalesced vector remap (rotating channel tables) plus a
    tail of < 16 bytes: the result must equal the planar transform of the same pixels, bit for bit."""
    from stainx_b200 import HistogramMatching

    g = torch.Generator(device=cuda).manual_seed(31)
    ref = (torch.rand((1, 3, 300, 300), device=cuda, generator=g).pow(1.6) * 255).round().to(torch.uint8)
    src = (torch.rand((2, 3, *hw), device=cuda, generator=g).pow(0.7) * 255).round().to(torch.uint8)
    planar = HistogramMatching(device=cuda, backend="torch_cuda").fit(ref)
    nhwc = HistogramMatching(device=cuda, backend="torch_cuda", channel_axis=-1).fit(ref.permute(0, 2, 3, 1).contiguous())
    want = planar.transform(src)
    got = nhwc.transform(src.permute(0, 2, 3, 1).contiguous())
    assert torch.equal(got.permute(0, 3, 1, 2), want)


def test_hm_misaligned_view_and_noncontiguous(cuda, ox):
    """Planes that do not start on 16-byte boundaries, and a non-contiguous input."""
    from stainx_b200 import HistogramMatching

    ref = noise_u8((1, 3, 33, 35), 5)
    big = noise_u8((3, 3, 33, 37), 6)
    src = big[:, :, :, 1:36]  # non-contiguous view
    n = HistogramMatching(device=cuda, backend="torch_cuda").fit(ref.to(cuda))
    out = n.transform(src.to(cuda))
    assert np.array_equal(_np(out), ox.hm_transform(np.ascontiguousarray(src.numpy()), ox.hm_fit(ref.numpy())))


def test_hm_empty_batch(cuda):
    from stainx_b200 import HistogramMatching

    n = HistogramMatching(device=cuda, backend="torch_cuda").fit(noise_u8((1, 3, 16, 16), 0).to(cuda))
    out = n.transform(torch.empty((0, 3, 16, 16), dtype=torch.uint8, device=cuda))
    assert tuple(out.shape) == (0, 3, 16, 16)


def test_hm_phases_equal_fused(cuda):
    from stainx_b200 import ops

    ref = noise_u8((1, 3, 96, 96), 3, 1.5).to(cuda)
    src = noise_u8((4, 3, 96, 96), 4, 0.7).to(cuda)
    ref_hist = ops.hm_fit(ref)
    fused = ops.hm_transform(src, ref_hist)
    counts = ops.hm_hist(src)
    # two half-batches accumulate into the same counts (what a sharded run all-reduces)
    halves = ops.hm_hist(src[:2].contiguous())
    ops.hm_hist(src[2:].contiguous(), counts=halves)
    assert torch.equal(counts, halves)
    assert int(counts.sum()) == src.numel()
    lut = ops.hm_build_lut(counts, src.numel() // 3, ops.hm_ref_cdf(ref_hist))
    lut_auto = ops.hm_build_lut(counts, -1, ops.hm_ref_cdf(ref_hist))
    assert torch.equal(lut, lut_auto)
    assert torch.equal(ops.hm_apply(src, lut), fused)


def test_hm_full_size_properties(cuda):
    """BASELINE config C2 (uint8 64x3x1024x1024): size-independent checks."""
    from stainx_b200 import ops

    g = torch.Generator(device=cuda).manual_seed(43)
    src = (torch.rand((64, 3, 1024, 1024), device=cuda, generator=g) * 255).round().to(torch.uint8)
    ref = (torch.rand((1, 3, 1024, 1024), device=cuda, generator=g).pow(2.0) * 255).round().to(torch.uint8)
    counts = ops.hm_hist(src)
    for c in range(3):  # histogram == bincount, and it accounts for every pixel
        want = torch.bincount(src[:, c].reshape(-1).long(), minlength=256)
        assert torch.equal(counts[c], want)
    ref_hist = ops.hm_fit(ref)
    lut = ops.hm_build_lut(counts, 64 * 1024 * 1024, ops.hm_ref_cdf(ref_hist))
    assert bool((lut[:, 1:] >= lut[:, :-1]).all()), "LUT must be monotone"
    out = ops.hm_transform(src, ref_hist)
    lut8 = lut.to(torch.uint8)
    for c in range(3):  # remap == gather through the LUT
        assert torch.equal(out[:, c], lut8[c][src[:, c].long()])
    # matching pulls the source histogram onto the reference: CDF distance shrinks
    oc = ops.hm_hist(out).double()
    rc = ops.hm_hist(ref).double()
    sc = counts.double()
    cdf = lambda x: torch.cumsum(x / x.sum(dim=1, keepdim=True), dim=1)  # noqa: E731
    assert (cdf(oc) - cdf(rc)).abs().max() < (cdf(sc) - cdf(rc)).abs().max()
    # a 256-entry LUT cannot split a grey level: the CDFs agree up to one bin's mass
    bin_mass = (rc / rc.sum(dim=1, keepdim=True)).max() + (sc / sc.sum(dim=1, keepdim=True)).max()
    assert (cdf(oc) - cdf(rc)).abs().max() <= bin_mass + 1e-9


def test_hm_dependent_launch_chain_back_to_back_and_in_a_graph(cuda):
    """sx_hm_transform is one chain of programmatic dependent launches (zero, histogram, LUT, remap; each
    kernel resident before the one in front of it has drained).  Back-to-back transforms of different
    batches whose scratch / output buffers are recycled by the caching allocator must stay bit-exact
    (no kernel of step k+1 may overtake a reader of step k), on a side stream and inside a CUDA graph."""
    from stainx_b200 import ops

    g = torch.Generator(device=cuda).manual_seed(11)
    batches = [(torch.rand((12, 3, 1024, 1024), device=cuda, generator=g).pow(p) * 255).round().to(torch.uint8) for p in (1.0, 2.0, 0.5)]
    ref = (torch.rand((1, 3, 256, 256), device=cuda, generator=g).pow(1.5) * 255).round().to(torch.uint8)
    ref_hist = ops.hm_fit(ref)
    ref_cdf = ops.hm_ref_cdf(ref_hist)
    want = []
    for b in batches:  # phase-level calls with a device synchronisation between the phases
        counts = ops.hm_hist(b)
        torch.cuda.synchronize()
        lut = ops.hm_build_lut(counts, b.numel() // 3, ref_cdf)
        torch.cuda.synchronize()
        want.append(ops.hm_apply(b, lut))
        torch.cuda.synchronize()
        for c in range(3):
            assert torch.equal(counts[c], torch.bincount(b[:, c].reshape(-1).long(), minlength=256))
    s = torch.cuda.Stream(cuda)
    with torch.cuda.stream(s):
        sums = []
        for i in range(30):
            out = ops.hm_transform(batches[i % 3], ref_hist)
            sums.append((i % 3, out.sum(dtype=torch.int64), (out != want[i % 3]).sum()))
            del out  # the next transform's output / workspace recycle this memory
    s.synchronize()
    for k, total, bad in sums:
        assert int(bad) == 0 and int(total) == int(want[k].sum(dtype=torch.int64))
    # small batches take the general kernels behind the same chain
    small = batches[1][:1, :, :200, :333].contiguous()
    counts = ops.hm_hist(small)
    assert torch.equal(ops.hm_transform(small, ref_hist), ops.hm_apply(small, ops.hm_build_lut(counts, small.numel() // 3, ref_cdf)))
    # graph capture of the chain
    import ctypes

    src = batches[0]
    static_out = torch.empty_like(src)
    ws = torch.empty(int(nv_lib().sx_hm_workspace_bytes()), dtype=torch.uint8, device=cuda)

    def enqueue():
        rc = nv_lib().sx_hm_transform(ctypes.c_void_p(src.data_ptr()), 0, 0, 12, 1024, 1024, ctypes.c_void_p(ref_hist.data_ptr()), ctypes.c_void_p(static_out.data_ptr()),
                                      ctypes.c_void_p(ws.data_ptr()), ws.numel(), ctypes.c_void_p(torch.cuda.current_stream(cuda).cuda_stream))
        assert rc == 0

    enqueue()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        enqueue()
        enqueue()
    static_out.zero_()
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(static_out, want[0])


# ======================================================================== Reinhard
@pytest.mark.parametrize("name", ["reinhard_u8", "reinhard_f32", "reinhard_he_u8"])
def test_reinhard_golden(cuda, name):
    from stainx_b200 import Reinhard

    g = golden(name)
    n = Reinhard(device=cuda, backend="torch_cuda").fit(torch.from_numpy(g["ref"]).to(cuda))
    assert np.abs(_np(n._reference_mean) - g["mean"]).max() <= 1e-3
    assert np.abs(_np(n._reference_std) - g["std"]).max() <= 1e-3
    # transform with the reference's own fitted parameters (isolates the transform)
    n._reference_mean = torch.from_numpy(g["mean"]).to(cuda)
    n._reference_std = torch.from_numpy(g["std"]).to(cuda)
    out = _np(n.transform(torch.from_numpy(g["src"]).to(cuda)))
    assert out.dtype == g["out"].dtype
    diff = np.abs(out.astype(np.float64) - g["out"].astype(np.float64))
    if out.dtype == np.uint8:
        assert diff.max() <= 1
        assert (diff > 0).mean() < 0.01
    else:
        assert diff.max() <= F32_TOL


@pytest.mark.parametrize("dtype", ["u8", "f32"])
@pytest.mark.parametrize("shape", [(1, 3, 1, 2), (2, 3, 9, 7), (2, 3, 64, 64), (1, 3, 321, 199), (3, 3, 128, 256), (10, 3, 512, 512)])
def test_reinhard_vs_oracle(cuda, ox, dtype, shape):
    """Last shape = BASELINE config C1 (README quick-start): fit 1x3x512x512, transform 10x3x512x512."""
    from stainx_b200 import Reinhard

    make = noise_u8 if dtype == "u8" else noise_f32
    ref = make((1, 3, shape[2], shape[3]), 42)
    src = make(shape, 43, 1.4)
    n = Reinhard(device=cuda, backend="torch_cuda").fit(ref.to(cuda))
    mean, std = ox.reinhard_fit(ref.numpy())
    assert np.abs(_np(n._reference_mean) - mean).max() <= 1e-3
    assert np.abs(_np(n._reference_std) - std).max() <= 1e-3
    out = _np(n.transform(src.to(cuda)))
    want = ox.reinhard_transform(src.numpy(), _np(n._reference_mean), _np(n._reference_std))
    diff = np.abs(out.astype(np.float64) - want.astype(np.float64))
    if dtype == "u8":
        assert diff.max() <= 1
        assert (diff > 0).mean() < 0.01
    else:
        assert diff.max() <= F32_TOL


def test_reinhard_identity_property(cuda):
    """fit_transform on the same batch maps LAB statistics onto themselves: output == input."""
    from stainx_b200 import Reinhard

    x = noise_f32((4, 3, 256, 256), 11).to(cuda)
    out = Reinhard(device=cuda, backend="torch_cuda").fit_transform(x)
    assert (out - x).abs().max().item() <= 2e-3


def test_reinhard_chain_back_to_back_and_in_a_graph(cuda, ox):
    """sx_reinhard_transform chains statistics -> finalize (a programmatic dependent launch) -> transform:
    back-to-back transforms of different batches with recycled scratch must each use their own statistics,
    eagerly on a side stream and replayed from a CUDA graph."""
    import ctypes

    from stainx_b200 import ops

    g = torch.Generator(device=cuda).manual_seed(21)
    batches = [torch.rand((6, 3, 512, 512), device=cuda, generator=g).pow(p) for p in (0.6, 1.0, 1.8)]
    mean = torch.tensor([160.0, 135.0, 120.0], device=cuda)
    std = torch.tensor([35.0, 9.0, 14.0], device=cuda)
    want = []
    for b in batches:  # phase-level calls with a device synchronisation between the phases
        sums = ops.reinhard_stats(b)
        torch.cuda.synchronize()
        m, s_ = ops.reinhard_finalize(sums)
        torch.cuda.synchronize()
        want.append(ops.reinhard_apply(b, m, s_, mean, std))
        torch.cuda.synchronize()
    ref = ox.reinhard_transform(_np(batches[0].cpu()), _np(mean.cpu()), _np(std.cpu()))
    assert np.abs(_np(want[0]) - ref).max() <= F32_TOL
    s = torch.cuda.Stream(cuda)
    with torch.cuda.stream(s):
        bad = []
        for i in range(24):
            out = ops.reinhard_transform(batches[i % 3], mean, std)
            bad.append((out != want[i % 3]).sum())
            del out
    s.synchronize()
    assert all(int(b) == 0 for b in bad)
    src = batches[1]
    static_out = torch.empty_like(src)
    ws = torch.empty(int(nv_lib().sx_reinhard_workspace_bytes()), dtype=torch.uint8, device=cuda)

    def enqueue():
        rc = nv_lib().sx_reinhard_transform(ctypes.c_void_p(src.data_ptr()), 1, 6, 512, 512, ctypes.c_void_p(mean.data_ptr()), ctypes.c_void_p(std.data_ptr()), ctypes.c_void_p(static_out.data_ptr()),
                                            ctypes.c_void_p(ws.data_ptr()), ws.numel(), ctypes.c_void_p(torch.cuda.current_stream(cuda).cuda_stream))
        assert rc == 0

    enqueue()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        enqueue()
        enqueue()
    static_out.zero_()
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(static_out, want[1])


def test_reinhard_phase_api_shards(cuda):
    """Statistics accumulate over shards (what the multi-GPU path all-reduces)."""
    from stainx_b200 import ops

    x = noise_f32((4, 3, 64, 64), 12).to(cuda)
    whole = ops.reinhard_stats(x)
    parts = ops.reinhard_stats(x[:1].contiguous())
    ops.reinhard_stats(x[1:].contiguous(), sums=parts)
    assert torch.allclose(whole, parts, rtol=1e-12, atol=1e-9)
    assert whole[6].item() == 4 * 64 * 64
    m, s = ops.reinhard_finalize(whole)
    m2, s2 = ops.reinhard_fit(x)
    assert torch.allclose(m, m2, rtol=0, atol=1e-5) and torch.allclose(s, s2, rtol=0, atol=1e-5)


# ======================================================================== Macenko
@pytest.mark.parametrize("name", ["macenko_he_u8", "macenko_he_f32", "macenko_he_pooled_u8"])
def test_macenko_golden(cuda, name):
    from stainx_b200 import Macenko

    g = golden(name)
    n = Macenko(device=cuda, backend="torch_cuda").fit(torch.from_numpy(g["ref"]).to(cuda))
    assert np.abs(_np(n._stain_matrix) - g["he"]).max() <= 1e-4
    assert np.abs(_np(n._target_max_conc) / g["maxc"] - 1).max() <= 1e-3
    n._stain_matrix = torch.from_numpy(g["he"]).to(cuda)
    n._target_max_conc = torch.from_numpy(g["maxc"]).to(cuda)
    out = _np(n.transform(torch.from_numpy(g["src"]).to(cuda)))
    assert out.dtype == g["out"].dtype and out.shape == g["out"].shape
    diff = np.abs(out.astype(np.float64) - g["out"].astype(np.float64))
    if out.dtype == np.uint8:
        assert diff.max() <= 1
        assert (diff > 0).mean() < 0.01
    else:
        assert diff.max() <= F32_TOL * 255.0  # float output stays in [0, 255]
    if "out01" in g.files:
        n.normalize_to_0_1 = True
        out01 = _np(n.transform(torch.from_numpy(g["src"]).to(cuda)))
        assert out01.dtype == np.float32
        d01 = np.abs(out01 - g["out01"])
        if g["src"].dtype == np.uint8:  # uint8 input is truncated to a grey level before the /255
            assert d01.max() <= 1.0 / 255.0 + 1e-7 and (d01 > 1e-7).mean() < 0.01
        else:
            assert d01.max() <= F32_TOL


def test_macenko_known_answer_512(cuda):
    """SURVEY.md section 8c known-answer vector: reference fit on the 512x512 seed-42 tile."""
    from stainx_b200 import Macenko

    g = golden("macenko_kat_512")
    tile = he_tile(512, 512, 42)
    if int(tile.long().sum()) != int(g["tile_sum"]):  # torch interpolate changed: fall back to the stored tile
        tile = torch.from_numpy(g["tile"])
    n = Macenko(device=cuda, backend="torch_cuda").fit(tile.to(cuda))
    he, maxc = _np(n._stain_matrix), _np(n._target_max_conc)
    assert np.abs(he - np.array([[0.50870, 0.34880], [0.74012, 0.78338], [0.43983, 0.51445]])).max() <= 1e-4
    assert np.abs(maxc - np.array([2.13133, 1.57859])).max() <= 2e-3
    assert np.abs(he - g["he"]).max() <= 1e-4
    assert np.abs(maxc / g["maxc"] - 1).max() <= 1e-3


def test_macenko_noise_golden(cuda, ox):
    """uint8 noise fixture: the golden vectors hold the reference evaluated with the middle
    eigenvector in canonical sign ('p') and flipped ('m').  On near-isotropic input that sign is not
    reproducible (SURVEY.md section 7 H-a): fit must match one of the two, and every transformed
    image must match the (reference-pinned) oracle under one of the two signs."""
    from stainx_b200 import Macenko

    g = golden("macenko_noise_u8")
    n = Macenko(device=cuda, backend="torch_cuda").fit(torch.from_numpy(g["ref"]).to(cuda))
    he = _np(n._stain_matrix)
    d = min(np.abs(he - g["he_p"]).max(), np.abs(he - g["he_m"]).max())
    assert d <= 1e-4, f"fit HE matches neither sign of the oracle: {d}"
    n._stain_matrix = torch.from_numpy(g["he_p"]).to(cuda)
    n._target_max_conc = torch.from_numpy(g["maxc_p"]).to(cuda)
    out = _np(n.transform(torch.from_numpy(g["src"]).to(cuda)))
    nimg = g["src"].shape[0]
    cand_p = ox.macenko_transform(g["src"], g["he_p"], g["maxc_p"], mid_signs=[1] * nimg)
    cand_m = ox.macenko_transform(g["src"], g["he_p"], g["maxc_p"], mid_signs=[-1] * nimg)
    assert np.abs(cand_p.astype(np.float64) - g["out_p"]).max() <= 1  # the oracle itself is pinned to the reference
    assert best_sign_diff(out, cand_p, cand_m).max() <= 1


@pytest.mark.parametrize("dtype", ["u8", "f32"])
@pytest.mark.parametrize("hw", [(64, 64), (96, 80), (321, 199), (256, 512)])
def test_macenko_vs_oracle_he_tiles(cuda, ox, dtype, hw):
    from stainx_b200 import Macenko

    h, w = hw
    ref = he_tile(h, w, 42)
    src = he_batch(3, h, w)
    if dtype == "f32":
        ref, src = ref.float() / 255.0, src.float() / 255.0
    n = Macenko(device=cuda, backend="torch_cuda").fit(ref.to(cuda))
    he, maxc = ox.macenko_fit(ref.numpy())
    assert np.abs(_np(n._stain_matrix) - he).max() <= 1e-4
    assert np.abs(_np(n._target_max_conc) / maxc - 1).max() <= 1e-3
    out = _np(n.transform(src.to(cuda)))
    want = ox.macenko_transform(src.numpy(), _np(n._stain_matrix), _np(n._target_max_conc))
    diff = np.abs(out.astype(np.float64) - want.astype(np.float64))
    if dtype == "u8":
        assert diff.max() <= 1
        assert (diff > 0).mean() < 0.01
    else:
        assert diff.max() <= F32_TOL * 255.0


def test_macenko_noise_vs_oracle_both_signs(cuda, ox):
    """torch.rand-style input (the bench distribution): per image, the CUDA output must match the
    oracle under one of the two middle-eigenvector signs."""
    from stainx_b200 import Macenko

    ref = noise_f32((1, 3, 128, 128), 42)
    src = noise_f32((4, 3, 128, 128), 43)
    n = Macenko(device=cuda, backend="torch_cuda", normalize_to_0_1=True).fit(ref.to(cuda))
    he_p, maxc_p = ox.macenko_fit(ref.numpy(), mid_sign=1)
    he_m, maxc_m = ox.macenko_fit(ref.numpy(), mid_sign=-1)
    he = _np(n._stain_matrix)
    if np.abs(he - he_p).max() <= np.abs(he - he_m).max():
        he_o, maxc_o = he_p, maxc_p
    else:
        he_o, maxc_o = he_m, maxc_m
    assert np.abs(he - he_o).max() <= 1e-4
    assert np.abs(_np(n._target_max_conc) / maxc_o - 1).max() <= 1e-3
    out = _np(n.transform(src.to(cuda)))
    cand_p = ox.macenko_transform(src.numpy(), he, _np(n._target_max_conc), mid_signs=[1] * 4) / 255.0
    cand_m = ox.macenko_transform(src.numpy(), he, _np(n._target_max_conc), mid_signs=[-1] * 4) / 255.0
    assert best_sign_diff(out, cand_p, cand_m).max() <= F32_TOL


def test_macenko_fallback_when_mask_empty(cuda, ox):
    """Bright image: no pixel has min OD >= 0.15, so all pixels are used (torch_backend.py L409-410)."""
    from stainx_b200 import Macenko

    ref = he_tile(64, 64, 42)
    g = torch.Generator().manual_seed(5)
    bright = (215 + 35 * torch.rand((2, 3, 48, 48), generator=g)).round().to(torch.uint8)
    # mix: one bright image, one normal tile, in the same batch
    src = torch.cat([bright[:1], he_tile(48, 48, 9, 1.0), bright[1:]])
    n = Macenko(device=cuda, backend="torch_cuda").fit(ref.to(cuda))
    he, maxc = _np(n._stain_matrix), _np(n._target_max_conc)
    out = _np(n.transform(src.to(cuda)))
    cand_p = ox.macenko_transform(src.numpy(), he, maxc, mid_signs=[1, 1, 1])
    cand_m = ox.macenko_transform(src.numpy(), he, maxc, mid_signs=[-1, -1, -1])
    assert best_sign_diff(out, cand_p, cand_m).max() <= 1


def test_macenko_brackets_always_hold_the_rank(cuda):
    """The subsample bracket must contain every wanted rank (status region stays zero), on noise,
    on stain-like tiles, on a periodic pattern that could alias with strided sampling, and on an
    image with a tiny tissue fraction."""
    from stainx_b200 import ops

    g = torch.Generator().manual_seed(3)
    noise = torch.rand((2, 3, 512, 512), generator=g)
    tiles = he_batch(2, 512, 512).float() / 255.0
    yy, xx = torch.meshgrid(torch.arange(512), torch.arange(512), indexing="ij")
    stripes = (0.25 + 0.5 * ((xx // 4 + yy // 64) % 2).float()).expand(1, 3, 512, 512).clone()
    stripes[:, 0] *= 0.8
    stripes += 0.05 * torch.rand((1, 3, 512, 512), generator=g)
    sparse = torch.full((1, 3, 512, 512), 0.97)
    sparse[:, :, 100:108, 200:232] = tiles[0, :, 100:108, 200:232]
    batch = torch.cat([noise, tiles, stripes.clamp(0, 1), sparse]).contiguous().to(cuda)
    he, maxc = ops.macenko_fit(tiles[:1].contiguous().to(cuda))
    ws = ops.MacenkoWorkspace(batch.shape[0], cuda)
    ws.begin()
    ws.moments(batch, False)
    ws.basis(0, batch.shape[0], True)
    ws.moments_fallback(batch)
    for stage in (0, 1):
        for level in (0, 1):
            ws.hist(batch, False, stage, level)
            ws.select(0, batch.shape[0], stage, level)
    assert int(ws.region("status").abs().sum()) == 0
    assert bool(torch.isfinite(ws.region("fit")).all())


def test_macenko_per_image_independence(cuda):
    """Every statistic is per image: transform(batch)[i] == transform(batch[i:i+1]) bit for bit."""
    from stainx_b200 import Macenko

    ref = he_tile(128, 128, 42)
    src = he_batch(5, 128, 128).float() / 255.0
    n = Macenko(device=cuda, backend="torch_cuda", normalize_to_0_1=True).fit(ref.to(cuda))
    whole = n.transform(src.to(cuda))
    for i in range(5):
        assert torch.equal(whole[i : i + 1], n.transform(src[i : i + 1].to(cuda)))


@pytest.mark.parametrize("dtype", ["u8", "f32", "f16"])
def test_macenko_per_image_independence_large_images(cuda, dtype):
    """The same at sizes where an image's rows are split over many CTAs and the split depends on the batch around it
    (1, 3 and 11 images): the moments are fixed-point sums of float32 partials over row blocks that are aligned within
    the image, so an image's statistics -- and every bit of its output -- do not depend on the batch it travels in."""
    from stainx_b200 import Macenko

    ref = he_tile(256, 256, 42)
    src = torch.cat([he_batch(8, 768, 1024), noise_u8((3, 3, 768, 1024), 9)])
    if dtype != "u8":
        ref, src = ref.float() / 255.0, src.float() / 255.0
    if dtype == "f16":
        ref, src = ref.half(), src.half()
    n = Macenko(device=cuda, backend="torch_cuda", normalize_to_0_1=dtype != "u8").fit(ref.to(cuda))
    dev_src = src.to(cuda)
    whole = n.transform(dev_src)
    for i in (0, 5, 10):
        assert torch.equal(whole[i : i + 1], n.transform(dev_src[i : i + 1].contiguous()))
    assert torch.equal(whole[4:7], n.transform(dev_src[4:7].contiguous()))


@pytest.mark.parametrize("dtype,n,h,w,unit", [("u8", 5, 256, 320, False), ("f32", 3, 255, 257, True), ("f16", 4, 256, 256, True), ("f32", 24, 1024, 1024, True), ("u8", 30, 1024, 1024, True)])
def test_macenko_fit_transform_is_fit_then_transform(cuda, dtype, n, h, w, unit):
    """sx_macenko_fit_transform reads the batch once for the moments of both steps (per-image fixed-point moments,
    summed, ARE the pooled moments): fitted parameters and output equal fit() followed by transform() bit for bit --
    single chain, unaligned (scalar kernels) and multi-chain batches (> 64 MB)."""
    from stainx_b200 import Macenko

    x = he_batch(min(n, 8), h, w).repeat((n + 7) // 8, 1, 1, 1)[:n]
    if dtype != "u8":
        x = x.float() / 255.0
    if dtype == "f16":
        x = x.half()
    x = x.to(cuda)
    one = Macenko(device=cuda, backend="torch_cuda", normalize_to_0_1=unit)
    out = one.fit_transform(x)
    two = Macenko(device=cuda, backend="torch_cuda", normalize_to_0_1=unit).fit(x)
    assert torch.equal(one._stain_matrix, two._stain_matrix) and torch.equal(one._target_max_conc, two._target_max_conc)
    want = two.transform(x)
    assert out.dtype == want.dtype and torch.equal(out, want)


def test_macenko_fit_transform_vs_oracle(cuda, ox):
    """... and against the CPU oracle's pooled fit + per-image transform (float32, [0, 1] output)."""
    from stainx_b200 import Macenko

    x = he_batch(4, 256, 256).float() / 255.0
    n = Macenko(device=cuda, backend="torch_cuda", normalize_to_0_1=True)
    out = n.fit_transform(x.to(cuda))
    he, maxc = ox.macenko_fit(x.numpy())
    assert np.abs(n._stain_matrix.cpu().numpy() - he).max() <= 1e-4
    assert np.abs(n._target_max_conc.cpu().numpy() / maxc - 1).max() <= 1e-3
    want = ox.macenko_transform(x.numpy(), he, maxc) / 255.0
    assert np.abs(out.cpu().numpy() - want).max() <= 1e-3


def test_macenko_fit_transform_native_argument_checks(cuda):
    """The C-ABI entry point rejects a workspace sized for n instead of n + 1 slots and an empty batch."""
    import ctypes

    from stainx_b200 import _native as nv

    x = (he_batch(2, 64, 64).float() / 255.0).to(cuda)
    he, maxc, out = torch.empty(6, device=cuda), torch.empty(2, device=cuda), torch.empty_like(x)
    nbytes = int(nv.lib().sx_macenko_workspace_bytes(2))
    ws = torch.empty(int(nv.lib().sx_macenko_workspace_bytes(3)), dtype=torch.uint8, device=cuda)
    p = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
    rc = nv.lib().sx_macenko_fit_transform(p(x), nv.SX_F32, 2, 64, 64, p(he), p(maxc), p(out), nv.SX_F32, ctypes.c_float(1.0), p(ws), nbytes, None)
    assert rc != 0 and b"workspace too small" in nv.lib().sx_last_error()
    rc = nv.lib().sx_macenko_fit_transform(p(x), nv.SX_F32, 0, 64, 64, p(he), p(maxc), p(out), nv.SX_F32, ctypes.c_float(1.0), p(ws), ws.numel(), None)
    assert rc != 0 and b"empty reference batch" in nv.lib().sx_last_error()
    rc = nv.lib().sx_macenko_fit_transform(p(x), nv.SX_F32, 2, 64, 64, p(he), p(maxc), p(out), nv.SX_U8, ctypes.c_float(1.0), p(ws), ws.numel(), None)
    assert rc != 0 and b"output dtype" in nv.lib().sx_last_error()


def test_macenko_self_reference_reconstructs(cuda):
    """With the image's own HE / maxC as target, the Beer-Lambert tile is reproduced (the stain
    plane projection of OD is OD itself up to 8-bit rounding)."""
    from stainx_b200 import Macenko

    x = he_tile(256, 256, 77, 1.05).to(cuda)
    n = Macenko(device=cuda, backend="torch_cuda")
    out = n.fit_transform(x)
    assert out.dtype == torch.uint8
    assert (out.int() - x.int()).abs().max().item() <= 2


def test_macenko_full_size_float(cuda, ox):
    """BASELINE config C3 shape class (float32 1024x1024): a 6-image batch against the oracle, and
    the full 64-image batch through batch-vs-single equality on sampled images."""
    from stainx_b200 import Macenko

    ref = he_tile(1024, 1024, 42).float() / 255.0
    src = he_batch(6, 1024, 1024).float() / 255.0
    n = Macenko(device=cuda, backend="torch_cuda", normalize_to_0_1=True).fit(ref.to(cuda))
    he, maxc = ox.macenko_fit(ref.numpy())
    assert np.abs(_np(n._stain_matrix) - he).max() <= 1e-4
    assert np.abs(_np(n._target_max_conc) / maxc - 1).max() <= 1e-3
    out = _np(n.transform(src.to(cuda)))
    want = ox.macenko_transform(src[:2].numpy(), _np(n._stain_matrix), _np(n._target_max_conc)) / 255.0
    assert np.abs(out[:2] - want).max() <= F32_TOL
    big = src.to(cuda).repeat(11, 1, 1, 1)[:64].contiguous()
    whole = n.transform(big)
    assert tuple(whole.shape) == (64, 3, 1024, 1024)
    # per-image statistics: an image's result must not depend on the batch around it (up to the
    # summation order of the moment partial sums, which follows the CTA split)
    for i in (0, 17, 63):
        assert (whole[i] - torch.from_numpy(out[i % 6]).to(cuda)).abs().max().item() <= 1e-5


def test_macenko_sharded_fit_emulation(cuda):
    """Pooled fit over two 'ranks' emulated on one GPU: per-shard phases + the reductions a
    process group would apply (SUM / MAX / MIN) give the single-device fit bit for bit."""
    from stainx_b200 import _native as nv
    from stainx_b200 import ops

    imgs = torch.cat([he_tile(96, 96, 42), he_tile(96, 96, 7, 1.1), he_tile(96, 96, 8, 0.9)]).to(cuda)
    he_ref, maxc_ref = ops.macenko_fit(imgs)
    shards = [imgs[:1].contiguous(), imgs[1:].contiguous()]
    wss = [ops.MacenkoWorkspace(1, cuda) for _ in shards]

    def reduce(name, op):
        stack = torch.stack([w.region(name) for w in wss])
        red = {"sum": stack.sum(0), "max": stack.max(0).values, "min": stack.min(0).values}[op]
        for w in wss:
            w.region(name).copy_(red.to(w.region(name).dtype))

    for w, s in zip(wss, shards):
        w.begin()
        w.moments(s, pooled=True)
    reduce("moments", "sum")
    reduce("odrange", "max")
    for w in wss:
        w.basis(0, 1, allow_fallback=False)
    for stage in (nv.SX_STAGE_ANGLE, nv.SX_STAGE_CONC):
        for w, s in zip(wss, shards):
            w.hist(s, True, stage, 0)
        reduce("hist1", "sum")
        reduce("counters", "sum")
        for w in wss:
            w.select(0, 1, stage, 0)
        for w, s in zip(wss, shards):
            w.hist(s, True, stage, 1)
        reduce("hist2", "sum")
        reduce("counters", "sum")
        reduce("vmin", "min")
        reduce("vmax", "max")
        for w in wss:
            w.select(0, 1, stage, 1)
    for w in wss:
        fit = w.region("fit")[0]
        assert torch.allclose(fit[:6].reshape(3, 2), he_ref, rtol=0, atol=1e-6)
        assert torch.allclose(fit[6:8], maxc_ref, rtol=1e-6, atol=0)
    assert torch.equal(wss[0].region("fit"), wss[1].region("fit"))
    assert int(wss[0].region("status").abs().sum()) == 0


def test_macenko_sharded_fit_with_ragged_ranks(cuda):
    """ranks of a sharded pooled fit may run different kernel variants (a shard whose planes are not
    16-byte aligned takes the scalar kernels: 1 pixel per sampled group instead of 16) or hold no images at all.  The
    sample brackets must still be identical on every rank -- the group size travels in the MAX-combined ODRANGE region --
    or the summed cell histograms would mix different cell definitions."""
    from stainx_b200 import _native as nv
    from stainx_b200 import ops

    imgs = torch.cat([he_tile(128, 160, 42), he_tile(128, 160, 7, 1.1), he_tile(128, 160, 8, 0.9)]).to(cuda)
    he_ref, maxc_ref = ops.macenko_fit(imgs)
    flat = torch.empty(imgs[1:].numel() + 1, dtype=torch.uint8, device=cuda)
    misaligned = flat[1:].view(2, 3, 128, 160)  # contiguous, but every plane starts on an odd address
    misaligned.copy_(imgs[1:])
    assert misaligned.data_ptr() % 16 != 0 and misaligned.is_contiguous()
    shards = [imgs[:1].contiguous(), misaligned, imgs[:0]]  # vector kernels / scalar kernels / no images
    wss = [ops.MacenkoWorkspace(1, cuda) for _ in shards]

    def reduce(name, op):
        stack = torch.stack([w.region(name) for w in wss])
        red = {"sum": stack.sum(0), "max": stack.max(0).values, "min": stack.min(0).values}[op]
        for w in wss:
            w.region(name).copy_(red.to(w.region(name).dtype))

    for w, s in zip(wss, shards):
        w.begin()
        if s.shape[0]:
            w.moments(s, pooled=True)
    reduce("moments", "sum")
    reduce("odrange", "max")
    for w in wss:
        w.basis(0, 1, allow_fallback=False)
    for stage in (nv.SX_STAGE_ANGLE, nv.SX_STAGE_CONC):
        for level in (0, 1):
            for w, s in zip(wss, shards):
                if s.shape[0]:
                    w.hist(s, True, stage, level)
            reduce("hist1" if level == 0 else "hist2", "sum")
            reduce("counters", "sum")
            if level == 1:
                reduce("vmin", "min")
                reduce("vmax", "max")
            for w in wss:
                w.select(0, 1, stage, level)
    for w in wss:
        assert torch.equal(w.region("fit"), wss[0].region("fit")), "ranks derived different fits"
        assert int(w.region("status")[0, 0]) == 0
    fit = wss[0].region("fit")[0]
    assert torch.allclose(fit[:6].reshape(3, 2), he_ref, rtol=0, atol=1e-6)
    assert torch.allclose(fit[6:8], maxc_ref, rtol=1e-5, atol=0)


@pytest.mark.gpu
def test_host_stream_matches_direct_transform(cuda):
    """ingest.HostStream (pinned host in/out, three-stream pipeline) returns exactly what the
    device-resident transform returns, batch after batch, also when staging buffers are recycled."""
    from stainx_b200 import HistogramMatching, Macenko
    from stainx_b200.ingest import HostStream

    g = torch.Generator().manual_seed(5)
    ref = (torch.rand(1, 3, 96, 128, generator=g) * 255).round().to(torch.uint8)
    batches = [(torch.rand(3, 3, 96, 128, generator=g) * 255).round().to(torch.uint8).pin_memory() for _ in range(5)]
    hm = HistogramMatching(device=cuda, backend="torch_cuda").fit(ref.to(cuda))
    pipe = HostStream(hm, depth=2)
    outs = list(pipe.map(batches))
    for b, o in zip(batches, outs):
        assert torch.equal(o, hm.transform(b.to(cuda)).cpu())
    gold = golden("macenko_he_u8")
    mk = Macenko(device=cuda, backend="torch_cuda").fit(torch.from_numpy(gold["ref"]).to(cuda))
    src = torch.from_numpy(gold["src"]).pin_memory()
    pipe = HostStream(mk, depth=3)
    tickets = [pipe.submit(src) for _ in range(4)]
    want = mk.transform(src.to(cuda)).cpu()
    for t in tickets:
        assert torch.equal(t.wait(), want)
    assert pipe.h2d_bytes == 4 * src.numel() * src.element_size()


@pytest.mark.gpu
def test_macenko_ties_overflow_the_hit_queue(cuda, ox):
    """Images made of a few flat colours put most pixels INSIDE the order-statistic brackets (ties),
    so the shared hit queue of the resolve pass overflows constantly and the direct path runs; the
    result must still match the oracle (uint8 and float32, pipeline and pooled fit)."""
    from stainx_b200 import Macenko

    g = torch.Generator().manual_seed(11)
    tile = he_tile(256, 256, 42)
    palette = tile[0, :, ::64, ::64].reshape(3, -1).T.contiguous()  # 16 stain-like colours
    idx = torch.randint(0, palette.shape[0], (3, 256, 256), generator=g)
    idx[:, :, :128] = idx[:, :1, :1]  # half of every image is one flat colour
    flat = palette[idx].permute(0, 3, 1, 2).contiguous()  # (3, 3, 256, 256) uint8
    ref = he_tile(256, 256, 7)
    for src in (flat, flat.float() / 255.0):
        n = Macenko(device=cuda, backend="torch_cuda").fit(ref.to(cuda) if src.dtype == torch.uint8 else (ref.float() / 255.0).to(cuda))
        out = _np(n.transform(src.to(cuda)))
        want = ox.macenko_transform(src.numpy(), _np(n._stain_matrix), _np(n._target_max_conc))
        diff = np.abs(out.astype(np.float64) - want.astype(np.float64))
        assert diff.max() <= (1 if src.dtype == torch.uint8 else F32_TOL * 255.0)
    he, maxc = ox.macenko_fit(flat.numpy())
    n = Macenko(device=cuda, backend="torch_cuda").fit(flat.to(cuda))
    assert np.abs(_np(n._stain_matrix) - he).max() <= 1e-4
    assert np.abs(_np(n._target_max_conc) / maxc - 1).max() <= 1e-3


def test_macenko_transform_in_cuda_graph_and_on_side_stream(cuda):
    """The multi-chain transform forks into library-owned side streams and joins back: it must stay
    ordered on a non-default caller stream and be capturable into a CUDA graph (replay == eager)."""
    from stainx_b200 import ops

    g = torch.Generator(device=cuda).manual_seed(9)
    src = (torch.rand((24, 3, 1024, 1024), device=cuda, generator=g) * 255).to(torch.uint8)  # 75 MB: several chains
    he, maxc = ops.macenko_fit(src[:1])
    want = ops.macenko_transform(src, he, maxc, unit=False)
    torch.cuda.synchronize()
    s = torch.cuda.Stream(cuda)
    with torch.cuda.stream(s):
        got = ops.macenko_transform(src, he, maxc, unit=False)
        total = got.float().sum()  # ordered after the join on the same stream
    s.synchronize()
    assert torch.equal(got, want) and float(total) == float(want.float().sum())
    # graph capture: static input/output/workspace buffers
    static_out = torch.empty_like(src)
    ws = torch.empty(int(nv_lib().sx_macenko_workspace_bytes(src.shape[0])), dtype=torch.uint8, device=cuda)
    import ctypes

    def enqueue():
        rc = nv_lib().sx_macenko_transform(ctypes.c_void_p(src.data_ptr()), 0, 24, 1024, 1024, ctypes.c_void_p(he.data_ptr()), ctypes.c_void_p(maxc.data_ptr()), ctypes.c_void_p(static_out.data_ptr()), 0,
                                           ctypes.c_float(1.0), ctypes.c_void_p(ws.data_ptr()), ws.numel(), ctypes.c_void_p(torch.cuda.current_stream(cuda).cuda_stream))
        assert rc == 0

    enqueue()  # warm-up outside capture (occupancy queries, function attributes)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        enqueue()
    static_out.zero_()
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(static_out, want)


def test_macenko_chain_layout_does_not_change_a_bit(cuda):
    """The multi-chain transform runs its per-image kernels on high-priority helper streams.  Chains and
    helper streams only change WHERE the same kernels run: one chain, three chains with the small kernels on the
    chains' own streams, and the default (three chains + helper streams) must agree bit for bit, uint8 and float32."""
    import os

    from stainx_b200 import _native as nv
    from stainx_b200 import ops

    os.environ["SX_ENABLE_TUNING"] = "1"  # read on the first call of a development hook
    lib = nv.lib()
    g = torch.Generator(device=cuda).manual_seed(21)
    src = torch.cat([he_batch(4, 1024, 1024).to(cuda).repeat(3, 1, 1, 1), (torch.rand((12, 3, 1024, 1024), device=cuda, generator=g) * 255).to(torch.uint8)])  # 75 MB
    he, maxc = ops.macenko_fit(src[:1])
    want = ops.macenko_transform(src, he, maxc, unit=False)
    srcf = src[:8].float() / 255.0  # 100 MB of float32: two chains of four images
    wantf = ops.macenko_transform(srcf, he, maxc, unit=True)
    assert lib.sx_macenko_set_tuning(-1, 0) == 0, "development hooks are not enabled"
    try:
        for bits in ((1 << 4), (3 << 4) | 4, (2 << 4), (4 << 4)):  # chains << 4 | 4: no helper streams
            assert lib.sx_macenko_set_tuning(-1, bits) == 0
            assert torch.equal(ops.macenko_transform(src, he, maxc, unit=False), want), bits
            assert torch.equal(ops.macenko_transform(srcf, he, maxc, unit=True), wantf), bits
    finally:
        lib.sx_macenko_set_tuning(-1, 0)


def test_macenko_fit_transform_in_cuda_graph_and_on_side_stream(cuda):
    """sx_macenko_fit_transform forks the pooled fit onto a library-owned stream beside the transform chains and joins
    before `apply`: ordered on a non-default caller stream, capturable into a CUDA graph, replay == eager."""
    import ctypes

    from stainx_b200 import ops

    src = (he_batch(8, 1024, 1024).repeat(3, 1, 1, 1).float() / 255.0).to(cuda)  # 24 x 12.6 MB: three chains + the fit stream
    he_w, maxc_w, want = ops.macenko_fit_transform(src, unit=True)
    torch.cuda.synchronize()
    s = torch.cuda.Stream(cuda)
    with torch.cuda.stream(s):
        he_g, maxc_g, got = ops.macenko_fit_transform(src, unit=True)
        total = got.sum() + he_g.sum()  # ordered after both joins on the same stream
    s.synchronize()
    assert torch.equal(got, want) and torch.equal(he_g, he_w) and torch.equal(maxc_g, maxc_w)
    assert float(total) == float(want.sum() + he_w.sum())
    static_out, he, maxc = torch.empty_like(src), torch.zeros(6, device=cuda), torch.zeros(2, device=cuda)
    ws = torch.empty(int(nv_lib().sx_macenko_workspace_bytes(src.shape[0] + 1)), dtype=torch.uint8, device=cuda)

    def enqueue():
        rc = nv_lib().sx_macenko_fit_transform(ctypes.c_void_p(src.data_ptr()), 1, 24, 1024, 1024, ctypes.c_void_p(he.data_ptr()), ctypes.c_void_p(maxc.data_ptr()), ctypes.c_void_p(static_out.data_ptr()), 1,
                                               ctypes.c_float(1.0 / 255.0), ctypes.c_void_p(ws.data_ptr()), ws.numel(), ctypes.c_void_p(torch.cuda.current_stream(cuda).cuda_stream))
        assert rc == 0

    enqueue()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        enqueue()
    static_out.zero_()
    he.zero_()
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(static_out, want) and torch.equal(he.reshape(3, 2), he_w) and torch.equal(maxc, maxc_w)


def nv_lib():
    from stainx_b200 import _native

    return _native.lib()


def test_real_he_tiles_all_methods(cuda):
    """Real H&E crops of the reference's example slides against the reference's own outputs
    (tests/golden/real_he_u8.npz): histogram matching bit-exact, Reinhard / Macenko within one grey
    level, fitted parameters within the parity bars."""
    from stainx_b200 import HistogramMatching, Macenko, Reinhard

    g = golden("real_he_u8")
    ref, src = torch.from_numpy(g["ref"]).to(cuda), torch.from_numpy(g["src"]).to(cuda)
    hm = HistogramMatching(device=cuda, backend="torch_cuda").fit(ref)
    assert np.array_equal(_np(torch.stack(hm._ref_histograms_256)), g["hm_ref_hist"])
    assert np.array_equal(_np(hm.transform(src)), g["hm_out"])
    rh = Reinhard(device=cuda, backend="torch_cuda").fit(ref)
    assert np.abs(_np(rh._reference_mean) - g["rh_mean"]).max() <= 1e-3 and np.abs(_np(rh._reference_std) - g["rh_std"]).max() <= 1e-3
    d = np.abs(_np(rh.transform(src)).astype(np.int32) - g["rh_out"].astype(np.int32))
    assert d.max() <= 1 and (d > 0).mean() < 0.01
    mk = Macenko(device=cuda, backend="torch_cuda").fit(ref)
    assert np.abs(_np(mk._stain_matrix) - g["mk_he"]).max() <= 1e-4
    assert np.abs(_np(mk._target_max_conc) / g["mk_maxc"] - 1).max() <= 1e-3
    d = np.abs(_np(mk.transform(src)).astype(np.int32) - g["mk_out"].astype(np.int32))
    assert d.max() <= 1 and (d > 0).mean() < 0.01


def test_macenko_missed_brackets_are_recovered_exactly(cuda, ox):
    """VERDICT r1 / ADVICE: a wanted rank outside its sample bracket must not end in a clamped answer.  The
    development hook replaces every sample bracket by an empty one far from the data, so every rank search
    misses: the transform, the pooled fit and the sharded-fit control flow (exact=True: level-2 histograms)
    must return what the unforced run returns, and the STATUS region must show recoveries, not misses."""
    import ctypes
    import os

    from stainx_b200 import _native as nv
    from stainx_b200 import ops
    from stainx_b200.backends.torch_cuda_backend import MacenkoCUDA

    os.environ["SX_ENABLE_TUNING"] = "1"  # read on the first call of a development hook
    lib = nv.lib()
    ref = he_tile(512, 512, 42)  # large enough that the sample is a true subsample (else the bracket is exact anyway)
    src = torch.cat([he_batch(2, 512, 512), noise_u8((1, 3, 512, 512), 5)]).to(cuda)
    srcf = (src.float() / 255.0).contiguous()
    he, maxc = ops.macenko_fit(ref.to(cuda))
    want_u8, want_f = ops.macenko_transform(src, he, maxc), ops.macenko_transform(srcf, he, maxc, unit=True)
    pooled = ops.macenko_fit(src)
    assert lib.sx_macenko_set_tuning(-1, 2) == 0, "development hooks are not enabled"
    try:
        n, _, h, w = src.shape
        nbytes = int(lib.sx_macenko_workspace_bytes(n))
        ws = torch.zeros(nbytes, dtype=torch.uint8, device=cuda)
        out = torch.empty_like(src)
        vp = ctypes.c_void_p
        rc = lib.sx_macenko_transform(vp(src.data_ptr()), 0, n, h, w, vp(he.data_ptr()), vp(maxc.data_ptr()), vp(out.data_ptr()), 0, ctypes.c_float(1.0), vp(ws.data_ptr()), nbytes, vp(torch.cuda.current_stream(cuda).cuda_stream))
        assert rc == 0
        off, size = ctypes.c_int64(), ctypes.c_int64()
        lib.sx_macenko_region(n, nv.REGIONS["status"], ctypes.byref(off), ctypes.byref(size))
        status = ws[off.value : off.value + size.value].view(torch.int32).view(n, 4).cpu()
        assert int(status[:, 0].abs().sum()) == 0, "an unrecovered miss"
        assert bool((status[:, 3] == 2).all()), f"both stages of every image must have taken the exact path: {status[:, 3].tolist()}"
        # The exact pass resolves a different bracket (the coarse bin of the rank) into its 4096 cells, so a selected
        # value may differ from the sampled path's inside one cell width (<= bracket / 4096): same bars as vs the oracle.
        d = (out.int() - want_u8.int()).abs()
        assert int(d.max()) <= 1 and float((d > 0).float().mean()) < 0.01
        got_f = ops.macenko_transform(srcf, he, maxc, unit=True)
        assert float((got_f - want_f).abs().max()) <= 1e-4
        he2, maxc2 = ops.macenko_fit(src)  # pooled fit, in-kernel recovery
        assert float((he2 - pooled[0]).abs().max()) <= 1e-5 and float((maxc2 / pooled[1] - 1).abs().max()) <= 1e-4
        # the phase-level protocol of the sharded fit: level 0 misses, exact=True (level 2) recovers
        impl = MacenkoCUDA(cuda)
        impl._reducer.enabled = False
        he3, maxc3 = impl._pooled_fit_sharded(src)
        assert float((he3 - pooled[0]).abs().max()) <= 1e-5 and float((maxc3 / pooled[1] - 1).abs().max()) <= 1e-4
        want = ox.macenko_transform(_np(src), _np(he), _np(maxc), mid_signs=[1, 1, 1])  # and the recovered transform meets the oracle bar
        alt = ox.macenko_transform(_np(src), _np(he), _np(maxc), mid_signs=[-1, -1, -1])
        assert best_sign_diff(_np(out), want, alt).max() <= 1
    finally:
        assert lib.sx_macenko_set_tuning(-1, 0) == 0


def test_macenko_is_bit_reproducible_run_to_run(cuda):
    """The per-image moments are combined as fixed-point integers, so the order in which the CTAs of the moments pass
    finish cannot change them: repeated transforms (and fits) of the same batch return the same bits, uint8 and float32,
    multi-chain batches included."""
    from stainx_b200 import ops

    g = torch.Generator(device=cuda).manual_seed(19)
    src8 = (torch.rand((24, 3, 1024, 1024), device=cuda, generator=g) * 255).to(torch.uint8)  # 75 MB: three chains
    srcf = torch.rand((8, 3, 1024, 1024), device=cuda, generator=g)
    he, maxc = ops.macenko_fit(src8[:1].contiguous())
    want8, wantf, want_fit = ops.macenko_transform(src8, he, maxc), ops.macenko_transform(srcf, he, maxc, unit=True), ops.macenko_fit(srcf)
    bad = []
    for _ in range(6):
        bad.append((ops.macenko_transform(src8, he, maxc) != want8).sum())
        bad.append((ops.macenko_transform(srcf, he, maxc, unit=True) != wantf).sum())
        fit = ops.macenko_fit(srcf)
        bad.append((fit[0] != want_fit[0]).sum() + (fit[1] != want_fit[1]).sum())
    assert all(int(b) == 0 for b in bad)
