"""Generate golden input/output vectors from the REFERENCE itself (run in the build container).

    python tests/golden/make_golden.py

Imports rendeirolab/stainx v0.1.4 from /root/reference/src (read-only), runs its torch CPU backend
(the numerical oracle of SURVEY.md section 8c: ``backend="torch"``, ``device="cpu"``) on small seeded
inputs and stores inputs + outputs in ``tests/golden/*.npz``.  /root/reference does not exist on
the GPU box, so the tests only ever read the committed .npz files.

Macenko on noise is evaluated with both signs of the middle eigenvector (SURVEY.md section 7 H-a):
``MacenkoTorch._eigh_torch`` (torch_backend.py L367-373) is wrapped to flip column 1.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F  # noqa: N812

sys.path.insert(0, "/root/reference/src")
import stainx  # noqa: E402
from stainx import HistogramMatching, Macenko, Reinhard  # noqa: E402
from stainx.backends.torch_backend import MacenkoTorch  # noqa: E402

OUT = Path(__file__).resolve().parent

HE_REF = torch.tensor([[0.5626, 0.2159], [0.7201, 0.8012], [0.4062, 0.5581]], dtype=torch.float32)


def he_tile(h: int, w: int, seed: int, he_scale: float = 1.0) -> torch.Tensor:
    """Beer-Lambert H&E tile, same construction as the reference's test fixture
    (tests/torch_interface/test_correctness_against_references.py L41-54)."""
    g = torch.Generator().manual_seed(seed)
    gh, gw = max(h // 8, 1), max(w // 8, 1)
    c_h = F.interpolate(torch.rand(1, 1, gh, gw, generator=g), size=(h, w), mode="bilinear", align_corners=False).squeeze()
    c_e = F.interpolate(torch.rand(1, 1, gh, gw, generator=g), size=(h, w), mode="bilinear", align_corners=False).squeeze()
    conc = torch.stack([0.3 + 1.8 * c_h, 0.2 + 1.0 * c_e], dim=0)
    od = torch.einsum("cs,shp->chp", HE_REF * he_scale, conc)
    return (240.0 * torch.exp(-od)).clamp(0, 255).round().to(torch.uint8).unsqueeze(0)


def noise_u8(shape, seed, gamma=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(*shape, generator=g).pow(gamma) * 255).round().to(torch.uint8)


def noise_f32(shape, seed, gamma=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(*shape, generator=g).pow(gamma)


def save(name, **arrays):
    conv = {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in arrays.items()}
    np.savez_compressed(OUT / f"{name}.npz", **conv)
    print(name, {k: (v.shape, str(v.dtype)) for k, v in conv.items()})


def hm_cases():
    cases = {
        "hm_u8": (noise_u8((1, 3, 40, 56), 42, 2.0), noise_u8((2, 3, 37, 41), 43, 0.5)),
        "hm_f32": (noise_f32((1, 3, 40, 56), 44, 0.7), noise_f32((2, 3, 37, 41), 45, 1.6)),
        "hm_u8_uniform": (noise_u8((1, 3, 64, 64), 46), noise_u8((3, 3, 64, 64), 47)),
    }
    for name, (ref, src) in cases.items():
        n = HistogramMatching(device="cpu", backend="torch", channel_axis=1).fit(ref)
        out = n.transform(src)
        save(name, ref=ref, src=src, ref_hist=torch.stack(n._ref_histograms_256), out=out)
    # knife edge: fit and transform on the same tensor (LUT lands on integers +- 1 ulp)
    x = noise_u8((2, 3, 48, 48), 48, 1.5)
    n = HistogramMatching(device="cpu", backend="torch", channel_axis=1)
    out = n.fit_transform(x)
    save("hm_u8_self", ref=x, src=x, ref_hist=torch.stack(n._ref_histograms_256), out=out)
    # NHWC (channel_axis=-1)
    ref, src = noise_u8((1, 40, 56, 3), 49, 2.0), noise_u8((2, 37, 41, 3), 50, 0.5)
    n = HistogramMatching(device="cpu", backend="torch", channel_axis=-1).fit(ref)
    save("hm_u8_nhwc", ref=ref, src=src, ref_hist=torch.stack(n._ref_histograms_256), out=n.transform(src))
    # sparse histogram (few distinct grey levels; exercises quantile_diff <= 1e-10 and edge rules)
    ref = (noise_u8((1, 3, 32, 32), 51) // 64) * 64
    src = (noise_u8((2, 3, 32, 32), 52) // 32) * 32
    n = HistogramMatching(device="cpu", backend="torch", channel_axis=1).fit(ref)
    save("hm_u8_sparse", ref=ref, src=src, ref_hist=torch.stack(n._ref_histograms_256), out=n.transform(src))


def reinhard_cases():
    cases = {
        "reinhard_u8": (noise_u8((1, 3, 32, 48), 60), noise_u8((2, 3, 33, 47), 61)),
        "reinhard_f32": (noise_f32((1, 3, 32, 48), 62), noise_f32((2, 3, 33, 47), 63)),
        "reinhard_he_u8": (he_tile(64, 64, 42), torch.cat([he_tile(64, 64, 123, 1.15), he_tile(64, 64, 124, 0.9)])),
    }
    for name, (ref, src) in cases.items():
        n = Reinhard(device="cpu", backend="torch").fit(ref)
        out = n.transform(src)
        save(name, ref=ref, src=src, mean=n._reference_mean, std=n._reference_std, out=out.contiguous())
    # colour-space round trip fixture (reference tests/test_torch_backend_color_space.py)
    from stainx.backends.torch_backend import TorchBackendBase

    x = noise_f32((1, 3, 16, 16), 64)
    lab = TorchBackendBase.rgb_to_lab_torch(x)
    save("colorspace_f32", rgb=x, lab=lab.contiguous(), back=TorchBackendBase.lab_to_rgb_torch(lab).contiguous())


def macenko_cases():
    # well-posed Beer-Lambert tiles (sign independent)
    ref = he_tile(96, 96, 42, 1.0)
    src = torch.cat([he_tile(96, 96, 123, 1.15), he_tile(96, 96, 124, 0.9)])
    for name, cast in (("macenko_he_u8", lambda t: t), ("macenko_he_f32", lambda t: t.float() / 255.0)):
        n = Macenko(device="cpu", backend="torch").fit(cast(ref))
        out = n.transform(cast(src))
        n01 = Macenko(device="cpu", backend="torch", normalize_to_0_1=True)
        n01._stain_matrix, n01._target_max_conc, n01._is_fitted = n._stain_matrix, n._target_max_conc, True
        save(name, ref=cast(ref), src=cast(src), he=n._stain_matrix, maxc=n._target_max_conc, out=out, out01=n01.transform(cast(src)))
    # pooled fit over several reference images + non-square tile
    ref = torch.cat([he_tile(80, 56, 42, 1.0), he_tile(80, 56, 7, 1.1)])
    src = he_tile(80, 56, 99, 0.95)
    n = Macenko(device="cpu", backend="torch").fit(ref)
    save("macenko_he_pooled_u8", ref=ref, src=src, he=n._stain_matrix, maxc=n._target_max_conc, out=n.transform(src))
    # known-answer vector quoted in SURVEY.md section 8c (512x512 tile, seed 42): store the answer and
    # a checksum of the tile; the tile itself is regenerated by the test helper.
    big = he_tile(512, 512, 42)
    n = Macenko(device="cpu", backend="torch").fit(big)
    save("macenko_kat_512", he=n._stain_matrix, maxc=n._target_max_conc, tile_sum=np.int64(big.long().sum().item()), tile=big)

    # noise: evaluate the reference with both middle-eigenvector signs
    ref = noise_u8((1, 3, 64, 64), 70)
    src = noise_u8((2, 3, 64, 64), 71)
    orig = MacenkoTorch._eigh_torch
    res = {}
    for sign in (1, -1):
        def flipped(cov, _s=sign):
            w, v = orig(cov)
            v = v.clone()
            # canonical sign (largest |component| positive) times _s, so the stored vectors do not
            # depend on this machine's LAPACK sign choice
            col = v[:, 1]
            big_i = col.abs().argmax()
            v[:, 1] = col * (1.0 if col[big_i] >= 0 else -1.0) * _s
            return w, v

        MacenkoTorch._eigh_torch = staticmethod(flipped)
        n = Macenko(device="cpu", backend="torch").fit(ref)
        res[f"he_{'p' if sign > 0 else 'm'}"] = n._stain_matrix
        res[f"maxc_{'p' if sign > 0 else 'm'}"] = n._target_max_conc
        res[f"out_{'p' if sign > 0 else 'm'}"] = n.transform(src)
    MacenkoTorch._eigh_torch = staticmethod(orig)
    n = Macenko(device="cpu", backend="torch").fit(ref)
    save("macenko_noise_u8", ref=ref, src=src, he_native=n._stain_matrix, maxc_native=n._target_max_conc, out_native=n.transform(src), **res)


def real_he_cases():
    """Real H&E tiles: 160x160 crops of the reference's example images (examples/data/target.png as
    the reference slide, test_1 / test_3 as sources), normalised by all three methods of the
    reference's torch CPU backend."""
    from PIL import Image

    data = Path("/root/reference/examples/data")

    def crop(name, y, x, size=160):
        im = np.asarray(Image.open(data / f"{name}.png").convert("RGB"))
        return torch.from_numpy(np.ascontiguousarray(im[y:y + size, x:x + size].transpose(2, 0, 1))).unsqueeze(0)

    ref = crop("target", 400, 400)
    src = torch.cat([crop("test_1", 300, 500), crop("test_3", 600, 200)])
    out = {}
    n = HistogramMatching(device="cpu", backend="torch", channel_axis=1).fit(ref)
    out["hm_ref_hist"], out["hm_out"] = torch.stack(n._ref_histograms_256), n.transform(src)
    n = Reinhard(device="cpu", backend="torch").fit(ref)
    out["rh_mean"], out["rh_std"], out["rh_out"] = n._reference_mean, n._reference_std, n.transform(src).contiguous()
    n = Macenko(device="cpu", backend="torch").fit(ref)
    out["mk_he"], out["mk_maxc"], out["mk_out"] = n._stain_matrix, n._target_max_conc, n.transform(src)
    save("real_he_u8", ref=ref, src=src, **out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "real":
        real_he_cases()
        sys.exit(0)
    print("reference stainx", stainx.__version__, "torch", torch.__version__, torch.backends.cpu.get_cpu_capability())
    hm_cases()
    reinhard_cases()
    macenko_cases()
    real_he_cases()
