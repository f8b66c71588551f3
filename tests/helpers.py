"""Input generators shared by the tests (mirrors the reference's fixtures)."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F  # noqa: N812

HE_REF = torch.tensor([[0.5626, 0.2159], [0.7201, 0.8012], [0.4062, 0.5581]], dtype=torch.float32)


def he_tile(h: int, w: int, seed: int, he_scale: float = 1.0) -> torch.Tensor:
    """Beer-Lambert H&E tile ``I = 240 exp(-(HE s) C)`` with low-frequency concentration maps:
    the reference's Macenko fixture (tests/torch_interface/test_correctness_against_references.py
    L41-54).  Returns uint8 (1, 3, h, w)."""
    g = torch.Generator().manual_seed(seed)
    gh, gw = max(h // 8, 1), max(w // 8, 1)
    c_h = F.interpolate(torch.rand(1, 1, gh, gw, generator=g), size=(h, w), mode="bilinear", align_corners=False).squeeze()
    c_e = F.interpolate(torch.rand(1, 1, gh, gw, generator=g), size=(h, w), mode="bilinear", align_corners=False).squeeze()
    conc = torch.stack([0.3 + 1.8 * c_h, 0.2 + 1.0 * c_e], dim=0)
    od = torch.einsum("cs,shp->chp", HE_REF * he_scale, conc)
    return (240.0 * torch.exp(-od)).clamp(0, 255).round().to(torch.uint8).unsqueeze(0)


def he_batch(n: int, h: int, w: int, seed0: int = 123) -> torch.Tensor:
    scales = [1.15, 0.9, 1.05, 0.85, 1.1, 0.95, 1.0, 1.2]
    return torch.cat([he_tile(h, w, seed0 + i, scales[i % len(scales)]) for i in range(n)])


def noise_u8(shape, seed: int, gamma: float = 1.0) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(*shape, generator=g).pow(gamma) * 255).round().to(torch.uint8)


def noise_f32(shape, seed: int, gamma: float = 1.0) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.rand(*shape, generator=g).pow(gamma)


def best_sign_diff(out: np.ndarray, cand_p: np.ndarray, cand_m: np.ndarray) -> np.ndarray:
    """Per-image max-abs difference against the better of two oracle evaluations (middle
    eigenvector sign +/-): the parity protocol for Macenko on near-isotropic inputs
    (SURVEY.md section 7 H-a)."""
    o = out.astype(np.float64)
    dp = np.abs(o - cand_p.astype(np.float64)).reshape(out.shape[0], -1).max(axis=1)
    dm = np.abs(o - cand_m.astype(np.float64)).reshape(out.shape[0], -1).max(axis=1)
    return np.minimum(dp, dm)
