"""Golden vectors for float16 / bfloat16 tensors, generated from the REFERENCE itself (build container only).

    python tests/golden/make_golden_half.py

The reference's torch CPU backend widens 16-bit float tensors to float32, computes, and casts the result back
(torch_backend.py:L103-131); Macenko's ``normalize_to_0_1`` then divides the cast-back tensor by 255 in the
same dtype (normalizers/_template.py:L111-112).  numpy has no bfloat16, so the 16-bit tensors are stored as
their raw bit patterns (int16) next to the dtype name; ``tests/test_gpu_half.py`` re-views them.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, "/root/reference/src")
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from stainx import HistogramMatching, Macenko, Reinhard  # noqa: E402

from tests.helpers import he_batch, he_tile, noise_f32  # noqa: E402

OUT = Path(__file__).resolve().parent


def bits(t: torch.Tensor) -> np.ndarray:
    return t.contiguous().view(torch.int16).numpy()


arrays = {}
for name, dt in (("f16", torch.float16), ("bf16", torch.bfloat16)):
    # histogram matching (skewed noise), NCHW, odd size: scalar head / tail paths
    ref, src = noise_f32((1, 3, 40, 56), 61, 0.7).to(dt), noise_f32((2, 3, 37, 41), 62, 1.6).to(dt)
    n = HistogramMatching(device="cpu", backend="torch", channel_axis=1).fit(ref)
    out = n.transform(src)
    assert out.dtype == dt
    arrays.update({f"hm_{name}_ref": bits(ref), f"hm_{name}_src": bits(src), f"hm_{name}_out": bits(out), f"hm_{name}_ref_hist": torch.stack(n._ref_histograms_256).numpy()})
    # Reinhard on stain-like tiles
    ref, src = (he_tile(64, 80, 42).float() / 255.0).to(dt), (he_batch(2, 64, 80).float() / 255.0).to(dt)
    n = Reinhard(device="cpu", backend="torch").fit(ref)
    out = n.transform(src)
    assert out.dtype == dt
    arrays.update({f"rh_{name}_ref": bits(ref), f"rh_{name}_src": bits(src), f"rh_{name}_out": bits(out), f"rh_{name}_mean": n._reference_mean.float().numpy(), f"rh_{name}_std": n._reference_std.float().numpy()})
    # Macenko on stain-like tiles, [0, 255] output and normalize_to_0_1
    n = Macenko(device="cpu", backend="torch").fit(ref)
    out = n.transform(src)
    n.normalize_to_0_1 = True
    out01 = n.transform(src)
    assert out.dtype == dt and out01.dtype == dt
    arrays.update({f"mk_{name}_out": bits(out), f"mk_{name}_out01": bits(out01), f"mk_{name}_he": n._stain_matrix.float().numpy(), f"mk_{name}_maxc": n._target_max_conc.float().numpy()})
np.savez_compressed(OUT / "half_io.npz", **arrays)
print({k: (v.shape, str(v.dtype)) for k, v in arrays.items()})
