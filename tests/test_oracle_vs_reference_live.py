"""CPU, build container only: the oracle against the REFERENCE ITSELF on random inputs.

``tests/test_oracle_golden.py`` pins the oracle with committed outputs of the reference; this file runs the
reference's torch CPU backend (``/root/reference/src``, ``backend="torch"``, ``device="cpu"`` -- the numerical oracle of
SURVEY.md section 8c) live, on seeded random shapes, dtypes and value distributions that no golden file holds, and
compares ``oracle/stainx_oracle.c`` with it at the same bars.  ``/root/reference`` does not exist on the GPU box: the
whole module is skipped there (nothing under ``-m gpu``, ``smoke()`` or ``bench.py`` reads the reference at run time).
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import pytest
import torch

from tests.helpers import he_tile

REFERENCE_SRC = Path("/root/reference/src")
if not (REFERENCE_SRC / "stainx" / "__init__.py").exists():
    pytest.skip("the reference is not present on this machine", allow_module_level=True)
if str(REFERENCE_SRC) not in sys.path:
    sys.path.append(str(REFERENCE_SRC))  # appended: nothing of this repo can be shadowed (the packages are `stainx` / `stainx_b200`)
try:
    from stainx import HistogramMatching, Macenko, Reinhard
except Exception as exc:  # noqa: BLE001
    pytest.skip(f"the reference does not import here: {exc}", allow_module_level=True)


def _rand(rng: np.random.Generator, dtype: str, shape) -> torch.Tensor:
    """Skewed noise: a power curve per channel, sometimes quantised to a few grey levels (sparse histograms)."""
    x = rng.random(shape, dtype=np.float32)
    for c in range(3):
        x[:, c] = x[:, c] ** rng.uniform(0.4, 2.5)
    if rng.random() < 0.25:
        levels = int(rng.integers(3, 40))
        x = np.floor(x * levels) / levels
    if dtype == "u8":
        return torch.from_numpy(np.round(x * 255).astype(np.uint8))
    return torch.from_numpy(x.astype(np.float32))


@pytest.mark.parametrize("seed", range(60))
def test_hm_bit_exact_vs_reference(ox, seed):
    rng = np.random.default_rng(1000 + seed)
    dtype = "u8" if seed % 3 else "f32"
    ref = _rand(rng, dtype, (int(rng.integers(1, 3)), 3, int(rng.integers(5, 70)), int(rng.integers(5, 70))))
    src = _rand(rng, dtype, (int(rng.integers(1, 4)), 3, int(rng.integers(5, 70)), int(rng.integers(5, 70))))
    n = HistogramMatching(device="cpu", backend="torch", channel_axis=1).fit(ref)
    want_hist = torch.stack(n._ref_histograms_256).numpy()
    want = n.transform(src).numpy()
    assert np.array_equal(ox.hm_fit(ref.numpy()), want_hist)
    got = ox.hm_transform(src.numpy(), want_hist)
    assert got.dtype == want.dtype and np.array_equal(got, want)


@pytest.mark.parametrize("seed", range(24))
def test_reinhard_vs_reference(ox, seed):
    rng = np.random.default_rng(2000 + seed)
    dtype = "u8" if seed % 2 else "f32"
    ref = _rand(rng, dtype, (int(rng.integers(1, 3)), 3, int(rng.integers(8, 90)), int(rng.integers(8, 90))))
    src = _rand(rng, dtype, (int(rng.integers(1, 4)), 3, int(rng.integers(8, 90)), int(rng.integers(8, 90))))
    n = Reinhard(device="cpu", backend="torch").fit(ref)
    mean, std = ox.reinhard_fit(ref.numpy())
    assert np.abs(mean - n._reference_mean.numpy().reshape(3)).max() <= 1e-4
    assert np.abs(std - n._reference_std.numpy().reshape(3)).max() <= 1e-4
    want = n.transform(src).numpy()
    got = ox.reinhard_transform(src.numpy(), n._reference_mean.numpy().reshape(3), n._reference_std.numpy().reshape(3))
    diff = np.abs(got.astype(np.float64) - want.astype(np.float64))
    if dtype == "u8":
        assert diff.max() <= 1 and (diff > 0).mean() < 2e-3  # truncation knife edge
    else:
        assert diff.max() <= 1e-4


@pytest.mark.parametrize("seed", range(16))
def test_macenko_vs_reference_on_stain_like_tiles(ox, seed):
    """Beer-Lambert tiles (the reference's own fixture; on noise the comparison is ill-posed, SURVEY.md section 7 H-a)."""
    rng = np.random.default_rng(3000 + seed)
    h, w = int(rng.integers(40, 160)), int(rng.integers(40, 160))
    ref = he_tile(h, w, 42 + seed, 1.0)
    src = torch.cat([he_tile(h, w, 500 + 3 * seed + i, float(rng.uniform(0.85, 1.2))) for i in range(int(rng.integers(1, 4)))])
    if seed % 2:
        ref, src = ref.float() / 255.0, src.float() / 255.0
    n = Macenko(device="cpu", backend="torch").fit(ref)
    he_w, maxc_w = n._stain_matrix.numpy(), n._target_max_conc.numpy().reshape(2)
    he, maxc = ox.macenko_fit(ref.numpy())
    assert np.abs(he - he_w).max() <= 1e-5
    assert np.abs(maxc / maxc_w - 1).max() <= 1e-4
    want = n.transform(src).numpy()
    got = ox.macenko_transform(src.numpy(), he_w, maxc_w)
    diff = np.abs(got.astype(np.float64) - want.astype(np.float64))
    if got.dtype == np.uint8:
        assert diff.max() <= 1 and (diff > 0).mean() < 2e-3
    else:
        assert diff.max() <= 1e-3 * 255
