"""CPU: the image-sharded (multi-GPU) control flow, run with gloo and world_size 2.

The product's sharded pipelines (stainx_b200/backends/torch_cuda_backend.py) are Python around kernel
phases and all-reduces.  Here the kernel layer is replaced by tests/cpu_ops.py (numpy + the CPU
oracle) so that two ranks on this box can check: sharded result == single-device result, all ranks
fit identically, reference-mode broadcast, batch-mode owner selection.
"""
from __future__ import annotations

import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from tests.conftest import golden  # noqa: E402
from tests.helpers import he_tile, noise_f32, noise_u8  # noqa: E402


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, case: str, out_dir: str):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from stainx_b200.backends.torch_cuda_backend import HistogramMatchingCUDA, MacenkoCUDA, ReinhardCUDA
        from stainx_b200.sharding import StatReducer, shard_range
        from tests import cpu_ops

        red = StatReducer.world()
        assert red.enabled and red.world_size == world and red.rank == rank
        result = {}
        if case == "hm":
            ref, src = noise_u8((2, 3, 40, 40), 1, 2.0), noise_u8((5, 3, 40, 40), 2, 0.6)
            lo, hi = shard_range(src.shape[0], rank, world)
            rlo, rhi = shard_range(ref.shape[0], rank, world)
            b = cpu_ops.cpu_backend(HistogramMatchingCUDA, "cpu", reducer=red)
            counts, hists = b.compute_reference_histograms(ref[rlo:rhi])  # pooled fit over the sharded reference
            result["ref_hist"] = torch.stack(hists).numpy()
            result["out"] = b.transform(src[lo:hi], hists).numpy()
            result["range"] = (lo, hi)
        elif case == "reinhard":
            ref, src = noise_f32((2, 3, 32, 32), 3), noise_f32((5, 3, 32, 32), 4, 1.5)
            lo, hi = shard_range(src.shape[0], rank, world)
            rlo, rhi = shard_range(ref.shape[0], rank, world)
            b = cpu_ops.cpu_backend(ReinhardCUDA, "cpu", reducer=red)
            mean, std = b.compute_reference_mean_std(ref[rlo:rhi])
            result["mean"], result["std"] = mean.numpy(), std.numpy()
            result["out"] = b.transform(src[lo:hi], mean, std).numpy()
            result["range"] = (lo, hi)
        elif case == "macenko":
            ref = torch.cat([he_tile(64, 64, 42), he_tile(64, 64, 7, 1.1), he_tile(64, 64, 8, 0.9)])
            rlo, rhi = shard_range(ref.shape[0], rank, world)
            b = cpu_ops.cpu_backend(MacenkoCUDA, "cpu", reducer=red)
            he, maxc = b.compute_reference_stain_matrix(ref[rlo:rhi])
            result["he"], result["maxc"] = he.numpy(), maxc.numpy()
        elif case == "macenko_fit_transform":
            from stainx_b200 import Macenko

            batch = torch.cat([he_tile(48, 48, 42), he_tile(48, 48, 7, 1.1), he_tile(48, 48, 8, 0.9)])
            lo, hi = shard_range(batch.shape[0], rank, world)
            n = Macenko(device="cpu", normalize_to_0_1=True, process_group="world")
            n._backend_impl = cpu_ops.cpu_backend(MacenkoCUDA, "cpu", reducer=n._make_reducer())
            result["out"] = n.fit_transform(batch[lo:hi]).numpy()  # pooled fit over BOTH ranks' tiles, then the own tiles
            result["he"], result["maxc"] = n._stain_matrix.numpy(), n._target_max_conc.numpy()
            result["range"] = (lo, hi)
        elif case == "broadcast":
            from stainx_b200 import Reinhard

            n = Reinhard(device="cpu", process_group="world")
            n._backend_impl = cpu_ops.cpu_backend(ReinhardCUDA, "cpu", reducer=n._make_reducer())
            ref = noise_f32((1, 3, 24, 24), 10 + rank)  # ranks hold DIFFERENT tensors: only src's counts
            n.fit_broadcast(ref if rank == 1 else None, src=1)
            result["mean"], result["std"] = n._reference_mean.numpy(), n._reference_std.numpy()
        elif case == "batch_mode":
            from stainx_b200 import StainNormalizerTransform

            t = StainNormalizerTransform("reinhard", mode="batch", batch_ref_index=3, process_group="world")
            t.normalizer._backend_impl = cpu_ops.cpu_backend(ReinhardCUDA, "cpu", reducer=t.normalizer._make_reducer())
            t._follow_device = lambda device: None  # CPU stand-in: skip the CUDA-only device sync
            full = noise_f32((5, 3, 16, 16), 21)
            lo, hi = shard_range(5, rank, world)  # rank 0: [0,3), rank 1: [3,5) -> global index 3 lives on rank 1
            out = t(full[lo:hi])
            result["mean"] = t.normalizer._reference_mean.numpy()
            result["out"] = out.numpy()
            result["range"] = (lo, hi)
        np.save(os.path.join(out_dir, f"{case}_{rank}.npy"), result, allow_pickle=True)
    finally:
        dist.destroy_process_group()


def _run(case: str, tmp_path, world: int = 2):
    mp.spawn(_worker, args=(world, _free_port(), case, str(tmp_path)), nprocs=world, join=True)
    return [np.load(tmp_path / f"{case}_{r}.npy", allow_pickle=True).item() for r in range(world)]


def test_hm_sharded_equals_single_device(tmp_path, ox):
    res = _run("hm", tmp_path)
    ref, src = noise_u8((2, 3, 40, 40), 1, 2.0), noise_u8((5, 3, 40, 40), 2, 0.6)
    ref_hist = ox.hm_fit(ref.numpy())
    want = ox.hm_transform(src.numpy(), ref_hist)
    for r in res:
        assert np.array_equal(r["ref_hist"], ref_hist)  # pooled fit over both ranks' references
        lo, hi = r["range"]
        assert np.array_equal(r["out"], want[lo:hi])     # batch-global histogram -> bit-exact shards
    assert np.array_equal(res[0]["ref_hist"], res[1]["ref_hist"])


def test_reinhard_sharded_equals_single_device(tmp_path, ox):
    res = _run("reinhard", tmp_path)
    ref, src = noise_f32((2, 3, 32, 32), 3), noise_f32((5, 3, 32, 32), 4, 1.5)
    mean, std = ox.reinhard_fit(ref.numpy())
    want = ox.reinhard_transform(src.numpy(), mean, std)
    for r in res:
        assert np.abs(r["mean"] - mean).max() <= 1e-4 and np.abs(r["std"] - std).max() <= 1e-4
        lo, hi = r["range"]
        assert np.abs(r["out"] - want[lo:hi]).max() <= 1e-4
    assert np.array_equal(res[0]["mean"], res[1]["mean"]) and np.array_equal(res[0]["std"], res[1]["std"])


def test_macenko_pooled_fit_sharded(tmp_path, ox):
    res = _run("macenko", tmp_path)
    ref = torch.cat([he_tile(64, 64, 42), he_tile(64, 64, 7, 1.1), he_tile(64, 64, 8, 0.9)])
    he, maxc = ox.macenko_fit(ref.numpy())
    for r in res:
        assert np.abs(r["he"] - he).max() <= 1e-4
        assert np.abs(r["maxc"] / maxc - 1).max() <= 1e-3
    assert np.array_equal(res[0]["he"], res[1]["he"]) and np.array_equal(res[0]["maxc"], res[1]["maxc"])


def test_macenko_fit_transform_sharded(tmp_path, ox):
    """Normalizer.fit_transform on a sharded batch = pooled fit over every rank's tiles + transform of the own tiles
    (the backend's one-call path; with the CPU stand-in it takes the two-step branch of the same method)."""
    res = _run("macenko_fit_transform", tmp_path)
    batch = torch.cat([he_tile(48, 48, 42), he_tile(48, 48, 7, 1.1), he_tile(48, 48, 8, 0.9)])
    he, maxc = ox.macenko_fit(batch.numpy())
    want = ox.macenko_transform(batch.numpy(), he, maxc) / 255.0
    assert np.array_equal(res[0]["he"], res[1]["he"]) and np.array_equal(res[0]["maxc"], res[1]["maxc"])
    for r in res:
        assert np.abs(r["he"] - he).max() <= 1e-4 and np.abs(r["maxc"] / maxc - 1).max() <= 1e-3
        lo, hi = r["range"]
        assert r["out"].dtype == np.float32 and r["out"].shape[0] == hi - lo
        assert np.abs(r["out"] - want[lo:hi]).max() <= 1.0 / 255.0 + 1e-6  # uint8 in: one grey level (truncation edge)
        assert (np.abs(r["out"] - want[lo:hi]) > 1e-6).mean() < 0.01


def test_reference_fit_is_broadcast_from_src(tmp_path, ox):
    res = _run("broadcast", tmp_path)
    mean, std = ox.reinhard_fit(noise_f32((1, 3, 24, 24), 11).numpy())  # rank 1's tensor
    for r in res:
        assert np.abs(r["mean"] - mean).max() <= 1e-4 and np.abs(r["std"] - std).max() <= 1e-4
    assert np.array_equal(res[0]["mean"], res[1]["mean"])


def test_batch_mode_owner_fits_and_broadcasts(tmp_path, ox):
    res = _run("batch_mode", tmp_path)
    full = noise_f32((5, 3, 16, 16), 21)
    mean, _ = ox.reinhard_fit(full[3:4].numpy())
    for r in res:
        assert np.abs(r["mean"] - mean).max() <= 1e-4
    assert np.array_equal(res[0]["mean"], res[1]["mean"])
    assert res[0]["out"].shape[0] == 3 and res[1]["out"].shape[0] == 2


def test_cpu_restatement_of_the_selection_matches_the_oracle(ox):
    """The two-level 24-bit order-statistic scheme (numpy restatement of the kernel protocol) returns
    the oracle's exact nearest-rank values: HE / maxC agree to float32 rounding."""
    from tests import cpu_ops

    g = golden("macenko_he_u8")
    he, maxc = cpu_ops.macenko_fit(torch.from_numpy(g["ref"]))
    assert np.abs(he.numpy() - g["he"]).max() <= 1e-5
    assert np.abs(maxc.numpy() / g["maxc"] - 1).max() <= 1e-5
    g = golden("macenko_he_pooled_u8")
    he, maxc = cpu_ops.macenko_fit(torch.from_numpy(g["ref"]))
    assert np.abs(he.numpy() - g["he"]).max() <= 1e-5
    assert np.abs(maxc.numpy() / g["maxc"] - 1).max() <= 1e-5
