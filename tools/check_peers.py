#!/usr/bin/env python
"""Sharded HistogramMatching on N GPUs of one node (run under torchrun): the fused NVLink exchange
(sx_hm_build_lut_peers) against the NCCL all-reduce path and against the single-device result of the
whole batch.  Exits non-zero on any mismatch.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_peers.py
"""
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from stainx_b200 import HistogramMatching, Macenko, Reinhard  # noqa: E402
from tests.helpers import he_tile  # noqa: E402
from stainx_b200.sharding import shard_range  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)

g = torch.Generator().manual_seed(7)
ref = (torch.rand(1, 3, 256, 256, generator=g) ** 1.7 * 255).round().to(torch.uint8)
n_total = 4 * world + 1  # uneven shards
whole = (torch.rand(n_total, 3, 256, 256, generator=g) ** 0.6 * 255).round().to(torch.uint8)
lo, hi = shard_range(n_total, rank, world)
mine = whole[lo:hi].to(dev)

hm = HistogramMatching(device=dev, backend="torch_cuda", process_group="world")
hm.fit_broadcast(ref.to(dev), src=0)
impl = hm._get_backend_impl()
ex = impl._peer_exchange()
print(f"rank {rank}: peer exchange {'ON' if ex is not None else 'unavailable (NCCL path)'}", flush=True)
outs = [hm.transform(mine) for _ in range(5)]  # several epochs: both count parities, flag reuse
torch.cuda.synchronize()

# NCCL path of the same sharded batch
impl._exchange = False
want_nccl = hm.transform(mine)
# single-device result of the whole batch
single = HistogramMatching(device=dev, backend="torch_cuda").fit(ref.to(dev))
want_whole = single.transform(whole.to(dev))[lo:hi]
ok = all(torch.equal(o, want_nccl) for o in outs) and torch.equal(want_nccl, want_whole)

# Reinhard: fused finalize over peers against the NCCL path and the single-device whole batch
reff, wholef = ref.float() / 255.0, whole.float() / 255.0
rh = Reinhard(device=dev, backend="torch_cuda", process_group="world")
rh.fit_broadcast(reff.to(dev), src=0)
rimpl = rh._get_backend_impl()
rex = rimpl._peer_exchange()
r_outs = [rh.transform(wholef[lo:hi].to(dev)) for _ in range(4)]
rimpl._exchange = False
r_nccl = rh.transform(wholef[lo:hi].to(dev))
r_single = Reinhard(device=dev, backend="torch_cuda").fit(reff.to(dev)).transform(wholef.to(dev))[lo:hi]
# peers == NCCL exactly (same kernels, same sums); vs the single-device transform of the whole batch 1e-4: a shard below
# 2 MP takes the SFU transfer curves and the whole batch the interpolated tables (reinhard.cu: use_tables), which differ
# by ~1.5e-5 -- both are within 1e-3 of the oracle (tests/test_gpu_parity.py)
r_ok = rex is not None and all(float((o - r_nccl).abs().max()) <= 1e-6 for o in r_outs) and float((r_nccl - r_single).abs().max()) <= 1e-4
print(f"rank {rank}: reinhard peers {'ON' if rex is not None else 'unavailable'} max|peers-nccl|={max(float((o - r_nccl).abs().max()) for o in r_outs):.2e} max|nccl-single|={float((r_nccl - r_single).abs().max()):.2e}", flush=True)
ok = ok and (r_ok or rex is None)

# Macenko pooled fit of a sharded batch: peer combine kernel against NCCL all-reduces and the single-device fit
tiles = torch.cat([he_tile(256, 256, 100 + i, 0.9 + 0.05 * (i % 5)) for i in range(n_total)])
mk = Macenko(device=dev, backend="torch_cuda", process_group="world")
mimpl = mk._get_backend_impl()
mex = mimpl._peer_exchange()
fits = []
for _ in range(3):
    mk.fit(tiles[lo:hi].to(dev))
    fits.append((mk._stain_matrix.clone(), mk._target_max_conc.clone()))
mimpl._exchange = False
mk.fit(tiles[lo:hi].to(dev))
he_nccl, mc_nccl = mk._stain_matrix.clone(), mk._target_max_conc.clone()
single_mk = Macenko(device=dev, backend="torch_cuda").fit(tiles.to(dev))
m_ok = all(torch.equal(h, he_nccl) and torch.equal(m, mc_nccl) for h, m in fits)
m_ok = m_ok and float((he_nccl - single_mk._stain_matrix).abs().max()) <= 1e-6 and float((mc_nccl / single_mk._target_max_conc - 1).abs().max()) <= 1e-6
print(f"rank {rank}: macenko peers {'ON' if mex is not None else 'unavailable'} fit==nccl: {m_ok} |he-single|={float((he_nccl - single_mk._stain_matrix).abs().max()):.2e}", flush=True)
ok = ok and (m_ok or mex is None)
mimpl._exchange = mex if mex is not None else False
bigf = torch.rand(16, 3, 1024, 1024, device=dev)
for label, use_peers in (("peers", True), ("nccl", False)):
    mimpl._exchange = (mex if mex is not None else False) if use_peers else False
    for _ in range(3):
        mk.fit(bigf)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(20):
        mk.fit(bigf)
    torch.cuda.synchronize()
    if rank == 0:
        print(f"macenko pooled fit 16x1024^2 f32 per rank, exchange={label}: {(time.perf_counter() - t0) / 20 * 1e6:.1f} us", flush=True)

# Macenko fit_transform of a sharded batch: ONE library call per rank (sx_macenko_fit_transform_peers: per-image moments
# once, their sum pooled over the ranks, five fused exchanges, transform of the own tiles).  Must equal fit() then
# transform() on the same shard bit for bit, and the rank's slice of the single-device fit_transform of the whole batch.
ft = Macenko(device=dev, backend="torch_cuda", normalize_to_0_1=True, process_group="world")
ft._backend_impl = mimpl  # share the exchange
mimpl._exchange = mex if mex is not None else False
shard = tiles[lo:hi].to(dev)
got_ft = ft.fit_transform(shard)
two_ft = Macenko(device=dev, backend="torch_cuda", normalize_to_0_1=True, process_group="world")
two_ft._backend_impl = mimpl
want_two = two_ft.fit(shard).transform(shard)
single_ft = Macenko(device=dev, backend="torch_cuda", normalize_to_0_1=True)
want_single = single_ft.fit_transform(tiles.to(dev))[lo:hi]
d_single = float((got_ft - want_single).abs().max()) if hi > lo else 0.0
ft_ok = torch.equal(got_ft, want_two) and torch.equal(ft._stain_matrix, two_ft._stain_matrix) and torch.equal(ft._target_max_conc, two_ft._target_max_conc) and d_single <= 1e-5
print(f"rank {rank}: macenko fit_transform (one call) == fit + transform: {torch.equal(got_ft, want_two)}; max|sharded-single| = {d_single:.2e}", flush=True)
ok = ok and (ft_ok or mex is None)
for label, fn in (("fit_transform (one call, shared moments pass)", lambda: ft.fit_transform(bigf)), ("fit + transform (two calls)", lambda: two_ft.fit(bigf).transform(bigf))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    if rank == 0:
        print(f"macenko 16x1024^2 f32 per rank, {label}: {(time.perf_counter() - t0) / 20 * 1e6:.1f} us", flush=True)
del bigf

# StainNormalizerTransform(mode="batch") on a sharded batch: the rank that owns global image `batch_ref_index` fits,
# the parameters are broadcast; every rank's result must equal the single-device module's on the gathered batch
from stainx_b200 import StainNormalizerTransform  # noqa: E402

b_ok = True
for method, tol in (("histogram_matching", 0.0), ("reinhard", 1e-6), ("macenko", 1e-6)):
    gidx = n_total - 2  # lives on the last rank
    tm = StainNormalizerTransform(method=method, mode="batch", device=dev, batch_ref_index=gidx, process_group="world")
    got = tm(tiles[lo:hi].to(dev))
    want = StainNormalizerTransform(method=method, mode="batch", device=dev, batch_ref_index=gidx)(tiles.to(dev))[lo:hi]
    diff = (got.float() - want.float()).abs()
    d = float(diff.max())
    if method == "macenko":
        # per-image statistics are fixed-point sums over row blocks aligned within the image: a tile's output does not
        # depend on the batch (shard) it travels in -- bit for bit
        frac = float((diff > 0).float().mean())
        b_ok = b_ok and d == 0.0
        print(f"rank {rank}: batch-mode module {method}: max|sharded-single| = {d:.2e} on {frac:.1e} of the values", flush=True)
        continue
    b_ok = b_ok and d <= tol
    print(f"rank {rank}: batch-mode module {method}: max|sharded-single| = {d:.2e}", flush=True)
ok = ok and b_ok

# timing of the exchange + LUT phase
impl._exchange = ex if ex is not None else False
big = (torch.rand(16, 3, 1024, 1024, device=dev) * 255).to(torch.uint8)
for label, use_peers in (("peers", True), ("nccl", False)):
    impl._exchange = (ex if ex is not None else False) if use_peers else False
    for _ in range(5):
        hm.transform(big)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(50):
        hm.transform(big)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 50 * 1e6
    if rank == 0:
        print(f"transform 16x1024^2 per rank, exchange={label}: {dt:.1f} us per step", flush=True)
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("PEERS CHECK", "OK" if int(flag.item()) else "MISMATCH", flush=True)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) else 1)
