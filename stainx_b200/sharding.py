"""Image-sharded execution: combine fit / source statistics across ranks.

The reference is single-process (no distributed code).  Batches shard
naturally by image, one process per GPU; what must be exchanged is tiny:

===========================  ===============================================  ==========
path                         payload                                          reduce op
===========================  ===============================================  ==========
HistogramMatching hist       int64 (3, 256) counts + pixel count              SUM
Reinhard stats               float64 (8,) shifted sums + count                SUM
Macenko pooled fit           moments / OD range / level-0 and level-1 hists    SUM/MAX/MIN
reference-mode fit           fitted parameters                                broadcast
===========================  ===============================================  ==========

``StatReducer`` wraps ``torch.distributed`` (NCCL over NVLink on the GPU box, gloo in the CPU
tests).  With no process group it is the identity, so the single-device path never touches
``torch.distributed``.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class StatReducer:
    """All-reduce helper bound to a process group (``None`` = single device, no-op)."""

    def __init__(self, group: dist.ProcessGroup | None = None, enabled: bool | None = None):
        self.group = group
        if enabled is None:
            enabled = group is not None
        self.enabled = bool(enabled) and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1

    @classmethod
    def world(cls) -> "StatReducer":
        """Reducer over the default process group when one is initialised."""
        if dist.is_available() and dist.is_initialized():
            return cls(group=None, enabled=True)
        return cls()

    @property
    def world_size(self) -> int:
        return dist.get_world_size(self.group) if self.enabled else 1

    @property
    def rank(self) -> int:
        return dist.get_rank(self.group) if self.enabled else 0

    def _reduce(self, t: torch.Tensor, op) -> torch.Tensor:
        if self.enabled:
            dist.all_reduce(t, op=op, group=self.group)
        return t

    def sum_(self, t: torch.Tensor) -> torch.Tensor:
        return self._reduce(t, dist.ReduceOp.SUM)

    def max_(self, t: torch.Tensor) -> torch.Tensor:
        return self._reduce(t, dist.ReduceOp.MAX)

    def min_(self, t: torch.Tensor) -> torch.Tensor:
        return self._reduce(t, dist.ReduceOp.MIN)

    def sum_int(self, value: int, device: torch.device) -> int:
        """Sum of a python int over ranks (host round trip; used for pixel counts only)."""
        if not self.enabled:
            return int(value)
        t = torch.tensor([int(value)], dtype=torch.int64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return int(t.item())

    def broadcast_(self, tensors: list[torch.Tensor], src: int = 0) -> list[torch.Tensor]:
        if self.enabled:
            for t in tensors:
                dist.broadcast(t, src=src, group=self.group)
        return tensors


class PeerExchange:
    """A small buffer that every rank of the group has mapped over NVLink (torch symmetric memory),
    for exchanges fused into kernels (``sx_hm_build_lut_peers``: the all-reduce of the histogram
    counts happens inside the LUT kernel with peer loads, no NCCL call).

    ``PeerExchange.create`` returns ``None`` when symmetric memory is unavailable (gloo groups, ranks
    on different nodes, old torch): callers then fall back to ``StatReducer``'s NCCL all-reduce.  All
    ranks must create it and step ``epoch`` in lockstep (collective semantics)."""

    def __init__(self, buf: torch.Tensor, handle, group):
        self.buf = buf
        self.handle = handle
        self.group = group
        self.rank = int(handle.rank)
        self.world = int(handle.world_size)
        self.ptrs_dev = int(handle.buffer_ptrs_dev)  # device array of `world` buffer pointers
        self.epoch = 0

    @classmethod
    def create(cls, reducer: StatReducer, device: torch.device, nbytes: int) -> "PeerExchange | None":
        if not reducer.enabled or torch.device(device).type != "cuda" or dist.get_backend(reducer.group) != "nccl":
            return None
        group = reducer.group if reducer.group is not None else dist.group.WORLD

        def agree(ok: bool) -> bool:  # all ranks or none (one small all-reduce)
            flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=reducer.group)
            return bool(int(flag.item()))

        # Step 1 is purely local (allocation): a failure here must not leave the other ranks inside a
        # collective, so the ranks agree on it BEFORE anyone enters the rendezvous.
        buf = handle = None
        try:
            import torch.distributed._symmetric_memory as symm_mem

            buf = symm_mem.empty(int(nbytes), dtype=torch.uint8, device=device)
            buf.zero_()
            local_ok = True
        except Exception:  # noqa: BLE001 - any failure means "no peer memory on this rank"
            local_ok = False
        if not agree(local_ok):
            return None
        # Step 2 is collective: every rank enters it, and they agree on its outcome afterwards.
        try:
            handle = symm_mem.rendezvous(buf, group)
            torch.cuda.synchronize(device)
            mapped = handle is not None
        except Exception:  # noqa: BLE001
            mapped = False
        if not agree(mapped):
            return None
        dist.barrier(group=reducer.group)  # every rank's zeros are in place before anyone signals
        return cls(buf, handle, reducer.group)

    def view(self, offset: int, shape: tuple[int, ...], dtype: torch.dtype) -> torch.Tensor:
        n = 1
        for d in shape:
            n *= d
        size = n * torch.empty((), dtype=dtype).element_size()
        return self.buf[offset : offset + size].view(dtype).view(*shape)


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous image range [lo, hi) of ``rank`` when ``n`` images are split over ``world`` ranks."""
    base, extra = divmod(int(n), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)
