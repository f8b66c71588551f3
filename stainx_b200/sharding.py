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


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous image range [lo, hi) of ``rank`` when ``n`` images are split over ``world`` ranks."""
    base, extra = divmod(int(n), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)
