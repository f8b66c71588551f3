"""Host ingest for whole-slide streaming: pinned host batches in, pinned
host results out, with the host->device copy of batch i+1 and the device->host copy of result i-1
overlapping the kernels of batch i.

The reference has no counterpart: its DataLoader example moves every batch with a blocking
``.to(device)`` before the transform (``examples/torch_transform_example.py:L43-64``), so PCIe
transfers and kernels serialise.  On a B200 the normalisation kernels take ~0.2-1 ms per 64 MP
batch while each direction of a PCIe Gen5 x16 link needs ~3.7 ms for the same uint8 batch: the link
is the whole-job bottleneck, and keeping BOTH directions busy at once doubles end-to-end throughput.

    stream = HostStream(normalizer, depth=2)           # any fitted normalizer / StainNormalizerTransform
    tickets = [stream.submit(batch, out) for batch, out in zip(pinned_batches, pinned_outputs)]
    for t in tickets:
        t.wait()                                       # out is complete

Three CUDA streams (copy-in, compute, copy-out) are chained per batch with events; ``depth`` device
staging buffers are recycled, each guarded by the event of its last reader.  Host code only
(torch streams and events); the pixel work is the normalizer's own ``transform``.
"""
from __future__ import annotations

from typing import Any, Callable

import torch

__all__ = ["DeviceStream", "HostStream", "Ticket", "bind_host_thread_to_device"]


def bind_host_thread_to_device(device: torch.device | int | None = None) -> list[int] | None:
    """Restrict the calling process to the CPU cores NVML reports as local to `device`'s PCIe root (its NUMA node),
    so that pinned host buffers allocated AFTERWARDS are first-touched on that node: a pinned buffer on the
    far socket crosses the inter-socket link on every copy and caps host<->device bandwidth.  Call it once per
    rank before allocating pinned memory.  Returns the core list, or None when NVML / the OS call is unavailable
    (nothing is changed then)."""
    import os

    try:
        import pynvml

        idx = torch.cuda.current_device() if device is None else (device if isinstance(device, int) else (torch.device(device).index or 0))
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        if visible:
            ids = [v.strip() for v in visible.split(",") if v.strip()]
            if idx < len(ids) and ids[idx].isdigit():
                idx = int(ids[idx])
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (ncpu + 63) // 64)
        cores = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1 and 64 * w + b < ncpu]
        allowed = sorted(set(cores) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:  # noqa: BLE001 - best effort: no NVML, no sched_setaffinity, restricted container
        return None


class Ticket:
    """Completion handle of one submitted batch."""

    def __init__(self, done: torch.cuda.Event, out: torch.Tensor):
        self._done = done
        self.out = out

    def ready(self) -> bool:
        return self._done.query()

    def wait(self) -> torch.Tensor:
        self._done.synchronize()
        return self.out


class HostStream:
    """Pipelined ``transform`` of host-resident batches.

    ``normalizer`` is anything with ``transform(images)`` or ``__call__`` that maps a device batch
    to a device batch (``Reinhard`` / ``Macenko`` / ``HistogramMatching`` after ``fit``, or a
    ``StainNormalizerTransform``).  Host tensors should be pinned; pageable memory still works but
    makes the copies synchronous (torch semantics)."""

    def __init__(self, normalizer: Any, device: torch.device | str | None = None, depth: int = 2):
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self._fn: Callable[[torch.Tensor], torch.Tensor] = getattr(normalizer, "transform", None) or normalizer
        dev = device if device is not None else getattr(normalizer, "device", None)
        self.device = torch.device(dev if dev is not None else "cuda")
        if self.device.type != "cuda":
            raise ValueError("HostStream requires a CUDA device")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self._s_in = torch.cuda.Stream(self.device)
        self._s_compute = torch.cuda.Stream(self.device)
        self._s_out = torch.cuda.Stream(self.device)
        self._slots: list[dict] = [{"buf": None, "free": None} for _ in range(depth)]
        self._inflight: list[torch.cuda.Event] = []  # completion events of the batches enqueued and not yet waited for
        self._max_inflight = depth + 1
        self._next = 0
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def submit(self, host_in: torch.Tensor, host_out: torch.Tensor | None = None) -> Ticket:
        """Enqueue one batch.  ``host_out`` (pinned, shape/dtype of the result) receives the result;
        when omitted a pinned tensor is allocated once the result's shape is known."""
        if host_in.device.type != "cpu":
            raise ValueError("HostStream.submit expects a host tensor")
        slot = self._slots[self._next % len(self._slots)]
        self._next += 1
        # Back-pressure: never more than depth + 1 batches enqueued.  Without it a producer that submits faster than
        # the link drains piles up results whose memory the caching allocator cannot recycle (each is held by a
        # pending copy-out): every further batch then costs a cudaMalloc, which synchronises the device (measured:
        # 14.8 ms instead of 1.3 ms per 31 MB batch with 32 batches enqueued at once).
        while len(self._inflight) >= self._max_inflight:
            self._inflight.pop(0).synchronize()
        with torch.cuda.device(self.device):
            # Whatever the caller enqueued on its current stream before this submit -- above all the
            # kernels of fit() / fit_reference(), whose fitted tensors the transform below reads -- must
            # have finished before this batch's kernels start: the compute stream is non-blocking and
            # shares no implicit ordering with it.
            self._s_compute.wait_stream(torch.cuda.current_stream(self.device))
            # ---- host -> device, into a recycled staging buffer
            with torch.cuda.stream(self._s_in):
                if slot["free"] is not None:
                    self._s_in.wait_event(slot["free"])  # the kernels that read this buffer last have finished
                buf = slot["buf"]
                if buf is None or buf.shape != host_in.shape or buf.dtype != host_in.dtype:
                    buf = torch.empty(host_in.shape, dtype=host_in.dtype, device=self.device)
                    buf.record_stream(self._s_compute)
                    slot["buf"] = buf
                buf.copy_(host_in, non_blocking=True)
                copied = torch.cuda.Event()
                copied.record(self._s_in)
            # ---- kernels
            with torch.cuda.stream(self._s_compute):
                self._s_compute.wait_event(copied)
                out = self._fn(buf)
                computed = torch.cuda.Event()
                computed.record(self._s_compute)
                slot["free"] = computed
            # ---- device -> host
            with torch.cuda.stream(self._s_out):
                self._s_out.wait_event(computed)
                if host_out is None:
                    host_out = torch.empty(out.shape, dtype=out.dtype).pin_memory()
                host_out.copy_(out, non_blocking=True)
                out.record_stream(self._s_out)
                done = torch.cuda.Event()
                done.record(self._s_out)
        self._inflight.append(done)
        self.h2d_bytes += host_in.numel() * host_in.element_size()
        self.d2h_bytes += host_out.numel() * host_out.element_size()
        return Ticket(done, host_out)

    def map(self, batches, outputs=None):
        """Generator over results in submission order, keeping ``depth`` batches in flight."""
        pending: list[Ticket] = []
        outs = iter(outputs) if outputs is not None else None
        for b in batches:
            pending.append(self.submit(b, next(outs) if outs is not None else None))
            if len(pending) > len(self._slots):
                yield pending.pop(0).wait()
        for t in pending:
            yield t.wait()

    def synchronize(self) -> None:
        for s in (self._s_in, self._s_compute, self._s_out):
            s.synchronize()


class DeviceStream:
    """Independent DEVICE-resident batches, issued round-robin on a small pool of CUDA streams.

    One ``transform`` is a chain of dependent phases that stress different parts of the SM -- histogram matching counts
    with the shared-memory atomic unit and then remaps at HBM speed, Reinhard's statistics pass is SFU / issue bound and
    its second pass HBM bound -- so two transforms of DIFFERENT batches overlap: measured on a B200 (64 x 3 x 1024 x 1024),
    HistogramMatching uint8 127.9 -> 115.3 us per batch with three streams, Reinhard float32 460.7 -> 414.3 us with two,
    Macenko float32 785 -> 774 us (its transform already runs three chains of its own).  The batches must not depend on
    each other; every call is ordered behind the caller's current stream, and ``collect`` (or waiting on the returned
    event) orders the caller behind the results.

        pool = DeviceStream(norm, streams=3)
        outs = pool.map(device_batches)          # list of results, in order, ready for the current stream
    """

    def __init__(self, normalizer: Any, device: torch.device | str | None = None, streams: int = 3):
        if streams < 1:
            raise ValueError("streams must be >= 1")
        self._fn: Callable[[torch.Tensor], torch.Tensor] = getattr(normalizer, "transform", None) or normalizer
        dev = device if device is not None else getattr(normalizer, "device", None)
        self.device = torch.device(dev if dev is not None else "cuda")
        if self.device.type != "cuda":
            raise ValueError("DeviceStream requires a CUDA device")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self._streams = [torch.cuda.Stream(self.device) for _ in range(streams)]
        self._next = 0

    def submit(self, batch: torch.Tensor) -> tuple[torch.Tensor, torch.cuda.Event]:
        """Enqueue one batch on the next stream of the pool; returns (result, event recorded behind it).  The result
        belongs to the pool stream: wait for the event (``event.wait()`` on the consuming stream, or ``join``) before
        using it elsewhere.  Dropping the result frees its memory for the next batch on the same stream."""
        s = self._streams[self._next % len(self._streams)]
        self._next += 1
        with torch.cuda.device(self.device):
            s.wait_stream(torch.cuda.current_stream(self.device))  # inputs (and fitted parameters) produced on the caller's stream
            with torch.cuda.stream(s):
                out = self._fn(batch)
                batch.record_stream(s)
                done = torch.cuda.Event()
                done.record(s)
        return out, done

    def join(self) -> None:
        """The caller's current stream waits for everything submitted so far."""
        cur = torch.cuda.current_stream(self.device)
        for s in self._streams:
            cur.wait_stream(s)

    def map(self, batches) -> list[torch.Tensor]:
        """Results of ``batches`` in order, all alive at once (mind the memory), ready for the current stream."""
        cur = torch.cuda.current_stream(self.device)
        outs = []
        for b in batches:
            out, _ = self.submit(b)
            out.record_stream(cur)
            outs.append(out)
        self.join()
        return outs
