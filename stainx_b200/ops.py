"""Tensor-level wrappers of the C ABI: validate, allocate with torch, pass raw pointers.

This is the layer the reference implements in C++ with ATen (``src/stainx_cuda_torch/csrc/*.cu``):
input checks, output allocation, current-stream lookup.  Here torch only provides device memory
and the stream; every byte of pixel work happens inside ``libstainx_b200.so``.

All functions require CUDA tensors and raise otherwise -- there is no CPU path.
"""
from __future__ import annotations

import ctypes

import torch

from stainx_b200 import _native as nv
from stainx_b200._native import SX_BF16, SX_F16, SX_F32, SX_NCHW, SX_NHWC, SX_U8, check

__all__ = [
    "MacenkoWorkspace",
    "hm_apply", "hm_build_lut", "hm_build_lut_peers", "hm_fit", "hm_hist", "hm_ref_cdf", "hm_ref_hist", "hm_transform", "hm_transform_peers",
    "macenko_fit", "macenko_fit_peers", "macenko_peer_combine", "macenko_transform",
    "reinhard_apply", "reinhard_finalize", "reinhard_finalize_peers", "reinhard_fit", "reinhard_stats", "reinhard_transform",
]


def _ptr(t: torch.Tensor) -> ctypes.c_void_p:
    return ctypes.c_void_p(t.data_ptr())


def _stream(device: torch.device) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.uint8:
        return SX_U8
    if t.dtype == torch.float32:
        return SX_F32
    if t.dtype == torch.float16:
        return SX_F16
    if t.dtype == torch.bfloat16:
        return SX_BF16
    raise TypeError(f"native kernels take uint8, float32, float16 or bfloat16 images, got {t.dtype}")


def _check_images(images: torch.Tensor, layout: int = SX_NCHW) -> tuple[int, int, int]:
    """Reference checks: CUDA, 4-D, 3 channels (src/stainx_cuda_torch/csrc/macenko.cu:L68-75)."""
    if not isinstance(images, torch.Tensor):
        raise TypeError(f"images must be a torch.Tensor, got {type(images)}")
    if not images.is_cuda:
        raise RuntimeError("input_images must be a CUDA tensor")
    if images.dim() != 4:
        raise RuntimeError(f"input_images must be 4D (N, C, H, W), got {images.dim()}D")
    if not images.is_contiguous():
        raise RuntimeError("input_images must be contiguous")
    if layout == SX_NHWC:
        n, h, w, c = images.shape
    else:
        n, c, h, w = images.shape
    if c != 3:
        raise RuntimeError(f"input_images must have 3 channels, got {c}")
    return int(n), int(h), int(w)


def _param(t: torch.Tensor, device: torch.device, shape: tuple[int, ...], name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor, got {type(t)}")
    t = t.to(device=device, dtype=torch.float32).contiguous()
    if tuple(t.shape) != shape:
        if t.numel() == int(torch.tensor(shape).prod()):
            t = t.reshape(shape)
        else:
            raise ValueError(f"{name} must have shape {shape}, got {tuple(t.shape)}")
    return t


# ----------------------------------------------------------------------------- histogram matching
def hm_hist(images: torch.Tensor, layout: int = SX_NCHW, counts: torch.Tensor | None = None) -> torch.Tensor:
    """Per-channel 256-bin counts of the whole batch, added into ``counts`` (int64 (3, 256))."""
    n, h, w = _check_images(images, layout)
    dev = images.device
    if counts is None:
        counts = torch.zeros((3, 256), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        check(nv.lib().sx_hm_hist(_ptr(images), _dtype_code(images), layout, n, h, w, _ptr(counts), _stream(dev)), "sx_hm_hist")
    return counts


def hm_ref_hist(counts: torch.Tensor) -> torch.Tensor:
    dev = counts.device
    out = torch.empty((3, 256), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(nv.lib().sx_hm_ref_hist(_ptr(counts), _ptr(out), _stream(dev)), "sx_hm_ref_hist")
    return out


def hm_ref_cdf(ref_hist: torch.Tensor) -> torch.Tensor:
    dev = ref_hist.device
    ref_hist = _param(ref_hist, dev, (3, 256), "ref_hist")
    out = torch.empty((3, 256), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(nv.lib().sx_hm_ref_cdf(_ptr(ref_hist), _ptr(out), _stream(dev)), "sx_hm_ref_cdf")
    return out


def hm_build_lut(counts: torch.Tensor, npix: int, ref_cdf: torch.Tensor) -> torch.Tensor:
    dev = counts.device
    out = torch.empty((3, 256), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(nv.lib().sx_hm_build_lut(_ptr(counts), int(npix), _ptr(ref_cdf), _ptr(out), _stream(dev)), "sx_hm_build_lut")
    return out


def hm_build_lut_peers(exchange, ref_cdf: torch.Tensor, counts_out: torch.Tensor | None = None) -> torch.Tensor:
    """LUT of a sharded batch with the all-reduce of the counts fused into the kernel (NVLink peer
    loads; ``exchange`` is a ``sharding.PeerExchange`` whose epoch the caller has advanced and whose
    counts[epoch & 1] hold this rank's histogram)."""
    dev = exchange.buf.device
    out = torch.empty((3, 256), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(nv.lib().sx_hm_build_lut_peers(ctypes.c_void_p(exchange.ptrs_dev), exchange.world, exchange.rank, exchange.epoch & 0xFFFFFFFF, _ptr(ref_cdf),
                                             _ptr(out), _ptr(counts_out) if counts_out is not None else None, _stream(dev)), "sx_hm_build_lut_peers")
    return out


def hm_transform_peers(images: torch.Tensor, exchange, ref_cdf: torch.Tensor, layout: int = SX_NCHW) -> torch.Tensor:
    """The sharded transform in one library call: advances the exchange's epoch, then zero / histogram /
    LUT with the all-reduce over NVLink peer memory / remap as one chain of dependent launches."""
    n, h, w = _check_images(images, layout)
    dev = images.device
    out = torch.empty_like(images)
    ws = torch.empty(768 * 4, dtype=torch.uint8, device=dev)
    epoch = exchange.epoch + 1  # committed only when the launch went through: a rank that raises here must not run ahead of its peers
    with torch.cuda.device(dev):
        check(nv.lib().sx_hm_transform_peers(_ptr(images), _dtype_code(images), layout, n, h, w, ctypes.c_void_p(exchange.ptrs_dev), _ptr(exchange.buf), exchange.world, exchange.rank,
                                             epoch & 0xFFFFFFFF, _ptr(ref_cdf), _ptr(out), _ptr(ws), ws.numel(), _stream(dev)), "sx_hm_transform_peers")
    exchange.epoch = epoch
    return out


def hm_apply(images: torch.Tensor, lut: torch.Tensor, layout: int = SX_NCHW) -> torch.Tensor:
    n, h, w = _check_images(images, layout)
    dev = images.device
    out = torch.empty_like(images)
    with torch.cuda.device(dev):
        check(nv.lib().sx_hm_apply(_ptr(images), _dtype_code(images), layout, n, h, w, _ptr(lut), _ptr(out), _stream(dev)), "sx_hm_apply")
    return out


def hm_transform(images: torch.Tensor, ref_hist: torch.Tensor, layout: int = SX_NCHW) -> torch.Tensor:
    """hist -> ref_cdf -> LUT -> remap on the current stream (single device)."""
    n, h, w = _check_images(images, layout)
    dev = images.device
    ref_hist = _param(ref_hist, dev, (3, 256), "ref_hist")
    out = torch.empty_like(images)
    nbytes = int(nv.lib().sx_hm_workspace_bytes())
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(nv.lib().sx_hm_transform(_ptr(images), _dtype_code(images), layout, n, h, w, _ptr(ref_hist), _ptr(out), _ptr(ws), nbytes, _stream(dev)), "sx_hm_transform")
    return out


def hm_fit(images: torch.Tensor, layout: int = SX_NCHW) -> torch.Tensor:
    n, h, w = _check_images(images, layout)
    dev = images.device
    out = torch.empty((3, 256), dtype=torch.float32, device=dev)
    nbytes = int(nv.lib().sx_hm_workspace_bytes())
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(nv.lib().sx_hm_fit(_ptr(images), _dtype_code(images), layout, n, h, w, _ptr(out), _ptr(ws), nbytes, _stream(dev)), "sx_hm_fit")
    return out


# ----------------------------------------------------------------------------- reinhard
def reinhard_stats(images: torch.Tensor, sums: torch.Tensor | None = None) -> torch.Tensor:
    """Adds the batch's LAB sums into ``sums`` (float64 (8,): 3 sums, 3 sums of squares, count, pad)."""
    n, h, w = _check_images(images)
    dev = images.device
    if sums is None:
        sums = torch.zeros(8, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        check(nv.lib().sx_reinhard_stats(_ptr(images), _dtype_code(images), n, h, w, _ptr(sums), _stream(dev)), "sx_reinhard_stats")
    return sums


def reinhard_finalize(sums: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    dev = sums.device
    mean = torch.empty(3, dtype=torch.float32, device=dev)
    std = torch.empty(3, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(nv.lib().sx_reinhard_finalize(_ptr(sums), _ptr(mean), _ptr(std), _stream(dev)), "sx_reinhard_finalize")
    return mean, std


def reinhard_finalize_peers(exchange, epoch: int | None = None) -> tuple[torch.Tensor, torch.Tensor]:
    """mean / std of a sharded batch with the all-reduce of the sums fused into the kernel (NVLink peer
    loads; ``exchange``: a ``sharding.PeerExchange`` whose sums[epoch & 1] hold this rank's statistics;
    ``epoch`` defaults to the exchange's current one)."""
    dev = exchange.buf.device
    epoch = exchange.epoch if epoch is None else int(epoch)
    mean = torch.empty(3, dtype=torch.float32, device=dev)
    std = torch.empty(3, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(nv.lib().sx_reinhard_finalize_peers(ctypes.c_void_p(exchange.ptrs_dev), exchange.world, exchange.rank, epoch & 0xFFFFFFFF, _ptr(mean), _ptr(std), _stream(dev)), "sx_reinhard_finalize_peers")
    return mean, std


def reinhard_apply(images: torch.Tensor, src_mean: torch.Tensor, src_std: torch.Tensor, ref_mean: torch.Tensor, ref_std: torch.Tensor) -> torch.Tensor:
    n, h, w = _check_images(images)
    dev = images.device
    src_mean, src_std = _param(src_mean, dev, (3,), "src_mean"), _param(src_std, dev, (3,), "src_std")
    ref_mean, ref_std = _param(ref_mean, dev, (3,), "target_mean"), _param(ref_std, dev, (3,), "target_std")
    out = torch.empty_like(images)
    with torch.cuda.device(dev):
        check(nv.lib().sx_reinhard_apply(_ptr(images), _dtype_code(images), n, h, w, _ptr(src_mean), _ptr(src_std), _ptr(ref_mean), _ptr(ref_std), _ptr(out), _stream(dev)), "sx_reinhard_apply")
    return out


def reinhard_transform(images: torch.Tensor, ref_mean: torch.Tensor, ref_std: torch.Tensor) -> torch.Tensor:
    n, h, w = _check_images(images)
    dev = images.device
    ref_mean, ref_std = _param(ref_mean, dev, (3,), "target_mean"), _param(ref_std, dev, (3,), "target_std")
    out = torch.empty_like(images)
    nbytes = int(nv.lib().sx_reinhard_workspace_bytes())
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(nv.lib().sx_reinhard_transform(_ptr(images), _dtype_code(images), n, h, w, _ptr(ref_mean), _ptr(ref_std), _ptr(out), _ptr(ws), nbytes, _stream(dev)), "sx_reinhard_transform")
    return out


def reinhard_fit(images: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    n, h, w = _check_images(images)
    dev = images.device
    mean = torch.empty(3, dtype=torch.float32, device=dev)
    std = torch.empty(3, dtype=torch.float32, device=dev)
    nbytes = int(nv.lib().sx_reinhard_workspace_bytes())
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(nv.lib().sx_reinhard_fit(_ptr(images), _dtype_code(images), n, h, w, _ptr(mean), _ptr(std), _ptr(ws), nbytes, _stream(dev)), "sx_reinhard_fit")
    return mean, std


# ----------------------------------------------------------------------------- macenko
_REGION_DTYPES = {"moments": torch.int64, "odrange": torch.float32, "hist1": torch.int32, "hist2": torch.int32, "vmin": torch.float32, "vmax": torch.float32, "fit": torch.float32, "counters": torch.int64, "status": torch.int32}
_REGION_SHAPES = {"moments": (12,), "odrange": (8,), "hist1": (2, 4096), "hist2": (2, 4096), "vmin": (2, 4096), "vmax": (2, 4096), "fit": (8,), "counters": (8,), "status": (4,)}


class MacenkoWorkspace:
    """Device scratch of the Macenko phases for ``slots`` statistic slots, plus typed views of the
    regions a sharded run all-reduces (``include/stainx_b200.h``: ``sx_macenko_region``)."""

    def __init__(self, slots: int, device: torch.device, buffer: torch.Tensor | None = None):
        self.slots = int(slots)
        self.device = torch.device(device)
        self.nbytes = int(nv.lib().sx_macenko_workspace_bytes(self.slots))
        # `buffer`: caller-provided storage (e.g. NVLink peer-mapped memory for the sharded fit)
        self.buffer = torch.empty(self.nbytes, dtype=torch.uint8, device=self.device) if buffer is None else buffer[: self.nbytes]
        self._views: dict[str, torch.Tensor] = {}

    def region(self, name: str) -> torch.Tensor:
        if name not in self._views:
            off, size = ctypes.c_int64(), ctypes.c_int64()
            check(nv.lib().sx_macenko_region(self.slots, nv.REGIONS[name], ctypes.byref(off), ctypes.byref(size)), "sx_macenko_region")
            raw = self.buffer[off.value : off.value + size.value]
            self._views[name] = raw.view(_REGION_DTYPES[name]).view(self.slots, *_REGION_SHAPES[name])
        return self._views[name]

    # -- phases (each enqueues on the current stream of the workspace device) --
    def _call(self, fn_name: str, *args) -> None:
        with torch.cuda.device(self.device):
            check(getattr(nv.lib(), fn_name)(*args, _stream(self.device)), fn_name)

    def begin(self) -> None:
        self._call("sx_macenko_begin", _ptr(self.buffer), self.slots)

    def moments(self, images: torch.Tensor, pooled: bool, slot0: int = 0) -> None:
        n, h, w = _check_images(images)
        self._call("sx_macenko_moments", _ptr(images), _dtype_code(images), n, h, w, int(pooled), slot0, _ptr(self.buffer), self.slots)

    def basis(self, slot0: int, count: int, allow_fallback: bool) -> None:
        self._call("sx_macenko_basis", _ptr(self.buffer), self.slots, slot0, count, int(allow_fallback))

    def moments_fallback(self, images: torch.Tensor, slot0: int = 0) -> None:
        n, h, w = _check_images(images)
        self._call("sx_macenko_moments_fallback", _ptr(images), _dtype_code(images), n, h, w, slot0, _ptr(self.buffer), self.slots)

    def hist(self, images: torch.Tensor, pooled: bool, stage: int, level: int, slot0: int = 0) -> None:
        n, h, w = _check_images(images)
        self._call("sx_macenko_hist", _ptr(images), _dtype_code(images), n, h, w, int(pooled), slot0, stage, level, _ptr(self.buffer), self.slots)

    def select(self, slot0: int, count: int, stage: int, level: int) -> None:
        self._call("sx_macenko_select", _ptr(self.buffer), self.slots, slot0, count, stage, level)

    def apply(self, images: torch.Tensor, he_ref: torch.Tensor, maxc_ref: torch.Tensor, out: torch.Tensor, unit: bool, slot0: int = 0) -> None:
        n, h, w = _check_images(images)
        scale = 1.0 / 255.0 if unit else 1.0
        self._call("sx_macenko_apply", _ptr(images), _dtype_code(images), n, h, w, slot0, _ptr(he_ref), _ptr(maxc_ref), _ptr(out), _dtype_code(out), ctypes.c_float(scale), _ptr(self.buffer), self.slots)


def macenko_peer_combine(exchange, which: int, scratch: torch.Tensor) -> None:
    """Combine slot 0's statistics of every rank in place, in one kernel over NVLink peer memory
    (``which``: 0 after moments, 1 after a sample pass, 2 after a resolve pass).  Advances the epoch."""
    epoch = exchange.epoch + 1
    dev = exchange.buf.device
    with torch.cuda.device(dev):
        check(nv.lib().sx_macenko_peer_combine(ctypes.c_void_p(exchange.ptrs_dev), exchange.world, exchange.rank, epoch & 0xFFFFFFFF, int(which), _ptr(scratch), _stream(dev)), "sx_macenko_peer_combine")
    exchange.epoch = epoch


def macenko_fit_peers(images: torch.Tensor, exchange, scratch: torch.Tensor, exact: bool = False) -> tuple[torch.Tensor, torch.Tensor]:
    """The sharded pooled fit of one NVLink node in one library call (five fused exchanges; advances the exchange's
    epoch by 5).  ``images``: this rank's shard (may hold zero images).  Returns HE (3, 2), maxC (2,)."""
    n, h, w = _check_images(images)
    dev = exchange.buf.device
    he = torch.empty((3, 2), dtype=torch.float32, device=dev)
    maxc = torch.empty(2, dtype=torch.float32, device=dev)
    first = exchange.epoch + 1
    with torch.cuda.device(dev):
        check(nv.lib().sx_macenko_fit_peers(_ptr(images) if n > 0 else None, _dtype_code(images), n, h, w, ctypes.c_void_p(exchange.ptrs_dev), _ptr(exchange.buf), exchange.world, exchange.rank,
                                            first & 0xFFFFFFFF, int(exact), _ptr(scratch), _ptr(he), _ptr(maxc), _stream(dev)), "sx_macenko_fit_peers")
    exchange.epoch += 5
    return he, maxc


def _macenko_out(images: torch.Tensor, unit: bool) -> torch.Tensor:
    # uint8 in -> uint8 out (reference preserves dtype, torch_backend.py:L560) unless the
    # normalize_to_0_1 division follows, which makes it float32 (_template.py:L111-112).
    if images.dtype == torch.uint8 and not unit:
        return torch.empty_like(images)
    if images.dtype in (torch.float16, torch.bfloat16):
        return torch.empty_like(images)  # 16-bit float in -> the same dtype out (the reference casts back, torch_backend.py:L131)
    return torch.empty(images.shape, dtype=torch.float32, device=images.device)


def macenko_transform(images: torch.Tensor, he_ref: torch.Tensor, maxc_ref: torch.Tensor, unit: bool = False) -> torch.Tensor:
    """Per-image Macenko transform on the current stream.  ``unit`` folds the /255 of
    ``normalize_to_0_1`` into the store (float32 output in [0, 1])."""
    n, h, w = _check_images(images)
    dev = images.device
    he_ref = _param(he_ref, dev, (3, 2), "stain_matrix")
    maxc_ref = _param(maxc_ref, dev, (2,), "target_max_conc")
    out = _macenko_out(images, unit)
    if n == 0 or h * w == 0:
        return out
    nbytes = int(nv.lib().sx_macenko_workspace_bytes(n))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    scale = 1.0 / 255.0 if unit else 1.0
    with torch.cuda.device(dev):
        check(nv.lib().sx_macenko_transform(_ptr(images), _dtype_code(images), n, h, w, _ptr(he_ref), _ptr(maxc_ref), _ptr(out), _dtype_code(out), ctypes.c_float(scale), _ptr(ws), nbytes, _stream(dev)), "sx_macenko_transform")
    return out


def macenko_fit(images: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    """Pooled fit over all reference images -> HE (3, 2), maxC (2,) float32 on the device."""
    n, h, w = _check_images(images)
    if n == 0 or h * w == 0:
        raise RuntimeError("Macenko fit needs at least one reference pixel")
    dev = images.device
    he = torch.empty((3, 2), dtype=torch.float32, device=dev)
    maxc = torch.empty(2, dtype=torch.float32, device=dev)
    nbytes = int(nv.lib().sx_macenko_workspace_bytes(1))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(nv.lib().sx_macenko_fit(_ptr(images), _dtype_code(images), n, h, w, _ptr(he), _ptr(maxc), _ptr(ws), nbytes, _stream(dev)), "sx_macenko_fit")
    return he, maxc


def macenko_fit_transform(images: torch.Tensor, unit: bool = False) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """``fit(images)`` then ``transform(images)`` with ONE moments pass over the batch (``sx_macenko_fit_transform``).
    Returns HE (3, 2), maxC (2,) and the normalised batch -- exactly what the two separate calls return."""
    n, h, w = _check_images(images)
    if n == 0 or h * w == 0:
        raise RuntimeError("Macenko fit needs at least one reference pixel")
    dev = images.device
    he = torch.empty((3, 2), dtype=torch.float32, device=dev)
    maxc = torch.empty(2, dtype=torch.float32, device=dev)
    out = _macenko_out(images, unit)
    nbytes = int(nv.lib().sx_macenko_workspace_bytes(n + 1))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    scale = 1.0 / 255.0 if unit else 1.0
    with torch.cuda.device(dev):
        check(nv.lib().sx_macenko_fit_transform(_ptr(images), _dtype_code(images), n, h, w, _ptr(he), _ptr(maxc), _ptr(out), _dtype_code(out), ctypes.c_float(scale), _ptr(ws), nbytes, _stream(dev)),
              "sx_macenko_fit_transform")
    return he, maxc, out


def macenko_fit_transform_peers(images: torch.Tensor, exchange, scratch: torch.Tensor, unit: bool = False, exact: bool = False) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """The sharded ``fit_transform`` of one NVLink node in one library call: pooled fit over every rank's images (five
    fused exchanges; advances the exchange's epoch by 5), then this rank's images transformed with it.  ``images`` may
    hold zero images."""
    n, h, w = _check_images(images)
    dev = exchange.buf.device
    he = torch.empty((3, 2), dtype=torch.float32, device=dev)
    maxc = torch.empty(2, dtype=torch.float32, device=dev)
    out = _macenko_out(images, unit)
    have = n > 0 and h * w > 0
    nbytes = int(nv.lib().sx_macenko_workspace_bytes(n)) if have else 0
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev) if have else None
    scale = 1.0 / 255.0 if unit else 1.0
    first = exchange.epoch + 1
    with torch.cuda.device(dev):
        check(nv.lib().sx_macenko_fit_transform_peers(_ptr(images) if have else None, _dtype_code(images), n, h, w, ctypes.c_void_p(exchange.ptrs_dev), _ptr(exchange.buf), exchange.world, exchange.rank,
                                                      first & 0xFFFFFFFF, int(exact), _ptr(scratch), _ptr(he), _ptr(maxc), _ptr(out) if have else None, _dtype_code(out), ctypes.c_float(scale),
                                                      _ptr(ws) if have else None, nbytes, _stream(dev)), "sx_macenko_fit_transform_peers")
    exchange.epoch += 5
    return he, maxc, out
