"""``backend="torch_cuda"`` implementations on top of ``libstainx_b200.so``.

Mirrors the reference's ``src/stainx/backends/torch_cuda_backend.py`` (class names, constructor
arguments, ``transform(images, *reference_params)`` signatures, error behaviour) and additionally
provides the fit entry points, which the reference only has in its torch backend
(``compute_reference_*_torch``, ``torch_backend.py:L143-179, L308-323, L463-519``): here fit runs on
the GPU with the same kernels as transform.

Sharded execution: every class takes an optional :class:`~stainx_b200.sharding.StatReducer`.  With
one, the batch passed to ``fit`` / ``transform`` is this rank's shard and whole-batch statistics
are all-reduced between kernel phases, so every rank sees the single-device result.
"""
from __future__ import annotations

import torch

from stainx_b200 import _native
from stainx_b200.sharding import StatReducer

CUDA_AVAILABLE = _native.FUNCTIONS_AVAILABLE

_CHANNELS_LAST = (-1, 3)
_NATIVE_DTYPES = (torch.uint8, torch.float32, torch.float16, torch.bfloat16)


class TorchCUDABackendBase:
    """Device / availability checks shared by the three backends
    (reference: ``torch_cuda_backend.py:L17-33``)."""

    def __init__(self, device: str | torch.device | None = None, reducer: StatReducer | None = None):
        self._ops = self._kernel_layer()
        self._reducer = reducer if reducer is not None else StatReducer()
        if device is None:
            if not torch.cuda.is_available():
                raise RuntimeError("CUDA is not available on this system")
            self.device = torch.device("cuda")
        else:
            self.device = torch.device(device)
        self._check_device()

    # The two hooks below are the only seams of this class: the CPU test-suite subclasses the backends
    # (tests/cpu_ops.py: ``cpu_backend``) to exercise the sharding control flow with gloo; the product
    # classes always bind the native library and always require a CUDA device.
    @staticmethod
    def _kernel_layer():
        if not _native.available():
            raise ImportError("libstainx_b200 is not built. CUDA backend is not available (python -m stainx_b200.build); there is no fallback backend.")
        from stainx_b200 import ops as native_ops

        return native_ops

    def _check_device(self) -> None:
        if self.device.type != "cuda":
            raise ValueError(f"CUDA backend requires CUDA device, got {self.device.type}")

    def _to_native(self, images: torch.Tensor) -> tuple[torch.Tensor, torch.dtype]:
        """Move to the device and to a dtype the kernels take: uint8, float32, float16 and bfloat16 are read
        natively (the 16-bit types are widened inside the kernels, 8 pixels per 128-bit load -- the reference
        widens them with a separate ``.float()`` pass, torch_backend.py:L103-113); any other dtype (float64)
        is converted to float32 in [0, 1] first and the result is cast back (L131)."""
        if not isinstance(images, torch.Tensor):
            raise TypeError(f"images must be a torch.Tensor, got {type(images)}")
        original = images.dtype
        images = images.to(self.device)
        if images.dtype not in _NATIVE_DTYPES:
            images = images.float()
        return images.contiguous(), original

    @staticmethod
    def _restore_dtype(result: torch.Tensor, original: torch.dtype) -> torch.Tensor:
        # torch_backend.py:L131: the result is cast back to the caller's dtype.
        if original in _NATIVE_DTYPES or result.dtype == original:
            return result
        return result.to(original)


class HistogramMatchingCUDA(TorchCUDABackendBase):
    def __init__(self, device: str | torch.device | None = None, channel_axis: int = 1, reducer: StatReducer | None = None):
        super().__init__(device, reducer)
        self.channel_axis = channel_axis
        self._exchange = None        # sharding.PeerExchange (NVLink peer memory) or False when unavailable
        self._ref_cdf_cache = None   # (ref_hist, ref_cdf): the reference CDF is a fit-time constant

    def _layout(self, images: torch.Tensor) -> int:
        if self.channel_axis == -1 or (self.channel_axis == 3 and images.ndim == 4):
            return _native.SX_NHWC
        return _native.SX_NCHW

    def _require_rgb(self, images: torch.Tensor) -> None:
        """The kernels are written for three channels ('Layout everywhere: C=3').
        The reference's torch backend loops over any channel count (torch_backend.py:L149-179,
        L228-286); that generality is deliberately not reproduced -- say so instead of failing deep
        inside the kernel layer."""
        if images.dim() != 4:
            raise ValueError(f"HistogramMatching expects a 4-D batch (NCHW, or NHWC with channel_axis=-1), got shape {tuple(images.shape)}")
        c = images.shape[-1] if self._layout(images) == _native.SX_NHWC else images.shape[1]
        if c != 3:
            raise ValueError(f"stainx_b200 HistogramMatching supports RGB images only (C=3 on channel_axis={self.channel_axis}), got C={c} with shape {tuple(images.shape)}; "
                             "grayscale / multi-channel inputs are a documented restriction of this backend.")

    def _stack_reference(self, reference_histogram: torch.Tensor | list) -> torch.Tensor:
        """(3, 256) float32 reference histograms from the list / single-tensor forms the reference
        accepts (``torch_cuda_backend.py:L51-75``)."""
        if isinstance(reference_histogram, (list, tuple)):
            if len(reference_histogram) == 0:
                raise ValueError("reference_histogram list cannot be empty")
            for i, h in enumerate(reference_histogram):
                if not isinstance(h, torch.Tensor):
                    raise TypeError(f"reference_histogram[{i}] must be a torch.Tensor, got {type(h)}")
                if h.dim() != 1 or h.size(0) != 256:
                    raise ValueError(f"Each histogram in reference_histogram list must be 1D with 256 elements. Got histogram at index {i} with shape {h.shape}")
            rows = [h.to(self.device) for h in reference_histogram[:3]]
            while len(rows) < 3:
                rows.append(rows[0])
            return torch.stack(rows, dim=0).float().contiguous()
        ref = reference_histogram.to(self.device)
        if ref.dim() == 2 and tuple(ref.shape) == (3, 256):
            return ref.float().contiguous()
        if ref.dim() != 1 or ref.size(0) != 256:
            raise ValueError(f"reference_histogram must be 1D with 256 elements. Got shape {ref.shape}")
        return ref.float().unsqueeze(0).repeat(3, 1).contiguous()

    def _reference(self, reference_histogram: torch.Tensor | list) -> tuple[torch.Tensor, torch.Tensor]:
        """(ref_hist (3, 256), ref_cdf (3, 256)) of the reference; both are fit-time constants, so they
        are cached for as long as the caller passes the SAME tensor objects, unmodified (the cache
        holds them, so their storage cannot be recycled under it)."""
        parts = list(reference_histogram) if isinstance(reference_histogram, (list, tuple)) else [reference_histogram]
        cached = self._ref_cdf_cache
        if cached is not None and len(cached[0]) == len(parts) and all(a is b for a, b in zip(cached[0], parts)) and cached[1] == [t._version for t in parts if isinstance(t, torch.Tensor)]:
            return cached[2], cached[3]
        ref_hist = self._stack_reference(reference_histogram)
        ref_cdf = self._ops.hm_ref_cdf(ref_hist)
        self._ref_cdf_cache = (parts, [t._version for t in parts], ref_hist, ref_cdf)
        return ref_hist, ref_cdf

    def _peer_exchange(self):
        """NVLink peer buffer for the fused counts exchange, created collectively on first use
        (``None`` when the group cannot map peer memory: the NCCL all-reduce is used instead)."""
        if self._exchange is None:
            from stainx_b200.sharding import PeerExchange

            nbytes = int(_native.lib().sx_hm_peer_buffer_bytes())
            self._exchange = PeerExchange.create(self._reducer, torch.device(self.device), nbytes) or False
        return self._exchange or None

    def compute_reference_counts(self, images: torch.Tensor) -> torch.Tensor:
        """Whole-reference-set per-channel counts, int64 (3, 256), summed over ranks."""
        images, _ = self._to_native(images)
        self._require_rgb(images)
        counts = self._ops.hm_hist(images, self._layout(images))
        return self._reducer.sum_(counts)

    def compute_reference_histograms(self, images: torch.Tensor) -> tuple[torch.Tensor, list[torch.Tensor]]:
        """H1 (``compute_reference_histograms_torch``): -> counts (3, 256) int64 and the list of three
        float32 (256,) histograms ``counts / (counts.sum() + 1e-8)``."""
        counts = self.compute_reference_counts(images)
        ref_hist = self._ops.hm_ref_hist(counts)
        return counts, [ref_hist[c] for c in range(3)]

    def transform(self, images: torch.Tensor, reference_histogram: torch.Tensor | list) -> torch.Tensor:
        images, original = self._to_native(images)
        self._require_rgb(images)
        layout = self._layout(images)
        if self._reducer.enabled:
            # H3 -> all-reduce -> H2 -> H4: the source histogram spans the whole sharded batch.
            ref_hist, ref_cdf = self._reference(reference_histogram)
            ex = self._peer_exchange()
            if ex is not None:
                # the all-reduce is fused into the LUT kernel (peer loads over NVLink, no NCCL call), and the
                # whole step is one library call = one chain of programmatic dependent launches
                return self._restore_dtype(self._ops.hm_transform_peers(images, ex, ref_cdf, layout), original)
            else:
                counts = self._reducer.sum_(self._ops.hm_hist(images, layout))
                # npix = -1: the LUT kernel takes the global pixel count from the reduced counts
                lut = self._ops.hm_build_lut(counts, -1, ref_cdf)
            result = self._ops.hm_apply(images, lut, layout)
        else:
            # one library call = one chain of programmatic dependent launches (zero, histogram, LUT, remap);
            # the stacked reference is cached, so a transform enqueues no torch kernel at all
            result = self._ops.hm_transform(images, self._reference(reference_histogram)[0], layout)
        return self._restore_dtype(result, original)


class ReinhardCUDA(TorchCUDABackendBase):
    def compute_reference_mean_std(self, images: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        """R5 (``compute_reference_mean_std_torch``): LAB mean / unbiased std over the whole set."""
        images, _ = self._to_native(images)
        if self._reducer.enabled:
            return self._ops.reinhard_finalize(self._reducer.sum_(self._ops.reinhard_stats(images)))
        return self._ops.reinhard_fit(images)

    _exchange = None  # sharding.PeerExchange (NVLink peer memory) or False when unavailable

    def _peer_exchange(self):
        if self._exchange is None:
            from stainx_b200.sharding import PeerExchange

            nbytes = int(_native.lib().sx_reinhard_peer_buffer_bytes())
            self._exchange = PeerExchange.create(self._reducer, torch.device(self.device), nbytes) or False
        return self._exchange or None

    def transform(self, images: torch.Tensor, target_mean: torch.Tensor, target_std: torch.Tensor) -> torch.Tensor:
        images, original = self._to_native(images)
        if self._reducer.enabled:
            ex = self._peer_exchange()
            if ex is not None:  # the all-reduce of the sums is fused into the finalize kernel (NVLink peer loads)
                epoch = ex.epoch + 1  # committed once every launch of the step went through (peers step in lockstep)
                sums = ex.view((epoch & 1) * 64, (8,), torch.float64)
                sums.zero_()
                self._ops.reinhard_stats(images, sums=sums)
                src_mean, src_std = self._ops.reinhard_finalize_peers(ex, epoch)
                ex.epoch = epoch
            else:
                src_mean, src_std = self._ops.reinhard_finalize(self._reducer.sum_(self._ops.reinhard_stats(images)))
            result = self._ops.reinhard_apply(images, src_mean, src_std, target_mean, target_std)
        else:
            result = self._ops.reinhard_transform(images, target_mean, target_std)
        return self._restore_dtype(result, original)


class MacenkoCUDA(TorchCUDABackendBase):
    """Macenko backend.

    ``precision`` is accepted for API compatibility (reference: ``"stable"`` = fp64 covariance +
    fp32 pixels, ``"fast"`` = fp16 pixel tensors).  This build has one path: fp32 pixels in
    registers, fp64 moment accumulation and eigen-decomposition, no materialised pixel tensors, so
    both values select the same kernels.
    """

    def __init__(self, device: str | torch.device | None = None, precision: str = "stable", reducer: StatReducer | None = None):
        if precision not in ("stable", "fast"):
            raise ValueError(f"precision must be 'stable' or 'fast', got {precision!r}")
        super().__init__(device, reducer)
        self._precision = precision

    def compute_reference_stain_matrix(self, images: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        """M11 (``compute_reference_stain_matrix_torch``): pooled fit -> HE (3, 2), maxC (2,)."""
        images, _ = self._to_native(images)
        if images.dim() != 4 or images.shape[1] != 3:
            raise ValueError(f"Macenko fit expects NCHW with C=3, got shape {tuple(images.shape)}")
        if not self._reducer.enabled:
            return self._ops.macenko_fit(images)
        return self._pooled_fit_sharded(images)

    _exchange = None  # sharding.PeerExchange (NVLink peer memory) or False when unavailable

    def _peer_exchange(self):
        if self._exchange is None:
            from stainx_b200.sharding import PeerExchange

            nbytes = int(_native.lib().sx_macenko_peer_buffer_bytes())
            self._exchange = PeerExchange.create(self._reducer, torch.device(self.device), nbytes) or False
            if self._exchange:
                self._scratch = torch.empty(int(_native.lib().sx_macenko_peer_scratch_bytes()), dtype=torch.uint8, device=self.device)
        return self._exchange or None

    def _pooled_fit_sharded(self, images: torch.Tensor, exact: bool = False) -> tuple[torch.Tensor, torch.Tensor]:
        """Pooled fit of a sharded reference batch: every rank streams its own images into slot 0 of
        its workspace; before each per-slot step the statistics of all ranks are combined -- in ONE
        kernel over NVLink peer memory when the group can map it (``sx_macenko_peer_combine``), else
        with NCCL all-reduces -- so every rank derives bit-identical HE / maxC.

        The brackets of the rank searches come from a ~1/64 subsample (``hist`` level 0).  Should a wanted
        rank fall outside its bracket (STATUS region, identical on every rank because it derives from the
        combined statistics), the fit is repeated with ``exact=True``: level 2 histograms every pixel group,
        which makes the bracket certain.  The check costs one host read per fit."""
        red = self._reducer
        ex = self._peer_exchange()
        ws = self._ops.MacenkoWorkspace(1, images.device, buffer=ex.buf) if ex is not None else self._ops.MacenkoWorkspace(1, images.device)
        coarse = 2 if exact else 0
        if ex is not None and hasattr(self._ops, "macenko_fit_peers"):
            # one library call: the phases below with the five exchanges fused in (no interpreter time between ~25 launches)
            he, maxc = self._ops.macenko_fit_peers(images, ex, self._scratch, exact=exact)
            if int(ws.region("status")[0, 0].item()) & 3:
                if exact:
                    raise _native.StainxNativeError("Macenko pooled fit: a rank fell outside an exact bracket (inconsistent statistics across ranks?)")
                return self._pooled_fit_sharded(images, exact=True)
            return he, maxc

        def combine(which: int) -> None:
            if ex is not None:
                self._ops.macenko_peer_combine(ex, which, self._scratch)
            elif which == 0:
                red.sum_(ws.region("moments"))
                red.max_(ws.region("odrange"))
            elif which == 1:
                red.sum_(ws.region("hist1"))
                red.sum_(ws.region("counters"))
            else:
                red.sum_(ws.region("hist2"))
                red.sum_(ws.region("counters"))
                red.min_(ws.region("vmin"))
                red.max_(ws.region("vmax"))

        ws.begin()
        if images.shape[0] > 0:
            ws.moments(images, pooled=True)
        combine(0)
        ws.basis(0, 1, allow_fallback=False)
        for stage in (_native.SX_STAGE_ANGLE, _native.SX_STAGE_CONC):
            if images.shape[0] > 0:
                ws.hist(images, True, stage, coarse)  # subsample pass (exact: every pixel group)
            combine(1)
            ws.select(0, 1, stage, 0)             # ranks + brackets, identical on every rank
            if images.shape[0] > 0:
                ws.hist(images, True, stage, 1)   # full pass
            combine(2)
            ws.select(0, 1, stage, 1)
        fit = ws.region("fit")[0]
        he, maxc = fit[:6].reshape(3, 2).clone(), fit[6:8].clone()
        missed = int(ws.region("status")[0, 0].item()) & 3
        if missed:
            if exact:
                raise _native.StainxNativeError("Macenko pooled fit: a rank fell outside an exact bracket (inconsistent statistics across ranks?)")
            return self._pooled_fit_sharded(images, exact=True)
        return he, maxc

    def fit_transform(self, images: torch.Tensor, normalize_to_0_1: bool = False) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """``fit(images)`` followed by ``transform(images)`` (base.py:L59-61) with the batch read ONCE for the moments:
        the per-image moments of the transform, summed, ARE the pooled moments of the fit (fixed-point integers, so bit
        for bit).  Returns HE, maxC and the normalised batch -- the same values as the two separate calls.  Sharded:
        the pooled fit spans every rank's images (``sx_macenko_fit_transform_peers``, one library call per rank)."""
        native, original = self._to_native(images)
        if native.dim() != 4 or native.shape[1] != 3:
            raise ValueError(f"Macenko fit expects NCHW with C=3, got shape {tuple(native.shape)}")
        unit = bool(normalize_to_0_1)
        ops = self._ops
        ex = self._peer_exchange() if self._reducer.enabled else None
        if not self._reducer.enabled and hasattr(ops, "macenko_fit_transform"):
            he, maxc, result = ops.macenko_fit_transform(native, unit=unit)
        elif ex is not None and hasattr(ops, "macenko_fit_transform_peers"):
            he, maxc, result = ops.macenko_fit_transform_peers(native, ex, self._scratch, unit=unit)
            status = ops.MacenkoWorkspace(1, native.device, buffer=ex.buf).region("status")
            if int(status[0, 0].item()) & 3:  # a wanted rank outside its sample bracket (identical on every rank): exact repeat
                he, maxc, result = ops.macenko_fit_transform_peers(native, ex, self._scratch, unit=unit, exact=True)
                if int(status[0, 0].item()) & 3:
                    raise _native.StainxNativeError("Macenko pooled fit: a rank fell outside an exact bracket (inconsistent statistics across ranks?)")
        else:
            he, maxc = self.compute_reference_stain_matrix(native)
            result = ops.macenko_transform(native, he, maxc.flatten(), unit=unit)
        if unit and original == torch.uint8:
            return he, maxc, result
        return he, maxc, self._restore_dtype(result, original)

    def transform(self, images: torch.Tensor, stain_matrix: torch.Tensor, target_max_conc: torch.Tensor, normalize_to_0_1: bool = False) -> torch.Tensor:
        images, original = self._to_native(images)
        if tuple(stain_matrix.shape) != (3, 2):
            raise ValueError(f"stain_matrix must have shape (3, 2), got {stain_matrix.shape}")
        if images.dim() != 4:
            raise ValueError(f"Macenko expects NCHW images, got shape {tuple(images.shape)}")
        if images.shape[1] != 3:
            raise ValueError(f"Macenko expects 3 channels in dim 1 (NCHW), got C={images.shape[1]} with shape {tuple(images.shape)}")
        # Per-image statistics: a sharded batch needs no exchange at transform time.
        result = self._ops.macenko_transform(images, stain_matrix, target_max_conc.flatten(), unit=bool(normalize_to_0_1))
        if normalize_to_0_1 and original == torch.uint8:
            return result  # the reference's `uint8 / 255.0` is float32
        return self._restore_dtype(result, original)
