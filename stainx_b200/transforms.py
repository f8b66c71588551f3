"""``StainNormalizerTransform``: stain normalization as an ``nn.Module`` for torch / torchvision
pipelines.

Same constructor, modes, validation messages and ``forward`` behaviour as the reference's
``src/stainx/transforms.py:L26-230``; the compute underneath is this package's CUDA path, so a
batch must end up on a CUDA device (CUDA inputs with ``device=None``, or an explicit
``device="cuda"``).

Modes: ``reference`` fits once on a fixed reference and transforms every batch; ``batch`` re-fits
on image ``batch_ref_index`` of every incoming batch (mutable state, not reproducible across steps).
Layout: Macenko / Reinhard take NCHW (or CHW); ``channel_axis`` is a histogram-matching option.
Value range: uint8 is [0, 255], float tensors are read as [0, 1] (never max-rescaled).  Macenko
built here defaults to ``normalize_to_0_1=True``.  Fitted parameters are plain attributes of the
inner normalizer: ``state_dict()`` does not carry them (call ``fit_reference`` after loading).

Sharded runs: pass ``process_group`` (a ``torch.distributed`` group or ``"world"``).  Each rank
then feeds its own image shard; reference-mode fitting happens on rank 0 and is broadcast,
``mode="batch"`` fits on the rank that owns global image ``batch_ref_index`` and broadcasts.
"""
from __future__ import annotations

from typing import Any, Literal

import torch
import torch.nn as nn

from stainx_b200.normalizers import HistogramMatching, Macenko, Reinhard

MethodName = Literal["macenko", "reinhard", "histogram_matching"]
ModeName = Literal["reference", "batch"]

_METHODS = {"macenko": Macenko, "reinhard": Reinhard, "histogram_matching": HistogramMatching}
_NCHW_AXES = frozenset({1, -3})
_NHWC_AXES = frozenset({-1, 3})
# tensors (or lists of tensors) that must follow the batch device
_FITTED_ATTRS = ("_stain_matrix", "_target_max_conc", "_reference_mean", "_reference_std", "_ref_counts", "_ref_histograms_256")


def _same_layout(a: int, b: int) -> bool:
    return (a in _NCHW_AXES and b in _NCHW_AXES) or (a in _NHWC_AXES and b in _NHWC_AXES)


class StainNormalizerTransform(nn.Module):
    def __init__(
        self,
        method: MethodName = "macenko",
        *,
        mode: ModeName = "reference",
        reference: torch.Tensor | None = None,
        device: str | torch.device | None = None,
        backend: str | None = None,
        channel_axis: int = 1,
        batch_ref_index: int = 0,
        normalize_to_0_1: bool | None = None,
        normalizer: Any | None = None,
        process_group: Any | None = None,
    ):
        super().__init__()
        if mode not in ("reference", "batch"):
            raise ValueError(f"Unsupported mode '{mode}'. Use 'reference' or 'batch'.")
        self.mode = mode
        self.channel_axis = channel_axis
        self.batch_ref_index = batch_ref_index
        self.device = None if device is None else torch.device(device)  # None: follow the input
        self._requested_backend = backend
        self._process_group = process_group

        if self.device is not None and self.device.type != "cuda":
            raise ValueError(f"backend='torch_cuda' requires a CUDA device, got {self.device}.")

        requested_unit = normalize_to_0_1
        if normalizer is not None:
            self.normalizer = self._adopt(normalizer, requested_unit, channel_axis)
        else:
            self.normalizer = self._build(method, backend, channel_axis, requested_unit)

        if mode == "reference":
            if reference is None and not getattr(self.normalizer, "_is_fitted", False):
                raise ValueError("mode='reference' requires a reference tensor (or a pre-fitted normalizer).")
            if reference is not None:
                self.fit_reference(reference)

    # ------------------------------------------------------------------ construction helpers
    def _adopt(self, normalizer: Any, requested_unit: bool | None, channel_axis: int) -> Any:
        """Use a pre-built normalizer; reconcile the transform's options with it."""
        if isinstance(normalizer, Macenko):
            if requested_unit is not None:
                normalizer.normalize_to_0_1 = bool(requested_unit)
        elif requested_unit:
            raise ValueError("normalize_to_0_1 only applies to Macenko normalizers.")
        if isinstance(normalizer, HistogramMatching):
            inner_axis = int(normalizer.channel_axis)
            if channel_axis != 1 and not _same_layout(channel_axis, inner_axis):
                raise ValueError(f"channel_axis={channel_axis} conflicts with prebuilt HistogramMatching(channel_axis={inner_axis}).")
            self.channel_axis = inner_axis  # layout checks follow the inner normalizer
        elif channel_axis not in _NCHW_AXES:
            raise ValueError(f"channel_axis={channel_axis} is only supported for histogram_matching; Macenko/Reinhard require NCHW (channel_axis=1).")
        return normalizer

    def _build(self, method: str, backend: str | None, channel_axis: int, requested_unit: bool | None) -> Any:
        if method not in _METHODS:
            raise ValueError(f"Unknown method '{method}'. Choose from {sorted(_METHODS)}")
        if method != "histogram_matching" and channel_axis not in _NCHW_AXES:
            raise ValueError(f"channel_axis={channel_axis} is only supported for histogram_matching; {method} requires NCHW (channel_axis=1).")
        if requested_unit and method != "macenko":
            raise ValueError("normalize_to_0_1 only applies to Macenko (method='macenko').")
        if self.device is not None:
            norm_device: Any = self.device
        else:
            # placeholder until the first tensor reveals where batches live
            norm_device = torch.device("cuda") if torch.cuda.is_available() else "cpu"
        common = {"device": norm_device, "backend": backend, "process_group": self._process_group}
        if method == "histogram_matching":
            return HistogramMatching(channel_axis=channel_axis, **common)
        if method == "macenko":
            unit = True if requested_unit is None else bool(requested_unit)  # training-safe default
            return Macenko(normalize_to_0_1=unit, **common)
        return Reinhard(**common)

    # ------------------------------------------------------------------ device / layout
    def _layout_axis(self) -> int:
        if isinstance(self.normalizer, HistogramMatching):
            return int(self.normalizer.channel_axis)
        return self.channel_axis

    def _follow_device(self, device: torch.device) -> None:
        """Point the inner normalizer (and its fitted tensors) at the batch device."""
        device = torch.device(device)
        if device.type != "cuda":
            raise ValueError(f"backend='torch_cuda' requires CUDA tensors when device=None; got {device}.")
        if torch.device(self.normalizer.device) == device:
            return
        self.normalizer.device = device
        self.normalizer._backend_impl = None
        for name in _FITTED_ATTRS:
            value = getattr(self.normalizer, name, None)
            if isinstance(value, torch.Tensor):
                setattr(self.normalizer, name, value.to(device))
            elif isinstance(value, (list, tuple)) and value and all(isinstance(v, torch.Tensor) for v in value):
                setattr(self.normalizer, name, type(value)(v.to(device) for v in value))

    def _prepare(self, images: torch.Tensor, for_fit: bool = False) -> torch.Tensor:
        if images.dim() == 3:
            images = images.unsqueeze(0)
        if images.dim() != 4:
            raise ValueError(f"Expected CHW/NCHW or HWC/NHWC image tensor, got shape {tuple(images.shape)}")
        if isinstance(self.normalizer, HistogramMatching) and self._layout_axis() in _NHWC_AXES:
            if images.shape[-1] != 3:
                raise ValueError(f"channels-last histogram matching expects shape (N, H, W, 3), got {tuple(images.shape)}")
        elif images.shape[1] != 3:
            raise ValueError(f"Expected NCHW with C=3 (got shape {tuple(images.shape)}). Macenko/Reinhard do not accept NHWC; use channel_axis=-1 only with histogram_matching, or permute to NCHW first.")
        target = self.device if self.device is not None else images.device
        if for_fit and self.device is None and target.type != "cuda" and torch.cuda.is_available():
            # The reference lets a device=None transform fit on a CPU reference and then follow CUDA batches
            # (tests/torch_interface/test_stain_normalizer_transform.py:L149-157).  There is no CPU compute path
            # here, so a host-resident REFERENCE is fitted on the CUDA device the normalizer currently points at;
            # batches must still arrive on a CUDA device.
            cur = torch.device(self.normalizer.device)
            target = cur if cur.type == "cuda" else torch.device("cuda", torch.cuda.current_device())
        self._follow_device(target)
        return images.to(target)

    # ------------------------------------------------------------------ public API
    def fit_reference(self, reference: torch.Tensor) -> "StainNormalizerTransform":
        """Fit the inner normalizer on a reference image or batch (rank 0 fits and broadcasts
        when a process group was given)."""
        ref = self._prepare(reference, for_fit=True)
        if self._process_group is not None:
            self.normalizer.fit_broadcast(ref, src=0)
        else:
            self.normalizer.fit(ref)
        return self

    def _fit_on_batch(self, batch: torch.Tensor) -> None:
        idx = self.batch_ref_index
        if self._process_group is None:
            if idx < 0 or idx >= batch.shape[0]:
                raise IndexError(f"batch_ref_index={idx} out of range for batch size {batch.shape[0]}")
            self.normalizer.fit(batch[idx : idx + 1])
            return
        # sharded: idx addresses the concatenation of all ranks' shards
        import torch.distributed as dist

        reducer = self.normalizer._make_reducer()
        sizes = [0] * reducer.world_size
        if reducer.enabled:
            dist.all_gather_object(sizes, int(batch.shape[0]), group=reducer.group)
        else:
            sizes = [int(batch.shape[0])]
        if idx < 0 or idx >= sum(sizes):
            raise IndexError(f"batch_ref_index={idx} out of range for global batch size {sum(sizes)}")
        owner, local = 0, idx
        while local >= sizes[owner]:
            local -= sizes[owner]
            owner += 1
        ref = batch[local : local + 1] if reducer.rank == owner else None
        self.normalizer.fit_broadcast(ref, src=owner)

    def forward(self, img: torch.Tensor) -> torch.Tensor:
        single = img.dim() == 3
        batch = self._prepare(img)
        if self.mode == "batch":
            self._fit_on_batch(batch)  # intentional: re-fits on every call
        out = self.normalizer.transform(batch)
        return out.squeeze(0) if single else out
