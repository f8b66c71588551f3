"""Histogram matching (reference: ``src/stainx/normalizers/histogram_matching.py``)."""
from __future__ import annotations

from typing import Any

import torch

from stainx_b200.normalizers._template import NormalizerTemplate


class HistogramMatching(NormalizerTemplate):
    def __init__(self, device: Any | None = None, backend: str | None = None, channel_axis: int = 1, process_group: Any | None = None):
        self.channel_axis = channel_axis
        super().__init__(device=device, backend=backend, process_group=process_group)

    def _init_algorithm_attributes(self) -> None:
        self._ref_counts = None           # int64 (3, 256) -- source of the lazy attributes below
        self._ref_histograms_256 = None   # list of 3 float32 (256,) tensors, what transform consumes
        self._lazy = None

    def _get_torch_cuda_class(self):
        from stainx_b200.backends.torch_cuda_backend import HistogramMatchingCUDA

        return HistogramMatchingCUDA

    def _get_backend_kwargs(self) -> dict:
        return {"channel_axis": self.channel_axis}

    def _compute_reference_params(self, images: Any) -> None:
        self._ref_counts, self._ref_histograms_256 = self._get_backend_impl().compute_reference_histograms(images)
        self._lazy = None

    def _allocate_reference_params(self, device) -> None:
        self._ref_counts = torch.zeros((3, 256), dtype=torch.int64, device=device)
        stacked = torch.empty((3, 256), dtype=torch.float32, device=device)
        self._ref_histograms_256 = [stacked[c] for c in range(3)]
        self._lazy = None

    def _fitted_tensors(self) -> list:
        return [self._ref_counts, *self._ref_histograms_256]

    def _get_reference_params(self) -> tuple:
        return (self._ref_histograms_256,)

    # The reference also stores per-channel (values, cdf) over the non-empty bins and
    # `_reference_histogram = ref_cdf[0]` (torch_backend.py:L162-179).  Nothing on the transform
    # path reads them, so they are derived on first access (a 256-element host-side view).
    def _legacy(self):
        if self._lazy is None and self._ref_counts is not None:
            vals, cdfs = [], []
            for c in range(3):
                counts = self._ref_counts[c].float()
                nz = torch.nonzero(counts, as_tuple=False).squeeze(-1)
                if nz.numel() > 0:
                    cdf = torch.cumsum(counts[nz], dim=0)
                    vals.append(nz.float())
                    cdfs.append(cdf / (cdf[-1] + 1e-8))
                else:
                    vals.append(torch.arange(256, dtype=torch.float32, device=counts.device))
                    cdfs.append(torch.zeros(256, dtype=torch.float32, device=counts.device))
            self._lazy = (vals, cdfs)
        return self._lazy

    @property
    def _ref_vals(self):
        lazy = self._legacy()
        return None if lazy is None else lazy[0]

    @property
    def _ref_cdf(self):
        lazy = self._legacy()
        return None if lazy is None else lazy[1]

    @property
    def _reference_histogram(self):
        lazy = self._legacy()
        return None if lazy is None else lazy[1][0]
