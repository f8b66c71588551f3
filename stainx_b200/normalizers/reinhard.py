"""Reinhard colour transfer (reference: ``src/stainx/normalizers/reinhard.py``)."""
from __future__ import annotations

from typing import Any

import torch

from stainx_b200.normalizers._template import NormalizerTemplate


class Reinhard(NormalizerTemplate):
    def _init_algorithm_attributes(self) -> None:
        self._reference_mean = None
        self._reference_std = None

    def _get_torch_cuda_class(self):
        from stainx_b200.backends.torch_cuda_backend import ReinhardCUDA

        return ReinhardCUDA

    def _compute_reference_params(self, images: Any) -> None:
        self._reference_mean, self._reference_std = self._get_backend_impl().compute_reference_mean_std(images)

    def _allocate_reference_params(self, device) -> None:
        self._reference_mean = torch.empty(3, dtype=torch.float32, device=device)
        self._reference_std = torch.empty(3, dtype=torch.float32, device=device)

    def _fitted_tensors(self) -> list:
        return [self._reference_mean, self._reference_std]

    def _get_reference_params(self) -> tuple:
        return (self._reference_mean, self._reference_std)
