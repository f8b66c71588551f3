"""Macenko stain normalization (reference: ``src/stainx/normalizers/macenko.py``)."""
from __future__ import annotations

from typing import Any

import torch

from stainx_b200.normalizers._template import NormalizerTemplate


class Macenko(NormalizerTemplate):
    """``normalize_to_0_1`` defaults to ``False`` (output ~``[0, 255]``);
    ``StainNormalizerTransform(method="macenko")`` defaults it to ``True``.

    ``precision`` is ``"stable"`` or ``"fast"``; both run the same fp32-pixel / fp64-statistics
    kernels in this build (see ``MacenkoCUDA``).  ``"fast"`` still requires the ``torch_cuda``
    backend, which is the only backend here.
    """

    def __init__(self, device: Any | None = None, backend: str | None = None, normalize_to_0_1: bool = False, precision: str = "stable", process_group: Any | None = None):
        if precision not in ("stable", "fast"):
            raise ValueError(f"precision must be 'stable' or 'fast', got {precision!r}")
        if precision == "fast" and backend not in (None, "torch_cuda"):
            raise ValueError(f"precision='fast' requires backend='torch_cuda', but backend is '{backend}'. Either set backend='torch_cuda' or use precision='stable'.")
        self._precision = precision
        self.normalize_to_0_1 = normalize_to_0_1
        super().__init__(device=device, backend=backend, process_group=process_group)

    def _init_algorithm_attributes(self) -> None:
        self._stain_matrix = None
        self._concentration_matrix = None
        self._target_max_conc = None

    def _get_torch_cuda_class(self):
        from stainx_b200.backends.torch_cuda_backend import MacenkoCUDA

        return MacenkoCUDA

    def _get_backend_kwargs(self) -> dict:
        return {"precision": self._precision} if self._precision != "stable" else {}

    def _compute_reference_params(self, images: Any) -> None:
        self._stain_matrix, self._target_max_conc = self._get_backend_impl().compute_reference_stain_matrix(images)
        self._concentration_matrix = None

    def _allocate_reference_params(self, device) -> None:
        self._stain_matrix = torch.empty((3, 2), dtype=torch.float32, device=device)
        self._target_max_conc = torch.empty(2, dtype=torch.float32, device=device)

    def _fitted_tensors(self) -> list:
        return [self._stain_matrix, self._target_max_conc]

    def _get_reference_params(self) -> tuple:
        return (self._stain_matrix, self._target_max_conc)

    def fit_transform(self, images: Any) -> Any:
        """``fit(images).transform(images)`` (base.py:L59-61) through the backend's one-call path: the batch is read
        once for the moments of both steps; fitted parameters and output are those of the two separate calls."""
        self._stain_matrix, self._target_max_conc, out = self._get_backend_impl().fit_transform(images, normalize_to_0_1=bool(self.normalize_to_0_1))
        self._concentration_matrix = None
        self._is_fitted = True
        return out

    def transform(self, images: Any) -> Any:
        if not self._is_fitted:
            raise ValueError("Must call fit() before transform()")
        # The reference divides the finished result by 255 in a separate pass
        # (_template.py:L111-112); here the division is folded into the kernel's store.
        return self._get_backend_impl().transform(images, self._stain_matrix, self._target_max_conc, normalize_to_0_1=bool(self.normalize_to_0_1))
