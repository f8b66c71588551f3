"""Shared fit/transform control flow of the normalizers.

Mirrors ``src/stainx/normalizers/_template.py:L15-142`` with one deliberate difference: there is
no backend dispatch.  ``backend`` may be ``None`` or ``"torch_cuda"``; ``"torch"`` (the
reference's pure-PyTorch backend) does not exist here and is rejected, and both fit and
transform run on the GPU through ``libstainx_b200.so``.
"""
from __future__ import annotations

from typing import Any

from stainx_b200.base import StainNormalizerBase
from stainx_b200.sharding import StatReducer

_VALID_BACKENDS = frozenset({"torch_cuda"})


class NormalizerTemplate(StainNormalizerBase):
    def __init__(self, device: str | Any | None = None, backend: str | None = None, process_group: Any | None = None):
        """
        Args:
            device: CUDA device (string or ``torch.device``); ``None`` picks ``cuda``.
            backend: ``None`` or ``"torch_cuda"``.
            process_group: optional ``torch.distributed`` group (or ``"world"``).  When given, the
                tensors passed to ``fit`` / ``transform`` are this rank's image shard and the
                whole-batch statistics are all-reduced, so all ranks fit identically.
        """
        super().__init__(device)
        if backend is not None and backend not in _VALID_BACKENDS:
            raise ValueError(f"Unsupported backend '{backend}'. Valid backends: {sorted(_VALID_BACKENDS)} (this build has no 'torch' backend and no fallback).")
        from stainx_b200.backends.torch_cuda_backend import CUDA_AVAILABLE

        if not CUDA_AVAILABLE:
            raise ImportError("Backend 'torch_cuda' requires the libstainx_b200 extension. Build it with `python -m stainx_b200.build`.")
        self.backend = "torch_cuda"
        self._backend_impl = None
        self._process_group = process_group
        self._init_algorithm_attributes()

    # -- hooks for subclasses -------------------------------------------------------------------
    def _init_algorithm_attributes(self) -> None:
        """Create the fitted-parameter attributes (all ``None`` before ``fit``)."""

    def _get_torch_cuda_class(self):
        raise NotImplementedError

    def _get_backend_kwargs(self) -> dict:
        return {}

    def _compute_reference_params(self, images: Any) -> None:
        raise NotImplementedError

    def _get_reference_params(self) -> tuple:
        raise NotImplementedError

    def _fitted_tensors(self) -> list:
        """Fitted parameters as a flat list of tensors (for broadcast)."""
        raise NotImplementedError

    # -- shared control flow --------------------------------------------------------------------
    def _select_backend(self) -> str:
        return "torch_cuda"

    def _make_reducer(self) -> StatReducer:
        if self._process_group is None:
            return StatReducer()
        if isinstance(self._process_group, str):
            if self._process_group != "world":
                raise ValueError("process_group must be a ProcessGroup, 'world' or None")
            return StatReducer.world()
        return StatReducer(self._process_group)

    def _get_backend_impl(self):
        if self._backend_impl is None:
            self._backend_impl = self._get_torch_cuda_class()(self.device, reducer=self._make_reducer(), **self._get_backend_kwargs())
        return self._backend_impl

    def fit(self, images: Any) -> "NormalizerTemplate":
        self._compute_reference_params(images)
        self._is_fitted = True
        return self

    def fit_broadcast(self, images: Any | None, src: int = 0) -> "NormalizerTemplate":
        """Reference-mode fit for sharded runs: rank ``src`` fits on ``images`` (single-device
        semantics, no pooling over ranks) and broadcasts the fitted parameters, so every rank
        transforms with bit-identical parameters.  Other ranks may pass ``None``."""
        import torch

        reducer = self._make_reducer()
        impl = self._get_backend_impl()
        if not reducer.enabled:
            return self.fit(images)
        saved = impl._reducer
        impl._reducer = StatReducer()  # local fit on src
        try:
            if reducer.rank == src:
                self._compute_reference_params(images)
            else:
                self._allocate_reference_params(torch.device(self.device))
        finally:
            impl._reducer = saved
        reducer.broadcast_(self._fitted_tensors(), src=src)
        self._is_fitted = True
        return self

    def _allocate_reference_params(self, device) -> None:
        raise NotImplementedError

    def transform(self, images: Any) -> Any:
        if not self._is_fitted:
            raise ValueError("Must call fit() before transform()")
        return self._get_backend_impl().transform(images, *self._get_reference_params())
