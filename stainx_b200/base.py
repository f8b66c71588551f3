"""Abstract normalizer interface (reference: ``src/stainx/base.py:L12-61``)."""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Any

from stainx_b200.utils import get_device


class StainNormalizerBase(ABC):
    """``fit`` / ``transform`` / ``fit_transform`` contract shared by all normalizers."""

    def __init__(self, device: str | Any | None = None):
        self.device = get_device(device)
        self._is_fitted = False

    @abstractmethod
    def fit(self, images: Any) -> "StainNormalizerBase":
        """Compute reference parameters from ``images``; returns ``self``."""

    @abstractmethod
    def transform(self, images: Any) -> Any:
        """Normalize ``images`` with the fitted parameters."""

    def fit_transform(self, images: Any) -> Any:
        return self.fit(images).transform(images)
