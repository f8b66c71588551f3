"""Device helper (reference: ``src/stainx/utils.py:L12-34``, ``get_device``)."""
from __future__ import annotations

from typing import Any

import torch


def get_device(device: str | Any | None) -> Any:
    """``None`` -> CUDA when present (else CPU placeholder); strings -> ``torch.device``; device
    objects pass through.  This build computes on CUDA only; a CPU device is accepted here (so
    that objects can be constructed on any host) and rejected when a backend is instantiated."""
    if device is None:
        return torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")
    if isinstance(device, str):
        return torch.device(device)
    return device
