"""Test configuration: `gpu` marker, repo root on sys.path, shared fixtures."""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden(name: str):
    return np.load(GOLDEN / f"{name}.npz")


@pytest.fixture(scope="session")
def ox():
    """The CPU oracle (compiled on first use)."""
    from oracle import oracle

    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def cuda():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("CUDA is not available")
    return torch.device("cuda:0")
