"""Multi-GPU checks (skipped on boxes with fewer than two GPUs): the fused NVLink peer exchange of
the sharded HistogramMatching transform against the NCCL path and the single-device result."""
from __future__ import annotations

import subprocess
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]
pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs on one node")
def test_hm_fused_peer_exchange_two_gpus():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29577", str(ROOT / "tools" / "check_peers.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "PEERS CHECK OK" in res.stdout
