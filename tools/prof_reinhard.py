#!/usr/bin/env python
"""Small Reinhard workload for ncu captures (development tool): one float32 64x3x1024x1024 transform."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from stainx_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(43)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
src = torch.rand((n, 3, 1024, 1024), device=dev, generator=g)
mean = torch.tensor([150.0, 140.0, 130.0], device=dev)
std = torch.tensor([40.0, 10.0, 12.0], device=dev)
for _ in range(2):
    out = ops.reinhard_transform(src, mean, std)
torch.cuda.synchronize()
print("done", float(out.sum()))
