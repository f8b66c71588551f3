#!/usr/bin/env python
"""Small Macenko / Reinhard workload for ncu captures (development tool)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from stainx_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(43)
which = sys.argv[1] if len(sys.argv) > 1 else "macenko"
n = 32
src = torch.rand((n, 3, 1024, 1024), device=dev, generator=g)
ref = torch.rand((1, 3, 1024, 1024), device=dev, generator=g)
if which == "macenko":
    he, maxc = ops.macenko_fit(ref)
    for _ in range(3):
        out = ops.macenko_transform(src, he, maxc, unit=True)
else:
    mean, std = ops.reinhard_fit(ref)
    for _ in range(3):
        out = ops.reinhard_transform(src, mean, std)
torch.cuda.synchronize()
print("done", float(out.mean()))
