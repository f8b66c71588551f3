#!/usr/bin/env python
"""uint8 Macenko transform workload for ncu captures (development tool)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from stainx_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(43)
n = 32
src8 = (torch.rand((n, 3, 1024, 1024), device=dev, generator=g) * 255).to(torch.uint8)
ref = torch.rand((1, 3, 1024, 1024), device=dev, generator=g)
he, maxc = ops.macenko_fit(ref)
for _ in range(2):
    out = ops.macenko_transform(src8, he, maxc, unit=True)
torch.cuda.synchronize()
print("done", float(out.mean()))
