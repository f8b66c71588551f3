#!/usr/bin/env python
"""One Macenko transform per dtype for an ncu launch list (development tool)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from stainx_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(43)
n = 64
src = torch.rand((n, 3, 1024, 1024), device=dev, generator=g)
ref = torch.rand((1, 3, 1024, 1024), device=dev, generator=g)
he, maxc = ops.macenko_fit(ref)
for _ in range(2):
    out = ops.macenko_transform(src, he, maxc, unit=True)
src8 = (src * 255).to(torch.uint8)
del src, out
for _ in range(2):
    out = ops.macenko_transform(src8, he, maxc, unit=True)
torch.cuda.synchronize()
print("done")
