#!/usr/bin/env python
"""Small HistogramMatching workload for ncu captures (development tool)."""
import os
import sys

os.environ["SX_ENABLE_TUNING"] = "1"
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from stainx_b200 import _native as nv  # noqa: E402
from stainx_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(43)
src = (torch.rand((64, 3, 1024, 1024), device=dev, generator=g) * 255).round().to(torch.uint8)
ref = (torch.rand((1, 3, 1024, 1024), device=dev, generator=g) * 255).round().to(torch.uint8)
ref_hist = ops.hm_fit(ref)
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 5
nv.lib().sx_hm_set_tuning(mode, -1, -1)
for _ in range(3):
    out = ops.hm_transform(src, ref_hist)
const = torch.full_like(src, 200)
counts = ops.hm_hist(const)
torch.cuda.synchronize()
print("done", int(out.sum()), int(counts.sum()))
