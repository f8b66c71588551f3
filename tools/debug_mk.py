#!/usr/bin/env python
"""Dump the per-slot Macenko workspace after every phase for a golden fixture (development tool)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from stainx_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
name = sys.argv[1] if len(sys.argv) > 1 else "macenko_noise_u8"
g = np.load(ROOT / "tests" / "golden" / f"{name}.npz")
src = torch.from_numpy(g["src"]).to(dev)
n = src.shape[0]
ws = ops.MacenkoWorkspace(n, dev)
ws.begin()
ws.moments(src, pooled=False)
ws.basis(0, n, True)
ws.moments_fallback(src)
torch.cuda.synchronize()
print("moments", ws.region("moments").cpu().numpy())
print("odrange", ws.region("odrange").cpu().numpy())
for stage in (0, 1):
    for level in (0, 1):
        ws.hist(src, False, stage, level)
        torch.cuda.synchronize()
        h = ws.region("hist1" if level == 0 else "hist2").cpu().numpy()
        print(f"stage {stage} level {level}: hist sums", h.reshape(n, 2, -1).sum(-1), "counters", ws.region("counters").cpu().numpy()[:, :4])
        if level == 1:
            vmin = ws.region("vmin").cpu().numpy()
            print("   finite vmin cells", np.isfinite(vmin).reshape(n, 2, -1).sum(-1))
        ws.select(0, n, stage, level)
        torch.cuda.synchronize()
        print("   fit", ws.region("fit").cpu().numpy(), "status", ws.region("status").cpu().numpy()[:, 0])
