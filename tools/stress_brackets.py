#!/usr/bin/env python
"""Repeat the bracket test scenario and report any miss / non-finite fit (development tool)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from stainx_b200 import ops  # noqa: E402
from tests.helpers import he_batch  # noqa: E402

cuda = torch.device("cuda:0")
g = torch.Generator().manual_seed(3)
noise = torch.rand((2, 3, 512, 512), generator=g)
tiles = he_batch(2, 512, 512).float() / 255.0
yy, xx = torch.meshgrid(torch.arange(512), torch.arange(512), indexing="ij")
stripes = (0.25 + 0.5 * ((xx // 4 + yy // 64) % 2).float()).expand(1, 3, 512, 512).clone()
stripes[:, 0] *= 0.8
stripes += 0.05 * torch.rand((1, 3, 512, 512), generator=g)
sparse = torch.full((1, 3, 512, 512), 0.97)
sparse[:, :, 100:108, 200:232] = tiles[0, :, 100:108, 200:232]
batch = torch.cat([noise, tiles, stripes.clamp(0, 1), sparse]).contiguous().to(cuda)
bad = 0
fits = []
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 200):
    ws = ops.MacenkoWorkspace(batch.shape[0], cuda)
    ws.begin()
    ws.moments(batch, False)
    ws.basis(0, batch.shape[0], True)
    ws.moments_fallback(batch)
    for stage in (0, 1):
        for level in (0, 1):
            ws.hist(batch, False, stage, level)
            ws.select(0, batch.shape[0], stage, level)
    st = ws.region("status").cpu()
    fit = ws.region("fit").cpu()
    fits.append(fit.clone())
    if int(st.abs().sum()) != 0 or not bool(torch.isfinite(fit).all()):
        bad += 1
        print("rep", rep, "status", st.tolist(), "fit", fit.tolist(), "counters", ws.region("counters").cpu().tolist())
f = torch.stack(fits)
print("bad", bad, "fit spread per slot:", (f.max(0).values - f.min(0).values).abs().max(1).values.tolist())
