#!/usr/bin/env python
"""Phase timeline of the fused Macenko kernel for the first image of team 0 (development tool)."""
import ctypes
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from stainx_b200 import _native as nv  # noqa: E402
from stainx_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(43)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
teams = int(sys.argv[2]) if len(sys.argv) > 2 else 0
src = torch.rand((n, 3, 1024, 1024), device=dev, generator=g)
ref = torch.rand((1, 3, 1024, 1024), device=dev, generator=g)
he, maxc = ops.macenko_fit(ref)
lib = nv.lib()
lib.sx_macenko_set_fused(1, teams)
for _ in range(2):
    ops.macenko_transform(src, he, maxc, unit=True)
buf = torch.zeros(64, dtype=torch.int64, device=dev)
lib.sx_macenko_set_timeline(ctypes.c_void_p(buf.data_ptr()))
ops.macenko_transform(src, he, maxc, unit=True)
torch.cuda.synchronize()
lib.sx_macenko_set_timeline(ctypes.c_void_p(0))
t = buf.cpu().tolist()
k = t[0]
stamps = t[1:1 + k]
names = ["start", "moments", "E1 basis", "sample A", "E2 bracket A", "resolve A", "E3 select A", "sample C", "E4 bracket C", "resolve C", "E5 select C", "apply"]
print(f"n={n} teams<={teams}: {k} stamps")
for i in range(1, min(k, 40)):
    nm = names[i % 12] if i % 12 < len(names) else "?"
    print(f"  {nm:14s} +{(stamps[i] - stamps[i - 1]) / 1e3:8.2f} us   (t={(stamps[i] - stamps[0]) / 1e3:8.2f})")
