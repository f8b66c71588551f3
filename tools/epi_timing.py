#!/usr/bin/env python
"""Epilogue timing of the Macenko pipeline (development tool; build with SX_EXTRA_NVCC_FLAGS=-DSX_MK_TIMING)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import ctypes  # noqa: E402

import torch  # noqa: E402

from stainx_b200 import _native as nv  # noqa: E402
from stainx_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(43)
n = 64
src = torch.rand((n, 3, 1024, 1024), device=dev, generator=g)
ref = torch.rand((1, 3, 1024, 1024), device=dev, generator=g)
he, maxc = ops.macenko_fit(ref)
ws = ops.MacenkoWorkspace(n, dev)
out = torch.empty_like(src)
lib = nv.lib()
for _ in range(3):
    nv.check(lib.sx_macenko_transform(ops._ptr(src), 1, n, 1024, 1024, ops._ptr(he), ops._ptr(maxc), ops._ptr(out), 1, ctypes.c_float(1 / 255.0), ops._ptr(ws.buffer), ws.nbytes, ops._stream(dev)), "t")
torch.cuda.synchronize()
c = ws.region("counters").cpu()
m = ws.region("moments").cpu()
t_sample = m[:, 11].view(torch.int64)
t0 = c[:, 4].min()
print("slot: epi_start  basis  sample  bracket  store   (us; start relative to the earliest epilogue)")
for i in range(0, n, 8):
    s = c[i]
    print(f"{i:3d}: {(s[4]-t0)/1e3:8.2f} {(s[5]-s[4])/1e3:7.2f} {(t_sample[i]-s[5])/1e3:7.2f} {(s[6]-t_sample[i])/1e3:7.2f} {(s[7]-s[6])/1e3:7.2f}")
print("last epilogue end - first epilogue start:", float(c[:, 7].max() - t0) / 1e3, "us")
