#!/usr/bin/env python
"""A/B of the high-priority helper streams of the Macenko transform (development tool; run on the GPU box).

    python tools/probe_helper.py

Times sx_macenko_transform / sx_macenko_fit_transform on the BASELINE shapes with the per-image kernels on the chains'
own streams (tuning bit 2) and on the helper streams (default), and checks that the outputs are equal
bit for bit (same kernels on the same data).
"""
from __future__ import annotations

import os
import sys
from pathlib import Path

os.environ["SX_ENABLE_TUNING"] = "1"
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

from stainx_b200 import _native as nv  # noqa: E402
from stainx_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
lib = nv.lib()


def timeit(fn, steps=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps * 1e3


def ab(name, fn, chains=(3,)):
    outs = {}
    for c in chains:
        for helper in (0, 1, 0, 1):  # twice: the run-to-run spread of these timings is 1-3 %
            lib.sx_macenko_set_tuning(-1, (c << 4) | (0 if helper else 4))
            us = timeit(fn)
            outs[(c, helper)] = fn()
            print(f"{name:44s} chains={c} helper={helper} {us:8.1f} us", flush=True)
    vals = list(outs.values())
    same = all(all(torch.equal(a, b) for a, b in zip(vals[0], v)) for v in vals[1:])
    print(f"{name:44s} outputs identical across settings: {same}", flush=True)
    lib.sx_macenko_set_tuning(-1, 0)
    return same


g = torch.Generator(device=dev).manual_seed(43)
ok = True
src = torch.rand((64, 3, 1024, 1024), device=dev, generator=g)
he, maxc = ops.macenko_fit(torch.rand((1, 3, 1024, 1024), device=dev, generator=g))
ok &= ab("c3 f32 64x1024^2 -> f32 [0,1]", lambda: (ops.macenko_transform(src, he, maxc, unit=True),), chains=(2, 3, 4))
ok &= ab("fit_transform f32 64x1024^2", lambda: ops.macenko_fit_transform(src, unit=True))
del src
u8 = (torch.rand((64, 3, 1024, 1024), device=dev, generator=g) * 255).round().to(torch.uint8)
he8, maxc8 = ops.macenko_fit(u8[:1])
ok &= ab("u8 64x1024^2 -> u8", lambda: (ops.macenko_transform(u8, he8, maxc8, unit=False),))
del u8
big = (torch.rand((32, 3, 2048, 2048), device=dev, generator=g) * 255).round().to(torch.uint8)
ok &= ab("c5 u8 32x2048^2 -> f32 [0,1]", lambda: (ops.macenko_transform(big, he8, maxc8, unit=True),), chains=(3, 4))
print("HELPER CHECK", "OK" if ok else "MISMATCH")
