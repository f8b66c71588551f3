#!/usr/bin/env python
"""Macenko persistent pipeline probe (development tool; run on the GPU box).

    python tools/probe_pipeline.py [quick]

Checks the pipeline against the phase-level chain (same arithmetic, one launch per step), exercises the exact
recovery path with forced bracket misses, and sweeps the schedule gap / rows per tile on the BASELINE shapes.
"""
from __future__ import annotations

import json
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
PEAK = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"]) if (ROOT / "MEASURED_PEAKS.json").exists() else 6540.2
lib = nv.lib()


def tune(phase=0, miss=0, gap=0, rows=0):
    assert lib.sx_macenko_set_tuning(-1, phase | (miss << 1) | (gap << 4) | (rows << 16)) == 0


def timeit(fn, steps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


def report(name, ms, nbytes):
    gbs = nbytes / (ms / 1e3) / 1e9
    print(f"{name:64s} {ms*1e3:9.1f} us  {gbs:8.1f} GB/s  {gbs/PEAK*100:5.1f}%", flush=True)


def check(name, src, he, maxc, unit):
    tune(phase=1)
    want = ops.macenko_transform(src, he, maxc, unit=unit)
    tune()
    got = ops.macenko_transform(src, he, maxc, unit=unit)
    tune(miss=1)
    rec = ops.macenko_transform(src, he, maxc, unit=unit)
    tune()
    torch.cuda.synchronize()
    d1 = float((got.float() - want.float()).abs().max())
    d2 = float((rec.float() - want.float()).abs().max())
    print(f"{name:40s} pipeline vs phase chain: max|d| = {d1:.3e}; forced-miss recovery vs phase chain: {d2:.3e}", flush=True)
    return d1, d2


quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
g = torch.Generator(device=dev).manual_seed(43)
ref = torch.rand((1, 3, 1024, 1024), device=dev, generator=g)
he, maxc = ops.macenko_fit(ref)

# ---- correctness on small / odd shapes first (a hang here is cheap)
for shape, dt in (((3, 3, 64, 64), "f32"), ((2, 3, 321, 199), "f32"), ((5, 3, 256, 512), "u8"), ((1, 3, 1, 7), "f32"), ((7, 3, 130, 131), "u8"), ((4, 3, 1024, 1024), "f32"), ((4, 3, 1024, 1024), "u8")):
    x = torch.rand(shape, device=dev, generator=g)
    if dt == "u8":
        x = (x * 255).to(torch.uint8)
    for unit in (False, True):
        check(f"{shape} {dt} unit={unit}", x, he, maxc, unit)

src = torch.rand((64, 3, 1024, 1024), device=dev, generator=g)
px = 64 * 1024 * 1024
check("64x1024^2 f32", src, he, maxc, True)
report("macenko transform f32 64x1024^2 pipeline (default)", timeit(lambda: ops.macenko_transform(src, he, maxc, unit=True)), 24 * px)
tune(phase=1)
report("macenko transform f32 64x1024^2 phase chain", timeit(lambda: ops.macenko_transform(src, he, maxc, unit=True)), 24 * px)
tune()
if not quick:
    for gap in (1, 3, 5, 7, 9):
        for rows in (1, 2, 4):
            tune(gap=gap, rows=rows)
            report(f"  f32 64x1024^2 gap={gap} rows/tile={rows}", timeit(lambda: ops.macenko_transform(src, he, maxc, unit=True), steps=6, warm=2), 24 * px)
    tune()
del src
src8 = (torch.rand((64, 3, 1024, 1024), device=dev, generator=g) * 255).to(torch.uint8)
report("macenko transform u8->u8 64x1024^2 pipeline", timeit(lambda: ops.macenko_transform(src8, he, maxc, unit=False)), 6 * px)
report("macenko transform u8->f32 unit 64x1024^2 pipeline", timeit(lambda: ops.macenko_transform(src8, he, maxc, unit=True)), 15 * px)
del src8
tiles = (torch.rand((32, 3, 2048, 2048), device=dev, generator=g) * 255).to(torch.uint8)
pxt = 32 * 2048 * 2048
report("macenko transform u8 2048^2 x32 -> f32 unit (C5 per GPU)", timeit(lambda: ops.macenko_transform(tiles, he, maxc, unit=True), steps=5), 15 * pxt)
if not quick:
    for gap in (1, 3, 5):
        for rows in (1, 2, 4):
            tune(gap=gap, rows=rows)
            report(f"  C5 gap={gap} rows/tile={rows}", timeit(lambda: ops.macenko_transform(tiles, he, maxc, unit=True), steps=4, warm=2), 15 * pxt)
    tune()
