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


def tune(phase=0, miss=0, gap=0, rows=0, timing=0):
    assert lib.sx_macenko_set_tuning(-1, phase | (miss << 1) | (timing << 3) | (gap << 4) | (rows << 16)) == 0


def timing_report(src, he, maxc, label, **kw):
    """One transform with the development counters on: where the service and streaming CTAs spend their clocks."""
    import ctypes

    n, _, h, w = src.shape
    tune(timing=1, **kw)
    nbytes = int(lib.sx_macenko_workspace_bytes(n))
    ws = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    out = torch.empty(src.shape, dtype=torch.float32, device=dev)
    vp = ctypes.c_void_p
    for _ in range(2):
        rc = lib.sx_macenko_transform(vp(src.data_ptr()), 0 if src.dtype == torch.uint8 else 1, n, h, w, vp(he.data_ptr()), vp(maxc.data_ptr()), vp(out.data_ptr()), 1, ctypes.c_float(1.0 / 255.0),
                                      vp(ws.data_ptr()), nbytes, vp(torch.cuda.current_stream().cuda_stream))
        assert rc == 0
    torch.cuda.synchronize()
    off, size = ctypes.c_int64(), ctypes.c_int64()
    lib.sx_macenko_region(n, 9, ctypes.byref(off), ctypes.byref(size))
    pad = ws[off.value : off.value + size.value].view(torch.int32)[3:].cpu().tolist()
    us = lambda c: c * 64 / 1965.0  # noqa: E731
    soff, ssize = ctypes.c_int64(), ctypes.c_int64()
    lib.sx_macenko_region(n, 8, ctypes.byref(soff), ctypes.byref(ssize))
    status = ws[soff.value : soff.value + ssize.value].view(torch.int32).view(n, 4).cpu()
    print(f"[{label}] service wait/work us: " + ", ".join(f"k{r}: {us(pad[2*r]):.0f}/{us(pad[2*r+1]):.0f}" for r in range(3)) +
          " | stage us (load, basis|select, sample+bracket, publish): " + "; ".join("k%d: " % r + "/".join(f"{us(pad[16+4*r+j]):.0f}" for j in range(4)) for r in range(2)) +
          " | sample_and_bracket us (zero, sample, prefix, bracket): " + "; ".join("/".join(f"{us(pad[24+4*r+j]):.0f}" for j in range(4)) for r in range(2)) +
          f" | streaming: {pad[11]} CTAs, {pad[10]} tiles, mean wait {us(pad[8])/max(pad[11],1):.0f} us of {us(pad[9])/max(pad[11],1):.0f} us | misses {int(status[:,0].ne(0).sum())}, recovered stages {int(status[:,1].sum())}", flush=True)
    tune()


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
timing_report(src, he, maxc, "f32 64x1024^2 default")
timing_report(src, he, maxc, "f32 64x1024^2 gap 7", gap=7)
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
