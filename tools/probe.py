#!/usr/bin/env python
"""Per-kernel timing probe (development tool; run on the GPU box).

    python tools/probe.py [hm] [reinhard] [macenko]

Times every kernel phase of each method with CUDA events on the BASELINE shapes and prints the
achieved algorithmic GB/s, sweeping the tuning knobs exposed by the library.
"""
from __future__ import annotations

import json
import os
import sys
from pathlib import Path

os.environ["SX_ENABLE_TUNING"] = "1"  # the launch-geometry hooks are inert without it

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

from stainx_b200 import _native as nv  # noqa: E402
from stainx_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
PEAK = 6540.2
try:
    PEAK = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
except Exception:
    pass


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
    print(f"{name:58s} {ms*1e3:9.1f} us  {gbs:8.1f} GB/s  {gbs/PEAK*100:5.1f}% of {PEAK:.0f}", flush=True)


def probe_hm():
    g = torch.Generator(device=dev).manual_seed(43)
    src = (torch.rand((64, 3, 1024, 1024), device=dev, generator=g) * 255).round().to(torch.uint8)
    ref = (torch.rand((1, 3, 1024, 1024), device=dev, generator=g) * 255).round().to(torch.uint8)
    px = 64 * 1024 * 1024
    ref_hist = ops.hm_fit(ref)
    lib = nv.lib()
    counts = torch.zeros((3, 256), dtype=torch.int64, device=dev)
    want = ops.hm_hist(src)
    for mode, ctas in ((0, 8), (5, 1), (9, 1), (6, 1)):
        lib.sx_hm_set_tuning(mode, ctas, -1)
        ok = torch.equal(ops.hm_hist(src), want)
        report(f"hm hist u8 planar mode={mode} ctas/sm={ctas} ok={ok}", timeit(lambda: ops.hm_hist(src, counts=counts)), 3 * px)
    const = torch.full_like(src, 200)
    lib.sx_hm_set_tuning(0, 8, -1)
    want_const = ops.hm_hist(const)
    for mode, ctas in ((0, 8), (5, 1)):
        lib.sx_hm_set_tuning(mode, ctas, -1)
        ok = torch.equal(ops.hm_hist(const), want_const)
        report(f"hm hist u8 CONSTANT image mode={mode} ok={ok}", timeit(lambda: ops.hm_hist(const, counts=counts)), 3 * px)
    del const
    lib.sx_hm_set_tuning(5, 8, -1)
    lut = ops.hm_build_lut(ops.hm_hist(src), px, ops.hm_ref_cdf(ref_hist))
    for ctas in (4, 8, 16):
        lib.sx_hm_set_tuning(-1, -1, ctas)
        report(f"hm apply u8 planar ctas/sm={ctas}", timeit(lambda: ops.hm_apply(src, lut)), 6 * px)
    lib.sx_hm_set_tuning(5, 8, 16)
    for mode in (5, 8):
        lib.sx_hm_set_tuning(mode, 8, 16)
        report(f"hm transform u8 (hist+lut+apply) hist mode={mode}", timeit(lambda: ops.hm_transform(src, ref_hist), steps=200, warm=10), 9 * px)
    lib.sx_hm_set_tuning(5, 8, 16)
    nhwc = src.permute(0, 2, 3, 1).contiguous()
    report("hm transform u8 NHWC", timeit(lambda: ops.hm_transform(nhwc, ref_hist, nv.SX_NHWC)), 9 * px)
    report("hm hist u8 NHWC", timeit(lambda: ops.hm_hist(nhwc, nv.SX_NHWC, counts=counts)), 3 * px)
    report("hm apply u8 NHWC", timeit(lambda: ops.hm_apply(nhwc, lut, nv.SX_NHWC)), 6 * px)
    del nhwc
    srcf = src[:32].float() / 255
    report("hm hist f32 planar (32 img)", timeit(lambda: ops.hm_hist(srcf, counts=counts)), 12 * px / 2)
    report("hm transform f32 (32 img)", timeit(lambda: ops.hm_transform(srcf, ref_hist)), 36 * px / 2)
    # reference points: plain device copy of the same bytes
    dst = torch.empty_like(src)
    report("torch copy_ u8 201MB (read+write)", timeit(lambda: dst.copy_(src)), 6 * px)


def probe_reinhard():
    g = torch.Generator(device=dev).manual_seed(43)
    lib = nv.lib()
    px = 64 * 1024 * 1024
    src = torch.rand((64, 3, 1024, 1024), device=dev, generator=g)
    mean = torch.tensor([150.0, 140.0, 130.0], device=dev)
    std = torch.tensor([40.0, 10.0, 12.0], device=dev)
    for ctas in (2, 3, 4):
        lib.sx_reinhard_set_tuning(ctas)
        report(f"reinhard stats f32 ctas/sm={ctas}", timeit(lambda: ops.reinhard_stats(src)), 12 * px)
        report(f"reinhard apply f32 ctas/sm={ctas}", timeit(lambda: ops.reinhard_apply(src, mean, std, mean, std)), 24 * px)
    lib.sx_reinhard_set_tuning(4)
    for ctas in (2, 3, 4):
        lib.sx_reinhard_set_tuning(ctas)
        report(f"reinhard transform f32 ctas/sm={ctas}", timeit(lambda: ops.reinhard_transform(src, mean, std), steps=20), 36 * px)
    lib.sx_reinhard_set_tuning(3)
    src8 = (src * 255).to(torch.uint8)
    del src
    report("reinhard stats u8", timeit(lambda: ops.reinhard_stats(src8)), 3 * px)
    report("reinhard apply u8", timeit(lambda: ops.reinhard_apply(src8, mean, std, mean, std)), 6 * px)
    report("reinhard transform u8", timeit(lambda: ops.reinhard_transform(src8, mean, std)), 9 * px)


def probe_macenko():
    g = torch.Generator(device=dev).manual_seed(43)
    lib = nv.lib()
    n = 64
    px = n * 1024 * 1024
    src = torch.rand((n, 3, 1024, 1024), device=dev, generator=g)
    ref = torch.rand((1, 3, 1024, 1024), device=dev, generator=g)
    he, maxc = ops.macenko_fit(ref)
    print("fit HE", he.flatten().tolist(), "maxC", maxc.tolist())
    ws = ops.MacenkoWorkspace(n, dev)
    out = torch.empty_like(src)

    def phases_until(stop_stage, stop_level):
        """Run the pipeline up to (not including) hist(stop_stage, stop_level)."""
        ws.begin()
        ws.moments(src, False)
        ws.basis(0, n, True)
        ws.moments_fallback(src)
        for stage in (0, 1):
            for level in (0, 1):
                if (stage, level) == (stop_stage, stop_level):
                    return
                ws.hist(src, False, stage, level)
                ws.select(0, n, stage, level)

    report("macenko begin (init workspace)", timeit(ws.begin), 0.001)
    report("macenko moments f32", timeit(lambda: ws.moments(src, False)), 12 * px)
    report("macenko basis", timeit(lambda: ws.basis(0, n, True)), 0.001)
    report("macenko moments_fallback (nothing flagged)", timeit(lambda: ws.moments_fallback(src)), 0.001)
    for stage in (0, 1):
        for level in (0, 1):
            phases_until(stage, level)
            report(f"macenko hist f32 stage={stage} level={level}", timeit(lambda: ws.hist(src, False, stage, level)), 12 * px)
            phases_until(stage, level)
            ws.hist(src, False, stage, level)
            report(f"macenko select stage={stage} level={level}", timeit(lambda: ws.select(0, n, stage, level)), 0.001)
    phases_until(2, 0)
    print("status flags:", int(ws.region("status").abs().sum()))
    for ctas in (2, 3, 4, 6, 8):
        lib.sx_macenko_set_tuning(ctas, -1)
        report(f"macenko apply f32 -> f32 unit ctas/sm={ctas}", timeit(lambda: ws.apply(src, he, maxc, out, True)), 24 * px)
        report(f"macenko transform f32 64x1024^2 ctas/sm={ctas}", timeit(lambda: ops.macenko_transform(src, he, maxc, unit=True), steps=10), 24 * px)
    lib.sx_macenko_set_tuning(4, -1)
    lib.sx_macenko_set_tuning(-1, 1)
    want = ops.macenko_transform(src, he, maxc, unit=True)
    report("macenko transform f32 64x1024^2 phase kernels", timeit(lambda: ops.macenko_transform(src, he, maxc, unit=True), steps=5), 24 * px)
    for chains in (1, 2, 3, 4):
        lib.sx_macenko_set_tuning(-1, chains << 4)
        report(f"macenko transform f32 64x1024^2 pipeline, {chains} chain(s)", timeit(lambda: ops.macenko_transform(src, he, maxc, unit=True), steps=10), 24 * px)
    lib.sx_macenko_set_tuning(-1, 0)
    got = ops.macenko_transform(src, he, maxc, unit=True)
    err = float((got - want).abs().max())
    report(f"macenko transform f32 64x1024^2 pipeline maxdiff={err:.2e}", timeit(lambda: ops.macenko_transform(src, he, maxc, unit=True), steps=10), 24 * px)
    # per-kernel times of the pipeline (events around each launch are not available through the
    # C ABI; time prefixes of the chain instead)
    del want, got
    src8 = (src * 255).to(torch.uint8)
    del src, out
    for phase in (1, 16, 32, 64):
        lib.sx_macenko_set_tuning(-1, phase)
        tag = "phase kernels" if phase == 1 else f"pipeline {phase >> 4} chain(s)"
        report(f"macenko transform u8 -> u8 {tag}", timeit(lambda: ops.macenko_transform(src8, he, maxc, unit=False), steps=5), 6 * px)
        report(f"macenko transform u8 -> f32 unit {tag}", timeit(lambda: ops.macenko_transform(src8, he, maxc, unit=True), steps=5), 15 * px)
    big = src8.reshape(16, 3, 2048, 2048)
    report("macenko transform u8 2048^2 x16 -> f32 unit pipeline", timeit(lambda: ops.macenko_transform(big, he, maxc, unit=True), steps=5), 15 * px)


if __name__ == "__main__":
    which = sys.argv[1:] or ["hm", "reinhard", "macenko"]
    print(torch.cuda.get_device_name(0), "peak", PEAK)
    if "hm" in which:
        probe_hm()
    if "reinhard" in which:
        probe_reinhard()
    if "macenko" in which:
        probe_macenko()
