#!/usr/bin/env python
"""Times the Macenko transform of BASELINE config c3 / c5 (development tool; run on the GPU box; environment knobs such
as SX_L1_PREF / SX_L1_PCT are read by the library once per process).

    python tools/probe_c3.py [label]
"""
import os
import sys
from pathlib import Path

os.environ["SX_ENABLE_TUNING"] = "1"
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from stainx_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
label = sys.argv[1] if len(sys.argv) > 1 else ""


def timeit(fn, steps=30, warm=5):
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


g = torch.Generator(device=dev).manual_seed(43)
src = torch.rand((64, 3, 1024, 1024), device=dev, generator=g)
he, maxc = ops.macenko_fit(src[:1])
t3 = timeit(lambda: ops.macenko_transform(src, he, maxc, unit=True))
ws = ops.MacenkoWorkspace(64, dev)
ws.begin()
tm = timeit(lambda: ws.moments(src, pooled=False))
tm21 = timeit(lambda: ws.moments(src[:21], pooled=False))
del src
big = (torch.rand((32, 3, 2048, 2048), device=dev, generator=g) * 255).round().to(torch.uint8)
t5 = timeit(lambda: ops.macenko_transform(big, he, maxc, unit=True))
print(f"{label:28s} c3 {t3:7.1f} us   c5 {t5:7.1f} us   moments f32 x64 {tm:6.1f} us, x21 {tm21:6.1f} us", flush=True)
