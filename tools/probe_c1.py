#!/usr/bin/env python
"""C1 (README quick-start: Reinhard fit 1x3x512x512 + transform 10x3x512x512 float32) latency probe (development tool)."""
import os
import sys
from pathlib import Path

os.environ["SX_ENABLE_TUNING"] = "1"
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from stainx_b200 import Reinhard, _native as nv, ops  # noqa: E402

dev = torch.device("cuda:0")
lib = nv.lib()
g = torch.Generator(device=dev).manual_seed(42)
ref = torch.rand((1, 3, 512, 512), device=dev, generator=g)
src = torch.rand((10, 3, 512, 512), device=dev, generator=g)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, steps=200, cold=True):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    evs = []
    for _ in range(steps):
        if cold:
            flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in evs) / steps * 1e3


mean, std = ops.reinhard_fit(ref)
for tables in (101, 100):
    lib.sx_reinhard_set_tuning(tables)
    for ctas in (1, 2, 3, 4, 6):
        lib.sx_reinhard_set_tuning(ctas)
        t_fit = timeit(lambda: ops.reinhard_fit(ref))
        t_tr = timeit(lambda: ops.reinhard_transform(src, mean, std))
        t_api = timeit(lambda: Reinhard(device=dev, backend="torch_cuda").fit(ref).transform(src))
        t_warm = timeit(lambda: Reinhard(device=dev, backend="torch_cuda").fit(ref).transform(src), cold=False)
        print(f"tables={'on' if tables == 101 else 'off'} ctas/SM={ctas}: fit {t_fit:6.1f} us  transform {t_tr:6.1f} us  API fit+transform {t_api:6.1f} us (L2 flushed) / {t_warm:6.1f} us (back to back)", flush=True)
