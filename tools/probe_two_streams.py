#!/usr/bin/env python
"""Independent batches on two streams (development tool): does the histogram of one transform overlap the remap of another?
HistogramMatching uint8 64x3x1024x1024, Reinhard / Macenko float32 64x3x1024x1024; K transforms issued round-robin on 1, 2, 3 streams."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from stainx_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(43)
src8 = [(torch.rand((64, 3, 1024, 1024), device=dev, generator=g) * 255).round().to(torch.uint8) for _ in range(2)]
ref_hist = ops.hm_fit(src8[0][:1].contiguous())
srcf = [torch.rand((64, 3, 1024, 1024), device=dev, generator=g) for _ in range(2)]
mean, std = ops.reinhard_fit(srcf[0][:1].contiguous())
he, maxc = ops.macenko_fit(srcf[0][:1].contiguous())
K = 24


def run(fn, nstreams):
    streams = [torch.cuda.Stream(dev) for _ in range(nstreams)]
    def once():
        cur = torch.cuda.current_stream(dev)
        for s in streams:
            s.wait_stream(cur)
        for i in range(K):
            with torch.cuda.stream(streams[i % nstreams]):
                fn(i)
        for s in streams:
            cur.wait_stream(s)
    once(); once()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); once(); once(); b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / (2 * K) * 1e3


for name, fn in (("HM uint8", lambda i: ops.hm_transform(src8[i % 2], ref_hist)), ("Reinhard f32", lambda i: ops.reinhard_transform(srcf[i % 2], mean, std)),
                 ("Macenko f32", lambda i: ops.macenko_transform(srcf[i % 2], he, maxc, unit=True))):
    print(name, " ".join(f"{n} stream(s): {run(fn, n):7.1f} us per transform" for n in (1, 2, 3)), flush=True)
