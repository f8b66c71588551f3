#!/usr/bin/env python
"""HostStream end-to-end probe (development tool): HM uint8 64x3x1024x1024, depth and in-flight limit sweep against the copy-only ceiling."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from stainx_b200 import HistogramMatching  # noqa: E402
from stainx_b200.ingest import HostStream, bind_host_thread_to_device  # noqa: E402

dev = torch.device("cuda:0")
print("bound to cores:", len(bind_host_thread_to_device(0) or []))
g = torch.Generator().manual_seed(1)
ref = (torch.rand(1, 3, 1024, 1024, generator=g) * 255).to(torch.uint8)
host_in = (torch.rand(64, 3, 1024, 1024, generator=g) * 255).to(torch.uint8).pin_memory()
outs = [torch.empty_like(host_in).pin_memory() for _ in range(3)]
hm = HistogramMatching(device=dev, backend="torch_cuda").fit(ref.to(dev))
K = 24


def run(depth, limit):
    pipe = HostStream(hm, device=dev, depth=depth)
    pipe._max_inflight = limit
    def go():
        ts = [pipe.submit(host_in, outs[i % 3]) for i in range(K)]
        for t in ts:
            t.wait()
    go()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); go(); pipe.synchronize(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / K


s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
dev_in, dev_out = torch.empty_like(host_in, device=dev), torch.empty_like(host_in, device=dev)
def copies():
    for _ in range(K):
        with torch.cuda.stream(s1):
            dev_in.copy_(host_in, non_blocking=True)
        with torch.cuda.stream(s2):
            outs[0].copy_(dev_out, non_blocking=True)
copies(); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); copies(); s1.synchronize(); s2.synchronize(); b.record(); torch.cuda.synchronize()
print(f"copy-only (both directions): {a.elapsed_time(b)/K:.3f} ms per batch")
for depth in (2, 3, 4):
    for limit in (depth + 1, 2 * depth + 2, 1000):
        print(f"depth {depth} in-flight limit {limit}: {run(depth, limit):.3f} ms per batch", flush=True)
