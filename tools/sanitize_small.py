#!/usr/bin/env python
"""Small run of every kernel family for compute-sanitizer (development tool; one tool per call)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from stainx_b200 import HistogramMatching, Macenko, Reinhard  # noqa: E402
from tests.helpers import he_tile  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(1)
ref8 = he_tile(96, 128, 42)
src8 = torch.cat([he_tile(96, 128, s, 1.1) for s in (1, 2, 3)])
flat = src8.clone()
flat[:, :, :, :96] = flat[:, :, :1, :1]  # ties: overflow the hit queue
odd = (torch.rand(2, 3, 33, 35, generator=g) * 255).round().to(torch.uint8)  # scalar (unaligned) path
for src in (src8, flat, odd, src8.float() / 255.0, flat.float() / 255.0, odd.float() / 255.0):
    ref = ref8 if src.dtype == torch.uint8 else ref8.float() / 255.0
    for cls in (HistogramMatching, Reinhard, Macenko):
        n = cls(device=dev, backend="torch_cuda").fit(ref.to(dev))
        out = n.transform(src.to(dev))
        n2 = cls(device=dev, backend="torch_cuda")
        out2 = n2.fit_transform(src.to(dev))
torch.cuda.synchronize()
# a batch large enough for the TMA-fed histogram kernel (>= 8 MB) and the multi-row pipeline
big = (torch.rand(4, 3, 1024, 1024, generator=g) * 255).round().to(torch.uint8).to(dev)
hm = HistogramMatching(device=dev, backend="torch_cuda").fit(big[:1])
o = hm.transform(big)
mk = Macenko(device=dev, backend="torch_cuda").fit(big[:1])
o = mk.transform(big[:2])
torch.cuda.synchronize()
print("sanitize run ok")
