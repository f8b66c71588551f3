#!/usr/bin/env python
"""Timeline of the Macenko transform chains (development tool; run on the GPU box): completion time of every kernel of
one sx_macenko_transform call, per chain, from timing events behind each launch (sx_macenko_trace).

    python tools/trace_mk.py [f32|u8|c5]
"""
from __future__ import annotations

import ctypes
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
what = sys.argv[1] if len(sys.argv) > 1 else "f32"
g = torch.Generator(device=dev).manual_seed(43)
if what == "f32":
    src = torch.rand((64, 3, 1024, 1024), device=dev, generator=g)
elif what == "u8":
    src = (torch.rand((64, 3, 1024, 1024), device=dev, generator=g) * 255).round().to(torch.uint8)
else:
    src = (torch.rand((32, 3, 2048, 2048), device=dev, generator=g) * 255).round().to(torch.uint8)
he, maxc = ops.macenko_fit(src[:1])
unit = what != "u8"
buf = ctypes.create_string_buffer(1 << 16)
for chains in (3, 4):
    for helper in (0, 1):
        lib.sx_macenko_set_tuning(-1, (chains << 4) | (0 if helper else 4))
        for _ in range(3):
            ops.macenko_transform(src, he, maxc, unit=unit)
        torch.cuda.synchronize()
        lib.sx_macenko_trace(1, None, 0)
        ops.macenko_transform(src, he, maxc, unit=unit)
        lib.sx_macenko_trace(0, buf, len(buf))
        rows = [r.split() for r in buf.value.decode().strip().splitlines()]
        print(f"--- {what} chains={chains} helper={helper}: completion times (us) per chain")
        names = []
        for r in rows:
            if r[1] not in names:
                names.append(r[1])
        print(f"{'kernel':16s}" + "".join(f"{'chain ' + str(c):>10s}" for c in range(chains)))
        for nme in names:
            print(f"{nme:16s}" + "".join(f"{next((float(r[2]) for r in rows if r[1] == nme and int(r[0]) == c), float('nan')):10.1f}" for c in range(chains)))
lib.sx_macenko_set_tuning(-1, 0)
