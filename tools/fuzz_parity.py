#!/usr/bin/env python
"""Randomised parity fuzz (development tool; run on the GPU box): random shapes, dtypes and layouts for all three methods
against the CPU oracle at the bars of the parity tests.

    python tools/fuzz_parity.py [iterations] [seed]
"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import oracle as ox  # noqa: E402
from stainx_b200 import HistogramMatching, Macenko, Reinhard  # noqa: E402
from tests.helpers import he_tile  # noqa: E402

dev = torch.device("cuda:0")
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
DT = {"u8": torch.uint8, "f32": torch.float32, "f16": torch.float16, "bf16": torch.bfloat16}
ULP = {"f16": 2.0**-11, "bf16": 2.0**-8, "f32": 0.0}
bad = 0
for it in range(iters):
    n, h, w = int(rng.integers(1, 6)), int(rng.integers(1, 300)), int(rng.integers(1, 300))
    dt = str(rng.choice(["u8", "f32", "f16", "bf16"]))
    method = str(rng.choice(["hm", "hm_nhwc", "reinhard", "macenko"]))
    if method == "macenko" and h * w < 256:
        h, w = h + 16, w + 16
    gamma = float(rng.uniform(0.4, 2.5))
    g = torch.Generator().manual_seed(int(rng.integers(1 << 30)))
    if method == "macenko":  # stain-like content (noise is ill-posed for Macenko)
        big = he_tile(max(h, 16), max(w, 16), int(rng.integers(1 << 20)), float(rng.uniform(0.8, 1.2)))
        src8 = torch.cat([he_tile(max(h, 16), max(w, 16), int(rng.integers(1 << 20)), float(rng.uniform(0.8, 1.2))) for _ in range(n)])[:, :, :h, :w].contiguous()
        ref8 = big[:, :, :h, :w].contiguous()
    else:
        src8 = (torch.rand((n, 3, h, w), generator=g).pow(gamma) * 255).round().to(torch.uint8)
        ref8 = (torch.rand((1, 3, h, w), generator=g).pow(1.0 / gamma) * 255).round().to(torch.uint8)
    src = src8 if dt == "u8" else (src8.float() / 255.0).to(DT[dt])
    ref = ref8 if dt == "u8" else (ref8.float() / 255.0).to(DT[dt])
    src_o = src.numpy() if dt in ("u8", "f32") else src.float().numpy()
    ref_o = ref.numpy() if dt in ("u8", "f32") else ref.float().numpy()
    try:
        if method.startswith("hm"):
            nhwc = method == "hm_nhwc"
            nm = HistogramMatching(device=dev, backend="torch_cuda", channel_axis=-1 if nhwc else 1)
            f = (lambda t: t.permute(0, 2, 3, 1).contiguous()) if nhwc else (lambda t: t)
            out = nm.fit(f(ref).to(dev)).transform(f(src).to(dev)).cpu()
            if nhwc:
                out = out.permute(0, 3, 1, 2)
            want = torch.from_numpy(ox.hm_transform(src_o, ox.hm_fit(ref_o))).to(DT[dt])
            ok = torch.equal(out, want)
        elif method == "reinhard":
            nm = Reinhard(device=dev, backend="torch_cuda").fit(ref.to(dev))
            out = nm.transform(src.to(dev)).cpu()
            want = ox.reinhard_transform(src_o, nm._reference_mean.cpu().numpy(), nm._reference_std.cpu().numpy())
            if dt == "u8":
                d = np.abs(out.numpy().astype(np.int32) - want.astype(np.int32))
                ok = d.max() <= 1 and ((d > 0).mean() < 0.01 or d.size < 400)
            else:
                ok = float((out.float() - torch.from_numpy(want).to(DT[dt]).float()).abs().max()) <= 1e-3 + ULP[dt]
        else:
            nm = Macenko(device=dev, backend="torch_cuda").fit(ref.to(dev))
            he, maxc = ox.macenko_fit(ref_o)
            ok = np.abs(nm._stain_matrix.cpu().numpy() - he).max() <= 1e-4
            out = nm.transform(src.to(dev)).cpu()
            want = ox.macenko_transform(src_o, nm._stain_matrix.cpu().numpy(), nm._target_max_conc.cpu().numpy())
            if dt == "u8":
                d = np.abs(out.numpy().astype(np.int32) - want.astype(np.int32))
                ok = ok and d.max() <= 1 and ((d > 0).mean() < 0.01 or d.size < 400)
            else:
                tol = 1e-3 * 255.0 + {"f32": 0.0, "f16": 0.125, "bf16": 1.0}[dt]
                ok = ok and float((out.float() - torch.from_numpy(want).to(DT[dt]).float()).abs().max()) <= tol
    except Exception as exc:  # noqa: BLE001
        ok = False
        print("EXC", type(exc).__name__, exc)
    if not ok:
        bad += 1
        print(f"MISMATCH it={it} method={method} dt={dt} shape=({n},3,{h},{w})", flush=True)
print(f"fuzz: {iters} cases, {bad} mismatches")
sys.exit(1 if bad else 0)
