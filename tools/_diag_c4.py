import sys, numpy as np, torch
sys.path.insert(0, ".")
from oracle import oracle as ox
from stainx_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(43)
src = torch.rand((16, 3, 512, 512), device=dev, generator=g)
for npool in (1, 4, 16):
    he, maxc = ops.macenko_fit(src[:npool].contiguous())
    fits = [ox.macenko_fit(src[:npool].cpu().numpy(), mid_sign=s) for s in (1, -1)]
    print("pool", npool, "HE gpu", he.flatten().tolist(), "maxc", maxc.tolist())
    for f in fits:
        print("   oracle HE diff %.2e maxc rel %.2e" % (np.abs(he.cpu().numpy() - f[0]).max(), np.abs(maxc.cpu().numpy() / f[1] - 1).max()), f[0].flatten().round(4).tolist(), f[1].tolist())
    out = ops.macenko_transform(src[:2].contiguous(), he, maxc, unit=True).cpu().numpy().astype(np.float64)
    cand = [ox.macenko_transform(src[:2].cpu().numpy(), he.cpu().numpy(), maxc.cpu().numpy(), mid_signs=[s, s]).astype(np.float64) / 255.0 for s in (1, -1)]
    for i in range(2):
        ds = [np.abs(out[i] - c[i]) for c in cand]
        print("   img", i, "max|d| per sign", [float(d.max()) for d in ds], "frac>1e-3", [float((d > 1e-3).mean()) for d in ds])
