"""CPU stand-in for ``stainx_b200.ops`` -- TEST-ONLY.

The sharded (multi-GPU) control flow in ``stainx_b200/backends/torch_cuda_backend.py`` is ordinary
Python around kernel phases and all-reduces.  To exercise it with ``gloo`` and world_size 2 on a
box without GPUs, the tests inject this module in place of the native ops layer.  It implements the
same phase functions on CPU tensors with numpy and the CPU oracle, including a numpy restatement of
the Macenko phase protocol (moments -> basis -> two-level order-statistic histograms), so the
regions that get all-reduced have the same meaning as in ``include/stainx_b200.h``.

Never imported by the product.
"""
from __future__ import annotations

import numpy as np
import torch

from oracle import oracle as ox

SX_NCHW, SX_NHWC = 0, 1
BINS = 4096
SHIFT = 0.75


def _nchw(images: torch.Tensor, layout: int) -> np.ndarray:
    a = images.numpy()
    return np.ascontiguousarray(a.transpose(0, 3, 1, 2)) if layout == SX_NHWC else np.ascontiguousarray(a)


# ------------------------------------------------------------------ histogram matching
def hm_hist(images, layout=SX_NCHW, counts=None):
    if counts is None:
        counts = torch.zeros((3, 256), dtype=torch.int64)
    if images.shape[0] > 0:
        counts += torch.from_numpy(ox.hm_counts(_nchw(images, layout)))
    return counts


def hm_ref_hist(counts):
    cf = counts.numpy().astype(np.float32)
    out = np.stack([cf[c] / np.float32(ox.sum_f32(cf[c]) + np.float32(1e-8)) for c in range(3)])
    return torch.from_numpy(out.astype(np.float32))


def hm_ref_cdf(ref_hist):
    return torch.from_numpy(ox.hm_ref_cdf(ref_hist.numpy()))


def hm_build_lut(counts, npix, ref_cdf):
    if npix < 0:
        npix = int(counts[0].sum())
    return torch.from_numpy(ox.hm_lut(counts.numpy(), int(npix), ref_cdf.numpy()))


def hm_apply(images, lut, layout=SX_NCHW):
    a = _nchw(images, layout)
    l = lut.numpy()
    out = np.empty_like(a)
    for c in range(3):
        if a.dtype == np.uint8:
            out[:, c] = l[c][a[:, c]].astype(np.uint8)
        else:
            q = np.clip(a[:, c] * np.float32(255.0), 0, 255).astype(np.uint8)
            out[:, c] = np.clip(l[c][q] / np.float32(255.0), 0, 1)
    if layout == SX_NHWC:
        out = np.ascontiguousarray(out.transpose(0, 2, 3, 1))
    return torch.from_numpy(out)


def hm_transform(images, ref_hist, layout=SX_NCHW):
    counts = hm_hist(images, layout)
    return hm_apply(images, hm_build_lut(counts, images.numel() // 3, hm_ref_cdf(ref_hist)), layout)


def hm_fit(images, layout=SX_NCHW):
    return hm_ref_hist(hm_hist(images, layout))


# ------------------------------------------------------------------ reinhard
def reinhard_stats(images, sums=None):
    if sums is None:
        sums = torch.zeros(8, dtype=torch.float64)
    if images.shape[0] > 0:
        lab = ox.rgb_to_lab(np.ascontiguousarray(images.numpy())).astype(np.float64) - 128.0
        sums[0:3] += torch.from_numpy(lab.sum(axis=(0, 2, 3)))
        sums[3:6] += torch.from_numpy((lab * lab).sum(axis=(0, 2, 3)))
        sums[6] += lab.shape[0] * lab.shape[2] * lab.shape[3]
    return sums


def reinhard_finalize(sums):
    s = sums.numpy()
    n = s[6]
    m = s[0:3] / n
    var = (s[3:6] - s[0:3] * m) / (n - 1.0)
    return torch.from_numpy((m + 128.0).astype(np.float32)), torch.from_numpy(np.sqrt(np.maximum(var, 0)).astype(np.float32))


def reinhard_apply(images, src_mean, src_std, ref_mean, ref_std):
    a = np.ascontiguousarray(images.numpy())
    lab = ox.rgb_to_lab(a)
    sm, ss, rm, rs = (t.numpy().astype(np.float32).reshape(1, 3, 1, 1) for t in (src_mean, src_std, ref_mean, ref_std))
    lab = ((lab - sm) / (ss + np.float32(1e-8))) * rs + rm
    rgb = ox.lab_to_rgb(lab.astype(np.float32))
    if a.dtype == np.uint8:
        rgb = np.clip(rgb * np.float32(255.0), 0, 255).astype(np.uint8)
    return torch.from_numpy(rgb)


def reinhard_transform(images, ref_mean, ref_std):
    m, s = reinhard_finalize(reinhard_stats(images))
    return reinhard_apply(images, m, s, ref_mean, ref_std)


def reinhard_fit(images):
    return reinhard_finalize(reinhard_stats(images))


# ------------------------------------------------------------------ macenko
def _od(images: torch.Tensor) -> np.ndarray:
    a = images.numpy()
    x = a.astype(np.float32) / np.float32(255.0) if a.dtype == np.uint8 else a.astype(np.float32)
    od = -np.log((x * np.float32(255.0) + np.float32(1.0)) / np.float32(240.0))
    return od.transpose(1, 0, 2, 3).reshape(3, -1).astype(np.float32)  # (3, N*H*W)


def _diamond(y, x):
    a = np.abs(x) + np.abs(y)
    r = np.where(a > 0, y / np.where(a > 0, a, 1), 0).astype(np.float32)
    return np.where(x >= 0, r, np.where(y >= 0, 2 - r, -2 - r)).astype(np.float32)


def _diamond_to_unit(p):
    p = float(p)
    if p > 1:
        x, y = -(p - 1), 2 - p
    elif p < -1:
        x, y = 1 + p, -2 - p
    else:
        x, y = 1 - abs(p), p
    h = np.hypot(x, y)
    return x / h, y / h


def _rank_index(q, n):
    return int(np.rint(0.01 * q * (n - 1)))


class MacenkoWorkspace:
    """numpy restatement of the pooled-fit phases (slot 0 only): sample histogram -> bracket ->
    count below + cells inside the bracket -> rank search.  The "sample" here is every pixel."""

    def __init__(self, slots: int, device=None):
        assert slots == 1
        self.slots = 1
        self._r = {}
        self.state = {}

    def region(self, name):
        return self._r[name]

    def begin(self):
        self._r = {
            "moments": torch.zeros((1, 12), dtype=torch.float64),
            "odrange": torch.full((1, 8), -np.inf, dtype=torch.float32),
            "counters": torch.zeros((1, 8), dtype=torch.int64),
            "hist1": torch.zeros((1, 2, BINS), dtype=torch.int32),
            "hist2": torch.zeros((1, 2, BINS), dtype=torch.int32),
            "vmin": torch.full((1, 2, BINS), np.inf, dtype=torch.float32),
            "vmax": torch.full((1, 2, BINS), -np.inf, dtype=torch.float32),
            "fit": torch.zeros((1, 8), dtype=torch.float32),
            "status": torch.zeros((1, 4), dtype=torch.int32),
        }
        self.state = {}

    def moments(self, images, pooled, slot0=0):
        od = _od(images)
        keep = od.min(axis=0) >= np.float32(0.15)
        x = (od[:, keep] - SHIFT).astype(np.float64)
        m = self._r["moments"][0]
        m[0] += x.shape[1]
        m[1:4] += torch.from_numpy(x.sum(axis=1))
        xx = x @ x.T
        m[4:10] += torch.tensor([xx[0, 0], xx[0, 1], xx[0, 2], xx[1, 1], xx[1, 2], xx[2, 2]])
        m[10] += od.shape[1]
        rg = self._r["odrange"][0]
        rg[0:3] = torch.maximum(rg[0:3], torch.from_numpy(-od.min(axis=1)))
        rg[3:6] = torch.maximum(rg[3:6], torch.from_numpy(od.max(axis=1)))

    def basis(self, slot0, count, allow_fallback):
        m = self._r["moments"][0].numpy()
        n = m[0]
        s = m[1:4]
        q = np.array([[m[4], m[5], m[6]], [m[5], m[7], m[8]], [m[6], m[8], m[9]]])
        cov = (q - np.outer(s, s) / n) / (n - 1.0)
        _, v = np.linalg.eigh(cov)
        for j in range(3):  # canonical sign: largest-magnitude component positive
            if v[np.abs(v[:, j]).argmax(), j] < 0:
                v[:, j] = -v[:, j]
        self.state.update(e=v[:, [1, 2]].astype(np.float32), n_sel=int(round(n)), n_all=int(round(m[10])))

    def _keys(self, images, stage):
        """-> per query (float key in [0, 2^24), value)."""
        od = _od(images)
        st = self.state
        if stage == 0:
            keep = od.min(axis=0) >= np.float32(0.15)
            t = st["e"].T @ od[:, keep]
            p = _diamond(t[1], t[0])
            key = np.clip((p + np.float32(2.0)) * np.float32(4194304.0), 0, 16777215).astype(np.float32)
            return [(key, p), (key, p)]
        c = st["pinv"] @ od
        out = []
        for j in range(2):
            u = (c[j].astype(np.float32) - st["c_lo"][j]) * st["c_scale"][j]
            out.append((np.clip(u, 0, 16777215).astype(np.float32), c[j].astype(np.float32)))
        return out

    def hist(self, images, pooled, stage, level, slot0=0):
        keys = self._keys(images, stage)
        cnt = self._r["counters"][0]
        if level in (0, 2):  # the stand-in histograms every pixel at level 0 already, so level 2 (exact coarse pass) is the same
            h = self._r["hist1"][0]
            for q in ((0,) if stage == 0 else (0, 1)):
                h[q] += torch.from_numpy(np.bincount(keys[q][0].astype(np.int64) >> 12, minlength=BINS).astype(np.int32))
                cnt[2 + q] += len(keys[q][0])
            return
        st = self.state
        for q in range(2):
            key, val = keys[q]
            lo, hi = st["lo_f"][q], st["hi_f"][q]
            inner = (key >= lo) & (key < hi)
            cell = np.full(key.shape, -1, dtype=np.int64)
            cell[inner] = 1 + np.minimum(((key[inner] - lo) * st["inv_w"][q]).astype(np.int64), BINS - 3)
            if st["open_lo"][q]:
                cell[key < lo] = 0
            else:
                cnt[q] += int((key < lo).sum())
            if st["open_hi"][q]:
                cell[key >= hi] = BINS - 1
            sel = cell >= 0
            self._r["hist2"][0, q] += torch.from_numpy(np.bincount(cell[sel], minlength=BINS).astype(np.int32))
            vmin, vmax = self._r["vmin"][0, q].numpy(), self._r["vmax"][0, q].numpy()
            np.minimum.at(vmin, cell[sel], val[sel])
            np.maximum.at(vmax, cell[sel], val[sel])

    @staticmethod
    def _bin_of_rank(h, k):
        cum = np.cumsum(h.astype(np.int64))
        return min(int(np.searchsorted(cum, k, side="right")), len(h) - 1)

    def select(self, slot0, count, stage, level):
        st = self.state
        cnt = self._r["counters"][0]
        if level == 0:  # wanted ranks and brackets
            st["rank"], st["lo_f"], st["hi_f"], st["inv_w"] = [0, 0], [0.0, 0.0], [0.0, 0.0], [0.0, 0.0]
            st["open_lo"], st["open_hi"] = [1, 1], [1, 1]
            for q in range(2):
                hq = 0 if stage == 0 else q
                n = st["n_sel"] if stage == 0 else st["n_all"]
                pct = (1.0 if q == 0 else 99.0) if stage == 0 else 99.0
                k = min(max(_rank_index(pct, n), 0), n - 1)
                st["rank"][q] = k
                m = int(cnt[2 + hq])
                h = self._r["hist1"][0, hq].numpy()
                b_lo, b_hi = 0, BINS - 1
                if m > 0 and n > 0:
                    if m >= n:
                        r_lo = r_hi = k
                        st["open_lo"][q] = st["open_hi"][q] = 0
                    else:
                        ks = k * m / n
                        sd = np.sqrt(m * (0.01 * pct) * (1 - 0.01 * pct))
                        r_lo, r_hi = int(np.floor(ks - 8 * sd)) - 2, int(np.ceil(ks + 8 * sd)) + 2
                        st["open_lo"][q], st["open_hi"][q] = int(r_lo <= 0), int(r_hi >= m - 1)
                    b_lo = self._bin_of_rank(h, min(max(r_lo, 0), m - 1))
                    b_hi = self._bin_of_rank(h, min(max(r_hi, 0), m - 1))
                st["lo_f"][q] = np.float32(b_lo * 4096)
                st["hi_f"][q] = np.float32((b_hi + 1) * 4096)
                st["inv_w"][q] = np.float32((BINS - 2) / ((b_hi - b_lo + 1) * 4096.0))
            return
        val = [0.0, 0.0]
        for q in range(2):
            h = self._r["hist2"][0, q].numpy().astype(np.int64)
            k = st["rank"][q] - int(cnt[q])
            assert 0 <= k < h.sum(), "rank outside the bracket"
            cell = self._bin_of_rank(h, k)
            before = int(h[:cell].sum())
            c = int(h[cell])
            lo, hi = float(self._r["vmin"][0, q, cell]), float(self._r["vmax"][0, q, cell])
            val[q] = lo + (hi - lo) * ((k - before) / (c - 1)) if c > 1 and hi > lo else lo
        fit = self._r["fit"][0]
        if stage == 0:
            e = st["e"].astype(np.float64)
            v = [e @ np.array(_diamond_to_unit(val[q])) for q in range(2)]
            vmin, vmax = v[0].astype(np.float32), v[1].astype(np.float32)
            he = np.stack([vmin, vmax], axis=1) if vmin[0] > vmax[0] else np.stack([vmax, vmin], axis=1)
            fit[0:6] = torch.from_numpy(he.reshape(-1))
            he64 = he.astype(np.float64)
            st["pinv"] = (np.linalg.inv(he64.T @ he64) @ he64.T).astype(np.float32)
            rg = self._r["odrange"][0].numpy().astype(np.float64)
            st["c_lo"], st["c_scale"] = [0.0, 0.0], [0.0, 0.0]
            for j in range(2):
                a, b = st["pinv"][j] * -rg[0:3], st["pinv"][j] * rg[3:6]
                lo, hi = np.minimum(a, b).sum(), np.maximum(a, b).sum()
                pad = 1e-6 * (abs(lo) + abs(hi)) + 1e-12
                st["c_lo"][j] = np.float32(lo - pad)
                st["c_scale"][j] = np.float32(16777216.0 / ((hi + pad) - (lo - pad)))
            for name, fill in (("hist1", 0), ("hist2", 0), ("vmin", np.inf), ("vmax", -np.inf), ("counters", 0)):
                self._r[name].fill_(fill)
        else:
            fit[6], fit[7] = val[0], val[1]


def macenko_fit(images):
    ws = MacenkoWorkspace(1)
    ws.begin()
    ws.moments(images, True)
    ws.basis(0, 1, False)
    for stage in (0, 1):
        for level in (0, 1):
            ws.hist(images, True, stage, level)
            ws.select(0, 1, stage, level)
    fit = ws.region("fit")[0]
    return fit[:6].reshape(3, 2).clone(), fit[6:8].clone()


def macenko_transform(images, he_ref, maxc_ref, unit=False):
    out = ox.macenko_transform(np.ascontiguousarray(images.numpy()), he_ref.numpy(), maxc_ref.numpy())
    if unit:
        out = out.astype(np.float32) / np.float32(255.0)
    return torch.from_numpy(out)


# ------------------------------------------------------------------ test-only backend subclasses
def cpu_backend(cls, *args, kernel_layer=None, **kwargs):
    """Instance of a product backend class (``HistogramMatchingCUDA`` / ``ReinhardCUDA`` /
    ``MacenkoCUDA``) whose two seams are overridden: the kernel layer is this module (or
    ``kernel_layer``) and CPU devices are accepted.  Lives in the tests on purpose -- the product
    classes take no substitute ops and refuse non-CUDA devices."""
    import sys

    layer = kernel_layer if kernel_layer is not None else sys.modules[__name__]

    class _CPU(cls):  # noqa: N801
        @staticmethod
        def _kernel_layer():
            return layer

        def _check_device(self) -> None:
            pass

    _CPU.__name__ = cls.__name__ + "OnCPU"
    return _CPU(*args, **kwargs)
