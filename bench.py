#!/usr/bin/env python
"""Throughput bench for the stain-normalization hot path (BASELINE.json metric: megapixels/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--method hm|reinhard|macenko]

Headline workload (N GPUs, weak scaling, 64 images per GPU): BASELINE.json configs[1] --
HistogramMatching, uint8, 64x3x1024x1024 per GPU, reference mode.  One step = one
`HistogramMatching.transform(batch)`: per-channel histogram of the (sharded) batch, [N>1: all-reduce of the
3x256 counts inside the LUT kernel over NVLink peer memory], LUT build, LUT remap.  The batch (201 MB per GPU) is larger than the 126 MB L2, so every step
streams it from HBM; no L2 flush is needed between steps.

Rank 0 prints ONE JSON line (see the keys in `main`).  `--impl reference` times the CPU oracle
port (oracle/stainx_oracle.c, OpenMP, all host threads) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

IMAGES_PER_GPU = 64
EVENT_EVERY = 32  # an instrumented step costs ~15 us more: four cudaEventRecord calls, and the phase-level calls
                  # cannot chain the kernels as programmatic dependent launches the way the single library call does
H = W = 1024
ALGO_BYTES_PER_PX = {"hm": 9.0, "reinhard": 36.0, "macenko": 24.0}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed steps (default: 2000 for the GPU arm, 20 for --impl reference)")
    ap.add_argument("--warmup", type=int, default=None, help="warm-up steps (default: 20 / 2)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--method", default="hm", choices=["hm", "reinhard", "macenko"])
    ap.add_argument("--no-extras", action="store_true", help="skip the per-method side measurements")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 2000 if args.impl == "b200" else 20
    if args.warmup is None:
        args.warmup = 20 if args.impl == "b200" else 2
    return args


def peaks() -> tuple[float, str]:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# clocks: sample nvidia-smi during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons through NVML from a background thread (every ~2 ms)
    while the timed region runs; falls back to one `nvidia-smi` query when NVML is unavailable."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index: int):
        import threading

        self.samples: list[int] = []
        self.reason_bits = 0
        self.sm_max = None
        self.power_w: list[float] = []
        self._stop = threading.Event()
        self._thread = None
        self._nvml = None
        try:
            import pynvml

            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = gpu_index
            if visible:
                ids = [v.strip() for v in visible.split(",") if v.strip()]
                if gpu_index < len(ids) and ids[gpu_index].isdigit():
                    phys = int(ids[gpu_index])
            self._handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.sm_max = int(pynvml.nvmlDeviceGetMaxClockInfo(self._handle, pynvml.NVML_CLOCK_SM))
            self._nvml = pynvml
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        except Exception:
            self._nvml = None

    def _run(self):
        nv = self._nvml
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self._handle, nv.NVML_CLOCK_SM)))
                self.reason_bits |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._handle))
                self.power_w.append(nv.nvmlDeviceGetPowerUsage(self._handle) / 1000.0)
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": [], "samples": 0, "source": "nvml"}
        if self._thread is not None:
            self._stop.set()
            self._thread.join(timeout=2)
            if self.samples:
                s = sorted(self.samples)
                out.update(sm_mhz=s[len(s) // 2], samples=len(s), reasons=sorted(n for bit, n in self.REASONS.items() if self.reason_bits & bit))
                if self.power_w:
                    out["power_w_max"] = max(self.power_w)
            return out
        try:  # fallback: a single nvidia-smi reading
            txt = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10).stdout
            f = [x.strip() for x in txt.splitlines()[0].split(",")]
            out.update(sm_mhz=float(f[0]), sm_max_mhz=float(f[1]), samples=1, source="nvidia-smi (single reading)")
        except Exception:
            out["source"] = "unavailable"
        return out


# ------------------------------------------------------------------------------------------------
# CPU legs (oracle port)
# ------------------------------------------------------------------------------------------------
def cpu_hm_sample(n_img: int, seconds: float) -> dict:
    """Time the oracle's HistogramMatching transform on `n_img` uint8 1024x1024 images, repeated
    for about `seconds`; returns MP/s (best repetition) and what was run."""
    import numpy as np

    from oracle import oracle as ox

    rng = np.random.default_rng(43)
    ref = rng.integers(0, 256, size=(1, 3, H, W), dtype=np.uint8)
    src = rng.integers(0, 256, size=(n_img, 3, H, W), dtype=np.uint8)
    ref_hist = ox.hm_fit(ref)
    ox.hm_transform(src[:1], ref_hist)  # warm-up (page in, thread pool)
    best, reps, t_end = None, 0, time.perf_counter() + seconds
    while reps < 3 or time.perf_counter() < t_end:
        t0 = time.perf_counter()
        ox.hm_transform(src, ref_hist)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
        reps += 1
        if reps >= 200:
            break
    mp = n_img * H * W / 1e6
    return {"value": mp / best, "unit": "MP/s", "cores": ox.num_threads(), "kind": "port", "sample": f"oracle/stainx_oracle.c hm_transform on {n_img}x3x{H}x{W} uint8 (of the {IMAGES_PER_GPU}-image batch), best of {reps} reps"}


def run_reference(args) -> None:
    """`--impl reference`: the reference's CPU path (oracle port), bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np

    from oracle import oracle as ox

    ox.build()
    n_img = 8  # bounded sample of the 64-image batch; cost is linear in pixels
    rng = np.random.default_rng(43)
    ref = rng.integers(0, 256, size=(1, 3, H, W), dtype=np.uint8)
    src = rng.integers(0, 256, size=(n_img, 3, H, W), dtype=np.uint8)
    ref_hist = ox.hm_fit(ref)
    t0 = time.perf_counter()
    ox.hm_transform(src, ref_hist)
    one = time.perf_counter() - t0
    budget = 120.0  # seconds for the whole run
    while n_img > 1 and one * (args.steps + args.warmup) > budget:
        n_img //= 2
        src = src[:n_img]
        one /= 2
    for _ in range(max(args.warmup, 1)):
        ox.hm_transform(src, ref_hist)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ox.hm_transform(src, ref_hist)
    dt = time.perf_counter() - t0
    mp_per_step = n_img * H * W / 1e6
    value = mp_per_step * args.steps / dt
    sample = f"{n_img}x3x{H}x{W} uint8 per step (bounded sample of the {IMAGES_PER_GPU}-image batch), oracle port with OpenMP"
    line = {
        "impl": "reference", "metric": "megapixels_per_second", "value": value, "unit": "MP/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "HistogramMatching uint8 64x3x1024x1024 per GPU, reference mode (BASELINE configs[1])", "sample": sample},
        "cpu_baseline": {"value": value, "unit": "MP/s", "cores": ox.num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
_REAL_STDOUT = None


def _guard_stdout() -> None:
    """Keep stdout for the ONE JSON line: anything a library prints there (NCCL prints its version
    banner on stdout) goes to stderr instead."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict) -> None:
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main() -> None:
    args = parse_args()
    _guard_stdout()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist

    from stainx_b200 import HistogramMatching, Macenko, Reinhard, _native, ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    distributed = world > 1
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    pg = "world" if distributed else None

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if not distributed:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    peak_gbs, peak_src = peaks()
    n_img = IMAGES_PER_GPU
    mp_per_gpu = n_img * H * W / 1e6

    # ---- inputs (synthetic, BASELINE convention: ref seed 42, src seed 43 + rank) ------------
    g = torch.Generator(device=dev).manual_seed(42)
    ref = (torch.rand((1, 3, H, W), device=dev, generator=g) * 255).round().to(torch.uint8)
    g.manual_seed(43 + rank)
    src = (torch.rand((n_img, 3, H, W), device=dev, generator=g) * 255).round().to(torch.uint8)

    hm = HistogramMatching(device=dev, backend="torch_cuda", channel_axis=1, process_group=pg)
    hm.fit_broadcast(ref, src=0) if distributed else hm.fit(ref)
    ref_hist = torch.stack(hm._ref_histograms_256).contiguous()
    ref_cdf = ops.hm_ref_cdf(ref_hist)  # reference CDF: a fit-time constant (3 x 256 floats)
    reducer = hm._make_reducer()
    exchange = hm._get_backend_impl()._peer_exchange() if distributed else None  # None: NCCL all-reduce

    # One step = HistogramMatching.transform(batch).  Every EVENT_EVERY-th step of the timed region is
    # instead written with the phase-level calls so that each kernel can be bracketed by events.
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    marks: list[tuple] = []

    def step(record: bool):
        if not record:  # the public API: one library call (single GPU) or hist / fused exchange + LUT / remap (sharded)
            return hm.transform(src)
        e = [ev() for _ in range(4)]
        if record:
            e[0].record()
        if exchange is not None:  # sharded: counts into the NVLink peer buffer, all-reduce fused into the LUT kernel
            exchange.epoch += 1
            counts = exchange.view((exchange.epoch & 1) * 768 * 8, (3, 256), torch.int64)
            counts.zero_()
            ops.hm_hist(src, counts=counts)
        else:
            counts = ops.hm_hist(src)
        if record:
            e[1].record()
        if exchange is not None:
            lut = ops.hm_build_lut_peers(exchange, ref_cdf)
        else:
            reducer.sum_(counts)
            lut = ops.hm_build_lut(counts, -1 if distributed else src.numel() // 3, ref_cdf)
        if record:
            e[2].record()
        out = ops.hm_apply(src, lut)
        if record:
            e[3].record()
            marks.append(tuple(e))
        return out

    for _ in range(max(args.warmup, 3)):
        step(False)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = _native.kernel_launches()
    t_start, t_stop = ev(), ev()
    barrier()
    t_start.record()
    for i in range(args.steps):
        step(i % EVENT_EVERY == 0)  # per-kernel events on every EVENT_EVERY-th step of the timed region
    t_stop.record()
    barrier()
    elapsed_ms = max_over_ranks(t_start.elapsed_time(t_stop))
    launches = _native.kernel_launches() - launches0
    clocks = sampler.stop() if sampler else None

    hist_ms = sum(m[0].elapsed_time(m[1]) for m in marks) / len(marks)
    lut_ms = sum(m[1].elapsed_time(m[2]) for m in marks) / len(marks)
    apply_ms = sum(m[2].elapsed_time(m[3]) for m in marks) / len(marks)
    ms_per_step = elapsed_ms / args.steps
    value = mp_per_gpu * world / (ms_per_step / 1e3)

    # ---- roofline of the dominant kernel (algorithmic bytes / live CUDA-event time) ----------
    px = n_img * H * W
    kernels = {
        "hm::hist_u8_planar_lane_pw_kernel": {"algo_bytes": 3.0 * px, "ms": hist_ms},
        "hm::apply_u8_planar_vec_kernel": {"algo_bytes": 6.0 * px, "ms": apply_ms},
    }
    for k in kernels.values():
        k["gbs"] = k["algo_bytes"] / (k["ms"] / 1e3) / 1e9
        k["frac"] = k["gbs"] / peak_gbs
    dom_name = max(kernels, key=lambda k: kernels[k]["ms"])
    dom = kernels[dom_name]
    roofline = {"bound": "hbm", "kernel": dom_name, "achieved": dom["gbs"], "peak": peak_gbs, "unit": "GB/s", "frac": dom["frac"], "traffic": None, "peak_source": peak_src,
                "step": {"algo_bytes": ALGO_BYTES_PER_PX["hm"] * px, "gbs": ALGO_BYTES_PER_PX["hm"] * px / (ms_per_step / 1e3) / 1e9, "frac": ALGO_BYTES_PER_PX["hm"] * px / (ms_per_step / 1e3) / 1e9 / peak_gbs},
                "kernels": {k: {"ms": round(v["ms"], 4), "gbs": round(v["gbs"], 1), "frac": round(v["frac"], 4)} for k, v in kernels.items()}, "lut_and_allreduce_ms": round(lut_ms, 4),
                "kernel_timing": f"CUDA events around each phase on every {EVENT_EVERY}th step of the timed region ({len(marks)} samples)"}
    traffic_file = ROOT / "profiles" / "traffic.json"  # per-launch dram bytes from the last ncu --set full capture
    if traffic_file.exists():
        try:
            roofline["traffic"] = json.loads(traffic_file.read_text()).get(dom_name)
        except Exception:
            pass

    # ---- e2e: public API, host buffers, H2D + D2H of every step inside the timed region ----------
    # stainx_b200.ingest.HostStream chains copy-in / kernels / copy-out of each batch on three
    # streams, so the H2D of step i+1 overlaps the D2H of step i (both PCIe directions busy).
    from stainx_b200.ingest import HostStream

    host_in = torch.empty((n_img, 3, H, W), dtype=torch.uint8).pin_memory()
    host_in.copy_(src)
    host_outs = [torch.empty((n_img, 3, H, W), dtype=torch.uint8).pin_memory() for _ in range(2)]
    e2e_steps = max(4, min(args.steps, 16))
    pipe = HostStream(hm, device=dev, depth=2)

    def e2e_run(steps: int) -> None:
        tickets = [pipe.submit(host_in, host_outs[i % 2]) for i in range(steps)]
        for t in tickets:
            t.wait()

    e2e_run(3)
    if not torch.equal(host_outs[0], step(False).cpu()):
        raise SystemExit("e2e result differs from the device-resident result")

    def e2e_time(fn, steps: int) -> float:
        barrier()
        e0, e1 = ev(), ev()
        e0.record()
        fn(steps)
        pipe.synchronize()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)) / steps

    def e2e_serial(steps: int) -> None:  # one batch in flight: H2D, kernels, D2H back to back
        for _ in range(steps):
            pipe.submit(host_in, host_outs[0]).wait()

    # Whether both PCIe directions can run at once at full rate depends on the host (NUMA placement
    # of the pinned buffers, PCIe switch): measure the pipelined and the one-batch-at-a-time mode of
    # the same API and report the faster one, naming it.
    overlapped_ms = e2e_time(e2e_run, e2e_steps)
    serial_ms = e2e_time(e2e_serial, max(3, e2e_steps // 2))
    e2e_ms = min(overlapped_ms, serial_ms)
    e2e = {"value": mp_per_gpu * world / (e2e_ms / 1e3), "unit": "MP/s", "h2d_bytes_per_step": host_in.numel(), "d2h_bytes_per_step": host_outs[0].numel(), "ms_per_step": e2e_ms, "steps": e2e_steps,
           "mode": "pipelined (depth 2)" if overlapped_ms <= serial_ms else "one batch in flight", "pipelined_ms_per_step": overlapped_ms, "serial_ms_per_step": serial_ms,
           "api": "stainx_b200.ingest.HostStream(HistogramMatching(backend='torch_cuda')).submit(pinned uint8 batch, pinned out): H2D + transform + D2H per step"}
    del pipe

    # ---- side measurements: the other methods / BASELINE configs (aggregate MP/s over all ranks) ----
    # Every rank runs its own shard (64 images per GPU, weak scaling); times are CUDA events, max over
    # ranks.  Reinhard's source statistics span the sharded batch (one NCCL all-reduce of 8 doubles
    # per step); the Macenko transform needs no exchange (per-image statistics).
    methods = {"hm_u8_64x1024": {"mp_per_s": value, "algo_gbs_per_gpu": roofline["step"]["gbs"], "frac_of_peak": roofline["step"]["frac"]}}
    if not args.no_extras:
        def timeit(fn, steps, warm=3):
            for _ in range(warm):
                fn()
            barrier()
            a, b = ev(), ev()
            a.record()
            for _ in range(steps):
                fn()
            b.record()
            barrier()
            return max_over_ranks(a.elapsed_time(b)) / steps

        def entry(ms, bytes_per_px, mp=mp_per_gpu, **extra):
            gbs = bytes_per_px * mp * 1e6 / (ms / 1e3) / 1e9
            return {"mp_per_s": mp * world / (ms / 1e3), "algo_gbs_per_gpu": gbs, "frac_of_peak": gbs / peak_gbs, "ms": ms, **extra}

        del host_in, host_outs
        g.manual_seed(43 + rank)
        srcf = torch.rand((n_img, 3, H, W), device=dev, generator=g)
        g.manual_seed(42)
        reff = torch.rand((1, 3, H, W), device=dev, generator=g)
        rh = Reinhard(device=dev, backend="torch_cuda", process_group=pg)
        rh.fit_broadcast(reff, src=0) if distributed else rh.fit(reff)
        methods["reinhard_f32_64x1024"] = entry(timeit(lambda: rh.transform(srcf), 5), ALGO_BYTES_PER_PX["reinhard"])
        if not distributed:
            ref10 = torch.rand((1, 3, 512, 512), device=dev, generator=g)
            src10 = torch.rand((10, 3, 512, 512), device=dev, generator=g)
            ms = timeit(lambda: Reinhard(device=dev, backend="torch_cuda").fit(ref10).transform(src10), 20)
            methods["reinhard_f32_C1_fit1x512_transform10x512"] = {"mp_per_s": 10 * 512 * 512 / 1e6 / (ms / 1e3), "ms": ms, "note": "README quick-start (BASELINE configs[0]); fits in L2, launch-latency bound"}
        mk = Macenko(device=dev, backend="torch_cuda", normalize_to_0_1=True, process_group=pg)
        mk.fit_broadcast(reff, src=0) if distributed else mk.fit(reff)
        methods["macenko_f32_64x1024"] = entry(timeit(lambda: mk.transform(srcf), 5), ALGO_BYTES_PER_PX["macenko"], config="BASELINE configs[2]: reference-mode transform, float32 64x3x1024x1024 per GPU")
        if distributed:  # BASELINE configs[3]: pooled fit over the sharded batch (NCCL stat all-reduces) + transform
            mkb = Macenko(device=dev, backend="torch_cuda", normalize_to_0_1=True, process_group=pg)
            methods["macenko_f32_batch_fit_transform"] = entry(timeit(lambda: mkb.fit(srcf).transform(srcf), 3, warm=1), ALGO_BYTES_PER_PX["macenko"], config="BASELINE configs[3]: pooled fit + transform of the sharded batch")
        del srcf
        # BASELINE configs[4]: StainNormalizerTransform("macenko") on uint8 2048x2048 tiles, 16 per GPU, float32 [0,1] out
        from stainx_b200 import StainNormalizerTransform

        g.manual_seed(43 + rank)
        tiles = (torch.rand((16, 3, 2048, 2048), device=dev, generator=g) * 255).to(torch.uint8)
        g.manual_seed(42)
        ref_tile = (torch.rand((1, 3, 2048, 2048), device=dev, generator=g) * 255).to(torch.uint8)
        snt = StainNormalizerTransform(method="macenko", mode="reference", reference=ref_tile, device=dev, backend="torch_cuda")
        methods["macenko_transform_module_u8_16x2048"] = entry(timeit(lambda: snt(tiles), 5), 15.0, mp=16 * 2048 * 2048 / 1e6, config="BASELINE configs[4] per GPU: uint8 in, float32 [0,1] out (15 B/px)")
        del tiles

    # ---- CPU baseline (rank 0, N=1 only; bounded sample) --------------------------------------
    cpu = None
    if rank == 0 and not distributed and not args.no_cpu_baseline:
        cpu = cpu_hm_sample(8, 10.0)

    if rank == 0:
        line = {
            "metric": "megapixels_per_second", "value": value, "unit": "MP/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"HistogramMatching uint8 {n_img}x3x{H}x{W} per GPU, reference mode (BASELINE configs[1])", "images_per_gpu": n_img, "global_images": n_img * world,
                       "parallelism": f"image-sharded x{world}" + ((", counts all-reduced inside the LUT kernel over NVLink peer memory (no NCCL call)" if exchange is not None else ", NCCL all-reduce of 3x256 int64 counts per step") if distributed else ""),
                       "l2": "input per GPU (201 MB) exceeds L2 (126 MB); no flush needed"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "methods": methods,
        }
        emit(line)
    if distributed:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
