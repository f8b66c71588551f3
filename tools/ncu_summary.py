#!/usr/bin/env python
"""Summarise an ncu report (``--set full``) or a launch list (``--metrics gpu__time_duration.sum``)
into the small text files kept under ``profiles/``.

    python tools/ncu_summary.py full   gpurun_out/prof_hm.ncu-rep  profiles/r01_hm_full.md
    python tools/ncu_summary.py launch gpurun_out/launches.csv     profiles/r01_hm_launches.md
"""
from __future__ import annotations

import csv
import io
import json
import re
import subprocess
import sys
from collections import OrderedDict
from pathlib import Path

METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__t_bytes.sum", "l2_bytes"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu_pipe_pct"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu_pipe_pct"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma_pipe_pct"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu_pipe_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__inst_executed.sum", "warp_insts"),
]


def short(name: str) -> str:
    name = re.sub(r"\(.*", "", name)
    name = name.replace("void ", "")
    return name.strip()


def full(rep: Path, out: Path) -> None:
    txt = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    head, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(head)}
    lines = [f"# ncu --set full summary of `{rep.name}` (per launch; cold-cache, serialised replays)", ""]
    names = [m for m, _ in METRICS if m in col]
    lines.append("| kernel | " + " | ".join(f"{dict(METRICS)[m]} [{units[col[m]]}]" for m in names) + " |")
    lines.append("|---|" + "---|" * len(names))
    traffic: dict[str, float] = OrderedDict()
    for r in rows[2:]:
        k = short(r[col["Kernel Name"]])
        if k.startswith("at::") or "distribution" in k or "elementwise" in k:
            continue
        lines.append("| " + k + " | " + " | ".join(r[col[m]] for m in names) + " |")
        try:
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            rd = float(r[col["dram__bytes_read.sum"]]) * scale[units[col["dram__bytes_read.sum"]]]
            wr = float(r[col["dram__bytes_write.sum"]]) * scale[units[col["dram__bytes_write.sum"]]]
            traffic[k] = rd + wr
        except Exception:
            pass
    out.write_text("\n".join(lines) + "\n")
    tj = out.with_suffix(".traffic.json")
    tj.write_text(json.dumps(traffic, indent=1) + "\n")
    print(out, tj)


def launch(csv_path: Path, out: Path) -> None:
    rows = [r for r in csv.reader(open(csv_path)) if len(r) > 10 and r[0].isdigit()]
    agg: dict[str, list[float]] = OrderedDict()
    for r in rows:
        if r[12] != "gpu__time_duration.sum":
            continue
        k = short(r[4])
        ns = float(r[14]) * {"ns": 1.0, "us": 1e3, "ms": 1e6}.get(r[13], 1.0)
        agg.setdefault(k, []).append(ns)
    total = sum(sum(v) for v in agg.values())
    lines = [f"# ncu launch list `{csv_path.name}`: device time per kernel (cold-cache, serialised: compare SHARES)", "", "| kernel | launches | mean us | total us | share |", "|---|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        lines.append(f"| {k[:110]} | {len(v)} | {sum(v) / len(v) / 1e3:.1f} | {sum(v) / 1e3:.1f} | {sum(v) / total * 100:.1f}% |")
    out.write_text("\n".join(lines) + "\n")
    print(out)


if __name__ == "__main__":
    mode, src, dst = sys.argv[1], Path(sys.argv[2]), Path(sys.argv[3])
    dst.parent.mkdir(exist_ok=True)
    (full if mode == "full" else launch)(src, dst)
