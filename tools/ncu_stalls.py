#!/usr/bin/env python
"""Print the headline metrics and warp-stall breakdown of every kernel in an ncu report (development tool)."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
h, units = rows[0], rows[1]
base = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]
for r in rows[2:]:
    print("==", r[h.index("Kernel Name")][:90])
    for m in base:
        if m in h:
            print(f"   {m:70s} {r[h.index(m)]:>14s} {units[h.index(m)]}")
    stalls = []
    for i, name in enumerate(h):
        if name.startswith("smsp__average_warps_issue_stalled_") and name.endswith("_per_issue_active.ratio"):
            try:
                stalls.append((float(r[i]), name[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    stalls.sort(reverse=True)
    print("   stalls (warps per issue):", ", ".join(f"{n}={v:.2f}" for v, n in stalls[:8]))
