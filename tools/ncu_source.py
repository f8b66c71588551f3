#!/usr/bin/env python
"""Top stall-sample instructions of a kernel from an ncu report's source page (development tool).

    python tools/ncu_source.py REPORT.ncu-rep KERNEL_REGEX [TOP]
"""
import csv
import io
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        name = rows[i][1]
        h = rows[i + 1]
        ia, isamp, iex = h.index("Source"), h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
        j = i + 2
        data = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            r = rows[j]
            if len(r) > iex and r[isamp].isdigit():
                data.append((int(r[isamp]), int(r[iex] or 0), j - i - 2, r[ia].strip()))
            j += 1
        tot = sum(d[0] for d in data) or 1
        print("==", name[:100], "samples", tot, "warp instr", sum(d[1] for d in data))
        for d in sorted(data, reverse=True)[:top]:
            print(f"{d[0]:6d} {d[0] / tot * 100:5.1f}%  ex={d[1]:9d}  #{d[2]:4d} {d[3]}")
        i = j
        break  # first instance only
    else:
        i += 1
