#!/usr/bin/env python3
"""Instructions executed and stall samples per CUDA source line of one kernel of an ncu capture (compiled with -lineinfo).
Usage: ncu -i REP --page source --csv --print-source cuda,sass > src.csv; python tools/ncu_lines.py src.csv [min_pct]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
hdr = None
lines = []   # (file, line, text, inst, samples)
cur_file = ""
for r in rows:
    if len(r) >= 2 and r[0] in ("File Path", "File Name"):
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) > 8 and r[0] == "Line No":
        hdr = r
        ci = hdr.index("Instructions Executed"); si = hdr.index("# Samples")
        continue
    if hdr is None or len(r) <= ci:
        continue
    if r[0] != "":   # a CUDA line with its aggregated counters
        try:
            lines.append((cur_file, int(r[0]), r[1].strip(), int(r[ci] or 0), int(r[si] or 0)))
        except ValueError:
            pass
tot = sum(l[3] for l in lines) or 1
stot = sum(l[4] for l in lines) or 1
print(f"total warp instructions {tot}, samples {stot}")
for f, n, t, c, s in lines:
    if 100.0 * c / tot >= minpct or 100.0 * s / stot >= minpct:
        print(f"{100.0 * c / tot:5.1f}% inst {100.0 * s / stot:5.1f}% smp  {f}:{n}  {t[:110]}")
