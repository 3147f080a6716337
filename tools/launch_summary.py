#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total time, share."""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
H = rows[hdr]
ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows[hdr + 1:]:
    name = r[ki].split("(")[0]
    v = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(r[ui], 1)
    tot[name] += v
    cnt[name] += 1
T = sum(tot.values())
print(" ".join(sys.argv[2:]))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{k:<22} launches {cnt[k]:3d}  total {v:9.1f} us  share {v / T * 100:5.1f}%")
