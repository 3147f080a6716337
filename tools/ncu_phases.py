#!/usr/bin/env python3
"""Per-phase (between BAR.SYNCs) executed-instruction breakdown of one kernel from an .ncu-rep
captured with --import-source on.  Usage: ncu_phases.py rep kernel_regex pixels_per_launch"""
import csv
import subprocess
import sys
from collections import Counter

rep, kern, px = sys.argv[1], sys.argv[2], float(sys.argv[3])
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = [r for r in csv.reader(raw.splitlines()) if len(r) > 6 and r[0].startswith("0x")]
# several launches may match: keep the first (addresses restart)
first = []
seen = set()
for r in rows:
    if r[0] in seen:
        break
    seen.add(r[0])
    first.append(r)
rows = first
tot = sum(int(r[5]) for r in rows)
print(f"{kern}: {tot} warp-inst, {tot * 32 / px:.1f} thread-inst per unit")
bar = [i for i, r in enumerate(rows) if r[1].strip().startswith("BAR")]
segs = [0] + bar + [len(rows)]
for a, b in zip(segs[:-1], segs[1:]):
    c = Counter()
    n = 0
    for r in rows[a:b]:
        parts = r[1].split()
        op = parts[1] if parts[0].startswith("@") else parts[0]
        c[op.split(".")[0]] += int(r[5])
        n += int(r[5])
    if n:
        print(f"  sass {a}-{b}: {n / tot * 100:5.1f}% = {n * 32 / px:6.1f}/unit ",
              [(k, round(v * 32 / px, 1)) for k, v in c.most_common(10)])
