#!/usr/bin/env python3
"""Reads bench.py's JSON line on stdin and prints the numbers that matter on one short line."""
import json
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else ""
for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    r = d.get("roofline") or {}
    st = {k: round(v, 3) for k, v in (r.get("stage_ms_per_step") or {}).items()}
    print(tag, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"], 3), st,
          "frac", r.get("frac"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
