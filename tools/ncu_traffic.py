#!/usr/bin/env python3
"""profiles/ncu_traffic.json from an `ncu --set full` capture of one chunk of the bench:
per-stage DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per frame.
Usage: python tools/ncu_traffic.py gpurun_out/prof.ncu-rep FRAMES_PER_LAUNCH > profiles/ncu_traffic.json"""
import csv
import json
import subprocess
import sys
from collections import defaultdict

rep, frames = sys.argv[1], float(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
stage = {"k_pyr_resize": "pyramid", "k_pyr_resize_generic": "pyramid", "k_fast_strips": "fast", "k_fast_bands": "fast", "k_blur7": "blur",
         "k_octree": "octree", "k_describe": "describe", "k_stereo_rows": "stereo", "k_stereo_match": "stereo",
         "k_stereo_cut": "stereo"}
tot = defaultdict(float)
for r in data:
    name = r[col["Kernel Name"]].split("(")[0].split("<")[0].replace("void ", "").strip()
    b = 0.0
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        b += float(r[col[m]].replace(",", "")) * mult.get(units[col[m]], 1)
    tot[stage.get(name, name)] += b
print(json.dumps({"source": rep.split("/")[-1], "frames_per_launch": frames,
                  "note": "one launch of every kernel of one chunk (the pyramid stage is its 7 launches together); cold-cache, serialised by ncu",
                  "dram_bytes_per_frame": {k: v / frames for k, v in tot.items()}}, indent=1))
