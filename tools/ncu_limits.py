#!/usr/bin/env python3
"""What bounds each extractor kernel, from an `ncu --set full` capture: profiles/r02_kernel_limits.json, which bench.py
quotes in `roofline.limiter` (so the sentence in the JSON line is data from the committed capture, not a literal).
Usage: ncu -i REP --page raw --csv > raw.csv; python tools/ncu_limits.py raw.csv FRAMES_PER_LAUNCH W H [source name] > profiles/r02_kernel_limits.json"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rows = list(csv.reader(open(sys.argv[1])))
frames = int(sys.argv[2]); w = int(sys.argv[3]); h = int(sys.argv[4])
source = sys.argv[5] if len(sys.argv) > 5 else os.path.basename(sys.argv[1])
hdr, units, data = rows[0], rows[1], rows[2:]
col = {k: i for i, k in enumerate(hdr)}
stage = {"k_pyr_resize": "pyramid", "k_pyr_resize_generic": "pyramid", "k_fast_bands": "fast", "k_blur7": "blur", "k_octree": "octree",
         "k_describe": "describe", "k_stereo_rows": "stereo", "k_stereo_match": "stereo", "k_stereo_cut": "stereo"}


def num(r, name):
    try:
        v = float(r[col[name]].replace(",", ""))
    except (KeyError, ValueError):
        return None
    u = units[col[name]]
    return v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)


import orb_slam2_chinesenotes_b200 as ob  # noqa: E402  (host-only call)
d = ob.plan_describe(1000, 1.2, 8, 20, 7, w, h)
P = int((d["level_w"].astype("int64") * d["level_h"].astype("int64")).sum())
agg = {}
for r in data:
    name = r[col["Kernel Name"]].split("(")[0].split("<")[0].replace("void ", "").strip()
    st = stage.get(name)
    if st is None:
        continue
    a = agg.setdefault(st, {"us": 0.0, "inst": 0.0, "ipc_t": 0.0, "alu_t": 0.0, "fma_t": 0.0, "lsu_t": 0.0, "xu_t": 0.0, "occ_t": 0.0, "dram": 0.0, "launches": 0,
                            "regs": 0, "kernels": set(), "stalls": {}})
    us = num(r, "gpu__time_duration.sum") or 0.0
    a["us"] += us; a["launches"] += 1; a["kernels"].add(name)
    a["inst"] += num(r, "smsp__inst_executed.sum") or 0.0
    for key, metric in (("ipc_t", "sm__inst_executed.avg.per_cycle_active"), ("alu_t", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                        ("fma_t", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"), ("lsu_t", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
                        ("xu_t", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"), ("occ_t", "sm__warps_active.avg.pct_of_peak_sustained_active")):
        a[key] += (num(r, metric) or 0.0) * us
    a["dram"] += (num(r, "dram__bytes_read.sum") or 0.0) + (num(r, "dram__bytes_write.sum") or 0.0)
    a["regs"] = max(a["regs"], int(num(r, "launch__registers_per_thread") or 0))
    for k, i in col.items():
        if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio") and r[i]:
            nm = k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")
            a["stalls"][nm] = a["stalls"].get(nm, 0.0) + float(r[i].replace(",", "")) * us
out = {}
for st, a in agg.items():
    us = a["us"] or 1.0
    ipc = a["ipc_t"] / us
    stalls = sorted(((k, v / us) for k, v in a["stalls"].items() if k != "selected"), key=lambda kv: -kv[1])[:3]
    e = {"source": "profiles/" + source, "kernels": sorted(a["kernels"]), "launches_per_chunk": a["launches"], "frames_per_launch": frames,
         "us_per_chunk_under_ncu": round(a["us"], 1), "warp_inst_per_clk_per_sm": round(ipc, 3), "issue_frac": round(ipc / 4.0, 3),
         "alu_pipe_pct": round(a["alu_t"] / us, 1), "fma_pipe_pct": round(a["fma_t"] / us, 1), "lsu_pipe_pct": round(a["lsu_t"] / us, 1),
         "xu_pipe_pct": round(a["xu_t"] / us, 1), "warps_active_pct": round(a["occ_t"] / us, 1), "registers": a["regs"],
         "dram_bytes_per_frame": round(a["dram"] / frames, 1), "top_stalls": [[k, round(v, 2)] for k, v in stalls]}
    if st in ("fast", "blur", "pyramid"):
        e["thread_inst_per_pixel_pair"] = round(a["inst"] * 32.0 / (P * frames / 2.0), 1)
    pipes = {"ALU pipe": e["alu_pipe_pct"], "FMA pipe": e["fma_pipe_pct"], "LSU pipe": e["lsu_pipe_pct"], "XU pipe": e["xu_pipe_pct"]}
    busiest = max(pipes, key=pipes.get)
    if e["issue_frac"] >= 0.6 or pipes[busiest] >= 60:
        e["binding"] = f"instruction issue ({ipc:.2f} of 4 warp instructions per clock and SM; busiest: {busiest} {pipes[busiest]:.0f} %)"
    else:
        e["binding"] = f"latency ({stalls[0][0]} is the top stall, {ipc:.2f} warp instructions per clock and SM)" if stalls else "latency"
    out[st] = e
json.dump(out, sys.stdout, indent=1)
print()
