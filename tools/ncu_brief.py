#!/usr/bin/env python3
"""Per-kernel digest of an `ncu --set full` capture (first launch of every kernel): duration, occupancy, issue rate,
pipe utilisation, shared-memory bank conflicts, DRAM bytes and the top warp stall reasons.
Usage: ncu -i REP --page raw --csv > raw.csv; python tools/ncu_brief.py raw.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum"]
seen = set()
for r in data:
    name = r[col["Kernel Name"]].split("(")[0]
    if name in seen:
        continue
    seen.add(name)
    print("----", r[col["Kernel Name"]][:60])
    for w in want:
        if w in col:
            print("  ", w, r[col[w]], units[col[w]])
    st = [(h, float(r[i].replace(",", ""))) for h, i in col.items()
          if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") and r[i]]
    for h, v in sorted(st, key=lambda x: -x[1])[:6]:
        print("     stall", h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), round(v, 2))
