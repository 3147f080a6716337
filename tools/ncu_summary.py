#!/usr/bin/env python3
"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into one line per kernel launch:
duration, DRAM bytes and throughput, SM / issue / pipe utilisation, occupancy, registers.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [> profiles/xyz.txt]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}


def g(r, name, default=""):
    i = col.get(name)
    return r[i] if i is not None and i < len(r) else default


def f(r, name):
    try:
        return float(g(r, name).replace(",", ""))
    except ValueError:
        return float("nan")


def scale(name, v):
    u = units[col[name]] if name in col else ""
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)
    return v * mult


print(f"{'kernel':<16}{'grid':>9}{'us':>9}{'dramRd MB':>10}{'dramWr MB':>10}{'GB/s':>8}{'dram%':>7}{'sm%':>6}{'issue%':>7}"
      f"{'alu%':>6}{'fma%':>6}{'lsu%':>6}{'l1hit%':>7}{'l2hit%':>7}{'occ%':>6}{'regs':>5}")
for r in data:
    name = g(r, "Kernel Name").split("(")[0]
    dur = scale("gpu__time_duration.sum", f(r, "gpu__time_duration.sum"))
    rd = scale("dram__bytes_read.sum", f(r, "dram__bytes_read.sum"))
    wr = scale("dram__bytes_write.sum", f(r, "dram__bytes_write.sum"))
    print(f"{name:<16}{g(r, 'launch__grid_size'):>9}{dur:9.1f}{rd / 1e6:10.2f}{wr / 1e6:10.2f}{(rd + wr) / dur / 1e3:8.0f}"
          f"{f(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):7.1f}"
          f"{f(r, 'sm__throughput.avg.pct_of_peak_sustained_elapsed'):6.1f}"
          f"{f(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):7.1f}"
          f"{f(r, 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active'):6.1f}"
          f"{f(r, 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active'):6.1f}"
          f"{f(r, 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active'):6.1f}"
          f"{f(r, 'l1tex__t_sector_hit_rate.pct'):7.1f}{f(r, 'lts__t_sector_hit_rate.pct'):7.1f}"
          f"{f(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):6.1f}{g(r, 'launch__registers_per_thread'):>5}")
