#!/usr/bin/env python3
"""Per-kernel SASS opcode histogram of the built library (cuobjdump -sass): which instructions each kernel is made of,
and whether the Blackwell-native ones are there (UTMALDG / UBLKCP = TMA, SYNCS = mbarrier, LDGSTS = cp.async).
Usage: python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "orb_slam2_chinesenotes_b200", "lib", "liborb_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
ver = subprocess.run(["nvcc", "--version"], capture_output=True, text=True).stdout.strip().splitlines()[-2]
print(f"cuobjdump -sass orb_slam2_chinesenotes_b200/lib/liborb_b200.so  ({ver})")
hist = collections.OrderedDict()
arch = {}
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        hist[cur] = collections.Counter()
        continue
    m = re.search(r"\.target\s+(\S+)|arch = (\S+)", line)
    if m and cur is None:
        pass
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(.*?);", line)
    if m and cur:
        toks = m.group(1).split()
        op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
        hist[cur][op] += 1
demangle = subprocess.run(["c++filt"] + list(hist), capture_output=True, text=True).stdout.splitlines()
for (name, h), pretty in sorted(zip(hist.items(), demangle), key=lambda t: t[1]):
    print(f"\n== {pretty.split('(')[0]}  ({sum(h.values())} instructions)")
    print("   " + "  ".join(f"{op} {n}" for op, n in h.most_common()))
    tma = [f"{op} x{n}" for op, n in h.items() if re.match(r"(UTMALDG|UTMASTG|UBLKCP|SYNCS|LDGSTS|UTMAPF)", op)]
    if tma:
        print("   async data path: " + "  ".join(sorted(tma)))
