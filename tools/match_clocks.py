#!/usr/bin/env python3
"""Where one SearchByProjection(Frame, MapPoints) block spends its time: clock64() stamps of block 0 (a library built with
-DORB_MATCH_CLOCKS: tools/build_variant.sh clk orb_match_batch.cu -DORB_MATCH_CLOCKS; run with ORB_B200_LIB=.../liborb_b200_clk.so)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import orb_slam2_chinesenotes_b200 as ob  # noqa: E402
from matcher_lib import extract_frame, projected_queries  # noqa: E402

W, H = 1241, 376
kps, desc, scale = extract_frame(W, H, 2000, 2)
q = projected_queries(kps, desc, 2000, 5)
F = ob.FrameView(kps, desc, (0.0, float(W), 0.0, float(H)))
M = ob.ORBmatcher(0.9, True)
for _ in range(3):
    nm, _ = M.SearchByProjection(F, scale, q, 3.0)
clk = (C.c_longlong * 16)()
assert ob.lib().orbm_debug_clocks(clk) == 0
c = list(clk)
names = {2: "grid sort", 3: "records + descriptors into position order"}
print("matches", nm, "keypoints", len(kps))
prev = c[0]
for i in (2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12):
    if c[i] <= prev:
        continue
    print(f"  {names.get(i, 'round %d' % (i - 4) if i < 12 else 'results')}: {(c[i] - prev) / 1.965e3:.1f} us")
    prev = c[i]
print(f"  total {(c[12] - c[0]) / 1.965e3:.1f} us")
