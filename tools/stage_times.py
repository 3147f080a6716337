#!/usr/bin/env python3
"""Per-stage device time of the extractor on a device-resident batch (profile mode: CUDA events around every launch).
Usage: python tools/stage_times.py [frames=512] [w=1241] [h=376] [nf=2000]   (ORB_B200_LIB selects the library build)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import orb_slam2_chinesenotes_b200 as ob  # noqa: E402
from synth import synth_frame  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 512
w = int(sys.argv[2]) if len(sys.argv) > 2 else 1241
h = int(sys.argv[3]) if len(sys.argv) > 3 else 376
nf = int(sys.argv[4]) if len(sys.argv) > 4 else 2000
base = np.stack([synth_frame(w, h, 100 + i) for i in range(8)])
imgs = torch.from_numpy(base).cuda().repeat((frames + 7) // 8, 1, 1)[:frames].contiguous()
ex = ob.ORBextractor(nf, 1.2, 8, 20, 7)
cap = ex.default_capacity()
kps = torch.zeros((frames, cap, 7), dtype=torch.float32, device="cuda")
desc = torch.zeros((frames, cap, 32), dtype=torch.uint8, device="cuda")
n = torch.zeros(frames, dtype=torch.int32, device="cuda")
ex.profile(True)
for _ in range(3):
    ex.extract_batch_raw(imgs, w * h, frames, w, h, w, kps, desc, cap, n)
ex.stage_ms()
reps = 5
for _ in range(reps):
    ex.extract_batch_raw(imgs, w * h, frames, w, h, w, kps, desc, cap, n)
ms, cnt = ex.stage_ms()
tot = sum(ms.values())
print(os.path.basename(os.environ.get("ORB_B200_LIB", "default")), frames, "frames:", {k: round(v / reps, 3) for k, v in ms.items() if cnt[k]}, "ms; sum", round(tot / reps, 3),
      "keypoints", int(n.sum().item()))
ex.close()
