#!/usr/bin/env python3
"""Per-stage device time of the extractor at small batch sizes (CUDA events around every launch, profile mode):
what one frame, one stereo pair and one 64-frame chunk cost stage by stage.  Development aid."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import orb_slam2_chinesenotes_b200 as ob  # noqa: E402
from synth import synth_frame  # noqa: E402

w, h, nf = 1241, 376, 2000
for batch in (1, 2, 16, 64):
    imgs = np.stack([synth_frame(w, h, 100 + i) for i in range(min(batch, 4))])
    imgs = np.ascontiguousarray(np.tile(imgs, ((batch + 3) // 4, 1, 1))[:batch])
    ex = ob.ORBextractor(nf, 1.2, 8, 20, 7)
    ex.profile(True)
    for _ in range(3):
        ex.extract_batch(imgs)
    ex.stage_ms()
    reps = 10
    for _ in range(reps):
        ex.extract_batch(imgs)
    ms, cnt = ex.stage_ms()
    print(batch, {k: round(v / reps * 1e3, 1) for k, v in ms.items() if cnt[k]}, "us per call")
    ex.close()
