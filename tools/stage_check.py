#!/usr/bin/env python3
"""Stage-by-stage comparison of the CUDA extractor with the oracle on a GPU box.

Prints, per level, how many pyramid / blur bytes, FAST candidates, quadtree keypoints and
final keypoints / descriptors differ.  Development aid (the pytest suite asserts the same).
"""
import sys
import os
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import orb_slam2_chinesenotes_b200 as ob  # noqa: E402
from oracle_lib import OracleExtractor  # noqa: E402
from synth import synth_frame  # noqa: E402


def check(w, h, nf, seed, verbose=True):
    img = synth_frame(w, h, seed)
    O = OracleExtractor(nf)
    n_o, k_o, d_o = O.extract(img)
    G = ob.ORBextractor(nf, 1.2, 8, 20, 7)
    k_g, d_g = G(img)
    ok = True
    for l in range(8):
        po, pg = O.pyramid(l), G.pyramid(l)
        pyr_bad = int((po != pg).sum()) if po.shape == pg.shape else -1
        pob, pgb = O.pyramid(l, True), G.pyramid(l, with_border=True)
        brd_bad = int((pob != pgb).sum()) if pob.shape == pgb.shape else -1
        bo = O.blurred(l)
        bg = G.blurred(l)
        blur_bad = int((bo != bg).sum()) if bo is not None else 0
        co = O.candidates(l)
        cg = G.candidates(l)
        so = set(map(tuple, np.stack([co["x"], co["y"], co["score"]], 1).tolist())) if len(co) else set()
        sg = set(map(tuple, cg.tolist()))
        ko = O.level_keypoints(l)
        kg = G.level_keypoints(l)
        lo = [(int(a["x"]) - 16, int(a["y"]) - 16, int(a["response"])) for a in ko]
        lg = [tuple(r) for r in kg.tolist()]
        line = (f"  L{l} {po.shape[1]}x{po.shape[0]}: pyr_bad={pyr_bad} border_bad={brd_bad} blur_bad={blur_bad} "
                f"cand {len(so)}/{len(sg)} missing={len(so - sg)} extra={len(sg - so)} "
                f"octree {len(lo)}/{len(lg)} equal_ordered={lo == lg} equal_set={set(lo) == set(lg)}")
        if verbose:
            print(line)
        ok &= pyr_bad == 0 and brd_bad == 0 and blur_bad == 0 and so == sg and lo == lg
    same_n = n_o == len(k_g)
    kp_eq = same_n and all((k_o[f] == k_g[f]).all() for f in k_o.dtype.names)
    ang_bad = int((k_o["angle"] != k_g["angle"]).sum()) if same_n else -1
    desc_bad = int((d_o != d_g).any(axis=1).sum()) if same_n else -1
    print(f"{w}x{h} nf={nf} seed={seed}: n oracle={n_o} gpu={len(k_g)} kps_equal={kp_eq} angle_bad={ang_bad} desc_bad={desc_bad} "
          f"stages_ok={ok}")
    G.close()
    O.close()
    return ok and kp_eq and desc_bad == 0


if __name__ == "__main__":
    t0 = time.time()
    cases = [(640, 480, 1000, 1), (1241, 376, 2000, 2), (752, 480, 1200, 3), (200, 150, 300, 5), (1920, 1080, 4000, 4)]
    res = [check(*c) for c in cases]
    print("ALL_OK" if all(res) else "MISMATCH", f"{time.time() - t0:.1f}s")
    sys.exit(0 if all(res) else 1)
