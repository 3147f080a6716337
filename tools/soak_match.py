"""Ad-hoc parity soak of the matchers: random frames / query sets / parameters through the single-problem entry points
(which run the block-per-problem kernels) against the oracle.  Usage (GPU box): python tools/soak_match.py [cases] [seed]"""
import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import orb_slam2_chinesenotes_b200 as ob
from bow_lib import bow_scene, search_by_bow
from matcher_lib import Matcher, extract_frame, perturbed_frame, projected_queries
from oracle_lib import KP_DTYPE

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
W, H = 1241, 376
kps, desc, scale = extract_frame(W, H, 2000, 2)
O = Matcher("oracle")
bad = 0
for c in range(cases):
    n = int(rng.integers(1, len(kps)))
    sel = rng.permutation(len(kps))[:n]
    k2, d2, _ = perturbed_frame(kps[sel], desc[sel], W, H, int(rng.integers(1 << 30)), shift=int(rng.integers(0, 40)), kmax=int(rng.integers(1, 120)))
    bounds = (float(rng.integers(0, 60)), float(W - rng.integers(0, 60)), float(rng.integers(0, 30)), float(H - rng.integers(0, 30)))
    nq = int(rng.integers(1, 4000))
    q = projected_queries(k2, d2, nq, int(rng.integers(1 << 30)))
    th, ratio = float(rng.choice([1.0, 2.0, 3.0, 5.0, 9.0])), float(rng.choice([0.6, 0.75, 0.9, 1.0]))
    ur = np.where(rng.random(n) < 0.5, k2["x"] - 20 * rng.random(n), -1).astype(np.float32) if rng.random() < 0.5 else None
    init = np.where(rng.random(n) < 0.1, rng.integers(0, nq, n), -1).astype(np.int32) if rng.random() < 0.5 else None
    a = O.search_by_projection_points(k2, d2, ur, scale, bounds, q, th, ratio, init)
    b = ob.ORBmatcher(ratio, True).SearchByProjection(ob.FrameView(k2, d2, bounds, ur), scale, q, th, init)
    ok1 = a[0] == b[0] and (a[1] == b[1]).all()
    # initialisation between the sub-frame and its perturbed copy
    k1s, d1s = kps[sel], desc[sel]
    prev = np.stack([k1s["x"] + rng.normal(0, 5, n), k1s["y"] + rng.normal(0, 5, n)], 1).astype(np.float32)
    win, co = int(rng.choice([10, 50, 100])), bool(rng.integers(0, 2))
    a = O.search_for_initialization(k1s, d1s, k2, d2, scale, bounds, prev, win, ratio, co)
    b = ob.ORBmatcher(ratio, co).SearchForInitialization(ob.FrameView(k1s, d1s, bounds), ob.FrameView(k2, d2, bounds), prev, win)
    ok2 = a[0] == b[0] and (a[1] == b[1]).all() and (a[2] == b[2]).all()
    # bag of words, both overloads
    s = bow_scene(k1s, d1s, int(rng.integers(1 << 30)), kmax=int(rng.integers(1, 80)), n2=int(rng.integers(1, n + 1)))
    ok3 = True
    for kf_kf in (False, True):
        want = search_by_bow("oracle", s, ratio, co, kf_kf)
        v1 = s["valid1"] & (1 - s["bad1"]); v2 = (s["valid2"] & (1 - s["bad2"])) if kf_kf else None
        got = ob.search_by_bow(ob.FrameView(s["k1"], s["d1"], bounds), s["fv1"], v1, ob.FrameView(s["k2"], s["d2"], bounds), s["fv2"], v2, kf_kf, ratio, co)
        ok3 = ok3 and got[0] == want[0] and (got[1] == want[1]).all()
    print(c, (n, nq, th, ratio, win, co), "ok" if ok1 and ok2 and ok3 else f"MISMATCH {ok1} {ok2} {ok3}", flush=True)
    bad += 0 if ok1 and ok2 and ok3 else 1
print("mismatches:", bad)
