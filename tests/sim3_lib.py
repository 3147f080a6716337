"""Scene + wrappers for the loop-closing overload ORBmatcher::SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th)
(src/ORBmatcher.cc:434-549).  Test infrastructure."""
import ctypes as C

import numpy as np

from matcher_lib import flip_bits
from oracle_lib import oracle, ref

vp, ci, cf = C.c_void_p, C.c_int, C.c_float


def sim3_scene(kps, desc, w, h, seed, K):
    """Key frame = (kps, desc); candidate points = its keypoints back-projected at synthetic depths, seen through a Sim3
    (scale 1.7, small rotation); some points behind the camera / out of range / seen from behind / already matched."""
    rng = np.random.default_rng(seed)
    n = len(kps)
    fx, fy, cx, cy = K
    s = 1.7
    ang = 0.03
    R = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]], np.float32)
    t = np.array([0.2, -0.1, 0.4], np.float32)
    Scw = np.eye(4, dtype=np.float32); Scw[:3, :3] = s * R; Scw[:3, 3] = s * t
    z = (4 + 30 * rng.random(n)).astype(np.float32)
    pc = np.stack([(kps["x"] + rng.normal(0, 2.0, n) - cx) / fx * z, (kps["y"] + rng.normal(0, 2.0, n) - cy) / fy * z, z], 1)
    xyz = ((pc - t) @ R).astype(np.float32)                     # p_w = R^T (p_c - t)
    behind = rng.random(n) < 0.04
    xyz[behind] = ((np.stack([pc[behind, 0], pc[behind, 1], -pc[behind, 2]], 1) - t) @ R).astype(np.float32)
    Ow = (-R.T @ t).astype(np.float32)
    PO = xyz - Ow
    dist = np.linalg.norm(PO, axis=1).astype(np.float32)
    normal = (PO / dist[:, None]).astype(np.float32)
    flip = rng.random(n) < 0.1                                  # seen from behind: fails the 60 degree test
    normal[flip] *= -1
    tilt = rng.random(n) < 0.2
    normal[tilt] = (normal[tilt] + rng.normal(0, 0.8, (int(tilt.sum()), 3))).astype(np.float32)
    perm = rng.permutation(n)                                   # vpPoints order is unrelated to the keypoint order
    matched = np.full(n, -1, np.int32)
    pre = rng.random(n) < 0.08
    matched[pre] = np.where(rng.random(int(pre.sum())) < 0.5, -2, rng.integers(0, n, int(pre.sum())))
    return dict(Scw=Scw, xyz=np.ascontiguousarray(xyz[perm]), normal=np.ascontiguousarray(normal[perm]),
                bad=(rng.random(n) < 0.03).astype(np.uint8), mp_desc=flip_bits(desc[perm], rng, 60),
                level=np.clip(kps["octave"][perm] + rng.integers(0, 2, n), 0, 7).astype(np.int32),
                min_dist=(dist[perm] * np.where(rng.random(n) < 0.05, 1.2, 0.5)).astype(np.float32),
                max_dist=(dist[perm] * 1.6).astype(np.float32), matched=matched)


def run_sim3(impl, kps, desc, s, scale, bounds, K, th, want_queries=False):
    L = ref() if impl == "ref" else oracle()
    n, npts = len(kps), len(s["bad"])
    out = np.zeros(n, np.int32)
    p = lambda a: a.ctypes.data
    head = [n, p(kps), p(desc), *bounds, p(scale)]
    tail = [p(K), p(s["Scw"]), npts, p(s["bad"]), p(s["xyz"]), p(s["normal"]), p(s["mp_desc"]), p(s["level"]), p(s["min_dist"]),
            p(s["max_dist"]), p(s["matched"]), p(out), int(th)]
    if impl == "ref":
        f = L.orbref_search_by_projection_sim3
        f.argtypes = [ci, vp, vp] + [cf] * 4 + [vp, ci] + [vp, vp, ci] + [vp] * 9 + [ci]
        return f(*head, len(scale), *tail), out
    f = L.orbo_search_by_projection_sim3
    f.argtypes = [ci, vp, vp] + [cf] * 4 + [vp] + [vp, vp, ci] + [vp] * 9 + [ci] + [vp] * 4
    uvr, minl, maxl, valid = np.zeros((npts, 3), np.float32), np.zeros(npts, np.int32), np.zeros(npts, np.int32), np.zeros(npts, np.uint8)
    nm = f(*head, *tail, p(uvr), p(minl), p(maxl), p(valid))
    return (nm, out, dict(uvr=uvr, minl=minl, maxl=maxl, valid=valid)) if want_queries else (nm, out)
