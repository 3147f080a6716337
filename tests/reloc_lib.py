"""Scene + wrappers for the relocalisation overload ORBmatcher::SearchByProjection(Frame&, KeyFrame*, set, th, ORBdist)
(src/ORBmatcher.cc:303-431).  Test infrastructure."""
import ctypes as C

import numpy as np

from matcher_lib import flip_bits
from oracle_lib import oracle, ref

vp, ci, cf = C.c_void_p, C.c_int, C.c_float


def reloc_scene(kps, desc, w, h, seed, K):
    """Key frame = (kps, desc) with synthetic depths; current frame = same points after a larger motion."""
    rng = np.random.default_rng(seed)
    n = len(kps)
    fx, fy, cx, cy = K
    z = (4 + 30 * rng.random(n)).astype(np.float32)
    xyz = np.stack([(kps["x"] - cx) / fx * z, (kps["y"] - cy) / fy * z, z], 1).astype(np.float32)
    ang = 0.02
    R = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]], np.float32)
    t = np.array([0.1, 0.03, -0.6], np.float32)
    Tcw = np.eye(4, dtype=np.float32); Tcw[:3, :3] = R; Tcw[:3, 3] = t
    pc = xyz @ R.T + t
    u = fx * pc[:, 0] / pc[:, 2] + cx + rng.normal(0, 2.5, n)
    v = fy * pc[:, 1] / pc[:, 2] + cy + rng.normal(0, 2.5, n)
    perm = rng.permutation(n)
    kc = kps[perm].copy()
    kc["x"] = np.clip(np.rint(u[perm]), 0, w - 1).astype(np.float32)
    kc["y"] = np.clip(np.rint(v[perm]), 0, h - 1).astype(np.float32)
    kc["angle"] = ((kc["angle"] + rng.normal(0, 5, n)) % 360).astype(np.float32)
    dist = np.linalg.norm(xyz, axis=1).astype(np.float32)
    return dict(cur_kps=kc, cur_desc=flip_bits(desc[perm], rng, 70), Tcw=Tcw,
                has_mp=(rng.random(n) < 0.85).astype(np.uint8), bad=(rng.random(n) < 0.03).astype(np.uint8),
                found=(rng.random(n) < 0.2).astype(np.uint8), xyz=np.ascontiguousarray(xyz), mp_desc=flip_bits(desc, rng, 40),
                level=np.clip(kps["octave"] + rng.integers(-1, 2, n), 0, 7).astype(np.int32),
                min_dist=(dist * np.where(rng.random(n) < 0.05, 1.2, 0.5)).astype(np.float32),
                max_dist=(dist * 1.6).astype(np.float32), kf_angle=kps["angle"].astype(np.float32).copy(),
                cur_taken=(rng.random(n) < 0.1).astype(np.uint8))


def run_reloc(impl, s, scale, bounds, K, th, orb_dist, check_ori, want_queries=False):
    L = ref() if impl == "ref" else oracle()
    n, nkf = len(s["cur_kps"]), len(s["has_mp"])
    out = np.zeros(n, np.int32)
    p = lambda a: a.ctypes.data
    common = [n, p(s["cur_kps"]), p(s["cur_desc"]), *bounds, p(scale)]
    tail = [nkf, p(s["has_mp"]), p(s["bad"]), p(s["found"]), p(s["xyz"]), p(s["mp_desc"]), p(s["level"]), p(s["min_dist"]),
            p(s["max_dist"]), p(s["kf_angle"]), p(s["Tcw"]), p(K), p(s["cur_taken"]), p(out), th, orb_dist, int(check_ori)]
    if impl == "ref":
        f = L.orbref_search_by_projection_reloc
        f.argtypes = [ci, vp, vp] + [cf] * 4 + [vp, ci] + [ci] + [vp] * 13 + [cf, ci, ci]
        return f(*common, len(scale), *tail), out
    f = L.orbo_search_by_projection_reloc
    f.argtypes = [ci, vp, vp] + [cf] * 4 + [vp] + [ci] + [vp] * 13 + [cf, ci, ci] + [vp] * 4
    uvr, minl, maxl, valid = np.zeros((nkf, 3), np.float32), np.zeros(nkf, np.int32), np.zeros(nkf, np.int32), np.zeros(nkf, np.uint8)
    nm = f(*common, *tail, p(uvr), p(minl), p(maxl), p(valid))
    return (nm, out, dict(uvr=uvr, minl=minl, maxl=maxl, valid=valid)) if want_queries else (nm, out)
