"""ORBmatcher::SearchForTriangulation (src/ORBmatcher.cc:1183-1361): synthetic key-frame pairs with feature vectors in CSR
form + one Python signature for the reference (oracle/_ref) and the plain-C oracle.  Test infrastructure."""
import ctypes as C

import numpy as np

from bow_lib import feature_vector
from matcher_lib import flip_bits
from oracle_lib import oracle, ref

vp, ci, cf = C.c_void_p, C.c_int, C.c_float
p = lambda a: None if a is None else a.ctypes.data


def triang_scene(kps, desc, w, h, seed, K, scale, n2=None, kmax=30):
    """Key frame 2 sees key frame 1's features after a mostly sideways motion: shifted along x by up to 40 px and off the
    (nearly horizontal) epipolar line by a few pixels times the level's scale; F12 is the skew matrix of a translation
    along x, slightly perturbed; the epipole is placed INSIDE the image so that the distance-to-epipole test bites."""
    rng = np.random.default_rng(seed)
    n = len(kps)
    perm = rng.permutation(n)[:n2 or n]
    k2 = kps[perm].copy()
    sc = scale[np.clip(k2["octave"], 0, len(scale) - 1)]
    k2["x"] = np.clip(k2["x"] + rng.integers(-40, 41, len(k2)), 0, w - 1).astype(np.float32)
    k2["y"] = np.clip(k2["y"] + rng.normal(0, 1.2, len(k2)) * sc, 0, h - 1).astype(np.float32)
    k2["angle"] = np.where(rng.random(len(k2)) < 0.8, (k2["angle"] + 15 + rng.normal(0, 5, len(k2))) % 360, rng.random(len(k2)) * 360).astype(np.float32)
    d2 = flip_bits(desc[perm], rng, kmax)
    d2[1::9] = d2[0::9][:len(d2[1::9])]                      # equal descriptors inside nodes: the LAST equally close candidate wins
    F12 = np.float32([[0, 0, 0], [0, 0, -1], [0, 1, 0]]) + rng.normal(0, 1e-6, (3, 3)).astype(np.float32)
    Cw = np.float32([0.55 * w - K[2], 0.45 * h - K[3], K[0]]) / np.float32(K[0]) * np.float32(3.0)   # projects to (0.55 w, 0.45 h) in KF2
    pose2 = np.concatenate([np.eye(3, dtype=np.float32).ravel(), np.zeros(3, np.float32)])
    invz = np.float32(1.0) / Cw[2]
    epi = np.float32([np.float32(np.float32(np.float32(K[0]) * Cw[0]) * invz) + np.float32(K[2]),
                      np.float32(np.float32(np.float32(K[1]) * Cw[1]) * invz) + np.float32(K[3])])
    k1, d1 = kps.copy(), desc.copy()
    m = len(k1[1::11])
    k1[1::11], d1[1::11] = k1[0::11][:m], d1[0::11][:m]       # twin features in key frame 1: both may take the same feature of key frame 2
    s = dict(k1=k1, d1=d1, k2=k2, d2=np.ascontiguousarray(d2),
             has1=(rng.random(n) < 0.3).astype(np.uint8), has2=(rng.random(len(k2)) < 0.3).astype(np.uint8),
             ur1=np.where(rng.random(n) < 0.5, kps["x"] - 20 * rng.random(n), -1).astype(np.float32),
             ur2=np.where(rng.random(len(k2)) < 0.5, k2["x"] - 20 * rng.random(len(k2)), -1).astype(np.float32),
             F12=np.ascontiguousarray(F12.astype(np.float32)), Cw=Cw.astype(np.float32), pose2=pose2, epipole=epi)
    s["fv1"] = feature_vector(s["d1"])
    s["fv2"] = feature_vector(s["d2"])
    return s


def search_for_triangulation(impl, s, K, scale, sigma2, only_stereo, check_ori, mono=False):
    """(nmatches, match12 [n1] -> feature of key frame 2 or -1).  mono: both key frames without right coordinates."""
    n1, n2 = len(s["k1"]), len(s["k2"])
    (id1, off1, f1), (id2, off2, f2) = s["fv1"], s["fv2"]
    ur1, ur2 = (None, None) if mono else (s["ur1"], s["ur2"])
    out = np.zeros(n1, np.int32)
    Kf = np.asarray(K, np.float32)
    side = [ci, vp, vp, vp, vp, ci, vp, vp, vp]
    a1 = [n1, p(s["k1"]), p(s["d1"]), p(s["has1"]), p(ur1), len(id1), p(id1), p(off1), p(f1)]
    a2 = [n2, p(s["k2"]), p(s["d2"]), p(s["has2"]), p(ur2), len(id2), p(id2), p(off2), p(f2)]
    if impl == "ref":
        f = ref().orbref_search_for_triangulation
        f.argtypes = side * 2 + [vp, vp, vp, vp, vp, vp, ci, ci, ci, vp]
        nm = f(*a1, *a2, p(s["F12"]), p(s["Cw"]), p(s["pose2"]), p(Kf), p(scale), p(sigma2), len(scale), int(only_stereo), int(check_ori), p(out))
    else:
        f = oracle().orbo_search_for_triangulation
        f.argtypes = side * 2 + [vp, vp, vp, vp, ci, ci, vp]
        nm = f(*a1, *a2, p(s["F12"]), p(s["epipole"]), p(scale), p(sigma2), int(only_stereo), int(check_ori), p(out))
    return nm, out
