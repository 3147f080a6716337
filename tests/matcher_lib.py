"""Matcher side of the checkers: ctypes wrappers with ONE Python signature for the reference
(oracle/_ref, prefix orbref_) and the plain-C oracle (prefix orbo_), plus the synthetic matcher
scenes of SURVEY.md App. E.2.  Test infrastructure only."""
import ctypes as C

import numpy as np

from oracle_lib import KP_DTYPE, OracleExtractor, oracle, ref
from synth import synth_frame

vp, ci, cf = C.c_void_p, C.c_int, C.c_float


def _p(a):
    return None if a is None else a.ctypes.data


class Matcher:
    """impl = 'ref' (reference's own ORBmatcher.cc / Frame.cc) or 'oracle' (C restatement)."""

    def __init__(self, impl):
        self.impl = impl
        self.L = ref() if impl == "ref" else oracle()
        assert self.L is not None
        self.pre = "orbref_" if impl == "ref" else "orbo_"

    def _fn(self, name, argtypes, restype=ci):
        f = getattr(self.L, self.pre + name)
        f.argtypes, f.restype = argtypes, restype
        return f

    def _scale_args(self, scale):
        # the reference harness also wants nlevels (it fills the Frame's tables)
        return (_p(scale), len(scale)) if self.impl == "ref" else (_p(scale),)

    def _scale_types(self):
        return [vp, ci] if self.impl == "ref" else [vp]

    def features_in_area(self, kps, scale, bounds, x, y, r, min_level=-1, max_level=-1):
        out = np.zeros(len(kps) + 1, np.int32)
        if self.impl == "ref":
            f = self._fn("features_in_area", [ci, vp, vp, ci] + [cf] * 4 + [cf, cf, cf, ci, ci, vp, ci])
            n = f(len(kps), _p(kps), _p(scale), len(scale), *bounds, x, y, r, min_level, max_level, _p(out), len(out))
        else:
            f = self._fn("features_in_area", [ci, vp] + [cf] * 4 + [cf, cf, cf, ci, ci, vp, ci])
            n = f(len(kps), _p(kps), *bounds, x, y, r, min_level, max_level, _p(out), len(out))
        return out[:n].copy()

    def search_for_initialization(self, kps1, desc1, kps2, desc2, scale, bounds, prev_matched, window, nnratio, check_ori):
        prev = np.ascontiguousarray(prev_matched, np.float32).copy()
        m12 = np.zeros(len(kps1), np.int32)
        st = [vp, ci] if self.impl == "ref" else []
        sa = (_p(scale), len(scale)) if self.impl == "ref" else ()
        f = self._fn("search_for_initialization", [ci, vp, vp, ci, vp, vp] + st + [cf] * 4 + [vp, vp, ci, cf, ci])
        n = f(len(kps1), _p(kps1), _p(desc1), len(kps2), _p(kps2), _p(desc2), *sa, *bounds, _p(prev), _p(m12), window, nnratio, int(check_ori))
        return n, m12, prev

    def search_by_projection_points(self, kps, desc, u_right, scale, bounds, q, th, nnratio, init_assign=None):
        out = np.zeros(len(kps), np.int32)
        f = self._fn("search_by_projection_points", [ci, vp, vp, vp] + self._scale_types() + [cf] * 4 +
                     [ci, vp, vp, vp, vp, vp, vp, vp, vp, vp, cf, cf])
        n = f(len(kps), _p(kps), _p(desc), _p(u_right), *self._scale_args(scale), *bounds, len(q["level"]), _p(q["proj"]), _p(q["level"]),
              _p(q["view_cos"]), _p(q["in_view"]), _p(q["bad"]), _p(q["obs"]), _p(q["desc"]), _p(init_assign), _p(out), th, nnratio)
        return n, out

    def search_by_projection_frame(self, cur, last, Tcw_cur, Tcw_last, K, bf, scale, bounds, th, mono, nnratio, check_ori, cur_init_obs=None):
        out = np.zeros(len(cur["kps"]), np.int32)
        f = self._fn("search_by_projection_frame", [ci, vp, vp, vp, ci, vp, vp, vp, vp, vp, vp, vp, vp, vp, cf] + self._scale_types() +
                     [cf] * 4 + [vp, vp, cf, ci, cf, ci])
        n = f(len(cur["kps"]), _p(cur["kps"]), _p(cur["desc"]), _p(cur.get("u_right")), len(last["kps"]), _p(last["kps"]), _p(last["has_mp"]),
              _p(last["outlier"]), _p(last["xyz"]), _p(last["mp_desc"]), _p(last["mp_obs"]), _p(Tcw_cur), _p(Tcw_last), _p(K), bf,
              *self._scale_args(scale), *bounds, _p(cur_init_obs), _p(out), th, int(mono), nnratio, int(check_ori))
        return n, out


# ---------------------------------------------------------------------------------------------- scenes
def extract_frame(w, h, nf, seed):
    O = OracleExtractor(nf)
    n, kps, desc = O.extract(synth_frame(w, h, seed))
    scale = O.tables()["scale"].copy()
    O.close()
    return kps, desc, scale


def flip_bits(desc, rng, kmax):
    out = desc.copy()
    for i in range(len(out)):
        k = int(rng.integers(0, kmax))
        if k:
            bits = rng.choice(256, k, replace=False)
            for b in bits:
                out[i, b >> 3] ^= np.uint8(1 << (b & 7))
    return out


def perturbed_frame(kps, desc, w, h, seed, shift=25, kmax=90):
    """F2 of SURVEY.md App. E.2: permuted, shifted by integer offsets, angles jittered, bits flipped."""
    rng = np.random.default_rng(seed)
    perm = rng.permutation(len(kps))
    k2 = kps[perm].copy()
    k2["x"] = np.clip(k2["x"] + rng.integers(-shift, shift + 1, len(k2)), 0, w - 1).astype(np.float32)
    k2["y"] = np.clip(k2["y"] + rng.integers(-shift, shift + 1, len(k2)), 0, h - 1).astype(np.float32)
    jitter = rng.random(len(k2)) < 0.8
    ang = np.where(jitter, (k2["angle"] + rng.normal(0, 6, len(k2))) % 360, rng.random(len(k2)) * 360)
    k2["angle"] = ang.astype(np.float32)
    d2 = flip_bits(desc[perm], rng, kmax)
    return k2, d2, perm


def projected_queries(kps2, desc2, nq, seed, nlevels=8):
    """Map-point queries of App. E.2 aimed at frame F2 (every 5th one at the first 40 keypoints: conflicts)."""
    rng = np.random.default_rng(seed)
    tgt = rng.integers(0, len(kps2), nq)
    tgt[::5] = rng.integers(0, min(40, len(kps2)), len(tgt[::5]))
    proj = np.zeros((nq, 3), np.float32)
    proj[:, 0] = kps2["x"][tgt] + rng.normal(0, 2, nq)
    proj[:, 1] = kps2["y"][tgt] + rng.normal(0, 2, nq)
    proj[:, 2] = proj[:, 0] - rng.random(nq).astype(np.float32) * 30
    level = np.minimum(kps2["octave"][tgt] + (rng.random(nq) < 0.3), nlevels - 1).astype(np.int32)
    q = dict(proj=np.ascontiguousarray(proj), level=level,
             view_cos=np.where(rng.random(nq) < 0.5, 0.9995, 0.9).astype(np.float32),
             in_view=(rng.random(nq) < 0.95).astype(np.uint8), bad=(rng.random(nq) < 0.03).astype(np.uint8),
             obs=(rng.random(nq) < 0.9).astype(np.int32), desc=flip_bits(desc2[tgt], rng, 80))
    return q


def two_view_scene(kps, desc, w, h, seed, K, n_cur_extra=0):
    """Last frame = (kps, desc) with synthetic depths; current frame = the same 3-D points seen after a small
    camera motion (keypoints re-projected and rounded, descriptors with bit flips)."""
    rng = np.random.default_rng(seed)
    n = len(kps)
    fx, fy, cx, cy = K
    z = (4 + 30 * rng.random(n)).astype(np.float32)
    xyz = np.stack([(kps["x"] - cx) / fx * z, (kps["y"] - cy) / fy * z, z], 1).astype(np.float32)   # last camera = world
    ang = 0.01
    R = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]], np.float32)
    t = np.array([0.05, -0.02, -0.4], np.float32)
    Tcw_last = np.eye(4, dtype=np.float32)
    Tcw_cur = np.eye(4, dtype=np.float32)
    Tcw_cur[:3, :3] = R
    Tcw_cur[:3, 3] = t
    pc = xyz @ R.T + t
    u = fx * pc[:, 0] / pc[:, 2] + cx + rng.normal(0, 1.5, n)
    v = fy * pc[:, 1] / pc[:, 2] + cy + rng.normal(0, 1.5, n)
    perm = rng.permutation(n)
    kc = kps[perm].copy()
    kc["x"] = np.clip(np.rint(u[perm]), 0, w - 1).astype(np.float32)
    kc["y"] = np.clip(np.rint(v[perm]), 0, h - 1).astype(np.float32)
    kc["angle"] = ((kc["angle"] + rng.normal(0, 5, n)) % 360).astype(np.float32)
    cur = dict(kps=kc, desc=flip_bits(desc[perm], rng, 70),
               u_right=np.where(rng.random(n) < 0.5, kc["x"] - 386.1448 / pc[perm, 2] + rng.normal(0, 1, n), -1).astype(np.float32))
    last = dict(kps=kps.copy(), has_mp=(rng.random(n) < 0.8).astype(np.uint8), outlier=(rng.random(n) < 0.05).astype(np.uint8),
                xyz=np.ascontiguousarray(xyz), mp_desc=flip_bits(desc, rng, 30), mp_obs=(rng.random(n) < 0.9).astype(np.int32))
    return cur, last, Tcw_cur, Tcw_last
