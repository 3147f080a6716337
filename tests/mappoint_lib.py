"""Map-point side of the checkers: MapPoint::ComputeDistinctiveDescriptors,
MapPoint::PredictScale, Frame::isInFrustum.  One Python signature for the reference (oracle/_ref/liborbref.so and
libmappointref.so, the reference's unmodified MapPoint.cc / Frame.cc) and the plain-C oracle.  Test infrastructure."""
import ctypes as C
import os

import numpy as np

from oracle_lib import ORACLE_DIR, REFERENCE_ROOT, build_ref, oracle, ref

vp, ci, cf = C.c_void_p, C.c_int, C.c_float
MPREF_SO = os.path.join(ORACLE_DIR, "_ref", "libmappointref.so")
_mpref = None


def mappoint_ref():
    """The reference's own MapPoint.cc, or None where it was not built."""
    global _mpref
    if _mpref is None:
        if os.path.isdir(REFERENCE_ROOT):
            build_ref()
        if not os.path.exists(MPREF_SO):
            return None
        L = C.CDLL(MPREF_SO)
        L.orbref_mp_distinctive.argtypes = [vp, ci, vp, vp]
        L.orbref_mp_predict_scale.argtypes = [cf, vp, ci, cf, ci, vp, vp]
        _mpref = L
    return _mpref


def _p(a):
    return None if a is None else a.ctypes.data


def descriptor_groups(seed, sizes=(1, 2, 3, 4, 5, 8, 9, 16, 17, 31, 32, 33, 64, 100, 257)):
    """Observation sets of map points: noisy copies of one descriptor (so medians tie often), some key frames bad."""
    rng = np.random.default_rng(seed)
    groups = []
    for n in sizes:
        base = rng.integers(0, 256, (1, 32), dtype=np.uint8)
        d = np.repeat(base, n, 0)
        flips = rng.integers(0, 256, (n, 32), dtype=np.uint8) & rng.integers(0, 256, (n, 32), dtype=np.uint8) & rng.integers(0, 256, (n, 32), dtype=np.uint8)
        d ^= flips
        d[:, 31] = np.arange(n) % 256                         # mostly distinct rows
        d[:, 30] = np.arange(n) // 256
        bad = (rng.random(n) < 0.15).astype(np.uint8)
        groups.append((np.ascontiguousarray(d), bad))
    return groups


def distinctive(impl, desc, bad=None):
    """(chosen descriptor bytes or None, index or None, median or None)"""
    if impl == "ref":
        out = np.zeros(32, np.uint8)
        ok = mappoint_ref().orbref_mp_distinctive(_p(desc), len(desc), _p(bad), _p(out))
        return (out if ok else None), None, None
    O = oracle()
    O.orbo_distinctive_descriptor.argtypes = [vp, ci, vp, vp]
    med = C.c_int(-1)
    i = O.orbo_distinctive_descriptor(_p(desc), len(desc), _p(bad), C.byref(med))
    return (desc[i].copy() if i >= 0 else None), (i if i >= 0 else None), (med.value if i >= 0 else None)


def predict_scale(impl, max_distance, cur_dist, scale_factor=1.2, nlevels=8):
    cur = np.ascontiguousarray(cur_dist, np.float32)
    out = np.zeros(len(cur), np.int32)
    if impl == "ref":
        mappoint_ref().orbref_mp_predict_scale(max_distance, _p(cur), len(cur), scale_factor, nlevels, _p(out), None)
        return out
    O = oracle()
    O.orbo_predict_scale.argtypes = [cf, cf, cf, ci]
    O.orbo_log_scale_factor.restype = cf
    O.orbo_log_scale_factor.argtypes = [cf]
    logs = O.orbo_log_scale_factor(scale_factor)
    for i, c in enumerate(cur):
        out[i] = O.orbo_predict_scale(max_distance, float(c), logs, nlevels)
    return out


K_KITTI = np.float32([718.856, 718.856, 607.1928, 185.2157])


def frustum_scene(seed, n=4000, w=1241, h=376):
    """A cloud of map points around a camera at a general pose: some behind it, outside the image, too near / too
    far for their scale-invariance range, or seen from the wrong side."""
    rng = np.random.default_rng(seed)
    ang = rng.normal(0, 0.2, 3)
    cx_, sx_ = np.cos(ang), np.sin(ang)
    Rx = np.array([[1, 0, 0], [0, cx_[0], -sx_[0]], [0, sx_[0], cx_[0]]])
    Ry = np.array([[cx_[1], 0, sx_[1]], [0, 1, 0], [-sx_[1], 0, cx_[1]]])
    Rz = np.array([[cx_[2], -sx_[2], 0], [sx_[2], cx_[2], 0], [0, 0, 1]])
    R = (Rz @ Ry @ Rx).astype(np.float32)
    t = rng.normal(0, 2, 3).astype(np.float32)
    Tcw = np.eye(4, dtype=np.float32)
    Tcw[:3, :3] = R
    Tcw[:3, 3] = t
    Ow = -(R.T.astype(np.float64) @ t.astype(np.float64))
    # points in camera coordinates, then to world
    z = rng.uniform(-5, 60, n)
    x = rng.normal(0, 0.6, n) * np.abs(z) * 1.2
    y = rng.normal(0, 0.25, n) * np.abs(z) * 1.2
    pc = np.stack([x, y, z], 1)
    pw = (pc - t.astype(np.float64)) @ R.astype(np.float64)          # R^T (pc - t)
    xyz = np.ascontiguousarray(pw, np.float32)
    view = xyz.astype(np.float64) - Ow
    dist = np.linalg.norm(view, axis=1)
    nrm = view / dist[:, None] + rng.normal(0, 0.5, (n, 3))
    nrm /= np.linalg.norm(nrm, axis=1)[:, None]
    flip = rng.random(n) < 0.1
    nrm[flip] *= -1
    max_d = (dist * rng.uniform(0.6, 3.0, n)).astype(np.float32)
    min_d = (max_d / np.float32(1.2 ** 7)).astype(np.float32)
    return dict(Tcw=Tcw, K=K_KITTI, bf=np.float32(386.1448), bounds=(0.0, float(w), 0.0, float(h)), xyz=xyz,
                normal=np.ascontiguousarray(nrm, np.float32), max_d=max_d, min_d=min_d)


def is_in_frustum(impl, s, scale_factor=1.2, nlevels=8, cos_limit=0.5):
    """(in_view u8 [n], proj [n,3], level [n], view_cos [n]); entries of points not in view are zero."""
    n = len(s["xyz"])
    iv, proj, lv, vc = np.zeros(n, np.uint8), np.zeros((n, 3), np.float32), np.zeros(n, np.int32), np.zeros(n, np.float32)
    L = ref() if impl == "ref" else oracle()
    f = L.orbref_is_in_frustum if impl == "ref" else L.orbo_is_in_frustum
    f.argtypes = [vp, vp, cf] + [cf] * 4 + [cf, ci, cf, ci] + [vp] * 8
    cnt = f(_p(s["Tcw"]), _p(s["K"]), float(s["bf"]), *s["bounds"], scale_factor, nlevels, cos_limit, n, _p(s["xyz"]), _p(s["normal"]),
            _p(s["max_d"]), _p(s["min_d"]), _p(iv), _p(proj), _p(lv), _p(vc))
    assert cnt == int(iv.sum())
    return iv, proj, lv, vc
