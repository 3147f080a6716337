"""Scenes + wrappers for the searches whose queries do not depend on one another:
  ORBmatcher::Fuse(KeyFrame*, vpMapPoints, th)                     src/ORBmatcher.cc:1364-1513
  ORBmatcher::Fuse(KeyFrame*, Scw, vpPoints, th, vpReplacePoint)   src/ORBmatcher.cc:1516-1633
  ORBmatcher::SearchBySim3(pKF1, pKF2, vpMatches12, s12, R12, t12, th)   src/ORBmatcher.cc:836-1052
One Python signature for the reference (oracle/_ref, orbref_*) and the plain-C restatement (orbo_*).  Test infrastructure."""
import ctypes as C

import numpy as np

from matcher_lib import flip_bits
from oracle_lib import oracle, ref

vp, ci, cf = C.c_void_p, C.c_int, C.c_float
p = lambda a: None if a is None else a.ctypes.data


def _pose(ang, t):
    R = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]], np.float32)
    return R, np.asarray(t, np.float32)


def _camera_centre(R, t):
    """Ow = -Rcw.t()*tcw in the shim's arithmetic (negation in double, products accumulated in float)."""
    Ow = np.zeros(3, np.float32)
    for i in range(3):
        s = np.float32(np.float32(np.float64(R[0, i]) * -1.0) * t[0])
        s = np.float32(s + np.float32(np.float32(np.float64(R[1, i]) * -1.0) * t[1]))
        s = np.float32(s + np.float32(np.float32(np.float64(R[2, i]) * -1.0) * t[2]))
        Ow[i] = s
    return Ow


def fuse_scene(kps, desc, w, h, seed, K, bf, sim3=False):
    """Key frame = (kps, desc), about half of its keypoints with a right coordinate and 40 % holding a map point already.
    Candidate points = the keypoints back-projected at synthetic depths (noisy), listed in random order with NULLs,
    duplicates, points that are already in the key frame, bad points, points behind the camera / out of range / seen from
    behind."""
    rng = np.random.default_rng(seed)
    n = len(kps)
    fx, fy, cx, cy = K
    R, t = _pose(0.03, [0.2, -0.1, 0.4])
    s = 1.7 if sim3 else 1.0
    z = (4 + 30 * rng.random(n)).astype(np.float32)
    noise = 1.2
    pc = np.stack([(kps["x"] + rng.normal(0, noise, n) - cx) / fx * z, (kps["y"] + rng.normal(0, noise, n) - cy) / fy * z, z], 1)
    xyz = ((pc - t) @ R).astype(np.float32)
    behind = rng.random(n) < 0.04
    xyz[behind] = ((np.stack([pc[behind, 0], pc[behind, 1], -pc[behind, 2]], 1) - t) @ R).astype(np.float32)
    Ow = _camera_centre(R, t)
    PO = xyz - Ow
    dist = np.linalg.norm(PO, axis=1).astype(np.float32)
    normal = (PO / dist[:, None]).astype(np.float32)
    flip = rng.random(n) < 0.1
    normal[flip] *= -1
    u_right = np.where(rng.random(n) < 0.55, kps["x"] - bf / z + rng.normal(0, 1.0, n), -1).astype(np.float32)
    # the universe: rows 0..n-1 = candidates (row i belongs to keypoint i), rows n..n+m-1 = points the key frame holds already
    holders = np.flatnonzero(rng.random(n) < 0.4)
    m = len(holders)
    npts = n + m
    kf_mp = np.full(n, -1, np.int32)
    kf_mp[holders] = n + np.arange(m)
    kf_idx = np.full(npts, -1, np.int32)
    kf_idx[n:] = holders
    # a few candidates ARE the key frame's points (IsInKeyFrame): swap them in for the holder rows
    own = holders[rng.random(m) < 0.1]
    for k in own:
        kf_idx[kf_mp[k]] = -1
        kf_mp[k] = k
        kf_idx[k] = k
    bad = (rng.random(npts) < 0.04).astype(np.uint8)
    nobs = rng.integers(0, 9, npts).astype(np.int32)
    full = lambda a, fill: np.concatenate([a, np.full((m,) + a.shape[1:], fill, a.dtype)])
    xyz_u = np.ascontiguousarray(np.concatenate([xyz, xyz[holders] + 0.01]).astype(np.float32))
    normal_u = np.ascontiguousarray(np.concatenate([normal, normal[holders]]).astype(np.float32))
    mp_desc = np.ascontiguousarray(np.concatenate([flip_bits(desc, rng, 60), desc[holders]]))
    level = full(np.clip(kps["octave"] + rng.integers(0, 2, n), 0, 7).astype(np.int32), 0)
    min_dist = full((dist * np.where(rng.random(n) < 0.05, 1.2, 0.5)).astype(np.float32), 0.0)
    max_dist = full((dist * 1.6).astype(np.float32), 1e9)
    lst = rng.permutation(n).astype(np.int32)
    lst = np.concatenate([lst, rng.integers(0, n, n // 10).astype(np.int32)])      # duplicates
    lst[rng.random(len(lst)) < 0.03] = -1                                          # NULL entries (first overload only)
    if sim3:
        lst = lst[lst >= 0]
    Scw = np.eye(4, dtype=np.float32)
    Scw[:3, :3] = s * R
    Scw[:3, 3] = s * t
    return dict(R=np.ascontiguousarray(R), t=t, Ow=Ow, Scw=Scw, u_right=u_right, kf_mp=kf_mp, kf_idx=kf_idx, bad=bad, nobs=nobs,
                xyz=xyz_u, normal=normal_u, mp_desc=mp_desc, level=np.ascontiguousarray(level), min_dist=np.ascontiguousarray(min_dist),
                max_dist=np.ascontiguousarray(max_dist), list=np.ascontiguousarray(lst), npts=npts)


def run_fuse(impl, kps, desc, s, scale, inv_sigma2, bounds, K, bf, th, sim3=False, stereo=True):
    """Returns (nFused, state) with state = dict(bad, nobs, kf_idx, kf_mp, replaced_by | replace) after the call; the oracle
    also returns the queries it searched."""
    L = ref() if impl == "ref" else oracle()
    n, npts, nlist = len(kps), s["npts"], len(s["list"])
    bad, nobs, kf_idx, kf_mp = s["bad"].copy(), s["nobs"].copy(), s["kf_idx"].copy(), s["kf_mp"].copy()
    replaced_by = np.full(npts, -1, np.int32)
    replace = np.full(nlist, -1, np.int32)
    ur = s["u_right"] if stereo else None
    Kf = np.asarray(K, np.float32)
    if impl == "ref":
        f = L.orbref_fuse
        f.argtypes = [ci, ci, vp, vp, vp] + [cf] * 4 + [vp, vp, ci, vp, cf, vp, vp, vp, vp, ci] + [vp] * 10 + [ci, vp, vp, vp, cf]
        nf = f(int(sim3), n, p(kps), p(desc), p(ur), *bounds, p(scale), p(inv_sigma2), len(scale), p(Kf), bf, p(s["R"]), p(s["t"]), p(s["Ow"]),
               p(s["Scw"]), npts, p(bad), p(s["xyz"]), p(s["normal"]), p(s["mp_desc"]), p(s["level"]), p(s["min_dist"]), p(s["max_dist"]),
               p(nobs), p(kf_idx), p(replaced_by), nlist, p(s["list"]), p(kf_mp), p(replace), th)
        q = None
    else:
        uvr, lvl, qur, valid = np.zeros((nlist, 3), np.float32), np.zeros(nlist, np.int32), np.zeros(nlist, np.float32), np.zeros(nlist, np.uint8)
        if sim3:
            f = L.orbo_fuse_sim3
            f.argtypes = [ci, vp, vp, vp] + [cf] * 4 + [vp, vp, vp, ci] + [vp] * 9 + [ci, vp, vp, vp, cf] + [vp] * 4
            nf = f(n, p(kps), p(desc), p(ur), *bounds, p(scale), p(Kf), p(s["Scw"]), npts, p(bad), p(s["xyz"]), p(s["normal"]), p(s["mp_desc"]),
                   p(s["level"]), p(s["min_dist"]), p(s["max_dist"]), p(nobs), p(kf_idx), nlist, p(s["list"]), p(kf_mp), p(replace), th,
                   p(uvr), p(lvl), p(qur), p(valid))
        else:
            f = L.orbo_fuse
            f.argtypes = [ci, vp, vp, vp] + [cf] * 4 + [vp, vp, vp, cf, vp, vp, vp, ci] + [vp] * 10 + [ci, vp, vp, cf] + [vp] * 4
            nf = f(n, p(kps), p(desc), p(ur), *bounds, p(scale), p(inv_sigma2), p(Kf), bf, p(s["R"]), p(s["t"]), p(s["Ow"]), npts, p(bad),
                   p(s["xyz"]), p(s["normal"]), p(s["mp_desc"]), p(s["level"]), p(s["min_dist"]), p(s["max_dist"]), p(nobs), p(kf_idx),
                   p(replaced_by), nlist, p(s["list"]), p(kf_mp), th, p(uvr), p(lvl), p(qur), p(valid))
        q = dict(uvr=uvr, level=lvl, ur=qur, valid=valid)
    st = dict(bad=bad, nobs=nobs, kf_idx=kf_idx, kf_mp=kf_mp)
    st["replace" if sim3 else "replaced_by"] = replace if sim3 else replaced_by
    return nf, st, q


def replay_fuse(s, q, best_idx, u_right, sim3):
    """What host/ORBmatcher_b200.hpp does with the device's per-point results: the reference's bookkeeping (Replace /
    AddObservation / AddMapPoint / vpReplacePoint) in list order, with the tests that depend on the changing state."""
    bad, nobs, kf_idx, kf_mp = s["bad"].copy(), s["nobs"].copy(), s["kf_idx"].copy(), s["kf_mp"].copy()
    npts, lst = s["npts"], s["list"]
    replaced_by = np.full(npts, -1, np.int32)
    replace = np.full(len(lst), -1, np.int32)
    found0 = np.zeros(npts, bool)
    found0[kf_mp[kf_mp >= 0]] = True

    def add_obs(pt, idx):
        if kf_idx[pt] >= 0:
            return
        kf_idx[pt] = idx
        nobs[pt] += 2 if (u_right is not None and u_right[idx] >= 0) else 1

    def replace_point(a, b):
        if a == b:
            return
        idx = kf_idx[a]
        kf_idx[a] = -1
        bad[a] = 1
        replaced_by[a] = b
        if idx >= 0:
            if kf_idx[b] < 0:
                kf_mp[idx] = b
                add_obs(b, idx)
            else:
                kf_mp[idx] = -1

    nf = 0
    for i, pt in enumerate(lst):
        if pt < 0 or not q["valid"][i]:
            continue
        if bad[pt] or (found0[pt] if sim3 else kf_idx[pt] >= 0):
            continue
        b = best_idx[i]
        if b < 0:
            continue
        inkf = kf_mp[b]
        if inkf >= 0:
            if not bad[inkf]:
                if sim3:
                    replace[i] = inkf
                elif nobs[inkf] > nobs[pt]:
                    replace_point(pt, inkf)
                else:
                    replace_point(inkf, pt)
        else:
            add_obs(pt, b)
            kf_mp[b] = pt
        nf += 1
    st = dict(bad=bad, nobs=nobs, kf_idx=kf_idx, kf_mp=kf_mp)
    st["replace" if sim3 else "replaced_by"] = replace if sim3 else replaced_by
    return nf, st


def same_state(a, b):
    return all((a[k] == b[k]).all() for k in a)


def window_best_free_oracle(kps, desc, u_right, bounds, q, qdesc, inv_sigma2, th_accept):
    L = oracle()
    nq = len(q["level"])
    bi, bd = np.zeros(nq, np.int32), np.zeros(nq, np.int32)
    f = L.orbo_window_best_free
    f.argtypes = [ci, vp, vp, vp] + [cf] * 4 + [ci] + [vp] * 6 + [ci, vp, vp]
    # the key frame's grid origin is its int-truncated bound
    b = (float(int(bounds[0])), bounds[1], float(int(bounds[2])), bounds[3])
    n = f(len(kps), p(kps), p(desc), p(u_right), *b, nq, p(q["uvr"]), p(q["level"]), p(q.get("ur")), p(q.get("valid")), p(qdesc), p(inv_sigma2),
          th_accept, p(bi), p(bd))
    return n, bi, bd


# ------------------------------------------------------------------------------------------------ SearchBySim3
def sim3_pair_scene(k1, d1, k2, d2, w, h, seed, K):
    """Two key frames of the same place: KF2's keypoints are KF1's shifted (tests/matcher_lib.perturbed_frame), so map points
    seen in KF1 project near their twins in KF2 through the relative similarity.  The universe holds KF1's points (rows
    0..n1-1, row i at keypoint i of KF1 when mp1[i] >= 0) and KF2's (rows n1..n1+n2-1)."""
    rng = np.random.default_rng(seed)
    n1, n2 = len(k1), len(k2)
    fx, fy, cx, cy = K
    s12 = np.float32(1.3)
    # Both key frames look along the same ray bundle and differ by the scale drift the loop closer is there to find: with
    # R1 = R2 and t2 = t1 / s12 the similarity p1 = s12 * R12 * p2 + t12 (R12 = R1 R2^T, t12 = t1 - s12 R12 t2) is a pure scaling
    # of camera coordinates, so a point keeps its pixel and lands on its twin keypoint in the other key frame.
    R1, t1 = _pose(0.02, [0.1, 0.05, 0.2])
    R2, t2 = R1.copy(), (t1 / s12).astype(np.float32)
    R12 = (R1 @ R2.T).astype(np.float32)
    t12 = (t1 - s12 * (R12 @ t2)).astype(np.float32) + np.float32([0.002, -0.001, 0.003])
    # KF1's points: back-project its keypoints in camera 1
    z1 = (4 + 30 * rng.random(n1)).astype(np.float32)
    pc1 = np.stack([(k1["x"] + rng.normal(0, 1.5, n1) - cx) / fx * z1, (k1["y"] + rng.normal(0, 1.5, n1) - cy) / fy * z1, z1], 1)
    xyz1 = ((pc1 - t1) @ R1).astype(np.float32)
    # KF2's points: back-project ITS keypoints in camera 2 (the points live at 1/s12 of the KF1 scale there)
    z2 = (4 + 30 * rng.random(n2)).astype(np.float32) / s12
    pc2 = np.stack([(k2["x"] + rng.normal(0, 1.5, n2) - cx) / fx * z2, (k2["y"] + rng.normal(0, 1.5, n2) - cy) / fy * z2, z2], 1)
    xyz2 = ((pc2 - t2) @ R2).astype(np.float32)
    npts = n1 + n2
    mp1 = np.where(rng.random(n1) < 0.85, np.arange(n1), -1).astype(np.int32)
    mp2 = np.where(rng.random(n2) < 0.85, n1 + np.arange(n2), -1).astype(np.int32)
    xyz = np.ascontiguousarray(np.concatenate([xyz1, xyz2]).astype(np.float32))
    mp_desc = np.ascontiguousarray(np.concatenate([flip_bits(d1, rng, 50), flip_bits(d2, rng, 50)]))
    level = np.concatenate([np.clip(k1["octave"] + rng.integers(0, 2, n1), 0, 7), np.clip(k2["octave"] + rng.integers(0, 2, n2), 0, 7)]).astype(np.int32)
    bad = (rng.random(npts) < 0.03).astype(np.uint8)
    min_dist = np.where(rng.random(npts) < 0.05, 1e3, 0.1).astype(np.float32)
    max_dist = np.full(npts, 1e3, np.float32)
    # matches found earlier (SearchByBoW): KF1 keypoint i already matched to a point of KF2
    m12 = np.full(n1, -1, np.int32)
    pre = np.flatnonzero(rng.random(n1) < 0.1) if n2 else np.zeros(0, np.int64)
    m12[pre] = n1 + rng.integers(0, max(n2, 1), len(pre))
    idx_in_kf2 = np.full(npts, -1, np.int32)
    idx_in_kf2[n1:] = np.where(mp2 >= 0, np.arange(n2), -1)
    return dict(R1=np.ascontiguousarray(R1), t1=t1, R2=np.ascontiguousarray(R2), t2=t2, s12=float(s12), R12=np.ascontiguousarray(R12), t12=t12,
                mp1=mp1, mp2=mp2, npts=npts, xyz=xyz, mp_desc=mp_desc, level=np.ascontiguousarray(level), bad=bad, min_dist=min_dist,
                max_dist=max_dist, m12=m12, idx_in_kf2=idx_in_kf2)


def run_search_by_sim3(impl, k1, d1, k2, d2, s, scale, bounds, K, th):
    L = ref() if impl == "ref" else oracle()
    m12 = s["m12"].copy()
    Kf = np.asarray(K, np.float32)
    head = [len(k1), p(k1), p(d1), p(s["mp1"]), len(k2), p(k2), p(d2), p(s["mp2"]), *bounds, p(scale)]
    mid = [p(Kf), p(s["R1"]), p(s["t1"]), p(s["R2"]), p(s["t2"]), s["s12"], p(s["R12"]), p(s["t12"]), s["npts"], p(s["bad"]), p(s["xyz"]),
           p(s["mp_desc"]), p(s["level"]), p(s["min_dist"]), p(s["max_dist"]), p(s["idx_in_kf2"]), p(m12), th]
    base = [ci, vp, vp, vp, ci, vp, vp, vp] + [cf] * 4 + [vp]
    tail = [vp] * 5 + [cf, vp, vp, ci] + [vp] * 8 + [cf]
    if impl == "ref":
        f = L.orbref_search_by_sim3
        f.argtypes = base + [ci] + tail
        return f(*head, len(scale), *mid), m12, None
    n1, n2 = len(k1), len(k2)
    q1 = dict(uvr=np.zeros((n1, 3), np.float32), level=np.zeros(n1, np.int32), valid=np.zeros(n1, np.uint8))
    q2 = dict(uvr=np.zeros((n2, 3), np.float32), level=np.zeros(n2, np.int32), valid=np.zeros(n2, np.uint8))
    f = L.orbo_search_by_sim3
    f.argtypes = base + tail + [vp] * 6
    nf = f(*head, *mid, p(q1["uvr"]), p(q1["level"]), p(q1["valid"]), p(q2["uvr"]), p(q2["level"]), p(q2["valid"]))
    return nf, m12, (q1, q2)


# ------------------------------------------------------------------------------------------------ the projection prologue alone
def fuse_project_oracle(sim3, pose24, K, bf, bounds, scale_factor, scale, th, xyz, normal, max_raw, min_raw, skip):
    """orbo_fuse_project: (uvr, level, ur, valid) per point."""
    n = len(max_raw)
    uvr, lvl, ur, valid = np.zeros((n, 3), np.float32), np.zeros(n, np.int32), np.zeros(n, np.float32), np.zeros(n, np.uint8)
    f = oracle().orbo_fuse_project
    f.argtypes, f.restype = [ci, vp, vp] + [cf] * 6 + [vp, ci, cf, ci] + [vp] * 9, None
    Kf = np.asarray(K, np.float32)
    b = (float(int(bounds[0])), float(int(bounds[1])), float(int(bounds[2])), float(int(bounds[3])))
    f(int(sim3), p(pose24), p(Kf), bf, *b, scale_factor, p(scale), len(scale), th, n, p(xyz), p(normal), p(max_raw), p(min_raw), p(skip),
      p(uvr), p(lvl), p(ur), p(valid))
    return dict(uvr=uvr, level=lvl, ur=ur, valid=valid)


def decompose_sim3(Scw):
    """R, t, Ow of src/ORBmatcher.cc:1524-1529 in the shim's arithmetic."""
    d = np.float64(0)
    for c in range(3):
        d += np.float64(Scw[0, c]) * np.float64(Scw[0, c])
    scw = np.float32(np.sqrt(d))
    inv = np.float64(1.0) / np.float64(scw)
    R = (Scw[:3, :3].astype(np.float64) * inv).astype(np.float32)
    t = (Scw[:3, 3].astype(np.float64) * inv).astype(np.float32)
    return np.ascontiguousarray(R), t, _camera_centre(R, t)


def raw_distances(s, rng, lo=1.0, hi=3.5):
    """mfMaxDistance / mfMinDistance for the points of a scene (the scenes above carry the INVARIANCE bounds the mocks return):
    raw values around the true distances, and the scene's bounds replaced by 1.2 * max / 0.8 * min as MapPoint computes them."""
    n = len(s["max_dist"])
    max_raw = (s["max_dist"] / np.float32(1.6) * (lo + (hi - lo) * rng.random(n))).astype(np.float32)
    min_raw = (max_raw / np.float32(1.2 ** 7) * np.where(rng.random(n) < 0.05, 8.0, 1.0)).astype(np.float32)
    s2 = dict(s)
    s2["max_dist"] = (np.float32(1.2) * max_raw).astype(np.float32)
    s2["min_dist"] = (np.float32(0.8) * min_raw).astype(np.float32)
    return s2, max_raw, min_raw


def fuse_pose24(s, sim3):
    R, t, Ow = decompose_sim3(s["Scw"]) if sim3 else (s["R"], s["t"], s["Ow"])
    return np.ascontiguousarray(np.concatenate([R.ravel(), t, Ow, np.zeros(9, np.float32)]).astype(np.float32))


def fuse_list_arrays(s, max_raw, min_raw, sim3):
    """Per LIST ENTRY arrays for the projection: the entry's point (a NULL entry = skipped dummy) and the reference's skip test at
    the start of the call (NULL, isBad(), IsInKeyFrame(pKF) / member of the already-found set)."""
    lst = s["list"]
    g = np.maximum(lst, 0)
    found0 = np.zeros(s["npts"], bool)
    found0[s["kf_mp"][s["kf_mp"] >= 0]] = True
    skip = (lst < 0) | (s["bad"][g] != 0) | (found0[g] if sim3 else (s["kf_idx"][g] >= 0))
    c = np.ascontiguousarray
    return dict(xyz=c(s["xyz"][g]), normal=c(s["normal"][g]), max_d=c(max_raw[g]), min_d=c(min_raw[g]), skip=c(skip.astype(np.uint8)),
                desc=c(s["mp_desc"][g]))


def sim3_poses(s):
    """pose24 of the two directions of SearchBySim3 in the reference's arithmetic (:854-856)."""
    s12, R12, t12 = np.float32(s["s12"]), s["R12"], s["t12"]
    inv_s = np.float64(1.0) / np.float64(s12)
    sR12 = (R12.astype(np.float64) * np.float64(s12)).astype(np.float32)
    sR21 = (R12.T.astype(np.float64) * inv_s).astype(np.float32)
    ns = (sR21.astype(np.float64) * -1.0).astype(np.float32)
    t21 = np.zeros(3, np.float32)
    for r in range(3):
        a = np.float32(ns[r, 0] * t12[0])
        a = np.float32(a + np.float32(ns[r, 1] * t12[1]))
        a = np.float32(a + np.float32(ns[r, 2] * t12[2]))
        t21[r] = a
    cat = lambda *a: np.ascontiguousarray(np.concatenate([np.asarray(x, np.float32).ravel() for x in a]))
    return cat(s["R1"], s["t1"], sR21, t21), cat(s["R2"], s["t2"], sR12, t12)
