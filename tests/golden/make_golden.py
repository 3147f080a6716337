#!/usr/bin/env python3
"""Generates the committed fixtures in tests/golden/ (run in the authoring container).

cv2_primitives.npz   outputs of Python cv2 4.13.0 / libm for one small image: pins oracle/cv_prims.c
ref_extract_*.npz    outputs of the reference's UNMODIFIED ORBextractor.cc (oracle/_ref/liborbref.so,
                     built from /root/reference against oracle/cvshim + bump arena): pins the oracle
                     restatement and, through it, the CUDA path
ref_match_*.npz      outputs of the reference's UNMODIFIED ORBmatcher.cc on mock frames
ref_fuse.npz         the same for ORBmatcher::Fuse (both overloads) and ORBmatcher::SearchBySim3
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle_lib import RefExtractor  # noqa: E402
from synth import synth_frame  # noqa: E402


def primitives():
    import cv2
    assert cv2.__version__.startswith("4.13"), cv2.__version__
    img = synth_frame(160, 120, 77)
    det = lambda t: cv2.FastFeatureDetector_create(threshold=t, nonmaxSuppression=True, type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    fast = lambda t: np.array([(int(k.pt[0]), int(k.pt[1]), int(k.response)) for k in det(t).detect(img)], np.int32).reshape(-1, 3)
    rng = np.random.default_rng(1)
    atan_in = rng.integers(-2 ** 20, 2 ** 20, (2000, 2)).astype(np.float32)
    atan_out = np.float32([cv2.fastAtan2(float(y), float(x)) for y, x in atan_in])
    libm = C.CDLL("libm.so.6")
    for f in ("sinf", "cosf"):
        getattr(libm, f).restype = C.c_float
        getattr(libm, f).argtypes = [C.c_float]
    angles = (rng.random(5000) * 2 * np.pi).astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "cv2_primitives.npz"), img=img,
                        resized=cv2.resize(img, (133, 100), interpolation=cv2.INTER_LINEAR),
                        blurred=cv2.GaussianBlur(img, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101),
                        bordered=cv2.copyMakeBorder(img, 19, 19, 19, 19, cv2.BORDER_REFLECT_101),
                        fast20=fast(20), fast7=fast(7), atan_in=atan_in, atan_out=atan_out, angles=angles,
                        sin=np.float32([libm.sinf(float(a)) for a in angles]),
                        cos=np.float32([libm.cosf(float(a)) for a in angles]))


def extractor():
    for name, (w, h, nf, seed) in {"tum": (640, 480, 1000, 1), "kitti": (1241, 376, 2000, 2), "small": (200, 150, 300, 5)}.items():
        R = RefExtractor(nf)
        n, kps, desc = R.extract(synth_frame(w, h, seed))
        t = R.tables()
        np.savez_compressed(os.path.join(HERE, f"ref_extract_{name}.npz"), w=w, h=h, nfeatures=nf, seed=seed,
                            kps=kps, desc=desc, **{"tab_" + k: v for k, v in t.items()})
        R.close()


def matchers():
    """Outputs of the reference's unmodified ORBmatcher.cc / Frame.cc on the deterministic scenes of
    tests/matcher_lib.py (inputs are regenerated from seeds; only results are stored)."""
    import ctypes as C
    from matcher_lib import Matcher, extract_frame, perturbed_frame, projected_queries, two_view_scene
    from oracle_lib import KP_DTYPE, ref
    from synth import stereo_pair
    W, H, NF = 1241, 376, 2000
    B = (0.0, float(W), 0.0, float(H))
    K = np.float32([718.856, 718.856, 607.1928, 185.2157])
    kps, desc, scale = extract_frame(W, H, NF, 2)
    k2, d2, _ = perturbed_frame(kps, desc, W, H, 11)
    R = Matcher("ref")
    out = {}
    prev = np.stack([kps["x"], kps["y"]], 1)
    n, m12, pv = R.search_for_initialization(kps, desc, k2, d2, scale, B, prev, 100, 0.9, True)
    out.update(init_n=n, init_m12=m12, init_prev=pv)
    q = projected_queries(k2, d2, 2000, 5)
    for th in (1.0, 3.0):
        n, a = R.search_by_projection_points(k2, d2, None, scale, B, q, th, 0.8, None)
        out[f"points_th{int(th)}_n"] = n; out[f"points_th{int(th)}_assign"] = a
    cur, last, Tc, Tl = two_view_scene(kps, desc, W, H, 21, K)
    n, a = R.search_by_projection_frame(cur, last, Tc, Tl, K, 386.1448, scale, B, 7.0, False, 0.9, True, None)
    out.update(frame_n=n, frame_assign=a)
    left, right = stereo_pair(W, H, 2)
    L = ref()
    cap = 2 * NF
    kl, kr = np.zeros(cap, KP_DTYPE), np.zeros(cap, KP_DTYPE)
    dl, dr = np.zeros((cap, 32), np.uint8), np.zeros((cap, 32), np.uint8)
    ur, dep, nr = np.zeros(cap, np.float32), np.zeros(cap, np.float32), C.c_int()
    vp, ci, cf = C.c_void_p, C.c_int, C.c_float
    L.orbref_stereo_frame.argtypes = [vp, vp, ci, ci, ci, cf, ci, ci, ci, cf, cf, cf, cf, cf, cf, vp, vp, ci, vp, vp, ci, C.POINTER(ci), vp, vp]
    n = L.orbref_stereo_frame(left.ctypes.data, right.ctypes.data, W, H, NF, 1.2, 8, 20, 7, float(K[0]), float(K[1]), float(K[2]), float(K[3]),
                              386.1448, 35.0, kl.ctypes.data, dl.ctypes.data, cap, kr.ctypes.data, dr.ctypes.data, cap, C.byref(nr),
                              ur.ctypes.data, dep.ctypes.data)
    out.update(stereo_n=n, stereo_nr=nr.value, stereo_uright=ur[:n], stereo_depth=dep[:n])
    np.savez_compressed(os.path.join(HERE, "ref_match_kitti.npz"), **out)


def projection_overloads():
    """Relocalisation (src/ORBmatcher.cc:303-431) and loop-closing (:434-549) overloads of the reference's unmodified
    ORBmatcher.cc on the scenes of tests/reloc_lib.py / tests/sim3_lib.py (regenerated from seeds; results stored)."""
    from matcher_lib import extract_frame, perturbed_frame
    from reloc_lib import reloc_scene, run_reloc
    from sim3_lib import run_sim3, sim3_scene
    W, H, NF = 1241, 376, 2000
    B = (0.0, float(W), 0.0, float(H))
    K = np.float32([718.856, 718.856, 607.1928, 185.2157])
    kps, desc, scale = extract_frame(W, H, NF, 2)
    k2, d2, _ = perturbed_frame(kps, desc, W, H, 11)
    out = {}
    s = reloc_scene(kps, desc, W, H, 31, K)
    for th, od, co in ((10.0, 100, True), (3.0, 64, True)):
        n, a = run_reloc("ref", s, scale, B, K, th, od, co)
        out[f"reloc_th{int(th)}_n"] = n; out[f"reloc_th{int(th)}_assign"] = a
    for th, seed in ((10, 41), (4, 42)):
        s3 = sim3_scene(k2, d2, W, H, seed, K)
        n, a = run_sim3("ref", k2, d2, s3, scale, B, K, th)
        out[f"sim3_th{th}_n"] = n; out[f"sim3_th{th}_assign"] = a
    np.savez_compressed(os.path.join(HERE, "ref_match_projection.npz"), **out)


def mappoint():
    """Outputs of the reference's unmodified MapPoint.cc (ComputeDistinctiveDescriptors, PredictScale) and
    Frame::isInFrustum on the deterministic scenes of tests/mappoint_lib.py."""
    from mappoint_lib import descriptor_groups, distinctive, frustum_scene, is_in_frustum, predict_scale
    out = dict(desc_seed=9, frustum_seed=7, predict_max=np.float32(12.5))
    for k, (d, bad) in enumerate(descriptor_groups(9)):
        for tag, b in (("all", None), ("bad", bad)):
            w = distinctive("ref", d, b)[0]
            out[f"distinctive_{tag}_{k}"] = np.zeros(0, np.uint8) if w is None else w
    rng = np.random.default_rng(3)
    cur = np.concatenate([np.float32(12.5) / np.float32(1.2) ** np.arange(-3, 12), rng.uniform(0.05, 200, 2000)]).astype(np.float32)
    out["predict_cur"] = cur
    out["predict_level"] = predict_scale("ref", 12.5, cur)
    iv, proj, lv, vc = is_in_frustum("ref", frustum_scene(7))
    out.update(frustum_in_view=iv, frustum_proj=proj, frustum_level=lv, frustum_view_cos=vc)
    np.savez_compressed(os.path.join(HERE, "ref_mappoint.npz"), **out)


def bow():
    """Outputs of the reference's unmodified ORBmatcher::SearchByBoW overloads on the scene of tests/bow_lib.py."""
    from bow_lib import bow_scene, search_by_bow
    from matcher_lib import extract_frame
    from test_bow_oracle import CASES
    kps, desc, _ = extract_frame(1241, 376, 2000, 2)
    s = bow_scene(kps, desc, 5)
    out = dict(seed=5)
    for k, (nnratio, check_ori, kf_kf) in enumerate(CASES):
        nm, m = search_by_bow("ref", s, nnratio, check_ori, kf_kf)
        out[f"nm_{k}"] = nm
        out[f"match_{k}"] = m
    np.savez_compressed(os.path.join(HERE, "ref_bow.npz"), **out)


def fuse():
    """Outputs of the reference's unmodified ORBmatcher::Fuse (both overloads, src/ORBmatcher.cc:1364-1633) and
    ORBmatcher::SearchBySim3 (:836-1052) on the scenes of tests/fuse_lib.py (regenerated from seeds; results stored)."""
    from fuse_lib import fuse_scene, run_fuse, run_search_by_sim3, sim3_pair_scene
    from matcher_lib import extract_frame, perturbed_frame
    from test_fuse_oracle import BF, BOUNDS, FUSE_CASES, H, K, NF, SIM3_CASES, W
    kps, desc, scale = extract_frame(W, H, NF, 2)
    k2, d2, _ = perturbed_frame(kps, desc, W, H, 11, shift=4, kmax=40)
    inv_sigma2 = (np.float32(1.0) / (scale * scale)).astype(np.float32)
    out = {}
    for th, seed, sim3, stereo in FUSE_CASES:
        s = fuse_scene(kps, desc, W, H, seed, K, BF, sim3)
        nf, st, _ = run_fuse("ref", kps, desc, s, scale, inv_sigma2, BOUNDS, K, BF, th, sim3, stereo)
        out[f"fuse_{seed}_n"] = nf
        for k, v in st.items():
            out[f"fuse_{seed}_{k}"] = v
    for th, seed in SIM3_CASES:
        s = sim3_pair_scene(kps, desc, k2, d2, W, H, seed, K)
        nf, m12, _ = run_search_by_sim3("ref", kps, desc, k2, d2, s, scale, BOUNDS, K, th)
        out[f"sim3_{seed}_n"] = nf
        out[f"sim3_{seed}_m12"] = m12
    np.savez_compressed(os.path.join(HERE, "ref_fuse.npz"), **out)


def triang():
    """Outputs of the reference's unmodified ORBmatcher::SearchForTriangulation (src/ORBmatcher.cc:1183-1361) on the scenes of
    tests/triang_lib.py (regenerated from seeds; results stored)."""
    from matcher_lib import extract_frame
    from test_triang_oracle import CASES, H, K, NF, W
    from triang_lib import search_for_triangulation, triang_scene
    kps, desc, scale = extract_frame(W, H, NF, 2)
    sigma2 = (scale * scale).astype(np.float32)
    out = {}
    for seed, only_stereo, check_ori, mono, n2 in CASES:
        s = triang_scene(kps, desc, W, H, seed, K, scale, n2)
        nm, m12 = search_for_triangulation("ref", s, K, scale, sigma2, only_stereo, check_ori, mono)
        out[f"triang_{seed}_n"] = nm
        out[f"triang_{seed}_m12"] = m12
    np.savez_compressed(os.path.join(HERE, "ref_triang.npz"), **out)


def frame_steps():
    """cv2.undistortPoints (OpenCV 4.13) on the keypoints / cameras of tests/test_frame_steps.py, and the reference's unmodified
    Frame::ComputeStereoFromRGBD on its RGB-D scene."""
    import ctypes as C
    import cv2
    from oracle_lib import ref
    from test_frame_steps import CAMS, depth_image, keypoints, undistort_oracle
    assert cv2.__version__.startswith("4.13"), cv2.__version__
    out = {}
    k = keypoints(11)
    for name, (K, D) in CAMS.items():
        Km = np.float32([[K[0], 0, K[2]], [0, K[1], K[3]], [0, 0, 1]])
        w = cv2.undistortPoints(np.stack([k["x"], k["y"]], 1).reshape(-1, 1, 2), Km, np.float32(D), None, Km).reshape(-1, 2)
        out[name] = np.stack([k["x"], k["y"]], 1) if D[0] == 0.0 else w
    k = keypoints(4)
    k["x"], k["y"] = np.clip(k["x"], 0, 639.9), np.clip(k["y"], 0, 479.9)
    ku = undistort_oracle(k, *CAMS["tum1"])
    d = depth_image(4)
    ur, dz = np.zeros(len(k), np.float32), np.zeros(len(k), np.float32)
    L = ref()
    L.orbref_stereo_from_rgbd.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p]
    L.orbref_stereo_from_rgbd(len(k), k.ctypes.data, ku.ctypes.data, d.ctypes.data, 640, 480, 40.0, ur.ctypes.data, dz.ctypes.data)
    out["rgbd_ur"], out["rgbd_depth"] = ur, dz
    np.savez_compressed(os.path.join(HERE, "cv2_undistort.npz"), **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["primitives", "extractor", "matchers", "projection_overloads", "mappoint", "bow", "fuse", "triang", "frame_steps"]
    for name in which:
        globals()[name]()
    print("golden fixtures written to", HERE)
