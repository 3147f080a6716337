"""The per-keypoint steps of the monocular / RGB-D Frame constructors (src/Frame.cc:127-240):
  Frame::UndistortKeyPoints      src/Frame.cc:436-468  = cv::undistortPoints of OpenCV 4.13 (third-party arithmetic): the CPU
                                 restatement against cv2 live (where importable) and against tests/golden/cv2_undistort.npz
  Frame::ComputeStereoFromRGBD   src/Frame.cc:702-727: the restatement against the reference's own Frame.cc (oracle/_ref)
and, on the GPU, orbx_undistort_keypoints_batch / orbx_stereo_from_rgbd_batch (csrc/orb_frame.cu) against the restatement."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle_lib import KP_DTYPE, oracle, ref

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cv2_undistort.npz")
vp, ci, cf = C.c_void_p, C.c_int, C.c_float
p = lambda a: None if a is None else a.ctypes.data
# fx fy cx cy | distortion: the TUM1 / TUM2 / TUM3 (no distortion) settings of the upstream examples, a fisheye-ish 8-coefficient
# model and a strong barrel where the inverse model's denominator changes sign for far-away points
CAMS = {
    "tum1": ([517.306408, 516.469215, 318.643040, 255.313989], [0.262383, -0.953104, -0.005358, 0.002628, 1.163314]),
    "tum2": ([520.908620, 521.007327, 325.141442, 249.701764], [0.231222, -0.784899, -0.003257, -0.000105, 0.917205]),
    "tum3": ([535.4, 539.2, 320.1, 247.6], [0.0, 0.0, 0.0, 0.0]),
    "four": ([458.654, 457.296, 367.215, 248.375], [-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05]),
    "eight": ([600.0, 600.0, 320.0, 240.0], [0.1, -0.05, 0.001, -0.002, 0.01, 0.2, -0.1, 0.03]),
    "twelve": ([600.0, 600.0, 320.0, 240.0], [0.1, -0.05, 0.001, -0.002, 0.01, 0.2, -0.1, 0.03, 0.002, -0.001, 0.0015, 0.0005]),
    "barrel": ([300.0, 300.0, 320.0, 240.0], [-2.5, 0.3, 0.0, 0.0, 0.0]),
}


def keypoints(seed, n=3000, w=640, h=480):
    rng = np.random.default_rng(seed)
    k = np.zeros(n, KP_DTYPE)
    k["x"] = (rng.random(n) * (w + 40) - 20).astype(np.float32)       # a few outside the image
    k["y"] = (rng.random(n) * (h + 40) - 20).astype(np.float32)
    k["size"], k["angle"], k["response"] = 31.0, rng.random(n) * 360, rng.integers(1, 255, n)
    k["octave"], k["class_id"] = rng.integers(0, 8, n), -1
    return k


def undistort_oracle(k, K, D):
    out = np.zeros(len(k), KP_DTYPE)
    f = oracle().orbo_undistort_keypoints
    f.argtypes, f.restype = [ci, vp, vp, vp, ci, vp], None
    Kf, Df = np.float32(K), np.float32(D)
    f(len(k), p(k), p(Kf), p(Df), len(Df), p(out))
    return out


def rgbd_oracle(k, ku, depth, bf):
    ur, dz = np.zeros(len(k), np.float32), np.zeros(len(k), np.float32)
    f = oracle().orbo_stereo_from_rgbd
    f.argtypes, f.restype = [ci, vp, vp, vp, ci, ci, cf, vp, vp], None
    f(len(k), p(k), p(ku), p(depth), depth.shape[1], depth.shape[0], bf, p(ur), p(dz))
    return ur, dz


def depth_image(seed, w=640, h=480):
    rng = np.random.default_rng(seed)
    d = (0.5 + 8 * rng.random((h, w))).astype(np.float32)
    d[rng.random((h, w)) < 0.2] = 0.0                                 # holes of the sensor
    d[rng.random((h, w)) < 0.01] = -1.0
    return np.ascontiguousarray(d)


def test_undistort_restatement_equals_stored_cv2_results():
    g = np.load(GOLDEN)
    for name, (K, D) in CAMS.items():
        k = keypoints(11)
        out = undistort_oracle(k, K, D)
        assert (out["x"].view(np.uint32) == g[name][:, 0].view(np.uint32)).all() and (out["y"].view(np.uint32) == g[name][:, 1].view(np.uint32)).all(), name
        for fld in ("size", "angle", "response", "octave", "class_id"):
            assert (out[fld] == k[fld]).all()
    assert (undistort_oracle(keypoints(11), *CAMS["tum3"])["x"] == keypoints(11)["x"]).all()      # mDistCoef[0] == 0: a copy


def test_undistort_restatement_equals_cv2_live():
    cv2 = pytest.importorskip("cv2")
    for seed in (1, 2, 3):
        for name, (K, D) in CAMS.items():
            k = keypoints(seed, 20000)
            Km = np.float32([[K[0], 0, K[2]], [0, K[1], K[3]], [0, 0, 1]])
            want = cv2.undistortPoints(np.stack([k["x"], k["y"]], 1).reshape(-1, 1, 2), Km, np.float32(D), None, Km).reshape(-1, 2)
            if D[0] == 0.0:
                want = np.stack([k["x"], k["y"]], 1)                  # the reference returns before calling OpenCV (:438-442)
            out = undistort_oracle(k, K, D)
            assert (out["x"].view(np.uint32) == want[:, 0].view(np.uint32)).all() and (out["y"].view(np.uint32) == want[:, 1].view(np.uint32)).all(), (seed, name)


@pytest.mark.skipif(ref() is None, reason="oracle/_ref/liborbref.so not built (needs /root/reference)")
def test_rgbd_restatement_equals_the_reference():
    L = ref()
    L.orbref_stereo_from_rgbd.argtypes = [ci, vp, vp, vp, ci, ci, cf, vp, vp]
    for seed in (4, 5):
        k = keypoints(seed)
        k["x"], k["y"] = np.clip(k["x"], 0, 639.9), np.clip(k["y"], 0, 479.9)      # the reference reads out of bounds otherwise
        ku = undistort_oracle(k, *CAMS["tum1"])
        d = depth_image(seed)
        ur, dz = np.zeros(len(k), np.float32), np.zeros(len(k), np.float32)
        L.orbref_stereo_from_rgbd(len(k), p(k), p(ku), p(d), 640, 480, 40.0, p(ur), p(dz))
        o_ur, o_dz = rgbd_oracle(k, ku, d, 40.0)
        assert (ur.view(np.uint32) == o_ur.view(np.uint32)).all() and (dz.view(np.uint32) == o_dz.view(np.uint32)).all()
        assert (dz > 0).sum() > 1000 and (dz < 0).sum() > 300


def test_rgbd_restatement_equals_stored_reference_results():
    g = np.load(GOLDEN)
    k = keypoints(4)
    k["x"], k["y"] = np.clip(k["x"], 0, 639.9), np.clip(k["y"], 0, 479.9)
    ku = undistort_oracle(k, *CAMS["tum1"])
    o_ur, o_dz = rgbd_oracle(k, ku, depth_image(4), 40.0)
    assert (g["rgbd_ur"].view(np.uint32) == o_ur.view(np.uint32)).all() and (g["rgbd_depth"].view(np.uint32) == o_dz.view(np.uint32)).all()


@pytest.mark.gpu
def test_device_steps_equal_the_restatement():
    import torch
    import orb_slam2_chinesenotes_b200 as ob
    P, cap = 5, 3100
    ks = [keypoints(20 + f, n) for f, n in enumerate((3000, 2500, 0, 3100, 17))]
    kp = np.zeros((P, cap), KP_DTYPE)
    for f, k in enumerate(ks):
        kp[f, :len(k)] = k
    d_k = torch.from_numpy(kp.view(np.uint8).reshape(P, cap, 28)).cuda()
    d_n = torch.from_numpy(np.int32([len(k) for k in ks])).cuda()
    depth = np.stack([depth_image(30 + f) for f in range(P)])
    d_depth_wide = torch.zeros((P, 480, 704), dtype=torch.float32, device="cuda")   # rows wider than the image: a pitch
    d_depth_wide[:, :, :640] = torch.from_numpy(depth).cuda()
    d_depth = d_depth_wide[:, :, :640]
    for name, (K, D) in CAMS.items():
        d_u = torch.full((P, cap, 28), 0xAB, dtype=torch.uint8, device="cuda")
        ob.undistort_keypoints_batch(d_k, d_u, d_n, K, D)
        ur = torch.full((P, cap), 7.0, device="cuda")
        dz = torch.full((P, cap), 7.0, device="cuda")
        ob.stereo_from_rgbd_batch(d_k, d_u, d_n, d_depth, 40.0, ur, dz)
        torch.cuda.synchronize()
        got = d_u.cpu().numpy().view(KP_DTYPE).reshape(P, cap)
        ur, dz = ur.cpu().numpy(), dz.cpu().numpy()
        for f, k in enumerate(ks):
            want = undistort_oracle(k, K, D)
            assert got[f, :len(k)].tobytes() == want.tobytes(), (name, f)
            assert (d_u[f, len(k):].cpu().numpy() == 0xAB).all()                    # nothing written behind a frame's keypoints
            w_ur, w_dz = rgbd_oracle(k, want, depth[f], 40.0)
            assert (ur[f, :len(k)].view(np.uint32) == w_ur.view(np.uint32)).all() and (dz[f, :len(k)].view(np.uint32) == w_dz.view(np.uint32)).all(), (name, f)
            assert (ur[f, len(k):] == -1).all() and (dz[f, len(k):] == -1).all()
    # in place, and without undistorted keypoints
    d_c = d_k.clone()
    ob.undistort_keypoints_batch(d_c, d_c, d_n, *CAMS["tum1"])
    ob.stereo_from_rgbd_batch(d_k, None, d_n, d_depth, 40.0, ur_t := torch.zeros((P, cap), device="cuda"), torch.zeros((P, cap), device="cuda"))
    torch.cuda.synchronize()
    assert d_c[0, :3000].cpu().numpy().tobytes() == undistort_oracle(ks[0], *CAMS["tum1"]).tobytes()
    assert (ur_t[0, :3000].cpu().numpy().view(np.uint32) == rgbd_oracle(ks[0], ks[0], depth[0], 40.0)[0].view(np.uint32)).all()
