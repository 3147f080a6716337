"""Ad-hoc parity soak of the searches added last: random key frames / point sets / parameters through the single-problem entry
points against the oracle -- the Fuse / SearchBySim3 search (orbm_window_best_free, with and without the chi-square gate),
SearchForTriangulation (orbm_search_for_triangulation), UndistortKeyPoints / ComputeStereoFromRGBD (batch of one).
Usage (GPU box): python tools/soak_fuse.py [cases] [seed]"""
import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import torch
import orb_slam2_chinesenotes_b200 as ob
from fuse_lib import window_best_free_oracle
from matcher_lib import extract_frame, flip_bits
from oracle_lib import KP_DTYPE
from test_frame_steps import depth_image, rgbd_oracle, undistort_oracle
from triang_lib import search_for_triangulation, triang_scene

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
W, H = 1241, 376
K = np.float32([718.856, 718.856, 607.1928, 185.2157])
kps, desc, scale = extract_frame(W, H, 2000, 2)
kbig, dbig, _ = extract_frame(1920, 1080, 4000, 7)
kbig = kbig.copy(); kbig["x"] *= W / 1920.0; kbig["y"] *= H / 1080.0
inv_sigma2 = (np.float32(1.0) / (scale * scale)).astype(np.float32)
sigma2 = (scale * scale).astype(np.float32)
bad = 0
for c in range(cases):
    src_k, src_d = (kbig, dbig) if rng.random() < 0.3 else (kps, desc)
    n = int(rng.integers(1, len(src_k) + 1))
    sel = rng.permutation(len(src_k))[:n]
    k, d = src_k[sel].copy(), src_d[sel].copy()
    if n > 8 and rng.random() < 0.5:
        d[1::5] = d[0]                                                            # ties
    bounds = (float(rng.integers(0, 60)), float(W - rng.integers(0, 60)), float(rng.integers(0, 30)), float(H - rng.integers(0, 30)))
    nq = int(rng.integers(1, 6000))
    tgt = rng.integers(0, n, nq)
    uvr = np.zeros((nq, 3), np.float32)
    uvr[:, 0] = k["x"][tgt] + rng.normal(0, rng.choice([0.5, 2.0, 6.0]), nq)
    uvr[:, 1] = k["y"][tgt] + rng.normal(0, rng.choice([0.5, 2.0, 6.0]), nq)
    level = np.clip(k["octave"][tgt] + rng.integers(-1, 2, nq), 0, 7).astype(np.int32)
    th = float(rng.choice([1.0, 2.5, 3.0, 4.0, 7.5, 20.0]))
    uvr[:, 2] = (th * scale[level]).astype(np.float32)
    far = rng.random(nq) < 0.05
    uvr[far, 0] += rng.choice([-1, 1], int(far.sum())) * W
    q = dict(uvr=np.ascontiguousarray(uvr), level=level, ur=(uvr[:, 0] - 30 * rng.random(nq)).astype(np.float32),
             valid=(rng.random(nq) < 0.9).astype(np.uint8))
    qd = flip_bits(d[tgt], rng, int(rng.integers(1, 120)))
    ur = np.where(rng.random(n) < 0.5, k["x"] - 30 * rng.random(n), -1).astype(np.float32) if rng.random() < 0.6 else None
    s2 = inv_sigma2 if rng.random() < 0.5 else None
    acc = int(rng.choice([50, 100, 30]))
    want = window_best_free_oracle(k, d, ur, bounds, q, qd, s2, acc)
    got = ob.window_best_free(ob.FrameView(k, d, bounds, ur), q["uvr"], q["level"], qd, acc, ur=q["ur"], valid=q["valid"], inv_sigma2=s2)
    ok1 = got[0] == want[0] and (got[1] == want[1]).all() and (got[2] == want[2]).all()
    # triangulation on a sub-frame of the KITTI frame
    n1 = int(rng.integers(2, len(kps) + 1))
    s1 = rng.permutation(len(kps))[:n1]
    sc = triang_scene(kps[s1], desc[s1], W, H, int(rng.integers(1 << 30)), K, scale, n2=int(rng.integers(1, n1 + 1)), kmax=int(rng.integers(1, 60)))
    sc["epipole"] = np.float32([rng.random() * W, rng.random() * H])
    only_stereo, co, mono = bool(rng.integers(0, 2)), bool(rng.integers(0, 2)), rng.random() < 0.3
    wt = search_for_triangulation("oracle", sc, K, scale, sigma2, only_stereo, co, mono)
    st1 = np.zeros(len(sc["k1"]), bool) if mono else sc["ur1"] >= 0
    st2 = np.zeros(len(sc["k2"]), bool) if mono else sc["ur2"] >= 0
    v1 = ((sc["has1"] == 0) & (st1 | (not only_stereo))).astype(np.uint8)
    v2 = ((sc["has2"] == 0) & (st2 | (not only_stereo))).astype(np.uint8)
    gt = ob.search_for_triangulation(ob.FrameView(sc["k1"], sc["d1"], bounds, None if mono else sc["ur1"]), sc["fv1"], v1,
                                     ob.FrameView(sc["k2"], sc["d2"], bounds, None if mono else sc["ur2"]), sc["fv2"], v2,
                                     sc["F12"], sc["epipole"], scale, sigma2, co)
    ok2 = gt[0] == wt[0] and (gt[1] == wt[1]).all()
    # undistortion + RGB-D on the first frame's keypoints with a random camera
    Kc = [float(400 + 300 * rng.random()), float(400 + 300 * rng.random()), float(W / 2 + rng.normal(0, 20)), float(H / 2 + rng.normal(0, 20))]
    D = [float(rng.normal(0, 0.3)), float(rng.normal(0, 0.5)), float(rng.normal(0, 0.005)), float(rng.normal(0, 0.005)), float(rng.normal(0, 0.5))][:int(rng.choice([4, 5]))]
    kk = k.copy()
    kk["x"] = np.clip(kk["x"], 0, 639); kk["y"] = np.clip(kk["y"], 0, 375)
    wu = undistort_oracle(kk, Kc, D)
    dep = depth_image(int(rng.integers(1 << 20)))
    w_ur, w_dz = rgbd_oracle(kk, wu, dep, 40.0)
    d_k = torch.from_numpy(kk.view(np.uint8).reshape(1, n, 28).copy()).cuda()
    d_u = torch.zeros_like(d_k)
    d_n = torch.tensor([n], dtype=torch.int32, device="cuda")
    g_ur, g_dz = torch.zeros((1, n), device="cuda"), torch.zeros((1, n), device="cuda")
    ob.undistort_keypoints_batch(d_k, d_u, d_n, Kc, D)
    ob.stereo_from_rgbd_batch(d_k, d_u, d_n, torch.from_numpy(dep[None]).cuda(), 40.0, g_ur, g_dz)
    torch.cuda.synchronize()
    ok3 = d_u.cpu().numpy().tobytes() == wu.tobytes() and (g_ur.cpu().numpy()[0].view(np.uint32) == w_ur.view(np.uint32)).all() \
        and (g_dz.cpu().numpy()[0].view(np.uint32) == w_dz.view(np.uint32)).all()
    print(c, (n, nq, th, acc, s2 is not None, ur is not None), (n1, only_stereo, co, mono), "ok" if ok1 and ok2 and ok3 else f"MISMATCH {ok1} {ok2} {ok3}",
          got[0], gt[0], flush=True)
    bad += 0 if ok1 and ok2 and ok3 else 1
print("mismatches:", bad)
