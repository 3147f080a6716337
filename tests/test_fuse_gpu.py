"""The order-free searches on the device (orbm_window_best_free[_batch], k_window_best_free in csrc/orb_match_batch.cu) through
the C ABI against the oracle: ORBmatcher::Fuse (both overloads, src/ORBmatcher.cc:1364-1513, :1516-1633) and
ORBmatcher::SearchBySim3 (:836-1052).  The host projects (the oracle's exported queries = what host/ORBmatcher_b200.hpp
computes), the GPU searches every point at once, the host replays the reference's map bookkeeping in list order."""
import numpy as np
import pytest

import orb_slam2_chinesenotes_b200 as ob
from fuse_lib import (fuse_scene, replay_fuse, run_fuse, run_search_by_sim3, same_state, sim3_pair_scene, window_best_free_oracle)
from matcher_lib import extract_frame, perturbed_frame
from oracle_lib import KP_DTYPE
from test_fuse_oracle import BF, BOUNDS, FUSE_CASES, H, K, NF, SIM3_CASES, W

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def scene():
    kps, desc, scale = extract_frame(W, H, NF, 2)
    k2, d2, _ = perturbed_frame(kps, desc, W, H, 11, shift=4, kmax=40)
    inv_sigma2 = (np.float32(1.0) / (scale * scale)).astype(np.float32)
    return dict(kps=kps, desc=desc, scale=scale, inv_sigma2=inv_sigma2, k2=k2, d2=d2)


@pytest.mark.parametrize("th,seed,sim3,stereo", FUSE_CASES)
def test_fuse_equals_oracle(scene, th, seed, sim3, stereo):
    s = fuse_scene(scene["kps"], scene["desc"], W, H, seed, K, BF, sim3)
    nf, st, q = run_fuse("oracle", scene["kps"], scene["desc"], s, scene["scale"], scene["inv_sigma2"], BOUNDS, K, BF, th, sim3, stereo)
    ur = s["u_right"] if stereo else None
    qdesc = np.ascontiguousarray(s["mp_desc"][np.maximum(s["list"], 0)])
    s2 = None if sim3 else scene["inv_sigma2"]
    F = ob.FrameView(scene["kps"], scene["desc"], BOUNDS, ur)
    g_n, g_bi, g_bd = ob.window_best_free(F, q["uvr"], q["level"], qdesc, 50, ur=q["ur"], valid=q["valid"], inv_sigma2=s2)
    o_n, o_bi, o_bd = window_best_free_oracle(scene["kps"], scene["desc"], ur, BOUNDS, q, qdesc, s2, 50)
    assert g_n == o_n and o_n > 300
    assert (g_bi == o_bi).all() and (g_bd == o_bd).all()
    nf2, st2 = replay_fuse(s, q, g_bi, ur, sim3)
    assert nf2 == nf and same_state(st, st2)
    # the same search with the key frame resident on the device
    R = ob.ResidentFrame(scene["kps"], scene["desc"], BOUNDS, ur)
    r_n, r_bi, r_bd = ob.window_best_free(R, q["uvr"], q["level"], qdesc, 50, ur=q["ur"], valid=q["valid"], inv_sigma2=s2)
    assert r_n == o_n and (r_bi == o_bi).all() and (r_bd == o_bd).all()


@pytest.mark.parametrize("th,seed", SIM3_CASES)
def test_search_by_sim3_equals_oracle(scene, th, seed):
    k1, d1, k2, d2 = scene["kps"], scene["desc"], scene["k2"], scene["d2"]
    s = sim3_pair_scene(k1, d1, k2, d2, W, H, seed, K)
    nf, m12, (q1, q2) = run_search_by_sim3("oracle", k1, d1, k2, d2, s, scene["scale"], BOUNDS, K, th)
    # direction 1: KF1's points searched in KF2; direction 2: KF2's points searched in KF1 (TH_HIGH)
    qd1 = np.ascontiguousarray(s["mp_desc"][np.maximum(s["mp1"], 0)])
    qd2 = np.ascontiguousarray(s["mp_desc"][np.maximum(s["mp2"], 0)])
    _, b1, _ = ob.window_best_free(ob.FrameView(k2, d2, BOUNDS), q1["uvr"], q1["level"], qd1, 100, valid=q1["valid"])
    _, b2, _ = ob.window_best_free(ob.FrameView(k1, d1, BOUNDS), q2["uvr"], q2["level"], qd2, 100, valid=q2["valid"])
    out = s["m12"].copy()
    found = 0
    for i1 in range(len(k1)):                                   # src/ORBmatcher.cc:1029-1042
        i2 = b1[i1]
        if i2 >= 0 and b2[i2] == i1:
            out[i1] = s["mp2"][i2]
            found += 1
    assert found == nf and nf > 200 and (out == m12).all()


def _dev(a):
    import torch
    a = np.ascontiguousarray(a)
    if a.dtype == KP_DTYPE:
        a = a.view(np.uint8).reshape(a.shape + (28,))
    return torch.from_numpy(a).cuda()


def _pad(rows, stride, fill=0):
    out = np.full((len(rows), stride) + rows[0].shape[1:], fill, rows[0].dtype)
    for i, r in enumerate(rows):
        out[i, :len(r)] = r
    return out


def _random_queries(rng, kps, desc, nq, nlevels=8):
    n = len(kps)
    tgt = rng.integers(0, max(n, 1), nq)
    uvr = np.zeros((nq, 3), np.float32)
    if n:
        uvr[:, 0] = kps["x"][tgt] + rng.normal(0, 3, nq)
        uvr[:, 1] = kps["y"][tgt] + rng.normal(0, 3, nq)
        level = np.clip(kps["octave"][tgt] + rng.integers(-1, 2, nq), 0, nlevels - 1).astype(np.int32)
        qdesc = desc[tgt].copy()
    else:
        uvr[:, 0], uvr[:, 1] = rng.random(nq) * W, rng.random(nq) * H
        level = rng.integers(0, nlevels, nq).astype(np.int32)
        qdesc = rng.integers(0, 256, (nq, 32)).astype(np.uint8)
    uvr[:, 2] = np.where(rng.random(nq) < 0.1, 60.0, 3.0 + 12 * rng.random(nq)).astype(np.float32)
    far = rng.random(nq) < 0.05                                  # windows off the image / hanging over its border
    uvr[far, 0] += rng.choice([-1, 1], int(far.sum())) * (W * 0.9)
    flips = rng.integers(0, 256, (nq, 32)).astype(np.uint8) & rng.integers(0, 256, (nq, 32)).astype(np.uint8) & rng.integers(0, 256, (nq, 32)).astype(np.uint8)
    qdesc ^= flips
    ur = (uvr[:, 0] - 30 * rng.random(nq)).astype(np.float32)
    valid = (rng.random(nq) < 0.9).astype(np.uint8)
    return dict(uvr=uvr, level=level, ur=ur, valid=valid), np.ascontiguousarray(qdesc)


@pytest.mark.parametrize("chi2,bounds", [(False, BOUNDS), (True, BOUNDS), (True, (40.0, W - 60.0, 25.0, H - 30.0))])
def test_batch_of_problems_equals_oracle(scene, chi2, bounds):
    """Many (key frame, point set) problems in one launch, device resident: frame sizes 0 .. 4000 (above ~2700 keypoints the
    sorted descriptors live in the global workspace), query counts 0 .. 5000, windows over the border, ties."""
    import torch
    rng = np.random.default_rng(300)
    kbig, dbig, _ = extract_frame(1920, 1080, 4000, 7)
    kbig = kbig.copy()
    kbig["x"] *= W / 1920.0
    kbig["y"] *= H / 1080.0                                       # 4000 keypoints crowded into the KITTI bounds
    k2, d2 = scene["k2"], scene["d2"]
    problems = []
    for i, (src, n, nq) in enumerate([(0, len(k2), 2000), (0, 1500, 2500), (1, len(kbig), 5000), (0, 0, 50), (0, 300, 0), (0, 64, 3000),
                                      (1, 3000, 1000)]):
        kk, dd = (kbig, dbig) if src else (k2, d2)
        sel = rng.permutation(len(kk))[:n]
        kp, de = kk[sel].copy(), dd[sel].copy()
        if n > 10:
            de[1::7] = de[0]                                     # equal descriptors: ties go to the first candidate in grid order
        q, qd = _random_queries(rng, kp, de, max(nq, 1))
        q = {k: v[:nq] for k, v in q.items()}
        ur = np.where(rng.random(n) < 0.5, kp["x"] - 30 * rng.random(n), -1).astype(np.float32)
        problems.append(dict(kps=kp, desc=de, ur=ur, q=q, qd=qd[:nq], nq=nq))
    cap, nqs = 4100, 5000
    kps = np.zeros((len(problems), cap), KP_DTYPE)
    for i, pr in enumerate(problems):
        kps[i, :len(pr["kps"])] = pr["kps"]
    d_kps, d_desc = _dev(kps), _dev(_pad([pr["desc"] for pr in problems], cap))
    d_n = _dev(np.int32([len(pr["kps"]) for pr in problems]))
    d_ur = _dev(_pad([pr["ur"] for pr in problems], cap, -1))
    F = ob.frames_batch(d_kps, d_desc, d_n, bounds, d_ur)
    dq = {k: _dev(_pad([pr["q"][k] for pr in problems], nqs)) for k in ("uvr", "level", "ur", "valid")}
    dq["desc"] = _dev(_pad([pr["qd"] for pr in problems], nqs))
    d_nq = _dev(np.int32([pr["nq"] for pr in problems]))
    bi = torch.full((len(problems), nqs), -7, dtype=torch.int32, device="cuda")
    bd = torch.full((len(problems), nqs), -7, dtype=torch.int32, device="cuda")
    nf = torch.full((len(problems),), -7, dtype=torch.int32, device="cuda")
    s2 = scene["inv_sigma2"] if chi2 else None
    ob.window_best_free_batch(F, dq, d_nq, nqs, bi, bd, nf, 50, inv_sigma2=s2)
    torch.cuda.synchronize()
    bi, bd, nf = bi.cpu().numpy(), bd.cpu().numpy(), nf.cpu().numpy()
    total = 0
    for i, pr in enumerate(problems):
        o_n, o_bi, o_bd = window_best_free_oracle(pr["kps"], pr["desc"], pr["ur"], bounds, pr["q"], pr["qd"], s2, 50)
        nq = pr["nq"]
        assert nf[i] == o_n, i
        assert (bi[i, :nq] == o_bi).all() and (bd[i, :nq] == o_bd).all(), i
        assert (bi[i, nq:] == -7).all()                           # nothing written behind a problem's queries
        total += o_n
    assert total > 1500


def test_degenerate_inputs(scene):
    F = ob.FrameView(scene["kps"], scene["desc"], BOUNDS)
    E = ob.FrameView(scene["kps"][:0], scene["desc"][:0], BOUNDS)
    rng = np.random.default_rng(5)
    q, qd = _random_queries(rng, scene["kps"], scene["desc"], 40)
    n, bi, bd = ob.window_best_free(E, q["uvr"], q["level"], qd, 50)              # a key frame without keypoints
    assert n == 0 and (bi == -1).all() and (bd == 256).all()
    n, bi, bd = ob.window_best_free(F, q["uvr"][:0], q["level"][:0], qd[:0], 50)  # no points
    assert n == 0 and len(bi) == 0
    n, bi, bd = ob.window_best_free(F, q["uvr"], q["level"], qd, 50, valid=np.zeros(40, np.uint8))
    assert n == 0 and (bi == -1).all() and (bd == 256).all()
    n, bi, bd = ob.window_best_free(F, q["uvr"], q["level"], qd, -1, valid=q["valid"])   # nothing passes the threshold; distances still reported
    o_n, o_bi, o_bd = window_best_free_oracle(scene["kps"], scene["desc"], None, BOUNDS, q, qd, None, -1)
    assert n == 0 and (bi == -1).all() and (bd == o_bd).all() and (bd < 256).any()


@pytest.mark.parametrize("sim3", [False, True])
def test_projection_then_search_stays_on_the_device(scene, sim3):
    """orbm_fuse_project_batch -> orbm_window_best_free_batch for several (key frame, point set) problems without touching the host
    in between: the projected queries equal the CPU restatement's bit for bit (real PredictScale arithmetic, both the Fuse gates
    and one direction of SearchBySim3), and so does the search on them."""
    import torch
    from fuse_lib import (fuse_list_arrays, fuse_pose24, fuse_project_oracle, fuse_scene, raw_distances, sim3_pair_scene, sim3_poses)
    kps, desc, scale = scene["kps"], scene["desc"], scene["scale"]
    th = 7.5 if sim3 else 3.0
    probs = []
    for seed in (91, 92, 93):
        if sim3:
            s = sim3_pair_scene(kps, desc, scene["k2"], scene["d2"], W, H, seed, K)
            rng = np.random.default_rng(seed)
            max_raw = (3.0 + 60 * rng.random(s["npts"])).astype(np.float32)
            min_raw = (max_raw / np.float32(1.2 ** 7)).astype(np.float32)
            pose = sim3_poses(s)[0]                               # KF1's points searched in KF2
            g = np.maximum(s["mp1"], 0)
            A = dict(xyz=np.ascontiguousarray(s["xyz"][g]), normal=None, max_d=np.ascontiguousarray(max_raw[g]), min_d=np.ascontiguousarray(min_raw[g]),
                     skip=((s["mp1"] < 0) | (s["m12"] >= 0) | (s["bad"][g] != 0)).astype(np.uint8), desc=np.ascontiguousarray(s["mp_desc"][g]))
            probs.append(dict(pose=pose, A=A, tk=scene["k2"], td=scene["d2"], ur=None))
        else:
            s = fuse_scene(kps, desc, W, H, seed, K, BF, False)
            s, max_raw, min_raw = raw_distances(s, np.random.default_rng(seed))
            probs.append(dict(pose=fuse_pose24(s, False), A=fuse_list_arrays(s, max_raw, min_raw, False), tk=kps, td=desc, ur=s["u_right"]))
    P = len(probs)
    nqs = max(len(pr["A"]["max_d"]) for pr in probs) + 5
    cap = max(len(pr["tk"]) for pr in probs) + 3
    pts = dict(xyz=_dev(_pad([pr["A"]["xyz"] for pr in probs], nqs)), max_d=_dev(_pad([pr["A"]["max_d"] for pr in probs], nqs, 1)),
               min_d=_dev(_pad([pr["A"]["min_d"] for pr in probs], nqs, 1)), skip=_dev(_pad([pr["A"]["skip"] for pr in probs], nqs, 1)))
    if not sim3:
        pts["normal"] = _dev(_pad([pr["A"]["normal"] for pr in probs], nqs))
    d_pose = _dev(np.stack([pr["pose"] for pr in probs]))
    d_nq = _dev(np.int32([len(pr["A"]["max_d"]) for pr in probs]))
    out = dict(uvr=torch.full((P, nqs, 3), 7.0, device="cuda"), level=torch.full((P, nqs), 7, dtype=torch.int32, device="cuda"),
               ur=torch.full((P, nqs), 7.0, device="cuda"), valid=torch.full((P, nqs), 7, dtype=torch.uint8, device="cuda"))
    bf = 0.0 if sim3 else BF
    ob.fuse_project_batch(d_pose, K, bf, BOUNDS, 1.2, scale, th, d_nq, nqs, pts, out, sim3=sim3)
    tk = np.zeros((P, cap), KP_DTYPE)
    for i, pr in enumerate(probs):
        tk[i, :len(pr["tk"])] = pr["tk"]
    d_k, d_d = _dev(tk), _dev(_pad([pr["td"] for pr in probs], cap))
    d_n = _dev(np.int32([len(pr["tk"]) for pr in probs]))
    d_ur = None if sim3 else _dev(_pad([pr["ur"] for pr in probs], cap, -1))
    F = ob.frames_batch(d_k, d_d, d_n, BOUNDS, d_ur)
    q = dict(uvr=out["uvr"], level=out["level"], ur=out["ur"], valid=out["valid"], desc=_dev(_pad([pr["A"]["desc"] for pr in probs], nqs)))
    bi = torch.zeros((P, nqs), dtype=torch.int32, device="cuda")
    bd = torch.zeros((P, nqs), dtype=torch.int32, device="cuda")
    nf = torch.zeros(P, dtype=torch.int32, device="cuda")
    ob.window_best_free_batch(F, q, d_nq, nqs, bi, bd, nf, 100 if sim3 else 50, inv_sigma2=None if sim3 else scene["inv_sigma2"])
    torch.cuda.synchronize()
    o = {k: v.cpu().numpy() for k, v in out.items()}
    bi, bd, nf = bi.cpu().numpy(), bd.cpu().numpy(), nf.cpu().numpy()
    for i, pr in enumerate(probs):
        A, n = pr["A"], len(pr["A"]["max_d"])
        w = fuse_project_oracle(sim3, pr["pose"], K, bf, BOUNDS, 1.2, scale, th, A["xyz"], A["normal"], A["max_d"], A["min_d"], A["skip"])
        m = w["valid"].astype(bool)
        assert (o["valid"][i, :n] == w["valid"]).all() and m.sum() > 500
        assert (o["uvr"][i, :n][m].view(np.uint32) == w["uvr"][m].view(np.uint32)).all() and (o["level"][i, :n][m] == w["level"][m]).all()
        assert (o["ur"][i, :n][m].view(np.uint32) == w["ur"][m].view(np.uint32)).all()
        assert (o["valid"][i, n:] == 0).all()                       # entries behind a problem's points are written as invalid
        want = window_best_free_oracle(pr["tk"], pr["td"], pr["ur"], BOUNDS, w, A["desc"], None if sim3 else scene["inv_sigma2"], 100 if sim3 else 50)
        assert nf[i] == want[0] and want[0] > 100 and (bi[i, :n] == want[1]).all() and (bd[i, :n] == want[2]).all()
