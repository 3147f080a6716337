"""Batched, device-resident window matchers (orbm_*_batch, csrc/orb_match_batch.cu) through the C ABI against
the oracle, problem by problem: match indices and counts identical, including the order-dependent "keypoint
already taken" bookkeeping that the kernel resolves by fixpoint rounds instead of walking the queries in order."""
import ctypes as C

import numpy as np
import pytest

import orb_slam2_chinesenotes_b200 as ob
from matcher_lib import Matcher, extract_frame, perturbed_frame, projected_queries
from oracle_lib import KP_DTYPE, oracle

pytestmark = pytest.mark.gpu
W, H, NF = 1241, 376, 2000


@pytest.fixture(scope="module")
def scene():
    kps, desc, scale = extract_frame(W, H, NF, 2)
    k2, d2, _ = perturbed_frame(kps, desc, W, H, 11)
    return dict(kps=kps, desc=desc, scale=scale, k2=k2, d2=d2)


def _dev(a):
    import torch
    a = np.ascontiguousarray(a)
    if a.dtype == KP_DTYPE:
        a = a.view(np.uint8).reshape(a.shape + (28,))
    return torch.from_numpy(a).cuda()


def _pad(rows, stride, fill=0):
    """list of [n_i, ...] arrays -> [P, stride, ...]"""
    out = np.full((len(rows), stride) + rows[0].shape[1:], fill, rows[0].dtype)
    for i, r in enumerate(rows):
        out[i, :len(r)] = r
    return out


def _frames(problems, bounds, cap, with_ur):
    import torch
    kps = np.zeros((len(problems), cap), KP_DTYPE)
    for i, p in enumerate(problems):
        kps[i, :len(p["kps"])] = p["kps"]
    d_kps = _dev(kps)
    d_desc = _dev(_pad([p["desc"] for p in problems], cap))
    d_n = _dev(np.int32([len(p["kps"]) for p in problems]))
    d_ur = _dev(_pad([p["ur"] for p in problems], cap, -1)) if with_ur else None
    keep = (d_kps, d_desc, d_n, d_ur)
    return ob.frames_batch(d_kps, d_desc, d_n, bounds, d_ur), keep


@pytest.mark.parametrize("th,nnratio,with_ur,bounds", [(3.0, 0.8, False, (0.0, float(W), 0.0, float(H))),
                                                      (5.0, 0.9, True, (0.0, float(W), 0.0, float(H))),
                                                      (1.0, 0.9, True, (40.0, W - 60.0, 25.0, H - 30.0))])
def test_points_batch_equals_oracle(scene, th, nnratio, with_ur, bounds):
    import torch
    O = Matcher("oracle")
    rng = np.random.default_rng(100)
    k2, d2 = scene["k2"], scene["d2"]
    problems = []
    for i, (n, nq) in enumerate([(len(k2), 2000), (1500, 2500), (len(k2), 700), (0, 50), (300, 0), (1999, 1999), (64, 3000)]):
        sel = rng.permutation(len(k2))[:n]
        kp, de = k2[sel].copy(), d2[sel].copy()
        q = projected_queries(kp, de, max(nq, 1), 200 + i) if n else projected_queries(k2, d2, max(nq, 1), 200 + i)
        q = {k: v[:nq] for k, v in q.items()}
        ur = np.where(rng.random(n) < 0.5, kp["x"] - 20 * rng.random(n), -1).astype(np.float32)
        init = np.where(rng.random(n) < 0.05, rng.integers(0, max(nq, 1), n), -1).astype(np.int32) if nq else np.full(n, -1, np.int32)
        problems.append(dict(kps=kp, desc=de, q=q, ur=ur, init=init, nq=nq))
    cap, nqs = 2100, 3000
    F, keep = _frames(problems, bounds, cap, with_ur)
    dq = {k: _dev(_pad([p["q"][k] for p in problems], nqs)) for k in ("proj", "level", "view_cos", "in_view", "bad", "obs", "desc")}
    d_nq = _dev(np.int32([p["nq"] for p in problems]))
    for use_init in (False, True):
        d_init = _dev(_pad([p["init"] for p in problems], cap, -1)) if use_init else None
        d_assign = torch.full((len(problems), cap), -7, dtype=torch.int32, device="cuda")
        d_nm = torch.zeros(len(problems), dtype=torch.int32, device="cuda")
        d_rounds = torch.zeros(len(problems), dtype=torch.int32, device="cuda")
        ob.search_by_projection_points_batch(F, scene["scale"], dq, d_nq, nqs, d_assign, d_nm, th, nnratio, d_init, d_rounds)
        torch.cuda.synchronize()
        assign, nm, rounds = d_assign.cpu().numpy(), d_nm.cpu().numpy(), d_rounds.cpu().numpy()
        total = 0
        for i, p in enumerate(problems):
            n = len(p["kps"])
            want = O.search_by_projection_points(p["kps"], p["desc"], p["ur"] if with_ur else None, scene["scale"], bounds, p["q"], th, nnratio,
                                                 p["init"] if use_init else None) if n and p["nq"] else (0, p["init"] if use_init else np.full(n, -1, np.int32))
            assert nm[i] == want[0], (i, nm[i], want[0])
            assert (assign[i, :n] == want[1]).all(), i
            assert (assign[i, n:] == -7).all()
            total += want[0]
        assert total > 800 and rounds.max() >= 2 and rounds.max() < 64, rounds   # conflicts were exercised and resolved in a few rounds


def test_best_batch_equals_oracle(scene):
    import torch
    rng = np.random.default_rng(12)
    k2, d2 = scene["k2"], scene["d2"]
    bounds = (0.0, float(W), 0.0, float(H))
    O = oracle()
    vp, ci, cf = C.c_void_p, C.c_int, C.c_float
    O.orbo_window_search_best.argtypes = [ci, vp, vp, vp] + [cf] * 4 + [ci] + [vp] * 11 + [ci, ci]
    problems = []
    for n, nq in [(len(k2), 1500), (1200, 2200), (len(k2), 1)]:
        sel = rng.permutation(len(k2))[:n]
        kp, de = k2[sel].copy(), d2[sel].copy()
        tgt = rng.integers(0, n, nq)
        tgt[::4] = rng.integers(0, 30, len(tgt[::4]))       # many queries compete for the same keypoints
        q = dict(uvr=np.stack([kp["x"][tgt] + rng.normal(0, 4, nq), kp["y"][tgt] + rng.normal(0, 4, nq), rng.choice([3.0, 8.0, 20.0, 60.0], nq)], 1).astype(np.float32),
                 min_level=rng.integers(-1, 4, nq).astype(np.int32))
        q["max_level"] = np.where(rng.random(nq) < 0.3, -1, q["min_level"] + rng.integers(0, 3, nq)).astype(np.int32)
        q["ur"] = (q["uvr"][:, 0] - 12 + rng.normal(0, 6, nq)).astype(np.float32)
        q["er_max"] = rng.choice([2.0, 6.0, 15.0], nq).astype(np.float32)
        q["valid"] = (rng.random(nq) < 0.9).astype(np.uint8)
        qd = de[tgt].copy(); qd[:, :4] ^= rng.integers(0, 256, (nq, 4), dtype=np.uint8)
        q["desc"] = qd
        q["q_angle"] = ((kp["angle"][tgt] + 40 + rng.normal(0, 8, nq)) % 360).astype(np.float32)
        q["q_obs"] = (rng.random(nq) < 0.8).astype(np.int32)
        problems.append(dict(kps=kp, desc=de, q=q, nq=nq, ur=np.where(rng.random(n) < 0.5, kp["x"] - 25 * rng.random(n), -1).astype(np.float32),
                             init=np.where(rng.random(n) < 0.1, rng.integers(0, 2, n), -1).astype(np.int32)))
    cap, nqs = 2048, 2200
    F, keep = _frames(problems, bounds, cap, True)
    dq = {k: _dev(_pad([p["q"][k] for p in problems], nqs)) for k in problems[0]["q"]}
    d_nq = _dev(np.int32([p["nq"] for p in problems]))
    d_init = _dev(_pad([p["init"] for p in problems], cap, -1))
    for th_accept, check_ori in ((100, True), (50, False), (64, True)):
        d_assign = torch.full((len(problems), cap), -7, dtype=torch.int32, device="cuda")
        d_nm = torch.zeros(len(problems), dtype=torch.int32, device="cuda")
        ob.window_search_best_batch(F, dq, d_nq, nqs, d_assign, d_nm, th_accept, check_ori, d_init)
        torch.cuda.synchronize()
        assign, nm = d_assign.cpu().numpy(), d_nm.cpu().numpy()
        for i, pr in enumerate(problems):
            n, nq, q = len(pr["kps"]), pr["nq"], pr["q"]
            ref_assign = np.zeros(n, np.int32)
            p = lambda a: a.ctypes.data
            want = O.orbo_window_search_best(n, p(pr["kps"]), p(pr["desc"]), p(pr["ur"]), *bounds, nq, p(q["uvr"]), p(q["min_level"]), p(q["max_level"]),
                                             p(q["ur"]), p(q["er_max"]), p(q["valid"]), p(q["desc"]), p(q["q_angle"]), p(q["q_obs"]), p(pr["init"]),
                                             p(ref_assign), th_accept, int(check_ori))
            assert nm[i] == want, (i, nm[i], want)
            assert (assign[i, :n] == ref_assign).all(), i
        assert nm[:2].min() > 100


def test_batch_equals_single_problem_entry_point(scene):
    """The batch kernel and the single-problem entry point (host arrays, three kernels) agree."""
    import torch
    bounds = (0.0, float(W), 0.0, float(H))
    k2, d2 = scene["k2"], scene["d2"]
    q = projected_queries(k2, d2, 2000, 5)
    F1 = ob.FrameView(k2, d2, bounds)
    want = ob.ORBmatcher(0.9, True).SearchByProjection(F1, scene["scale"], q, 3.0)
    pr = [dict(kps=k2, desc=d2)]
    F, keep = _frames(pr, bounds, len(k2), False)
    dq = {k: _dev(q[k][None]) for k in q}
    d_assign = torch.zeros((1, len(k2)), dtype=torch.int32, device="cuda")
    d_nm = torch.zeros(1, dtype=torch.int32, device="cuda")
    ob.search_by_projection_points_batch(F, scene["scale"], dq, _dev(np.int32([2000])), 2000, d_assign, d_nm, 3.0, 0.9)
    torch.cuda.synchronize()
    assert int(d_nm[0]) == want[0] and (d_assign[0].cpu().numpy() == want[1]).all()


def test_batch_rejects_oversized_problems(scene):
    import torch
    bounds = (0.0, float(W), 0.0, float(H))
    pr = [dict(kps=scene["k2"][:100], desc=scene["d2"][:100])]
    F, keep = _frames(pr, bounds, 128, False)
    # a bound the frames respect sizes the shared memory; one they exceed marks the problem
    for max_n, ok in ((100, True), (64, False)):
        Fb = ob.frames_batch(keep[0], keep[1], keep[2], bounds, None, max_n)
        q0 = projected_queries(scene["k2"][:100], scene["d2"][:100], 10, 1)
        dq0 = {k: _dev(q0[k][None]) for k in q0}
        a0 = torch.zeros((1, 128), dtype=torch.int32, device="cuda")
        n0 = torch.zeros(1, dtype=torch.int32, device="cuda")
        ob.search_by_projection_points_batch(Fb, scene["scale"], dq0, _dev(np.int32([10])), 10, a0, n0, 3.0, 0.9)
        torch.cuda.synchronize()
        assert (int(n0[0]) >= 0) == ok
    keep[2].fill_(500)                                   # n > kp_stride
    q = projected_queries(scene["k2"][:100], scene["d2"][:100], 10, 1)
    dq = {k: _dev(q[k][None]) for k in q}
    d_assign = torch.zeros((1, 128), dtype=torch.int32, device="cuda")
    d_nm = torch.zeros(1, dtype=torch.int32, device="cuda")
    ob.search_by_projection_points_batch(F, scene["scale"], dq, _dev(np.int32([10])), 10, d_assign, d_nm, 3.0, 0.9)
    torch.cuda.synchronize()
    assert int(d_nm[0]) == -1
    with pytest.raises(ob.OrbError):
        ob.search_by_projection_points_batch(F, scene["scale"], {k: v.cpu() for k, v in dq.items()}, _dev(np.int32([10])), 10, d_assign, d_nm, 3.0, 0.9)


@pytest.mark.parametrize("th,mono,check_ori", [(7.0, False, True), (15.0, True, True), (7.0, False, False)])
def test_frame_to_frame_batch_equals_oracle(scene, th, mono, check_ori):
    """SearchByProjection(cur, last) with the projection of the last frame's map points on the device too."""
    import torch
    from matcher_lib import two_view_scene
    O = Matcher("oracle")
    K = np.float32([718.856, 718.856, 607.1928, 185.2157])
    bounds = (0.0, float(W), 0.0, float(H))
    rng = np.random.default_rng(8)
    probs = []
    for seed, n_use in ((21, None), (22, 1500), (23, None)):
        kps, desc = (scene["kps"], scene["desc"]) if n_use is None else (scene["kps"][:n_use], scene["desc"][:n_use])
        cur, last, Tc, Tl = two_view_scene(kps, desc, W, H, seed, K)
        if seed == 23:                                       # backward motion: the level range flips (:219-224)
            Tc = Tc.copy(); Tc[2, 3] = 0.9
        if mono:
            cur = dict(cur); cur["u_right"] = None
        init_obs = np.where(rng.random(len(cur["kps"])) < 0.05, rng.integers(0, 2, len(cur["kps"])), -1).astype(np.int32)
        probs.append(dict(cur=cur, last=last, Tc=Tc, Tl=Tl, init=init_obs))
    cap = 2100
    fr = [dict(kps=p["cur"]["kps"], desc=p["cur"]["desc"], ur=p["cur"].get("u_right")) for p in probs]
    F, keep = _frames(fr, bounds, cap, not mono)
    d_last = dict(kps=_dev(np.stack([np.concatenate([p["last"]["kps"], np.zeros(cap - len(p["last"]["kps"]), KP_DTYPE)]) for p in probs])),
                  has_mp=_dev(_pad([p["last"]["has_mp"] for p in probs], cap)), outlier=_dev(_pad([p["last"]["outlier"] for p in probs], cap)),
                  xyz=_dev(_pad([p["last"]["xyz"] for p in probs], cap)), mp_desc=_dev(_pad([p["last"]["mp_desc"] for p in probs], cap)),
                  mp_obs=_dev(_pad([p["last"]["mp_obs"] for p in probs], cap)))
    d_Tc = _dev(np.stack([p["Tc"].reshape(16) for p in probs])); d_Tl = _dev(np.stack([p["Tl"].reshape(16) for p in probs]))
    d_nl = _dev(np.int32([len(p["last"]["kps"]) for p in probs]))
    for use_init in (False, True):
        d_init = _dev(_pad([p["init"] for p in probs], cap, -1)) if use_init else None
        d_assign = torch.full((len(probs), cap), -7, dtype=torch.int32, device="cuda")
        d_nm = torch.zeros(len(probs), dtype=torch.int32, device="cuda")
        ob.search_by_projection_frame_batch(F, d_Tc, d_Tl, K, 386.1448, scene["scale"], d_nl, cap, d_last, d_assign, d_nm, th, mono, check_ori, d_init)
        torch.cuda.synchronize()
        assign, nm = d_assign.cpu().numpy(), d_nm.cpu().numpy()
        for i, p in enumerate(probs):
            want = O.search_by_projection_frame(p["cur"], p["last"], p["Tc"], p["Tl"], K, 386.1448, scene["scale"], bounds, th, mono, 0.9, check_ori,
                                                p["init"] if use_init else None)
            n = len(p["cur"]["kps"])
            assert nm[i] == want[0] and want[0] > 100, (i, nm[i], want[0])
            assert (assign[i, :n] == want[1]).all()


def _adversarial_problem(kind, rng):
    """(kps, desc, queries) built to stress the order resolution and the grid."""
    nlevels = 8
    if kind == "pile":            # every keypoint and every query in one grid cell, a handful of distinct descriptors:
        n, nq = 600, 900          # long chains of "taken by an earlier query" -> many fixpoint rounds
        kp = np.zeros(n, KP_DTYPE)
        kp["x"] = 300 + rng.integers(0, 6, n); kp["y"] = 150 + rng.integers(0, 4, n); kp["octave"] = rng.integers(0, 2, n)
        base = rng.integers(0, 256, (5, 32), dtype=np.uint8)
        de = base[rng.integers(0, 5, n)].copy(); de[:, 0] ^= rng.integers(0, 4, n).astype(np.uint8)
        proj = np.stack([300 + rng.random(nq) * 6, 150 + rng.random(nq) * 4, 280 + rng.random(nq) * 6], 1).astype(np.float32)
        q = dict(proj=proj, level=rng.integers(0, 2, nq).astype(np.int32), view_cos=np.full(nq, 0.9, np.float32),
                 in_view=np.ones(nq, np.uint8), bad=np.zeros(nq, np.uint8), obs=(rng.random(nq) < 0.9).astype(np.int32),
                 desc=base[rng.integers(0, 5, nq)].copy())
    elif kind == "big":           # the largest frame the kernels take
        n, nq = 8192, 8192
        kp = np.zeros(n, KP_DTYPE)
        kp["x"] = rng.random(n) * W; kp["y"] = rng.random(n) * H; kp["octave"] = rng.integers(0, nlevels, n)
        de = rng.integers(0, 256, (n, 32), dtype=np.uint8)
        q = projected_queries(kp, de, nq, 3)
    else:                         # "outside": a third of the keypoints lie outside the image bounds (not in the grid)
        n, nq = 1500, 1500
        kp = np.zeros(n, KP_DTYPE)
        kp["x"] = rng.random(n) * (W + 600) - 300; kp["y"] = rng.random(n) * (H + 300) - 150; kp["octave"] = rng.integers(0, nlevels, n)
        de = rng.integers(0, 256, (n, 32), dtype=np.uint8)
        q = projected_queries(kp, de, nq, 4)
    return kp, de, q


@pytest.mark.parametrize("kind", ["pile", "big", "outside"])
def test_adversarial_problems_equal_oracle(scene, kind):
    import torch
    rng = np.random.default_rng(77)
    bounds = (0.0, float(W), 0.0, float(H))
    kp, de, q = _adversarial_problem(kind, rng)
    n, nq = len(kp), len(q["level"])
    want = Matcher("oracle").search_by_projection_points(kp, de, None, scene["scale"], bounds, q, 3.0, 0.9, None)
    F, keep = _frames([dict(kps=kp, desc=de)], bounds, n, False)
    dq = {k: _dev(q[k][None]) for k in q}
    d_assign = torch.zeros((1, n), dtype=torch.int32, device="cuda")
    d_nm = torch.zeros(1, dtype=torch.int32, device="cuda")
    d_rounds = torch.zeros(1, dtype=torch.int32, device="cuda")
    ob.search_by_projection_points_batch(F, scene["scale"], dq, _dev(np.int32([nq])), nq, d_assign, d_nm, 3.0, 0.9, None, d_rounds)
    torch.cuda.synchronize()
    assert int(d_nm[0]) == want[0] and (d_assign[0].cpu().numpy() == want[1]).all(), (kind, int(d_nm[0]), want[0], int(d_rounds[0]))
    if kind == "pile":
        assert want[0] > 20 and int(d_rounds[0]) > 8          # the chain really was long
    # the best-only mode on the same windows
    uvr = np.ascontiguousarray(np.concatenate([q["proj"][:, :2], np.full((nq, 1), 9.0, np.float32)], 1))
    qb = dict(uvr=uvr, min_level=np.maximum(q["level"] - 1, -1).astype(np.int32), max_level=(q["level"] + 1).astype(np.int32),
              desc=q["desc"], q_angle=(rng.random(nq) * 360).astype(np.float32), q_obs=q["obs"])
    O = oracle()
    vp, ci, cf = C.c_void_p, C.c_int, C.c_float
    O.orbo_window_search_best.argtypes = [ci, vp, vp, vp] + [cf] * 4 + [ci] + [vp] * 11 + [ci, ci]
    ref_assign = np.zeros(n, np.int32)
    p = lambda a: None if a is None else a.ctypes.data
    wantb = O.orbo_window_search_best(n, p(kp), p(de), None, *bounds, nq, p(qb["uvr"]), p(qb["min_level"]), p(qb["max_level"]), None, None, None,
                                      p(qb["desc"]), p(qb["q_angle"]), p(qb["q_obs"]), None, p(ref_assign), 100, 1)
    dqb = {k: _dev(v[None]) for k, v in qb.items()}
    ob.window_search_best_batch(F, dqb, _dev(np.int32([nq])), nq, d_assign, d_nm, 100, True, None, d_rounds)
    torch.cuda.synchronize()
    assert int(d_nm[0]) == wantb and (d_assign[0].cpu().numpy() == ref_assign).all(), (kind, int(d_nm[0]), wantb)


def test_search_for_initialization_batch_equals_oracle(scene):
    """Monocular initialisation as a batch: ordinary frame pairs (parallel rounds), a pile of identical keypoints (more
    claimants per keypoint than the rounds keep: in-order walk inside the kernel), empty sides."""
    import torch
    O = Matcher("oracle")
    bounds = (0.0, float(W), 0.0, float(H))
    rng = np.random.default_rng(31)
    k1, d1, k2, d2 = scene["kps"], scene["desc"], scene["k2"], scene["d2"]
    pile1 = np.zeros(300, KP_DTYPE); pile1["x"] = 200 + rng.integers(0, 5, 300); pile1["y"] = 100 + rng.integers(0, 5, 300)
    pile2 = np.zeros(120, KP_DTYPE); pile2["x"] = 200 + rng.integers(0, 5, 120); pile2["y"] = 100 + rng.integers(0, 5, 120)
    pile1["angle"] = rng.random(300) * 360; pile2["angle"] = rng.random(120) * 360
    pd2 = rng.integers(0, 256, (120, 32), dtype=np.uint8)                     # distinct targets ...
    pd1 = pd2[rng.integers(0, 40, 300)].copy()                                # ... 300 queries aimed at 40 of them
    pd1[:, 0] ^= rng.integers(0, 256, 300).astype(np.uint8); pd1[:, 1] ^= rng.integers(0, 16, 300).astype(np.uint8)
    probs = [dict(k1=k1, d1=d1, k2=k2, d2=d2, prev=np.stack([k1["x"], k1["y"]], 1)),
             dict(k1=k2[:1500], d1=d2[:1500], k2=k1, d2=d1, prev=np.stack([k2["x"][:1500] + 10, k2["y"][:1500] - 5], 1).astype(np.float32)),
             dict(k1=pile1, d1=pd1, k2=pile2, d2=pd2, prev=np.stack([pile1["x"], pile1["y"]], 1)),
             dict(k1=k1[:0], d1=d1[:0], k2=k2, d2=d2, prev=np.zeros((0, 2), np.float32)),
             dict(k1=k1[:200], d1=d1[:200], k2=k2[:0], d2=d2[:0], prev=np.stack([k1["x"][:200], k1["y"][:200]], 1))]
    cap = 2100
    FA, keepA = _frames([dict(kps=p["k1"], desc=p["d1"]) for p in probs], bounds, cap, False)
    FB, keepB = _frames([dict(kps=p["k2"], desc=p["d2"]) for p in probs], bounds, cap, False)
    for nnratio, check_ori, window in ((0.9, True, 100), (0.6, True, 50), (0.9, False, 100)):
        d_prev = _dev(_pad([np.ascontiguousarray(p["prev"], np.float32) for p in probs], cap))
        d_m12 = torch.full((len(probs), cap), -7, dtype=torch.int32, device="cuda")
        d_nm = torch.zeros(len(probs), dtype=torch.int32, device="cuda")
        d_rounds = torch.zeros(len(probs), dtype=torch.int32, device="cuda")
        ob.search_for_initialization_batch(FA, FB, d_prev, d_m12, d_nm, window, nnratio, check_ori, d_rounds)
        torch.cuda.synchronize()
        m12, nm, prev, rounds = d_m12.cpu().numpy(), d_nm.cpu().numpy(), d_prev.cpu().numpy(), d_rounds.cpu().numpy()
        for i, p in enumerate(probs):
            n1 = len(p["k1"])
            if n1 and len(p["k2"]):
                want = O.search_for_initialization(p["k1"], p["d1"], p["k2"], p["d2"], scene["scale"], bounds, p["prev"], window, nnratio, check_ori)
            else:
                want = (0, np.full(n1, -1, np.int32), np.ascontiguousarray(p["prev"], np.float32))
            assert nm[i] == want[0], (i, nm[i], want[0], rounds[i])
            assert (m12[i, :n1] == want[1]).all() and (prev[i, :n1] == want[2]).all(), i
        assert nm[0] > 50 and rounds[0] > 0 and nm[2] > 5 and rounds[2] < 0, (nm, rounds)   # pair 0 by parallel rounds, the pile in order
