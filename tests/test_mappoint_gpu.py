"""Map-point side of the path on the GPU (csrc/orb_mappoint.cu) through the C ABI against the oracle and the
committed reference fixtures: Frame::isInFrustum + MapPoint::PredictScale (float outputs compared by bit pattern),
MapPoint::ComputeDistinctiveDescriptors (index and median), and projection -> window matching without leaving the GPU."""
import os

import numpy as np
import pytest

import orb_slam2_chinesenotes_b200 as ob
from mappoint_lib import descriptor_groups, distinctive, frustum_scene, is_in_frustum

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_mappoint.npz")


def _t(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _project(scenes, cos_limit, shared=False, nlevels=8, sf=1.2):
    import torch
    P, n = len(scenes), len(scenes[0]["xyz"])
    stride = n + 37
    pad = lambda key, w: np.stack([np.concatenate([s[key].reshape(n, -1), np.zeros((stride - n, w), np.float32)]) for s in scenes]).reshape((P, stride, w) if w > 1 else (P, stride))
    if shared:
        pts = {k: _t(pad(k, w)[0]) for k, w in (("xyz", 3), ("normal", 3), ("max_d", 1), ("min_d", 1))}
    else:
        pts = {k: _t(pad(k, w)) for k, w in (("xyz", 3), ("normal", 3), ("max_d", 1), ("min_d", 1))}
    Tcw = _t(np.stack([s["Tcw"].reshape(16) for s in scenes]))
    nq = _t(np.full(P, n, np.int32))
    out = dict(in_view=torch.full((P, stride), 9, dtype=torch.uint8, device="cuda"), proj=torch.zeros((P, stride, 3), device="cuda"),
               level=torch.zeros((P, stride), dtype=torch.int32, device="cuda"), view_cos=torch.zeros((P, stride), device="cuda"))
    cnt = torch.zeros(P, dtype=torch.int32, device="cuda")
    s0 = scenes[0]
    ob.project_points_batch(Tcw, s0["K"], s0["bf"], s0["bounds"], sf, nlevels, nq, stride, pts, out, cos_limit, shared, cnt)
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in out.items()}, cnt.cpu().numpy(), n


@pytest.mark.parametrize("cos_limit", [0.5, 0.0, 0.8])
def test_projection_equals_oracle(cos_limit):
    scenes = [frustum_scene(s) for s in (1, 2, 3, 4, 5)]
    got, cnt, n = _project(scenes, cos_limit)
    for p, s in enumerate(scenes):
        iv, proj, lv, vc = is_in_frustum("oracle", s, cos_limit=cos_limit)
        assert (got["in_view"][p, :n] == iv).all() and cnt[p] == iv.sum() and iv.sum() > 300
        assert (got["in_view"][p, n:] == 9).all()                     # entries past nq are not touched
        m = iv.astype(bool)
        assert (got["level"][p, :n][m] == lv[m]).all()
        assert (got["proj"][p, :n][m].view(np.uint32) == proj[m].view(np.uint32)).all()
        assert (got["view_cos"][p, :n][m].view(np.uint32) == vc[m].view(np.uint32)).all()


def test_projection_shared_points_and_reference_fixture():
    g = np.load(GOLDEN)
    base = frustum_scene(int(g["frustum_seed"]))
    other = dict(base)
    other["Tcw"] = frustum_scene(11)["Tcw"]
    got, cnt, n = _project([base, other], 0.5, shared=True)
    m = g["frustum_in_view"].astype(bool)
    assert (got["in_view"][0, :n] == g["frustum_in_view"]).all() and (got["level"][0, :n][m] == g["frustum_level"][m]).all()
    assert (got["proj"][0, :n][m].view(np.uint32) == g["frustum_proj"][m].view(np.uint32)).all()
    assert (got["view_cos"][0, :n][m].view(np.uint32) == g["frustum_view_cos"][m].view(np.uint32)).all()
    iv, proj, lv, vc = is_in_frustum("oracle", other)
    assert (got["in_view"][1, :n] == iv).all() and (got["proj"][1, :n][iv.astype(bool)].view(np.uint32) == proj[iv.astype(bool)].view(np.uint32)).all()


def test_projection_other_pyramids():
    s = frustum_scene(21)
    for sf, nl in ((1.5, 5), (1.1, 12)):
        got, cnt, n = _project([s], 0.5, nlevels=nl, sf=sf)
        iv, proj, lv, vc = is_in_frustum("oracle", s, scale_factor=sf, nlevels=nl)
        m = iv.astype(bool)
        assert (got["in_view"][0, :n] == iv).all() and (got["level"][0, :n][m] == lv[m]).all() and len(set(lv[m])) >= 4


@pytest.mark.parametrize("seed", [1, 9])
def test_distinctive_descriptors_equal_oracle(seed):
    import torch
    groups = descriptor_groups(seed) + [(np.zeros((0, 32), np.uint8), np.zeros(0, np.uint8))]
    rng = np.random.default_rng(seed)
    big = rng.integers(0, 256, (700, 32), dtype=np.uint8)            # unrelated descriptors: medians around 128
    groups.append((big, (rng.random(700) < 0.1).astype(np.uint8)))
    allbad = (groups[3][0], np.ones(len(groups[3][0]), np.uint8))
    groups.append(allbad)
    desc = _t(np.concatenate([g[0] for g in groups]))
    bad = _t(np.concatenate([g[1] for g in groups]))
    off = _t(np.concatenate([[0], np.cumsum([len(g[0]) for g in groups])]).astype(np.int32))
    for use_bad in (False, True):
        idx = torch.zeros(len(groups), dtype=torch.int32, device="cuda")
        med = torch.zeros(len(groups), dtype=torch.int32, device="cuda")
        ob.distinctive_descriptors(desc, off, idx, med, bad if use_bad else None)
        torch.cuda.synchronize()
        idx, med = idx.cpu().numpy(), med.cpu().numpy()
        for k, (d, b) in enumerate(groups):
            _, wi, wm = distinctive("oracle", d, b if use_bad else None) if len(d) else (None, None, None)
            if wi is None:
                assert idx[k] == -1
            else:
                assert idx[k] == wi and med[k] == wm, (k, len(d))
    if seed == 9:                                                      # the reference's own choice (fixture)
        g = np.load(GOLDEN)
        idx = torch.zeros(len(groups), dtype=torch.int32, device="cuda")
        ob.distinctive_descriptors(desc, off, idx, None, bad)
        torch.cuda.synchronize()
        for k, (d, b) in enumerate(descriptor_groups(9)):
            want = g[f"distinctive_bad_{k}"]
            i = int(idx[k])
            assert (want.size == 0 and i == -1) or (d[i] == want).all()


def test_projection_feeds_window_matching_on_device():
    """Tracking::SearchLocalPoints end to end on the GPU: isInFrustum for every map point, then
    SearchByProjection(Frame, MapPoints) on the arrays the projection wrote; equal to oracle projection + oracle matcher."""
    import torch
    from matcher_lib import Matcher, extract_frame, flip_bits
    W, H = 1241, 376
    kps, desc, scale = extract_frame(W, H, 2000, 2)
    rng = np.random.default_rng(4)
    n = len(kps)
    fx, fy, cx, cy = [float(v) for v in np.float32([718.856, 718.856, 607.1928, 185.2157])]
    s = frustum_scene(31, n=10)                                         # pose only
    Tcw = s["Tcw"]
    R, t = Tcw[:3, :3].astype(np.float64), Tcw[:3, 3].astype(np.float64)
    z = rng.uniform(4, 40, n)
    pc = np.stack([(kps["x"] + rng.normal(0, 1.5, n) - cx) / fx * z, (kps["y"] + rng.normal(0, 1.5, n) - cy) / fy * z, z], 1)
    xyz = ((pc - t) @ R).astype(np.float32)
    Ow = -(R.T @ t)
    view = xyz - Ow
    dist = np.linalg.norm(view, axis=1)
    normal = (view / dist[:, None]).astype(np.float32)
    max_d = (dist * np.float32(1.2) ** kps["octave"]).astype(np.float32)
    min_d = (max_d / np.float32(1.2 ** 7)).astype(np.float32)
    sc = dict(Tcw=Tcw, K=s["K"], bf=s["bf"], bounds=(0.0, float(W), 0.0, float(H)), xyz=xyz, normal=normal, max_d=max_d, min_d=min_d)
    qdesc = flip_bits(desc, rng, 60)
    iv, proj, lv, vc = is_in_frustum("oracle", sc)
    q = dict(proj=proj, level=lv, view_cos=vc, in_view=iv, bad=np.zeros(n, np.uint8), obs=np.ones(n, np.int32), desc=qdesc)
    want = Matcher("oracle").search_by_projection_points(kps, desc, None, scale, sc["bounds"], q, 3.0, 0.8, None)
    assert want[0] > 800
    # device: project, then match on the same buffers
    pts = {k: _t(sc[k][None]) for k in ("xyz", "normal", "max_d", "min_d")}
    out = dict(in_view=torch.zeros((1, n), dtype=torch.uint8, device="cuda"), proj=torch.zeros((1, n, 3), device="cuda"),
               level=torch.zeros((1, n), dtype=torch.int32, device="cuda"), view_cos=torch.zeros((1, n), device="cuda"))
    nq = _t(np.int32([n]))
    d_Tcw = _t(Tcw.reshape(1, 16))
    ob.project_points_batch(d_Tcw, sc["K"], sc["bf"], sc["bounds"], 1.2, 8, nq, n, pts, out)
    d_kps = _t(kps.view(np.uint8).reshape(1, n, 28))
    d_desc = _t(desc[None])                                              # frames_batch keeps raw pointers: hold the tensors
    F = ob.frames_batch(d_kps, d_desc, nq, sc["bounds"])
    dq = dict(out, bad=_t(q["bad"][None]), obs=_t(q["obs"][None]), desc=_t(qdesc[None]))
    assign = torch.zeros((1, n), dtype=torch.int32, device="cuda")
    nm = torch.zeros(1, dtype=torch.int32, device="cuda")
    ob.search_by_projection_points_batch(F, scale, dq, nq, n, assign, nm, 3.0, 0.8)
    torch.cuda.synchronize()
    assert int(nm[0]) == want[0] and (assign[0].cpu().numpy() == want[1]).all()


def test_single_problem_host_arrays():
    """orbm_project_points / orbm_distinctive_descriptor (what the C++ forwarders call): host arrays in."""
    s = frustum_scene(41)
    iv, proj, lv, vc = ob.project_points(s["Tcw"], s["K"], s["bf"], s["bounds"], 1.2, 8, s["xyz"], s["normal"], s["max_d"], s["min_d"])
    wiv, wproj, wlv, wvc = is_in_frustum("oracle", s)
    m = wiv.astype(bool)
    assert (iv == wiv).all() and (lv[m] == wlv[m]).all() and (proj[m].view(np.uint32) == wproj[m].view(np.uint32)).all()
    assert (vc[m].view(np.uint32) == wvc[m].view(np.uint32)).all() and (proj[~m] == 0).all()
    for d, bad in descriptor_groups(4, sizes=(1, 7, 40, 300)):
        for b in (None, bad):
            _, wi, wm = distinctive("oracle", d, b)
            gi, gm = ob.distinctive_descriptor(d, b)
            assert (wi is None and gi == -1) or (gi == wi and gm == wm)
    assert ob.distinctive_descriptor(np.zeros((0, 32), np.uint8)) == (-1, -1)
