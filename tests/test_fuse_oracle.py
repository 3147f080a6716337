"""T1 for the order-free searches: the C restatement (oracle/orb_fuse_oracle.c) of ORBmatcher::Fuse (both overloads,
src/ORBmatcher.cc:1364-1513, :1516-1633) and ORBmatcher::SearchBySim3 (:836-1052) against the reference's own unmodified
ORBmatcher.cc (oracle/_ref, ref_fuse_harness.cc), and against the results stored from it in tests/golden/ref_fuse.npz."""
import os

import numpy as np
import pytest

from fuse_lib import (fuse_list_arrays, fuse_pose24, fuse_project_oracle, fuse_scene, raw_distances, replay_fuse, run_fuse, run_search_by_sim3,
                      same_state, sim3_pair_scene, sim3_poses, window_best_free_oracle)
from matcher_lib import extract_frame, perturbed_frame
from oracle_lib import ref

needs_ref = pytest.mark.skipif(ref() is None, reason="oracle/_ref/liborbref.so not built (needs /root/reference)")
W, H, NF = 1241, 376, 2000
BOUNDS = (0.0, float(W), 0.0, float(H))
K = np.float32([718.856, 718.856, 607.1928, 185.2157])
BF = 386.1448
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_fuse.npz")
FUSE_CASES = [(3.0, 51, False, True), (3.0, 52, False, False), (2.5, 53, False, True), (4.0, 54, True, True), (3.0, 55, True, False)]
SIM3_CASES = [(7.5, 61), (4.0, 62)]


@pytest.fixture(scope="module")
def scene():
    kps, desc, scale = extract_frame(W, H, NF, 2)
    k2, d2, _ = perturbed_frame(kps, desc, W, H, 11, shift=4, kmax=40)
    inv_sigma2 = (np.float32(1.0) / (scale * scale)).astype(np.float32)        # src/ORBextractor.cc:511-518
    return dict(kps=kps, desc=desc, scale=scale, inv_sigma2=inv_sigma2, k2=k2, d2=d2)


@needs_ref
@pytest.mark.parametrize("th,seed,sim3,stereo", FUSE_CASES)
def test_fuse_matches_the_reference(scene, th, seed, sim3, stereo):
    s = fuse_scene(scene["kps"], scene["desc"], W, H, seed, K, BF, sim3)
    a = run_fuse("ref", scene["kps"], scene["desc"], s, scene["scale"], scene["inv_sigma2"], BOUNDS, K, BF, th, sim3, stereo)
    b = run_fuse("oracle", scene["kps"], scene["desc"], s, scene["scale"], scene["inv_sigma2"], BOUNDS, K, BF, th, sim3, stereo)
    assert a[0] == b[0] and a[0] > 300
    assert same_state(a[1], b[1])
    if sim3:
        assert (b[1]["replace"] >= 0).sum() > 50
    else:
        assert (b[1]["replaced_by"] >= 0).sum() > 50 and (b[1]["bad"] != s["bad"]).sum() > 50


@needs_ref
@pytest.mark.parametrize("th,seed", SIM3_CASES)
def test_search_by_sim3_matches_the_reference(scene, th, seed):
    s = sim3_pair_scene(scene["kps"], scene["desc"], scene["k2"], scene["d2"], W, H, seed, K)
    a = run_search_by_sim3("ref", scene["kps"], scene["desc"], scene["k2"], scene["d2"], s, scene["scale"], BOUNDS, K, th)
    b = run_search_by_sim3("oracle", scene["kps"], scene["desc"], scene["k2"], scene["d2"], s, scene["scale"], BOUNDS, K, th)
    assert a[0] == b[0] and a[0] > 200
    assert (a[1] == b[1]).all()


def test_oracle_against_stored_reference_results(scene):
    """Always runs (also where /root/reference does not exist): results of the reference's unmodified code, stored by
    tests/golden/make_golden.py."""
    g = np.load(GOLDEN)
    for th, seed, sim3, stereo in FUSE_CASES:
        s = fuse_scene(scene["kps"], scene["desc"], W, H, seed, K, BF, sim3)
        nf, st, _ = run_fuse("oracle", scene["kps"], scene["desc"], s, scene["scale"], scene["inv_sigma2"], BOUNDS, K, BF, th, sim3, stereo)
        key = f"fuse_{seed}"
        assert nf == int(g[key + "_n"])
        for k, v in st.items():
            assert (v == g[f"{key}_{k}"]).all(), (key, k)
    for th, seed in SIM3_CASES:
        s = sim3_pair_scene(scene["kps"], scene["desc"], scene["k2"], scene["d2"], W, H, seed, K)
        nf, m12, _ = run_search_by_sim3("oracle", scene["kps"], scene["desc"], scene["k2"], scene["d2"], s, scene["scale"], BOUNDS, K, th)
        assert nf == int(g[f"sim3_{seed}_n"]) and (m12 == g[f"sim3_{seed}_m12"]).all()


@pytest.mark.parametrize("th,seed,sim3,stereo", FUSE_CASES)
def test_search_then_replay_equals_the_sequential_loop(scene, th, seed, sim3, stereo):
    """The split the device path makes: every point's search is independent of the map bookkeeping, so searching all points
    first (orbo_window_best_free = the device contract) and replaying the bookkeeping in list order afterwards gives the
    sequential loop's result."""
    s = fuse_scene(scene["kps"], scene["desc"], W, H, seed, K, BF, sim3)
    nf, st, q = run_fuse("oracle", scene["kps"], scene["desc"], s, scene["scale"], scene["inv_sigma2"], BOUNDS, K, BF, th, sim3, stereo)
    ur = s["u_right"] if stereo else None
    qdesc = np.ascontiguousarray(s["mp_desc"][np.maximum(s["list"], 0)])
    _, bi, bd = window_best_free_oracle(scene["kps"], scene["desc"], ur, BOUNDS, q, qdesc, None if sim3 else scene["inv_sigma2"], 50)
    nf2, st2 = replay_fuse(s, q, bi, ur, sim3)
    assert nf2 == nf and same_state(st, st2)
    assert ((bd <= 50) == (bi >= 0)).all()


def _same_queries(a, b, with_ur):
    m = a["valid"].astype(bool)
    ok = (a["valid"] == b["valid"]).all() and (a["uvr"][m].view(np.uint32) == b["uvr"][m].view(np.uint32)).all() and (a["level"][m] == b["level"][m]).all()
    return ok and (not with_ur or (a["ur"][m].view(np.uint32) == b["ur"][m].view(np.uint32)).all())


@pytest.mark.parametrize("th,seed,sim3,stereo", FUSE_CASES)
def test_projection_prologue_is_the_loops_prologue(scene, th, seed, sim3, stereo):
    """orbo_fuse_project (the contract of the device's orbm_fuse_project_batch, with the real PredictScale arithmetic) against the
    queries the Fuse restatement -- pinned against the reference -- emits when its preset PredictScale returns the same levels."""
    s = fuse_scene(scene["kps"], scene["desc"], W, H, seed, K, BF, sim3)
    s, max_raw, min_raw = raw_distances(s, np.random.default_rng(seed))
    L = fuse_list_arrays(s, max_raw, min_raw, sim3)
    got = fuse_project_oracle(False, fuse_pose24(s, sim3), K, 0.0 if sim3 else BF, BOUNDS, 1.2, scene["scale"], th, L["xyz"], L["normal"], L["max_d"],
                              L["min_d"], L["skip"])
    lvl = s["level"].copy()
    lvl[s["list"][got["valid"] == 1]] = got["level"][got["valid"] == 1]             # the preset PredictScale of the restatement
    s["level"] = lvl
    _, _, q = run_fuse("oracle", scene["kps"], scene["desc"], s, scene["scale"], scene["inv_sigma2"], BOUNDS, K, BF, th, sim3, stereo)
    assert _same_queries(q, got, not sim3) and got["valid"].sum() > 500 and len(np.unique(got["level"][got["valid"] == 1])) >= 6


@pytest.mark.parametrize("th,seed", SIM3_CASES)
def test_sim3_projection_prologue(scene, th, seed):
    k1, d1, k2, d2 = scene["kps"], scene["desc"], scene["k2"], scene["d2"]
    s = sim3_pair_scene(k1, d1, k2, d2, W, H, seed, K)
    rng = np.random.default_rng(seed)
    max_raw = (3.0 + 60 * rng.random(s["npts"])).astype(np.float32)
    min_raw = (max_raw / np.float32(1.2 ** 7)).astype(np.float32)
    s["max_dist"], s["min_dist"] = (np.float32(1.2) * max_raw).astype(np.float32), (np.float32(0.8) * min_raw).astype(np.float32)
    p1, p2 = sim3_poses(s)
    am1 = s["m12"] >= 0
    am2 = np.zeros(len(k2), bool)
    idx2 = s["idx_in_kf2"][s["m12"][am1]]
    am2[idx2[(idx2 >= 0) & (idx2 < len(k2))]] = True
    outs = []
    for pose, mp, am in ((p1, s["mp1"], am1), (p2, s["mp2"], am2)):
        g = np.maximum(mp, 0)
        skip = ((mp < 0) | am | (s["bad"][g] != 0)).astype(np.uint8)
        outs.append(fuse_project_oracle(True, pose, K, 0.0, BOUNDS, 1.2, scene["scale"], th, np.ascontiguousarray(s["xyz"][g]), None,
                                        np.ascontiguousarray(max_raw[g]), np.ascontiguousarray(min_raw[g]), skip))
    lvl = s["level"].copy()
    for o, mp in zip(outs, (s["mp1"], s["mp2"])):
        v = o["valid"] == 1
        lvl[mp[v]] = o["level"][v]
    s["level"] = lvl
    _, _, (q1, q2) = run_search_by_sim3("oracle", k1, d1, k2, d2, s, scene["scale"], BOUNDS, K, th)
    assert _same_queries(q1, outs[0], False) and _same_queries(q2, outs[1], False)
    assert outs[0]["valid"].sum() > 500 and outs[1]["valid"].sum() > 500


@needs_ref
def test_degenerate_inputs_match_the_reference(scene):
    """Empty point list, all points bad, a key frame without keypoints, an empty partner key frame: restatement = reference."""
    kps, desc, scale, s2 = scene["kps"], scene["desc"], scene["scale"], scene["inv_sigma2"]
    base = fuse_scene(kps, desc, W, H, 77, K, BF, False)
    variants = []
    v = dict(base); v["list"] = base["list"][:0].copy(); variants.append((v, kps, desc))
    v = dict(base); v["bad"] = np.ones_like(base["bad"]); variants.append((v, kps, desc))
    v = dict(base); v["list"] = np.full(50, -1, np.int32); variants.append((v, kps, desc))
    for v, k, d in variants:
        for sim3 in (False, True):
            if sim3 and (v["list"] < 0).any():
                continue                                                    # the Sim3 overload dereferences every entry
            a = run_fuse("ref", k, d, v, scale, s2, BOUNDS, K, BF, 3.0, sim3, True)
            b = run_fuse("oracle", k, d, v, scale, s2, BOUNDS, K, BF, 3.0, sim3, True)
            assert a[0] == b[0] == 0 and same_state(a[1], b[1])
    # a key frame without keypoints: nothing to fuse into
    e = fuse_scene(kps[:0], desc[:0], W, H, 78, K, BF, False)
    e["list"] = np.zeros(0, np.int32)
    a = run_fuse("ref", kps[:0], desc[:0], e, scale, s2, BOUNDS, K, BF, 3.0, False, False)
    b = run_fuse("oracle", kps[:0], desc[:0], e, scale, s2, BOUNDS, K, BF, 3.0, False, False)
    assert a[0] == b[0] == 0
    # SearchBySim3 against an empty partner
    s = sim3_pair_scene(kps, desc, scene["k2"][:0], scene["d2"][:0], W, H, 79, K)
    a = run_search_by_sim3("ref", kps, desc, scene["k2"][:0], scene["d2"][:0], s, scale, BOUNDS, K, 7.5)
    b = run_search_by_sim3("oracle", kps, desc, scene["k2"][:0], scene["d2"][:0], s, scale, BOUNDS, K, 7.5)
    assert a[0] == b[0] == 0 and (a[1] == b[1]).all()
