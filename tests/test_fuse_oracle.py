"""T1 for the order-free searches: the C restatement (oracle/orb_fuse_oracle.c) of ORBmatcher::Fuse (both overloads,
src/ORBmatcher.cc:1364-1513, :1516-1633) and ORBmatcher::SearchBySim3 (:836-1052) against the reference's own unmodified
ORBmatcher.cc (oracle/_ref, ref_fuse_harness.cc), and against the results stored from it in tests/golden/ref_fuse.npz."""
import os

import numpy as np
import pytest

from fuse_lib import (fuse_scene, replay_fuse, run_fuse, run_search_by_sim3, same_state, sim3_pair_scene, window_best_free_oracle)
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
