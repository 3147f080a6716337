"""T1 for the matchers: the C restatement (oracle/orb_match_oracle.c) against the reference's own
unmodified ORBmatcher.cc / Frame.cc (oracle/_ref) on the synthetic scenes of SURVEY.md App. E.2."""
import ctypes as C

import numpy as np
import pytest

from matcher_lib import Matcher, extract_frame, perturbed_frame, projected_queries, two_view_scene
from oracle_lib import KP_DTYPE, OracleExtractor, oracle, ref
from synth import stereo_pair

needs_ref = pytest.mark.skipif(ref() is None, reason="oracle/_ref/liborbref.so not built (needs /root/reference)")
W, H, NF = 1241, 376, 2000
BOUNDS = (0.0, float(W), 0.0, float(H))
K = np.float32([718.856, 718.856, 607.1928, 185.2157])


@pytest.fixture(scope="module")
def scene():
    kps, desc, scale = extract_frame(W, H, NF, 2)
    k2, d2, perm = perturbed_frame(kps, desc, W, H, 11)
    return dict(kps=kps, desc=desc, scale=scale, k2=k2, d2=d2)


@needs_ref
def test_descriptor_distance(scene):
    R, O = ref(), oracle()
    R.orbref_descriptor_distance.argtypes = [C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(0)
    for _ in range(300):
        a, b = rng.integers(0, len(scene["desc"]), 2)
        da, db = scene["desc"][a], scene["d2"][b]
        assert R.orbref_descriptor_distance(da.ctypes.data, db.ctypes.data) == O.orbo_descriptor_distance(da.ctypes.data, db.ctypes.data) \
            == int(np.unpackbits(da ^ db).sum())


@needs_ref
def test_features_in_area(scene):
    R, O = Matcher("ref"), Matcher("oracle")
    rng = np.random.default_rng(1)
    for _ in range(200):
        x, y = float(rng.uniform(-20, W + 20)), float(rng.uniform(-20, H + 20))
        r = float(rng.choice([2.5, 7.0, 15.0, 40.0, 100.0]))
        lv = [(-1, -1), (0, 0), (0, 3), (2, -1), (3, 4), (0, -1)][int(rng.integers(0, 6))]
        a = R.features_in_area(scene["k2"], scene["scale"], BOUNDS, x, y, r, *lv)
        b = O.features_in_area(scene["k2"], scene["scale"], BOUNDS, x, y, r, *lv)
        assert a.tolist() == b.tolist()


@needs_ref
@pytest.mark.parametrize("nnratio,check_ori,window", [(0.9, True, 100), (0.6, True, 50), (0.9, False, 100)])
def test_search_for_initialization(scene, nnratio, check_ori, window):
    R, O = Matcher("ref"), Matcher("oracle")
    prev = np.stack([scene["kps"]["x"], scene["kps"]["y"]], 1)
    a = R.search_for_initialization(scene["kps"], scene["desc"], scene["k2"], scene["d2"], scene["scale"], BOUNDS, prev, window, nnratio, check_ori)
    b = O.search_for_initialization(scene["kps"], scene["desc"], scene["k2"], scene["d2"], scene["scale"], BOUNDS, prev, window, nnratio, check_ori)
    assert a[0] == b[0] and a[0] > 50
    assert (a[1] == b[1]).all() and (a[2] == b[2]).all()


@needs_ref
@pytest.mark.parametrize("th,nnratio,with_uright", [(1.0, 0.8, False), (3.0, 0.8, False), (5.0, 0.9, True)])
def test_search_by_projection_points(scene, th, nnratio, with_uright):
    R, O = Matcher("ref"), Matcher("oracle")
    q = projected_queries(scene["k2"], scene["d2"], 2000, 5)
    rng = np.random.default_rng(6)
    ur = np.where(rng.random(len(scene["k2"])) < 0.5, scene["k2"]["x"] - 20 * rng.random(len(scene["k2"])), -1).astype(np.float32) if with_uright else None
    init = np.where(rng.random(len(scene["k2"])) < 0.05, rng.integers(0, 2000, len(scene["k2"])), -1).astype(np.int32)
    for ia in (None, init):
        a = R.search_by_projection_points(scene["k2"], scene["d2"], ur, scene["scale"], BOUNDS, q, th, nnratio, ia)
        b = O.search_by_projection_points(scene["k2"], scene["d2"], ur, scene["scale"], BOUNDS, q, th, nnratio, ia)
        assert a[0] == b[0] and a[0] > 300
        assert (a[1] == b[1]).all()
    # nmatches can exceed the number of attached keypoints
    assert a[0] >= (a[1] >= 0).sum() - (init >= 0).sum()


@needs_ref
@pytest.mark.parametrize("th,mono,check_ori", [(7.0, False, True), (15.0, True, True), (7.0, False, False)])
def test_search_by_projection_frame(scene, th, mono, check_ori):
    R, O = Matcher("ref"), Matcher("oracle")
    cur, last, Tc, Tl = two_view_scene(scene["kps"], scene["desc"], W, H, 21, K)
    if mono:
        cur = dict(cur); cur["u_right"] = None
    rng = np.random.default_rng(8)
    init_obs = np.where(rng.random(len(cur["kps"])) < 0.05, rng.integers(0, 2, len(cur["kps"])), -1).astype(np.int32)
    for io in (None, init_obs):
        a = R.search_by_projection_frame(cur, last, Tc, Tl, K, 386.1448, scene["scale"], BOUNDS, th, mono, 0.9, check_ori, io)
        b = O.search_by_projection_frame(cur, last, Tc, Tl, K, 386.1448, scene["scale"], BOUNDS, th, mono, 0.9, check_ori, io)
        assert a[0] == b[0] and a[0] > 200, (a[0], b[0])
        assert (a[1] == b[1]).all()


@needs_ref
def test_stereo_frame_vs_oracle():
    """The reference's own stereo Frame constructor (two extractors on two threads + ComputeStereoMatches)
    against oracle extraction + orbo_stereo_matches: keypoints, descriptors, uRight, depth bit-identical."""
    R = ref()
    left, right = stereo_pair(W, H, 2)
    cap = 2 * NF
    kl, kr = np.zeros(cap, KP_DTYPE), np.zeros(cap, KP_DTYPE)
    dl, dr = np.zeros((cap, 32), np.uint8), np.zeros((cap, 32), np.uint8)
    ur, dep, nr = np.zeros(cap, np.float32), np.zeros(cap, np.float32), C.c_int()
    vp, ci, cf = C.c_void_p, C.c_int, C.c_float
    R.orbref_stereo_frame.argtypes = [vp, vp, ci, ci, ci, cf, ci, ci, ci, cf, cf, cf, cf, cf, cf, vp, vp, ci, vp, vp, ci, C.POINTER(ci), vp, vp]
    bf, fx = 386.1448, float(K[0])
    n = R.orbref_stereo_frame(left.ctypes.data, right.ctypes.data, W, H, NF, 1.2, 8, 20, 7, fx, float(K[1]), float(K[2]), float(K[3]), bf, 35.0,
                              kl.ctypes.data, dl.ctypes.data, cap, kr.ctypes.data, dr.ctypes.data, cap, C.byref(nr), ur.ctypes.data, dep.ctypes.data)
    OL, OR = OracleExtractor(NF), OracleExtractor(NF)
    nl, okl, odl = OL.extract(left)
    nr2, okr, odr = OR.extract(right)
    assert n == nl and nr.value == nr2
    assert all((kl[:n][f] == okl[f]).all() for f in KP_DTYPE.names) and (dl[:n] == odl).all()
    assert all((kr[:nr2][f] == okr[f]).all() for f in KP_DTYPE.names) and (dr[:nr2] == odr).all()
    our, odep = np.zeros(nl, np.float32), np.zeros(nl, np.float32)
    O = oracle()
    O.orbo_stereo_matches.argtypes = [vp, vp, ci, vp, vp, ci, vp, vp, cf, cf, vp, vp]
    O.orbo_stereo_matches(OL.h, OR.h, nl, okl.ctypes.data, odl.ctypes.data, nr2, okr.ctypes.data, odr.ctypes.data, bf, fx, our.ctypes.data, odep.ctypes.data)
    assert (our.view(np.uint32) == ur[:n].view(np.uint32)).all() and (odep.view(np.uint32) == dep[:n].view(np.uint32)).all()
    assert (ur[:n] >= 0).sum() > 200
    OL.close(); OR.close()


@needs_ref
def test_stereo_bench_harness_is_deterministic():
    """orbref_stereo_bench (the CPU baseline of bench.py: several Frames in flight, persistent extractors) gives every
    pair the depths of the single-frame harness whatever the worker count -- the reference reads Frame::mb before the
    constructor sets it (src/Frame.cc:93 vs :121); the harness pins the steady-state value mbf/fx."""
    R = ref()
    w, h, nf, pairs, cap = 640, 480, 1000, 3, 1200
    frames = np.stack([im for p in range(pairs) for im in stereo_pair(w, h, 70 + p)])
    vp, ci, cf = C.c_void_p, C.c_int, C.c_float
    R.orbref_stereo_frame.argtypes = [vp, vp, ci, ci, ci, cf, ci, ci, ci, cf, cf, cf, cf, cf, cf, vp, vp, ci, vp, vp, ci, C.POINTER(ci), vp, vp]
    bf, fx = 386.1448, float(K[0])
    want = np.full((pairs, cap), -2, np.float32)
    for p in range(pairs):
        kl, dl = np.zeros(cap, KP_DTYPE), np.zeros((cap, 32), np.uint8)
        n = R.orbref_stereo_frame(frames[2 * p].ctypes.data, frames[2 * p + 1].ctypes.data, w, h, nf, 1.2, 8, 20, 7, fx, float(K[1]), float(K[2]),
                                  float(K[3]), bf, 35.0, kl.ctypes.data, dl.ctypes.data, cap, None, None, 0, None, None, want[p].ctypes.data)
        assert 0 < n <= cap
    for workers in (1, 2):
        got = np.full((pairs, cap), -2, np.float32)
        tot = C.c_longlong()
        R.orbref_stereo_bench(nf, 1.2, 8, 20, 7, frames.ctypes.data, pairs, w, h, workers, fx, float(K[1]), float(K[2]), float(K[3]), bf, 35.0,
                              C.byref(tot), got.ctypes.data, cap)
        assert (got.view(np.uint32) == want.view(np.uint32)).all(), workers
        assert tot.value == int((want > 0).sum()) > 100


def test_oracle_vs_matcher_fixtures(scene):
    """Always runs: results stored from the reference's unmodified ORBmatcher.cc / Frame.cc
    (tests/golden/make_golden.py) against the C restatement."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_match_kitti.npz"))
    O = Matcher("oracle")
    prev = np.stack([scene["kps"]["x"], scene["kps"]["y"]], 1)
    n, m12, pv = O.search_for_initialization(scene["kps"], scene["desc"], scene["k2"], scene["d2"], scene["scale"], BOUNDS, prev, 100, 0.9, True)
    assert n == int(g["init_n"]) and (m12 == g["init_m12"]).all() and (pv == g["init_prev"]).all()
    q = projected_queries(scene["k2"], scene["d2"], 2000, 5)
    for th in (1.0, 3.0):
        n, a = O.search_by_projection_points(scene["k2"], scene["d2"], None, scene["scale"], BOUNDS, q, th, 0.8, None)
        assert n == int(g[f"points_th{int(th)}_n"]) and (a == g[f"points_th{int(th)}_assign"]).all()
    cur, last, Tc, Tl = two_view_scene(scene["kps"], scene["desc"], W, H, 21, K)
    n, a = O.search_by_projection_frame(cur, last, Tc, Tl, K, 386.1448, scene["scale"], BOUNDS, 7.0, False, 0.9, True, None)
    assert n == int(g["frame_n"]) and (a == g["frame_assign"]).all()
    left, right = stereo_pair(W, H, 2)
    OL, OR = OracleExtractor(NF), OracleExtractor(NF)
    nl, okl, odl = OL.extract(left)
    nr, okr, odr = OR.extract(right)
    assert nl == int(g["stereo_n"]) and nr == int(g["stereo_nr"])
    our, odep = np.zeros(nl, np.float32), np.zeros(nl, np.float32)
    Lo = oracle()
    vp, ci, cf = C.c_void_p, C.c_int, C.c_float
    Lo.orbo_stereo_matches.argtypes = [vp, vp, ci, vp, vp, ci, vp, vp, cf, cf, vp, vp]
    Lo.orbo_stereo_matches(OL.h, OR.h, nl, okl.ctypes.data, odl.ctypes.data, nr, okr.ctypes.data, odr.ctypes.data, 386.1448, float(K[0]),
                           our.ctypes.data, odep.ctypes.data)
    assert (our.view(np.uint32) == g["stereo_uright"].view(np.uint32)).all()
    assert (odep.view(np.uint32) == g["stereo_depth"].view(np.uint32)).all()
    OL.close(); OR.close()


@needs_ref
@pytest.mark.parametrize("th,orb_dist,check_ori", [(10.0, 100, True), (3.0, 64, True), (10.0, 100, False)])
def test_relocalisation_overload(scene, th, orb_dist, check_ori):
    """src/ORBmatcher.cc:303-431 (the th/ORBdist pairs are the ones Tracking::Relocalization passes)."""
    from reloc_lib import reloc_scene, run_reloc
    s = reloc_scene(scene["kps"], scene["desc"], W, H, 31, K)
    a = run_reloc("ref", s, scene["scale"], BOUNDS, K, th, orb_dist, check_ori)
    b = run_reloc("oracle", s, scene["scale"], BOUNDS, K, th, orb_dist, check_ori)
    assert a[0] == b[0] and a[0] > 100, (a[0], b[0])
    assert (a[1] == b[1]).all()


def test_oracle_vs_projection_overload_fixtures(scene):
    """Always runs: relocalisation and loop-closing overloads, results stored from the reference's unmodified code."""
    import os
    from reloc_lib import reloc_scene, run_reloc
    from sim3_lib import run_sim3, sim3_scene
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_match_projection.npz"))
    s = reloc_scene(scene["kps"], scene["desc"], W, H, 31, K)
    for th, od, co in ((10.0, 100, True), (3.0, 64, True)):
        n, a = run_reloc("oracle", s, scene["scale"], BOUNDS, K, th, od, co)
        assert n == int(g[f"reloc_th{int(th)}_n"]) and (a == g[f"reloc_th{int(th)}_assign"]).all()
    for th, seed in ((10, 41), (4, 42)):
        s3 = sim3_scene(scene["k2"], scene["d2"], W, H, seed, K)
        n, a = run_sim3("oracle", scene["k2"], scene["d2"], s3, scene["scale"], BOUNDS, K, th)
        assert n == int(g[f"sim3_th{th}_n"]) and (a == g[f"sim3_th{th}_assign"]).all()


@needs_ref
@pytest.mark.parametrize("th,seed", [(10, 41), (4, 42), (20, 43)])
def test_loop_closing_overload(scene, th, seed):
    """src/ORBmatcher.cc:434-549 (KeyFrame, Sim3): the restatement against the reference's unmodified code."""
    from sim3_lib import run_sim3, sim3_scene
    s = sim3_scene(scene["k2"], scene["d2"], W, H, seed, K)
    a = run_sim3("ref", scene["k2"], scene["d2"], s, scene["scale"], BOUNDS, K, th)
    b = run_sim3("oracle", scene["k2"], scene["d2"], s, scene["scale"], BOUNDS, K, th)
    assert a[0] == b[0] and a[0] > 100
    assert (a[1] == b[1]).all()
    assert ((a[1] != s["matched"]) & (s["matched"] != -1)).sum() == 0          # pre-matched keypoints are never overwritten
