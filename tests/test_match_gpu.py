"""GPU matchers through the C ABI against the oracle (and the reference fixtures): brute-force
Hamming, the three window searches incl. their order-dependent bookkeeping and the rotation
histogram, and the stereo row matcher.  Match indices and float outputs are bit-exact."""
import os

import numpy as np
import pytest

import orb_slam2_chinesenotes_b200 as ob
from matcher_lib import Matcher, extract_frame, perturbed_frame, projected_queries, two_view_scene
from oracle_lib import KP_DTYPE, OracleExtractor, oracle
from synth import stereo_pair

pytestmark = pytest.mark.gpu
W, H, NF = 1241, 376, 2000
BOUNDS = (0.0, float(W), 0.0, float(H))
K = np.float32([718.856, 718.856, 607.1928, 185.2157])


@pytest.fixture(scope="module")
def scene():
    kps, desc, scale = extract_frame(W, H, NF, 2)
    k2, d2, perm = perturbed_frame(kps, desc, W, H, 11)
    return dict(kps=kps, desc=desc, scale=scale, k2=k2, d2=d2)


def test_hamming_bf_matches_oracle(scene):
    O = oracle()
    q, t = scene["desc"][:700], scene["d2"]
    bi, bd, bd2 = ob.hamming_bf(q, t)
    oi, od, od2 = (np.zeros(len(q), np.int32) for _ in range(3))
    O.orbo_hamming_bf(q.ctypes.data, len(q), t.ctypes.data, len(t), oi.ctypes.data, od.ctypes.data, od2.ctypes.data)
    assert (bi == oi).all() and (bd == od).all() and (bd2 == od2).all()
    # edge cases: one train descriptor, duplicates (first wins), DescriptorDistance
    bi, bd, bd2 = ob.hamming_bf(q[:5], t[:1])
    assert (bi == 0).all() and (bd2 == 256).all()
    tt = np.concatenate([t[:10], t[:10]])
    bi, bd, bd2 = ob.hamming_bf(t[:10], tt)
    assert (bi == np.arange(10)).all() and (bd == 0).all() and (bd2 == 0).all()
    assert ob.ORBmatcher.DescriptorDistance(q[0], t[0]) == int(np.unpackbits(q[0] ^ t[0]).sum())


def test_hamming_bf_async_on_a_stream(scene):
    import torch
    rng = np.random.default_rng(4)
    q = torch.from_numpy(rng.integers(0, 256, (3, 130, 32), dtype=np.uint8)).cuda()
    t = torch.from_numpy(rng.integers(0, 256, (3, 200, 32), dtype=np.uint8)).cuda()
    outs = [torch.zeros(3 * 130, dtype=torch.int32, device="cuda") for _ in range(3)]
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    ob.hamming_bf_async(q, t, outs, nprob=3, stream=st.cuda_stream)
    st.synchronize()
    want = ob.hamming_bf(q.cpu().numpy(), t.cpu().numpy(), nprob=3)
    assert all((o.cpu().numpy() == w).all() for o, w in zip(outs, want))
    with pytest.raises(ob.OrbError):
        ob.hamming_bf_async(q.cpu(), t, outs, nprob=3)


def test_hamming_bf_batched_problems(scene):
    rng = np.random.default_rng(3)
    q = rng.integers(0, 256, (4, 300, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (4, 257, 32), dtype=np.uint8)
    bi, bd, bd2 = ob.hamming_bf(q, t, nprob=4)
    for p in range(4):
        d = np.unpackbits(q[p][:, None, :] ^ t[p][None, :, :], axis=2).sum(2)
        assert (bi[p * 300:(p + 1) * 300] == d.argmin(1)).all() and (bd[p * 300:(p + 1) * 300] == d.min(1)).all()


@pytest.mark.parametrize("th,nnratio,with_uright", [(1.0, 0.8, False), (3.0, 0.8, False), (5.0, 0.9, True)])
def test_search_by_projection_points(scene, th, nnratio, with_uright):
    O = Matcher("oracle")
    q = projected_queries(scene["k2"], scene["d2"], 2000, 5)
    rng = np.random.default_rng(6)
    n = len(scene["k2"])
    ur = np.where(rng.random(n) < 0.5, scene["k2"]["x"] - 20 * rng.random(n), -1).astype(np.float32) if with_uright else None
    init = np.where(rng.random(n) < 0.05, rng.integers(0, 2000, n), -1).astype(np.int32)
    F = ob.FrameView(scene["k2"], scene["d2"], BOUNDS, ur)
    M = ob.ORBmatcher(nnratio, True)
    for ia in (None, init):
        a = O.search_by_projection_points(scene["k2"], scene["d2"], ur, scene["scale"], BOUNDS, q, th, nnratio, ia)
        b = M.SearchByProjection(F, scene["scale"], q, th, ia)
        assert a[0] == b[0] and a[0] > 300
        assert (a[1] == b[1]).all()


@pytest.mark.parametrize("th,mono,check_ori", [(7.0, False, True), (15.0, True, True), (7.0, False, False)])
def test_search_by_projection_frame(scene, th, mono, check_ori):
    O = Matcher("oracle")
    cur, last, Tc, Tl = two_view_scene(scene["kps"], scene["desc"], W, H, 21, K)
    if mono:
        cur = dict(cur); cur["u_right"] = None
    rng = np.random.default_rng(8)
    init_obs = np.where(rng.random(len(cur["kps"])) < 0.05, rng.integers(0, 2, len(cur["kps"])), -1).astype(np.int32)
    F = ob.FrameView(cur["kps"], cur["desc"], BOUNDS, cur.get("u_right"))
    M = ob.ORBmatcher(0.9, check_ori)
    for io in (None, init_obs):
        a = O.search_by_projection_frame(cur, last, Tc, Tl, K, 386.1448, scene["scale"], BOUNDS, th, mono, 0.9, check_ori, io)
        b = M.SearchByProjectionFrame(F, last, Tc, Tl, K, 386.1448, scene["scale"], th, mono, io)
        assert a[0] == b[0] and a[0] > 200
        assert (a[1] == b[1]).all()


@pytest.mark.parametrize("nnratio,check_ori,window", [(0.9, True, 100), (0.6, True, 50), (0.9, False, 100)])
def test_search_for_initialization(scene, nnratio, check_ori, window):
    O = Matcher("oracle")
    prev = np.stack([scene["kps"]["x"], scene["kps"]["y"]], 1)
    a = O.search_for_initialization(scene["kps"], scene["desc"], scene["k2"], scene["d2"], scene["scale"], BOUNDS, prev, window, nnratio, check_ori)
    F1 = ob.FrameView(scene["kps"], scene["desc"], BOUNDS)
    F2 = ob.FrameView(scene["k2"], scene["d2"], BOUNDS)
    b = ob.ORBmatcher(nnratio, check_ori).SearchForInitialization(F1, F2, prev, window)
    assert a[0] == b[0] and a[0] > 50
    assert (a[1] == b[1]).all() and (a[2] == b[2]).all()


def test_empty_and_degenerate_inputs(scene):
    M = ob.ORBmatcher(0.9, True)
    F0 = ob.FrameView(np.zeros(0, KP_DTYPE), np.zeros((0, 32), np.uint8), BOUNDS)
    F2 = ob.FrameView(scene["k2"], scene["d2"], BOUNDS)
    assert M.SearchForInitialization(F0, F2, np.zeros((0, 2), np.float32), 100)[0] == 0
    n, m12, _ = M.SearchForInitialization(F2, F0, np.zeros((len(scene["k2"]), 2), np.float32), 100)
    assert n == 0 and (m12 == -1).all()
    q = projected_queries(scene["k2"], scene["d2"], 50, 5)
    q["in_view"][:] = 0                                   # nothing in view: no matches, nothing attached
    n, a = M.SearchByProjection(F2, scene["scale"], q, 3.0)
    assert n == 0 and (a == -1).all()


def test_stereo_pair_end_to_end():
    """BASELINE.json configs[1]: KITTI-shape stereo pair, 2000 features per image, left/right extraction
    + stereo row matching; everything against the oracle (which equals the reference's stereo Frame)."""
    left, right = stereo_pair(W, H, 2)
    GL, GR = ob.ORBextractor(NF, 1.2, 8, 20, 7), ob.ORBextractor(NF, 1.2, 8, 20, 7)
    kl, dl = GL(left)
    kr, dr = GR(right)
    OL, OR = OracleExtractor(NF), OracleExtractor(NF)
    nl, okl, odl = OL.extract(left)
    nr, okr, odr = OR.extract(right)
    assert nl == len(kl) and nr == len(kr) and (dl == odl).all() and (dr == odr).all()
    bf, fx = 386.1448, float(K[0])
    ur, dep, nm = ob.stereo_matches(GL, GR, kl, dl, kr, dr, bf, fx)
    our, odep = np.zeros(nl, np.float32), np.zeros(nl, np.float32)
    import ctypes as C
    O = oracle()
    O.orbo_stereo_matches.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                      C.c_float, C.c_float, C.c_void_p, C.c_void_p]
    nd = O.orbo_stereo_matches(OL.h, OR.h, nl, okl.ctypes.data, odl.ctypes.data, nr, okr.ctypes.data, odr.ctypes.data, bf, fx,
                               our.ctypes.data, odep.ctypes.data)
    assert nm == nd and nd > 200
    assert (ur.view(np.uint32) == our.view(np.uint32)).all() and (dep.view(np.uint32) == odep.view(np.uint32)).all()
    for x in (GL, GR, OL, OR):
        x.close()


def _oracle_stereo(left, right, nf, bf, fx):
    import ctypes as C
    OL, OR = OracleExtractor(nf), OracleExtractor(nf)
    nl, okl, odl = OL.extract(left)
    nr, okr, odr = OR.extract(right)
    our, odep = np.zeros(nl, np.float32), np.zeros(nl, np.float32)
    O = oracle()
    O.orbo_stereo_matches.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                      C.c_float, C.c_float, C.c_void_p, C.c_void_p]
    nd = O.orbo_stereo_matches(OL.h, OR.h, nl, okl.ctypes.data, odl.ctypes.data, nr, okr.ctypes.data, odr.ctypes.data, bf, fx,
                               our.ctypes.data, odep.ctypes.data)
    OL.close(); OR.close()
    return (nl, okl, odl), (nr, okr, odr), our, odep, nd


@pytest.mark.parametrize("w,h,nf,pairs,chunk", [(1241, 376, 2000, 3, 64), (640, 480, 1000, 5, 4), (752, 480, 1200, 2, 3)])
def test_stereo_batch_device_resident(w, h, nf, pairs, chunk):
    """orbx_extract_stereo_batch: L/R extraction + ComputeStereoMatches without leaving the GPU, pairs cut into
    chunks (a pair never straddles two); every pair equals the oracle's stereo Frame bit for bit."""
    bf, fx = 386.1448, float(K[0])
    frames = np.stack([im for p in range(pairs) for im in stereo_pair(w, h, 40 + p)])
    G = ob.ORBextractor(nf, 1.2, 8, 20, 7)
    G.set_chunk(chunk)
    kps, desc, n, ur, dep, ns = G.extract_stereo_batch(frames, bf, fx)
    for p in range(pairs):
        (nl, okl, odl), (nr, okr, odr), our, odep, nd = _oracle_stereo(frames[2 * p], frames[2 * p + 1], nf, bf, fx)
        assert n[2 * p] == nl and n[2 * p + 1] == nr
        assert (kps[2 * p, :nl] == okl).all() and (desc[2 * p, :nl] == odl).all() and (desc[2 * p + 1, :nr] == odr).all()
        assert ns[p] == nd and nd > 50
        assert (ur[p, :nl].view(np.uint32) == our.view(np.uint32)).all(), p
        assert (dep[p, :nl].view(np.uint32) == odep.view(np.uint32)).all(), p
    G.close()


def test_stereo_batch_device_pointers_async():
    """The same through the _async entry point with every buffer in device memory (torch tensors)."""
    import torch
    w, h, nf, pairs = 640, 480, 1000, 4
    bf, fx = 386.1448, float(K[0])
    frames = np.stack([im for p in range(pairs) for im in stereo_pair(w, h, 60 + p)])
    G = ob.ORBextractor(nf, 1.2, 8, 20, 7)
    want = G.extract_stereo_batch(frames, bf, fx)
    cap = want[0].shape[1]
    dev = torch.device("cuda:0")
    d_img = torch.from_numpy(frames).to(dev)
    d_kps = torch.zeros((2 * pairs, cap, 28), dtype=torch.uint8, device=dev)
    d_desc = torch.zeros((2 * pairs, cap, 32), dtype=torch.uint8, device=dev)
    d_n = torch.zeros(2 * pairs, dtype=torch.int32, device=dev)
    d_ur = torch.full((pairs, cap), -1.0, dtype=torch.float32, device=dev)
    d_dep = torch.full((pairs, cap), -1.0, dtype=torch.float32, device=dev)
    d_ns = torch.zeros(pairs, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    G.extract_stereo_batch_raw(d_img, h * w, pairs, w, h, w, d_kps, d_desc, cap, d_n, bf, fx, d_ur, d_dep, d_ns, asynchronous=True)
    G.sync()
    n = d_n.cpu().numpy()
    assert (n == want[2]).all() and (d_ns.cpu().numpy() == want[5]).all()
    ur, dep = d_ur.cpu().numpy(), d_dep.cpu().numpy()
    for p in range(pairs):
        nl = n[2 * p]
        assert (ur[p, :nl].view(np.uint32) == want[3][p, :nl].view(np.uint32)).all()
        assert (dep[p, :nl].view(np.uint32) == want[4][p, :nl].view(np.uint32)).all()
    assert (d_desc.cpu().numpy() == want[1]).all()
    G.close()


@pytest.mark.parametrize("th,orb_dist,check_ori", [(10.0, 100, True), (3.0, 64, True), (10.0, 100, False)])
def test_relocalisation_window_search(scene, th, orb_dist, check_ori):
    """src/ORBmatcher.cc:303-431 through the generic device primitive: the host projects (oracle's exported
    queries = what the C++ forwarder computes), the GPU does windows, distances, ordered claims, histogram."""
    from reloc_lib import reloc_scene, run_reloc
    s = reloc_scene(scene["kps"], scene["desc"], W, H, 31, K)
    nm, assign, q = run_reloc("oracle", s, scene["scale"], BOUNDS, K, th, orb_dist, check_ori, want_queries=True)
    F = ob.FrameView(s["cur_kps"], s["cur_desc"], BOUNDS)
    init_obs = np.where(s["cur_taken"] > 0, 1, -1).astype(np.int32)
    g_nm, g_assign = ob.window_search_best(F, q["uvr"], q["minl"], q["maxl"], s["mp_desc"], orb_dist, check_ori, q_angle=s["kf_angle"],
                                           init_obs=init_obs, valid=q["valid"])
    assert g_nm == nm and nm > 100 and (g_assign == assign).all()


@pytest.mark.parametrize("th,seed", [(10, 41), (4, 42)])
def test_loop_closing_window_search(scene, th, seed):
    """src/ORBmatcher.cc:434-549 through the generic device primitive: the host decomposes the Sim3 and projects (the
    oracle's exported queries = what the C++ forwarder computes); the GPU does windows, distances and ordered claims."""
    from sim3_lib import run_sim3, sim3_scene
    s = sim3_scene(scene["k2"], scene["d2"], W, H, seed, K)
    nm, assign, q = run_sim3("oracle", scene["k2"], scene["d2"], s, scene["scale"], BOUNDS, K, th, want_queries=True)
    F = ob.FrameView(scene["k2"], scene["d2"], BOUNDS)
    init_obs = np.where(s["matched"] != -1, 1, -1).astype(np.int32)
    g_nm, g_assign = ob.window_search_best(F, q["uvr"], q["minl"], q["maxl"], s["mp_desc"], 50, False, init_obs=init_obs, valid=q["valid"])
    g_assign = np.where(g_assign == -2, s["matched"], g_assign)
    assert g_nm == nm and nm > 100 and (g_assign == assign).all()


def test_window_search_generic_random_queries(scene):
    """Random windows / level ranges / right-image gates / pre-attached points against the oracle primitive."""
    import ctypes as C
    rng = np.random.default_rng(12)
    k2, d2 = scene["k2"], scene["d2"]
    n, nq = len(k2), 1500
    tgt = rng.integers(0, n, nq)
    uvr = np.stack([k2["x"][tgt] + rng.normal(0, 4, nq), k2["y"][tgt] + rng.normal(0, 4, nq), rng.choice([3.0, 8.0, 20.0, 60.0], nq)], 1).astype(np.float32)
    minl = rng.integers(-1, 4, nq).astype(np.int32)
    maxl = np.where(rng.random(nq) < 0.3, -1, minl + rng.integers(0, 3, nq)).astype(np.int32)
    u_right = np.where(rng.random(n) < 0.5, k2["x"] - 25 * rng.random(n), -1).astype(np.float32)
    ur = (uvr[:, 0] - 12 + rng.normal(0, 6, nq)).astype(np.float32)
    er_max = rng.choice([2.0, 6.0, 15.0], nq).astype(np.float32)
    valid = (rng.random(nq) < 0.9).astype(np.uint8)
    qd = d2[tgt].copy(); qd[:, :4] ^= rng.integers(0, 256, (nq, 4), dtype=np.uint8)
    q_angle = (rng.random(nq) * 360).astype(np.float32)
    q_obs = (rng.random(nq) < 0.8).astype(np.int32)
    init_obs = np.where(rng.random(n) < 0.1, rng.integers(0, 2, n), -1).astype(np.int32)
    O = oracle()
    vp, ci, cf = C.c_void_p, C.c_int, C.c_float
    O.orbo_window_search_best.argtypes = [ci, vp, vp, vp] + [cf] * 4 + [ci] + [vp] * 11 + [ci, ci]
    for th_accept, check_ori in ((100, True), (50, False)):
        ref_assign = np.zeros(n, np.int32)
        p = lambda a: a.ctypes.data
        nm = O.orbo_window_search_best(n, p(k2), p(d2), p(u_right), *BOUNDS, nq, p(uvr), p(minl), p(maxl), p(ur), p(er_max), p(valid), p(qd),
                                       p(q_angle), p(q_obs), p(init_obs), p(ref_assign), th_accept, int(check_ori))
        F = ob.FrameView(k2, d2, BOUNDS, u_right)
        g_nm, g_assign = ob.window_search_best(F, uvr, minl, maxl, qd, th_accept, check_ori, q_angle=q_angle, q_obs=q_obs, init_obs=init_obs,
                                               ur=ur, er_max=er_max, valid=valid)
        assert g_nm == nm and nm > 100 and (g_assign == ref_assign).all()   # random query angles: the histogram prunes most


def test_very_large_query_set_takes_the_list_path(scene):
    """More queries than the block-per-problem kernel holds in shared memory: the entry point falls back to the
    candidate-list kernels with the in-order resolve; same results."""
    O = Matcher("oracle")
    k, d = scene["k2"][:600].copy(), scene["d2"][:600].copy()
    q = projected_queries(k, d, 30000, 77)
    a = O.search_by_projection_points(k, d, None, scene["scale"], BOUNDS, q, 3.0, 0.9, None)
    b = ob.ORBmatcher(0.9, True).SearchByProjection(ob.FrameView(k, d, BOUNDS), scene["scale"], q, 3.0)
    assert a[0] == b[0] and a[0] > 100 and (a[1] == b[1]).all()


def test_matcher_entry_points_are_reentrant(scene):
    """The reference's tracking, mapping and loop-closing threads construct matchers concurrently: the
    single-problem entry points keep their staging memory per thread.  Four threads, different problems, 20 calls each."""
    import threading
    O = Matcher("oracle")
    F = ob.FrameView(scene["k2"], scene["d2"], BOUNDS)
    jobs = []
    for t in range(4):
        q = projected_queries(scene["k2"], scene["d2"], 1200 + 100 * t, 50 + t)
        jobs.append((q, O.search_by_projection_points(scene["k2"], scene["d2"], None, scene["scale"], BOUNDS, q, 3.0, 0.9, None)))
    errors = []

    def work(q, want):
        try:
            M = ob.ORBmatcher(0.9, True)
            for _ in range(20):
                nm, a = M.SearchByProjection(F, scene["scale"], q, 3.0)
                if nm != want[0] or not (a == want[1]).all():
                    errors.append(nm)
        except Exception as e:                       # an exception in a thread must fail the test too
            errors.append(repr(e))

    threads = [threading.Thread(target=work, args=j) for j in jobs]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:3]


def test_very_large_best_only_query_set_takes_the_list_path(scene):
    """orbm_window_search_best with more queries than the block-per-problem kernel holds: candidate lists + in-order
    resolve (k_window_candidates / k_resolve_frame), same results as the oracle."""
    import ctypes as C
    rng = np.random.default_rng(5)
    k, d = scene["k2"][:500].copy(), scene["d2"][:500].copy()
    n, nq = len(k), 30000
    tgt = rng.integers(0, n, nq)
    uvr = np.stack([k["x"][tgt] + rng.normal(0, 3, nq), k["y"][tgt] + rng.normal(0, 3, nq), rng.choice([4.0, 12.0], nq)], 1).astype(np.float32)
    minl = np.full(nq, -1, np.int32); maxl = np.full(nq, -1, np.int32)
    qd = d[tgt].copy(); qd[:, :2] ^= rng.integers(0, 256, (nq, 2), dtype=np.uint8)
    q_angle = ((k["angle"][tgt] + rng.normal(0, 5, nq)) % 360).astype(np.float32)
    q_obs = (rng.random(nq) < 0.5).astype(np.int32)
    O = oracle()
    vp, ci, cf = C.c_void_p, C.c_int, C.c_float
    O.orbo_window_search_best.argtypes = [ci, vp, vp, vp] + [cf] * 4 + [ci] + [vp] * 11 + [ci, ci]
    p = lambda a: None if a is None else a.ctypes.data
    ref_assign = np.zeros(n, np.int32)
    nm = O.orbo_window_search_best(n, p(k), p(d), None, *BOUNDS, nq, p(uvr), p(minl), p(maxl), None, None, None, p(qd), p(q_angle), p(q_obs), None,
                                   p(ref_assign), 100, 1)
    g_nm, g_assign = ob.window_search_best(ob.FrameView(k, d, BOUNDS), uvr, minl, maxl, qd, 100, True, q_angle=q_angle, q_obs=q_obs)
    assert g_nm == nm and nm > 100 and (g_assign == ref_assign).all()


def test_three_maxima_on_random_histograms():
    """a16, ORBmatcher::ComputeThreeMaxima (src/ORBmatcher.cc:1663-1707) as the matcher kernels run it (one shared device
    function): random bin counts incl. ties, empty histograms and the 10 % cut, against a direct restatement."""
    import ctypes as C
    rng = np.random.default_rng(16)
    n = 4000
    sizes = rng.integers(0, 40, (n, 30)).astype(np.int32)
    sizes[:200] = rng.integers(0, 3, (200, 30))                  # many ties
    sizes[200:300] = 0                                          # empty
    sizes[300:600] = (rng.random((300, 30)) < 0.1) * rng.integers(1, 200, (300, 30))   # sparse: the 10 % rule bites
    sizes[600:700, :] = 7                                       # all equal
    ind = np.zeros((n, 3), np.int32)
    L = ob.lib()
    L.orbm_debug_three_maxima.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    assert L.orbm_debug_three_maxima(sizes.ctypes.data, n, ind.ctypes.data, 0) == 0

    def three(s):
        m1 = m2 = m3 = 0
        i1 = i2 = i3 = -1
        for i, v in enumerate(s):
            if v > m1:
                m3, m2, m1, i3, i2, i1 = m2, m1, v, i2, i1, i
            elif v > m2:
                m3, m2, i3, i2 = m2, v, i2, i
            elif v > m3:
                m3, i3 = v, i
        if np.float32(m2) < np.float32(0.1) * np.float32(m1):
            i2 = i3 = -1
        elif np.float32(m3) < np.float32(0.1) * np.float32(m1):
            i3 = -1
        return i1, i2, i3

    want = np.array([three(s.tolist()) for s in sizes], np.int32)
    assert (ind == want).all()


def test_device_grid_returns_candidates_in_reference_order(scene):
    """a17, Frame::AssignFeaturesToGrid + GetFeaturesInArea (src/Frame.cc:243-259, 348-422) on the device grid: the
    candidate LIST of a window -- not just the match it leads to -- equals the oracle's, element by element (grid
    column, grid row, insertion order), with and without level bounds, incl. windows hanging over the image border."""
    import ctypes as C
    kps, desc = scene["k2"], scene["d2"]
    rng = np.random.default_rng(17)
    nq = 600
    xyr = np.stack([rng.uniform(-30, W + 30, nq), rng.uniform(-30, H + 30, nq), rng.choice([3.0, 7.5, 15.0, 40.0, 100.0], nq)], 1).astype(np.float32)
    lo = rng.choice([-1, 0, 1, 3], nq).astype(np.int32)
    hi = np.where(rng.random(nq) < 0.5, -1, lo + rng.integers(0, 4, nq)).astype(np.int32)
    hi = np.where((lo < 0) & (hi >= 0), hi + 1, hi).astype(np.int32)
    cap = len(kps)
    idx = np.zeros((nq, cap), np.int32)
    cnt = np.zeros(nq, np.int32)
    F = ob.FrameView(kps, desc, BOUNDS)
    L = ob.lib()
    L.orbm_debug_features_in_area.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    st = F.struct()
    assert L.orbm_debug_features_in_area(C.byref(st), nq, xyr.ctypes.data, lo.ctypes.data, hi.ctypes.data, cap, idx.ctypes.data, cnt.ctypes.data, 0) == 0
    M = Matcher("oracle")
    nonempty = 0
    for q in range(nq):
        want = M.features_in_area(kps, scene["scale"], BOUNDS, float(xyr[q, 0]), float(xyr[q, 1]), float(xyr[q, 2]), int(lo[q]), int(hi[q]))
        assert cnt[q] == len(want) and (idx[q, :cnt[q]] == want).all(), q
        nonempty += len(want) > 1
    assert nonempty > 200


def test_resident_frame_gives_the_same_results_as_host_frames(scene):
    """orbm_frame_upload / orbm_frame_view: a Frame uploaded once serves every search of that Frame (Tracking runs two to
    four per Frame) with results identical to the host-pointer path and to the CPU checker; mono views of a stereo frame
    (u_right nulled in the caller's copy of the view), an empty frame and repeated use of one handle are covered."""
    O = Matcher("oracle")
    rng = np.random.default_rng(6)
    n = len(scene["k2"])
    ur = np.where(rng.random(n) < 0.5, scene["k2"]["x"] - 20 * rng.random(n), -1).astype(np.float32)
    R = ob.ResidentFrame(scene["k2"], scene["d2"], BOUNDS, ur)
    H = ob.FrameView(scene["k2"], scene["d2"], BOUNDS, ur)
    M = ob.ORBmatcher(0.8, True)
    # 1. SearchByProjection(Frame, MapPoints), twice on the same handle with different query sets
    for seed, th in ((5, 3.0), (9, 5.0)):
        q = projected_queries(scene["k2"], scene["d2"], 2000, seed)
        a = O.search_by_projection_points(scene["k2"], scene["d2"], ur, scene["scale"], BOUNDS, q, th, 0.8, None)
        b = M.SearchByProjection(R, scene["scale"], q, th)
        c = M.SearchByProjection(H, scene["scale"], q, th)
        assert a[0] == b[0] == c[0] and a[0] > 300 and (a[1] == b[1]).all() and (b[1] == c[1]).all()
    # 2. the best-candidate-only search on the same handle (relocalisation / loop closing shape: no right coordinates)
    k, d = scene["k2"], scene["d2"]
    nq = 1200
    pick = rng.integers(0, n, nq)
    uvr = np.stack([k["x"][pick] + rng.normal(0, 2, nq), k["y"][pick] + rng.normal(0, 2, nq), np.full(nq, 12.0)], 1).astype(np.float32)
    minl = np.maximum(k["octave"][pick] - 1, 0).astype(np.int32)
    maxl = (k["octave"][pick] + 1).astype(np.int32)
    qd = d[pick].copy()
    g1 = ob.window_search_best(R, uvr, minl, maxl, qd, 100, False)
    g2 = ob.window_search_best(H, uvr, minl, maxl, qd, 100, False)
    assert g1[0] == g2[0] and g1[0] > 200 and (g1[1] == g2[1]).all()
    # 3. SearchForInitialization with both frames resident
    prev = np.stack([scene["kps"]["x"], scene["kps"]["y"]], 1)
    R1 = ob.ResidentFrame(scene["kps"], scene["desc"], BOUNDS)
    a = O.search_for_initialization(scene["kps"], scene["desc"], scene["k2"], scene["d2"], scene["scale"], BOUNDS, prev, 100, 0.9, True)
    b = ob.ORBmatcher(0.9, True).SearchForInitialization(R1, ob.ResidentFrame(scene["k2"], scene["d2"], BOUNDS), prev, 100)
    assert a[0] == b[0] and (a[1] == b[1]).all() and (a[2] == b[2]).all()
    # 4. an empty frame can be uploaded and searched (0 matches, as the reference on a black image)
    R0 = ob.ResidentFrame(np.zeros(0, KP_DTYPE), np.zeros((0, 32), np.uint8), BOUNDS)
    q = projected_queries(scene["k2"], scene["d2"], 50, 5)
    assert M.SearchByProjection(R0, scene["scale"], q, 3.0)[0] == 0
    for r in (R, R1, R0):
        r.close()


def test_crowded_grid_cell(scene):
    """The device grid is a counting sort with an insertion sort inside each cell; a frame that crowds more than 32 keypoints
    into one cell takes the bitonic-sort path instead.  Both give the reference's candidate order: 150 keypoints moved into
    one 19 x 8 px cell, queries aimed at them."""
    O = Matcher("oracle")
    k, d = scene["k2"].copy(), scene["d2"]
    rng = np.random.default_rng(12)
    n = len(k)
    move = rng.choice(n, 150, replace=False)
    k["x"][move] = 400.0 + rng.random(150).astype(np.float32) * 15.0
    k["y"][move] = 200.0 + rng.random(150).astype(np.float32) * 6.0
    q = projected_queries(k, d, 2000, 5)
    F = ob.FrameView(k, d, BOUNDS)
    for th in (1.0, 3.0):
        a = O.search_by_projection_points(k, d, None, scene["scale"], BOUNDS, q, th, 0.9, None)
        b = ob.ORBmatcher(0.9, True).SearchByProjection(F, scene["scale"], q, th)
        assert a[0] == b[0] and a[0] > 200 and (a[1] == b[1]).all()
