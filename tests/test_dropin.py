"""The drop-in C++ boundary: orb_slam2_chinesenotes_b200/host/ORBextractor.{h,cc} (same class interface
as the reference's include/ORBextractor.h) and the ORBmatcher forwarders of host/ORBmatcher_b200.hpp.
CPU: they compile and link against liborb_b200.so (OpenCV is played by oracle/cvshim in this image), and
the forwarders also compile against the reference's REAL Frame.h when the checkout is present.
GPU: the C++ class is called the way Frame::ExtractORB calls it and checked against the oracle."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle_lib import KP_DTYPE, OracleExtractor
from synth import synth_frame

HERE = os.path.dirname(os.path.abspath(__file__))
CPP = os.path.join(HERE, "cpp")


@pytest.fixture(scope="module")
def dropin(built_lib):
    r = subprocess.run(["make", "-C", CPP, "all", "matcher"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    L = C.CDLL(os.path.join(CPP, "_build", "libdropin.so"))
    L.dropin_create.restype = C.c_void_p
    L.dropin_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
    L.dropin_destroy.argtypes = [C.c_void_p]
    L.dropin_accessors.argtypes = [C.c_void_p] * 6
    L.dropin_call.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int]
    L.dropin_pyramid.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    return L


def test_dropin_builds_and_exports(dropin):
    for s in ("dropin_create", "dropin_call", "dropin_pyramid", "dropin_accessors"):
        assert hasattr(dropin, s)
    M = C.CDLL(os.path.join(CPP, "_build", "libmatcher_fwd.so"))
    assert M.matcher_forwarders_instantiate(0) == 0


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference checkout not present")
def test_matcher_forwarders_compile_against_reference_headers():
    r = subprocess.run(["make", "-C", CPP, "matcher_ref"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_dropin_fails_loudly_without_gpu(dropin):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert not dropin.dropin_create(1000, 1.2, 8, 20, 7)     # constructor threw: no CPU fallback


@pytest.mark.gpu
def test_dropin_class_matches_oracle(dropin):
    nf, w, h = 1000, 640, 480
    ex = dropin.dropin_create(nf, 1.2, 8, 20, 7)
    assert ex
    O = OracleExtractor(nf)
    sc, inv, s2, is2 = (np.zeros(8, np.float32) for _ in range(4))
    sf = C.c_float()
    assert dropin.dropin_accessors(ex, sc.ctypes.data, inv.ctypes.data, s2.ctypes.data, is2.ctypes.data, C.byref(sf)) == 8
    t = O.tables()
    assert (sc == t["scale"]).all() and (inv == t["inv_scale"]).all() and (s2 == t["sigma2"]).all() and (is2 == t["inv_sigma2"]).all()
    assert np.float32(sf.value) == np.float32(1.2)
    for seed in (1, 7):
        img = synth_frame(w, h, seed)
        n_o, k_o, d_o = O.extract(img)
        kps, desc = np.zeros(2 * nf, KP_DTYPE), np.zeros((2 * nf, 32), np.uint8)
        n = dropin.dropin_call(ex, img.ctypes.data, w, h, w, kps.ctypes.data, desc.ctypes.data, 2 * nf)
        assert n == n_o and all((kps[:n][f] == k_o[f]).all() for f in KP_DTYPE.names) and (desc[:n] == d_o).all()
        for l in range(8):                                    # public mvImagePyramid incl. the 19-px border
            ref = O.pyramid(l, True)
            lw, lh = C.c_int(), C.c_int()
            out = np.zeros_like(ref)
            assert dropin.dropin_pyramid(ex, l, 19, out.ctypes.data, out.strides[0], C.byref(lw), C.byref(lh)) == 0
            assert (lw.value + 38, lh.value + 38) == (ref.shape[1], ref.shape[0]) and (out == ref).all()
    kps, desc = np.zeros(8, KP_DTYPE), np.zeros((8, 32), np.uint8)
    assert dropin.dropin_call(ex, None, 0, 0, 0, kps.ctypes.data, desc.ctypes.data, 8) == -1   # empty image: outputs untouched
    dropin.dropin_destroy(ex)
    O.close()


@pytest.mark.gpu
def test_matcher_forwarders_run_like_the_patched_reference(dropin):
    """host/ORBmatcher_b200.hpp at run time: Frame / KeyFrame / MapPoint objects (stand-in types with the reference's
    member names) are filled from arrays, handed to the forwarders the way the patched ORBmatcher.cc, Tracking.cc and
    MapPoint.cc would, and what the forwarders write back into the objects is compared with the oracle."""
    from bow_lib import bow_scene, search_by_bow
    from mappoint_lib import descriptor_groups, distinctive, frustum_scene, is_in_frustum
    from matcher_lib import Matcher, extract_frame, perturbed_frame, projected_queries
    M = C.CDLL(os.path.join(CPP, "_build", "libmatcher_fwd.so"))
    vp, ci, cf = C.c_void_p, C.c_int, C.c_float
    p = lambda a: None if a is None else a.ctypes.data
    W, H = 1241, 376
    bounds = (0.0, float(W), 0.0, float(H))
    kps, desc, scale = extract_frame(W, H, 2000, 2)
    k2, d2, _ = perturbed_frame(kps, desc, W, H, 11)
    # SearchByProjection(Frame, MapPoints)
    q = projected_queries(k2, d2, 1500, 5)
    rng = np.random.default_rng(6)
    n = len(k2)
    ur = np.where(rng.random(n) < 0.5, k2["x"] - 20 * rng.random(n), -1).astype(np.float32)
    init = np.where(rng.random(n) < 0.05, rng.integers(0, 1500, n), -1).astype(np.int32)
    M.fwd_search_by_projection_points.argtypes = [ci, vp, vp, vp, vp, ci] + [cf] * 4 + [ci] + [vp] * 9 + [cf, cf]
    for ia in (None, init):
        out = np.zeros(n, np.int32)
        nm = M.fwd_search_by_projection_points(n, p(k2), p(d2), p(ur), p(scale), len(scale), *bounds, 1500, p(q["proj"]), p(q["level"]),
                                               p(q["view_cos"]), p(q["in_view"]), p(q["bad"]), p(q["obs"]), p(q["desc"]), p(ia), p(out), 3.0, 0.8)
        want = Matcher("oracle").search_by_projection_points(k2, d2, ur, scale, bounds, q, 3.0, 0.8, ia)
        assert nm == want[0] and nm > 300 and (out == want[1]).all()
    # a frame without keypoints (a black image): the reference returns 0 matches and tracking goes on to "lost"; the
    # forwarder must do the same, not throw (the C ABI accepts assign_out = NULL when the frame is empty)
    nm = M.fwd_search_by_projection_points(0, None, None, None, p(scale), len(scale), *bounds, 1500, p(q["proj"]), p(q["level"]),
                                           p(q["view_cos"]), p(q["in_view"]), p(q["bad"]), p(q["obs"]), p(q["desc"]), None, None, 3.0, 0.8)
    assert nm == 0
    # SearchByBoW x2
    s = bow_scene(kps, desc, 8, n2=1700)
    (id1, off1, f1), (id2, off2, f2) = s["fv1"], s["fv2"]
    M.fwd_search_by_bow.argtypes = [ci] + [ci, vp, vp, vp, vp, ci, vp, vp, vp] * 2 + [cf, ci, vp]
    for kf_kf in (0, 1):
        m12 = np.zeros(len(s["k1"]), np.int32)
        nm = M.fwd_search_by_bow(kf_kf, len(s["k1"]), p(s["k1"]), p(s["d1"]), p(s["valid1"]), p(s["bad1"]), len(id1), p(id1), p(off1), p(f1),
                                 len(s["k2"]), p(s["k2"]), p(s["d2"]), p(s["valid2"]), p(s["bad2"]), len(id2), p(id2), p(off2), p(f2), 0.75, 1, p(m12))
        want = search_by_bow("oracle", s, 0.75, True, bool(kf_kf))
        assert nm == want[0] and nm > 100 and (m12 == want[1]).all()
    # isInFrustum loop
    fs = frustum_scene(3)
    nf = len(fs["xyz"])
    iv, proj, lv, vc = np.zeros(nf, np.uint8), np.zeros((nf, 3), np.float32), np.zeros(nf, np.int32), np.zeros(nf, np.float32)
    M.fwd_is_in_frustum.argtypes = [vp, vp, cf] + [cf] * 4 + [cf, ci, cf, ci] + [vp] * 8
    cnt = M.fwd_is_in_frustum(p(fs["Tcw"]), p(fs["K"]), float(fs["bf"]), *fs["bounds"], 1.2, 8, 0.5, nf, p(fs["xyz"]), p(fs["normal"]),
                              p(fs["max_d"]), p(fs["min_d"]), p(iv), p(proj), p(lv), p(vc))
    wiv, wproj, wlv, wvc = is_in_frustum("oracle", fs)
    m = wiv.astype(bool)
    assert cnt == wiv.sum() and (iv == wiv).all() and (lv[m] == wlv[m]).all()
    assert (proj[m].view(np.uint32) == wproj[m].view(np.uint32)).all() and (vc[m].view(np.uint32) == wvc[m].view(np.uint32)).all()
    # ComputeDistinctiveDescriptors' median search
    M.fwd_distinctive.argtypes = [vp, ci]
    for d, _ in descriptor_groups(2, sizes=(1, 5, 33, 120)):
        assert M.fwd_distinctive(p(d), len(d)) == distinctive("oracle", d)[1]


@pytest.mark.gpu
def test_matcher_forwarders_keep_a_frame_on_the_device(dropin):
    """The forwarders upload a Frame once (b200::ResidentFrames, keyed by Frame::mnId) and every further search of that
    Frame reuses the device copy: one upload for five SearchByProjection calls, identical results, and the CPU checker's."""
    from matcher_lib import Matcher, extract_frame, perturbed_frame, projected_queries
    M = C.CDLL(os.path.join(CPP, "_build", "libmatcher_fwd.so"))
    vp, ci, cf = C.c_void_p, C.c_int, C.c_float
    p = lambda a: a.ctypes.data
    W, H = 1241, 376
    bounds = (0.0, float(W), 0.0, float(H))
    kps, desc, scale = extract_frame(W, H, 2000, 2)
    k2, d2, _ = perturbed_frame(kps, desc, W, H, 11)
    q = projected_queries(k2, d2, 1500, 5)
    q["in_view"][:] = 1; q["bad"][:] = 0; q["obs"][:] = 1
    nm, secs = C.c_int(), (C.c_double * 2)()
    M.fwd_resident_reuse.argtypes = [ci, ci, vp, vp, vp, ci] + [cf] * 4 + [ci] + [vp] * 6
    uploads = M.fwd_resident_reuse(5, len(k2), p(k2), p(d2), p(scale), len(scale), *bounds, 1500, p(q["proj"]), p(q["level"]),
                                   p(q["view_cos"]), p(q["desc"]), C.addressof(nm), C.addressof(secs))
    want = Matcher("oracle").search_by_projection_points(k2, d2, np.full(len(k2), -1, np.float32), scale, bounds, q, 3.0, 0.8, None)
    assert uploads == 1 and nm.value == want[0] and nm.value > 300
    print(f"SearchByProjection through the forwarder: first call {secs[0] * 1e3:.3f} ms, later calls {secs[1] * 1e3:.3f} ms")


@pytest.mark.gpu
def test_resident_frame_cache_bookkeeping(dropin):
    """b200::ResidentFrames: capacity 8 with least-recently-used eviction, one entry per frame, an entry without right
    coordinates superseded when they are needed."""
    from matcher_lib import extract_frame
    M = C.CDLL(os.path.join(CPP, "_build", "libmatcher_fwd.so"))
    kps, desc, _ = extract_frame(640, 480, 500, 3)
    M.fwd_resident_cache_lru.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    frames = 10
    r = M.fwd_resident_cache_lru(frames, len(kps), kps.ctypes.data, desc.ctypes.data)
    # 10 uploads, + 1 for the evicted first frame, + 0 for the last frame (a hit), + 1 when the last frame (cached without
    # right coordinates: index 9 is odd) is asked for with them; 8 entries held
    assert r == (frames + 1 + 1) * 100 + 8, r


@pytest.mark.gpu
def test_fuse_and_sim3_forwarders_run_like_the_patched_reference(dropin):
    """b200::Fuse (both overloads) and b200::SearchBySim3 at run time: KeyFrame / MapPoint stand-ins (the reference's member
    names, its Replace / AddObservation bookkeeping) are filled from arrays and handed to the forwarders the way the patched
    ORBmatcher.cc would; what they leave in the objects is compared with the CPU checker's sequential loop."""
    from fuse_lib import fuse_scene, run_fuse, run_search_by_sim3, same_state, sim3_pair_scene
    from matcher_lib import extract_frame, perturbed_frame
    from test_fuse_oracle import BF, BOUNDS, FUSE_CASES, K, SIM3_CASES
    M = C.CDLL(os.path.join(CPP, "_build", "libmatcher_fwd.so"))
    vp, ci, cf = C.c_void_p, C.c_int, C.c_float
    p = lambda a: None if a is None else a.ctypes.data
    W, H = 1241, 376
    kps, desc, scale = extract_frame(W, H, 2000, 2)
    k2, d2, _ = perturbed_frame(kps, desc, W, H, 11, shift=4, kmax=40)
    inv_sigma2 = (np.float32(1.0) / (scale * scale)).astype(np.float32)
    Kf = np.asarray(K, np.float32)
    M.fwd_fuse.argtypes = [ci, ci, vp, vp, vp] + [cf] * 4 + [vp, vp, ci, vp, cf, vp, vp, vp, vp, ci] + [vp] * 10 + [ci, vp, vp, vp, cf]
    for th, seed, sim3, stereo in FUSE_CASES:
        s = fuse_scene(kps, desc, W, H, seed, K, BF, sim3)
        want_n, want, _ = run_fuse("oracle", kps, desc, s, scale, inv_sigma2, BOUNDS, K, BF, th, sim3, stereo)
        bad, nobs, kf_idx, kf_mp = s["bad"].copy(), s["nobs"].copy(), s["kf_idx"].copy(), s["kf_mp"].copy()
        replaced_by, replace = np.full(s["npts"], -1, np.int32), np.full(len(s["list"]), -1, np.int32)
        ur = s["u_right"] if stereo else None
        nf = M.fwd_fuse(int(sim3), len(kps), p(kps), p(desc), p(ur), *BOUNDS, p(scale), p(inv_sigma2), len(scale), p(Kf), BF, p(s["R"]), p(s["t"]),
                        p(s["Ow"]), p(s["Scw"]), s["npts"], p(bad), p(s["xyz"]), p(s["normal"]), p(s["mp_desc"]), p(s["level"]), p(s["min_dist"]),
                        p(s["max_dist"]), p(nobs), p(kf_idx), p(replaced_by), len(s["list"]), p(s["list"]), p(kf_mp), p(replace), th)
        got = dict(bad=bad, nobs=nobs, kf_idx=kf_idx, kf_mp=kf_mp)
        got["replace" if sim3 else "replaced_by"] = replace if sim3 else replaced_by
        assert nf == want_n and nf > 300 and same_state(got, want), (th, seed, sim3, stereo)
    M.fwd_search_by_sim3.argtypes = [ci, vp, vp, vp, ci, vp, vp, vp] + [cf] * 4 + [vp, ci] + [vp] * 5 + [cf, vp, vp, ci] + [vp] * 8 + [cf]
    for th, seed in SIM3_CASES:
        s = sim3_pair_scene(kps, desc, k2, d2, W, H, seed, K)
        want_n, want_m12, _ = run_search_by_sim3("oracle", kps, desc, k2, d2, s, scale, BOUNDS, K, th)
        m12 = s["m12"].copy()
        nf = M.fwd_search_by_sim3(len(kps), p(kps), p(desc), p(s["mp1"]), len(k2), p(k2), p(d2), p(s["mp2"]), *BOUNDS, p(scale), len(scale), p(Kf),
                                  p(s["R1"]), p(s["t1"]), p(s["R2"]), p(s["t2"]), s["s12"], p(s["R12"]), p(s["t12"]), s["npts"], p(s["bad"]),
                                  p(s["xyz"]), p(s["mp_desc"]), p(s["level"]), p(s["min_dist"]), p(s["max_dist"]), p(s["idx_in_kf2"]), p(m12), th)
        assert nf == want_n and nf > 200 and (m12 == want_m12).all()


@pytest.mark.gpu
def test_triangulation_forwarder_runs_like_the_patched_reference(dropin):
    """b200::SearchForTriangulation at run time: key-frame stand-ins filled from arrays, the epipole computed by the forwarder,
    vMatchedPairs compared with the CPU checker."""
    from matcher_lib import extract_frame
    from test_triang_oracle import CASES, K
    from triang_lib import search_for_triangulation, triang_scene
    M = C.CDLL(os.path.join(CPP, "_build", "libmatcher_fwd.so"))
    vp, ci = C.c_void_p, C.c_int
    p = lambda a: None if a is None else a.ctypes.data
    W, H = 1241, 376
    kps, desc, scale = extract_frame(W, H, 2000, 2)
    sigma2 = (scale * scale).astype(np.float32)
    Kf = np.asarray(K, np.float32)
    side = [ci, vp, vp, vp, vp, ci, vp, vp, vp]
    M.fwd_search_for_triangulation.argtypes = side * 2 + [vp, vp, vp, vp, vp, vp, ci, ci, ci, vp]
    for seed, only_stereo, check_ori, mono, n2 in CASES:
        s = triang_scene(kps, desc, W, H, seed, K, scale, n2)
        want = search_for_triangulation("oracle", s, K, scale, sigma2, only_stereo, check_ori, mono)
        (id1, off1, f1), (id2, off2, f2) = s["fv1"], s["fv2"]
        ur1, ur2 = (None, None) if mono else (s["ur1"], s["ur2"])
        out = np.zeros(len(s["k1"]), np.int32)
        nm = M.fwd_search_for_triangulation(len(s["k1"]), p(s["k1"]), p(s["d1"]), p(s["has1"]), p(ur1), len(id1), p(id1), p(off1), p(f1),
                                            len(s["k2"]), p(s["k2"]), p(s["d2"]), p(s["has2"]), p(ur2), len(id2), p(id2), p(off2), p(f2),
                                            p(s["F12"]), p(s["Cw"]), p(s["pose2"]), p(Kf), p(scale), p(sigma2), len(scale), int(only_stereo),
                                            int(check_ori), p(out))
        assert nm == want[0] and nm > 40 and (out == want[1]).all(), (seed, only_stereo, check_ori, mono)
