"""T1 for ORBmatcher::SearchForTriangulation (src/ORBmatcher.cc:1183-1361): the C restatement (oracle/orb_match_oracle.c)
against the reference's own unmodified ORBmatcher.cc (oracle/_ref, ref_bow_harness.cc) and against the results stored from
it in tests/golden/ref_triang.npz."""
import os

import numpy as np
import pytest

from matcher_lib import extract_frame
from oracle_lib import ref
from triang_lib import search_for_triangulation, triang_scene

needs_ref = pytest.mark.skipif(ref() is None, reason="oracle/_ref/liborbref.so not built (needs /root/reference)")
W, H, NF = 1241, 376, 2000
K = np.float32([718.856, 718.856, 607.1928, 185.2157])
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_triang.npz")
# (seed, only_stereo, check_ori, mono, n2)
CASES = [(71, False, True, False, None), (72, True, True, False, None), (73, False, False, True, None), (74, False, True, True, 1500)]


@pytest.fixture(scope="module")
def scene():
    kps, desc, scale = extract_frame(W, H, NF, 2)
    sigma2 = (scale * scale).astype(np.float32)                                     # src/ORBextractor.cc:505-510
    return dict(kps=kps, desc=desc, scale=scale, sigma2=sigma2)


@needs_ref
@pytest.mark.parametrize("seed,only_stereo,check_ori,mono,n2", CASES)
def test_matches_the_reference(scene, seed, only_stereo, check_ori, mono, n2):
    s = triang_scene(scene["kps"], scene["desc"], W, H, seed, K, scene["scale"], n2)
    a = search_for_triangulation("ref", s, K, scene["scale"], scene["sigma2"], only_stereo, check_ori, mono)
    b = search_for_triangulation("oracle", s, K, scene["scale"], scene["sigma2"], only_stereo, check_ori, mono)
    assert a[0] == b[0] and a[0] > (40 if only_stereo else 150)
    assert (a[1] == b[1]).all()
    m = b[1][b[1] >= 0]
    if not only_stereo:
        assert len(np.unique(m)) < len(m)          # key-frame-2 features ARE shared: the reference never sets vbMatched2


def test_oracle_against_stored_reference_results(scene):
    g = np.load(GOLDEN)
    for seed, only_stereo, check_ori, mono, n2 in CASES:
        s = triang_scene(scene["kps"], scene["desc"], W, H, seed, K, scene["scale"], n2)
        nm, m12 = search_for_triangulation("oracle", s, K, scene["scale"], scene["sigma2"], only_stereo, check_ori, mono)
        assert nm == int(g[f"triang_{seed}_n"]) and (m12 == g[f"triang_{seed}_m12"]).all()


@needs_ref
def test_degenerate_inputs_match_the_reference(scene):
    """A key frame without feature vector, a partner of three features, every feature already holding a map point."""
    s = triang_scene(scene["kps"], scene["desc"], W, H, 75, K, scene["scale"])
    empty = dict(s); empty["fv2"] = (np.zeros(0, np.int32), np.zeros(1, np.int32), np.zeros(0, np.int32))
    full = dict(s); full["has1"] = np.ones_like(s["has1"])
    tiny = triang_scene(scene["kps"], scene["desc"], W, H, 76, K, scene["scale"], n2=3)
    for v in (empty, full, tiny):
        a = search_for_triangulation("ref", v, K, scene["scale"], scene["sigma2"], False, True, False)
        b = search_for_triangulation("oracle", v, K, scene["scale"], scene["sigma2"], False, True, False)
        assert a[0] == b[0] and (a[1] == b[1]).all()
    assert search_for_triangulation("oracle", empty, K, scene["scale"], scene["sigma2"], False, True, False)[0] == 0
