"""ORBmatcher::SearchForTriangulation on the GPU (k_bow_fixpoint<., TRI> in csrc/orb_match_bow.cu) through the C ABI against
the oracle and the committed reference fixtures (tests/golden/ref_triang.npz): matched feature indices and counts identical."""
import numpy as np
import pytest

import orb_slam2_chinesenotes_b200 as ob
from matcher_lib import extract_frame
from test_bow_gpu import _pad, _t
from test_triang_oracle import CASES, GOLDEN, H, K, NF, W
from triang_lib import search_for_triangulation, triang_scene

pytestmark = pytest.mark.gpu
BOUNDS = (0.0, float(W), 0.0, float(H))


@pytest.fixture(scope="module")
def scene():
    kps, desc, scale = extract_frame(W, H, NF, 2)
    return dict(kps=kps, desc=desc, scale=scale, sigma2=(scale * scale).astype(np.float32))


def _valid(s, only_stereo, mono):
    """a_valid / b_valid as host/ORBmatcher_b200.hpp builds them: no map point yet, and a stereo feature when bOnlyStereo."""
    st1 = np.zeros(len(s["k1"]), bool) if mono else s["ur1"] >= 0
    st2 = np.zeros(len(s["k2"]), bool) if mono else s["ur2"] >= 0
    v1 = (s["has1"] == 0) & (st1 | (not only_stereo))
    v2 = (s["has2"] == 0) & (st2 | (not only_stereo))
    return v1.astype(np.uint8), v2.astype(np.uint8)


@pytest.mark.parametrize("seed,only_stereo,check_ori,mono,n2", CASES)
def test_single_pair_equals_oracle_and_reference_fixture(scene, seed, only_stereo, check_ori, mono, n2):
    s = triang_scene(scene["kps"], scene["desc"], W, H, seed, K, scene["scale"], n2)
    want = search_for_triangulation("oracle", s, K, scene["scale"], scene["sigma2"], only_stereo, check_ori, mono)
    v1, v2 = _valid(s, only_stereo, mono)
    A = ob.FrameView(s["k1"], s["d1"], BOUNDS, None if mono else s["ur1"])
    B = ob.FrameView(s["k2"], s["d2"], BOUNDS, None if mono else s["ur2"])
    nm, m12 = ob.search_for_triangulation(A, s["fv1"], v1, B, s["fv2"], v2, s["F12"], s["epipole"], scene["scale"], scene["sigma2"], check_ori)
    assert nm == want[0] and nm > (40 if only_stereo else 150) and (m12 == want[1]).all()
    g = np.load(GOLDEN)
    assert nm == int(g[f"triang_{seed}_n"]) and (m12 == g[f"triang_{seed}_m12"]).all()


@pytest.mark.parametrize("only_stereo,check_ori,mono", [(False, True, False), (True, False, False), (False, True, True)])
def test_batch_equals_oracle(scene, only_stereo, check_ori, mono):
    import torch
    kps, desc, scale, sigma2 = scene["kps"], scene["desc"], scene["scale"], scene["sigma2"]
    scenes = [triang_scene(kps, desc, W, H, 81, K, scale), triang_scene(kps, desc, W, H, 82, K, scale, n2=1500),
              triang_scene(kps, desc, W, H, 83, K, scale, kmax=10), triang_scene(kps, desc, W, H, 84, K, scale, n2=40)]
    empty = triang_scene(kps, desc, W, H, 85, K, scale, n2=300)
    empty["fv2"] = (np.zeros(0, np.int32), np.zeros(1, np.int32), np.zeros(0, np.int32))      # a key frame without feature vector
    scenes.append(empty)
    cap, ns = 2100, 300

    def side(k, d, fv, ur):
        kp = np.zeros((len(scenes), cap), scenes[0][k].dtype)
        for i, s in enumerate(scenes):
            kp[i, :len(s[k])] = s[k]
        T = dict(kps=_t(kp), desc=_t(_pad([s[d] for s in scenes], cap)), n=_t(np.int32([len(s[k]) for s in scenes])),
                 node_id=_t(_pad([s[fv][0] for s in scenes], ns)), node_off=_t(_pad([s[fv][1] for s in scenes], ns + 1)),
                 n_nodes=_t(np.int32([len(s[fv][0]) for s in scenes])), feat=_t(_pad([s[fv][2] for s in scenes], cap)),
                 ur=None if mono else _t(_pad([s[ur] for s in scenes], cap, -1)))
        return ob.frames_batch(T["kps"], T["desc"], T["n"], BOUNDS, T["ur"]), (T["node_id"], T["node_off"], T["n_nodes"], T["feat"]), T

    A, VA, keepA = side("k1", "d1", "fv1", "ur1")
    B, VB, keepB = side("k2", "d2", "fv2", "ur2")
    vs = [_valid(s, only_stereo, mono) for s in scenes]
    v1, v2 = _t(_pad([v[0] for v in vs], cap)), _t(_pad([v[1] for v in vs], cap))
    F12 = _t(np.stack([s["F12"].ravel() for s in scenes]).astype(np.float32))
    epi = _t(np.stack([s["epipole"] for s in scenes]).astype(np.float32))
    P = len(scenes)
    m12 = torch.full((P, cap), -7, dtype=torch.int32, device="cuda")
    nm = torch.full((P,), -7, dtype=torch.int32, device="cuda")
    ob.search_for_triangulation_batch(A, VA, v1, B, VB, v2, F12, epi, scale, sigma2, check_ori, m12, nm)
    torch.cuda.synchronize()
    m12, nm = m12.cpu().numpy(), nm.cpu().numpy()
    for i, s in enumerate(scenes):
        want = search_for_triangulation("oracle", s, K, scale, sigma2, only_stereo, check_ori, mono)
        n1 = len(s["k1"])
        assert nm[i] == want[0], (i, nm[i], want[0])
        assert (m12[i, :n1] == want[1]).all() and (m12[i, n1:] == -7).all()
    assert nm[0] > (40 if only_stereo else 150) and nm[4] == 0
