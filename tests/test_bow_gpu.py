"""ORBmatcher::SearchByBoW on the GPU (csrc/orb_match_bow.cu) through the C ABI against the oracle and the committed
reference fixtures: matched feature indices and counts identical for both overloads."""
import os

import numpy as np
import pytest

import orb_slam2_chinesenotes_b200 as ob
from bow_lib import bow_scene, search_by_bow
from matcher_lib import extract_frame
from test_bow_oracle import CASES, GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def frame():
    kps, desc, _ = extract_frame(1241, 376, 2000, 2)
    return kps, desc


def _t(a):
    import torch
    a = np.ascontiguousarray(a)
    if a.dtype.names:
        a = a.view(np.uint8).reshape(a.shape + (28,))
    return torch.from_numpy(a).cuda()


def _pad(rows, stride, fill=0):
    out = np.full((len(rows), stride) + rows[0].shape[1:], fill, rows[0].dtype)
    for i, r in enumerate(rows):
        out[i, :len(r)] = r
    return out


def _side(scenes, k, d, fv, cap, nstride):
    kp = np.zeros((len(scenes), cap), scenes[0][k].dtype)
    for i, s in enumerate(scenes):
        kp[i, :len(s[k])] = s[k]
    T = dict(kps=_t(kp), desc=_t(_pad([s[d] for s in scenes], cap)), n=_t(np.int32([len(s[k]) for s in scenes])),
             node_id=_t(_pad([s[fv][0] for s in scenes], nstride)), node_off=_t(_pad([s[fv][1] for s in scenes], nstride + 1)),
             n_nodes=_t(np.int32([len(s[fv][0]) for s in scenes])), feat=_t(_pad([s[fv][2] for s in scenes], cap)))
    F = ob.frames_batch(T["kps"], T["desc"], T["n"], (0.0, 1241.0, 0.0, 376.0))
    return F, (T["node_id"], T["node_off"], T["n_nodes"], T["feat"]), T


@pytest.mark.parametrize("nnratio,check_ori,kf_kf", CASES)
def test_bow_batch_equals_oracle(frame, nnratio, check_ori, kf_kf):
    import torch
    scenes = [bow_scene(*frame, 1), bow_scene(*frame, 2, n2=1500), bow_scene(*frame, 3, kmax=15), bow_scene(*frame, 4, n2=40)]
    empty = bow_scene(*frame, 6, n2=300)
    empty["fv2"] = (np.zeros(0, np.int32), np.zeros(1, np.int32), np.zeros(0, np.int32))      # a frame without feature vector
    scenes.append(empty)
    cap, ns = 2100, 300
    A, VA, keepA = _side(scenes, "k1", "d1", "fv1", cap, ns)
    B, VB, keepB = _side(scenes, "k2", "d2", "fv2", cap, ns)
    v1 = _t(_pad([s["valid1"] & (1 - s["bad1"]) for s in scenes], cap))
    v2 = _t(_pad([s["valid2"] & (1 - s["bad2"]) for s in scenes], cap)) if kf_kf else None
    P = len(scenes)
    m12 = torch.full((P, cap), -7, dtype=torch.int32, device="cuda")
    m21 = torch.full((P, cap), -7, dtype=torch.int32, device="cuda")
    nm = torch.zeros(P, dtype=torch.int32, device="cuda")
    rounds = torch.zeros(P, dtype=torch.int32, device="cuda")
    ob.search_by_bow_batch(A, VA, v1, B, VB, v2, kf_kf, nnratio, check_ori, m12, nm, m21, rounds)
    torch.cuda.synchronize()
    m12, m21, nm, rounds = m12.cpu().numpy(), m21.cpu().numpy(), nm.cpu().numpy(), rounds.cpu().numpy()
    for i, s in enumerate(scenes):
        want = search_by_bow("oracle", s, nnratio, check_ori, kf_kf)
        n1, n2 = len(s["k1"]), len(s["k2"])
        assert nm[i] == want[0], (i, nm[i], want[0])
        assert (m12[i, :n1] == want[1]).all() and (m12[i, n1:] == -7).all()
        inv = np.full(n2, -1, np.int32)
        inv[want[1][want[1] >= 0]] = np.nonzero(want[1] >= 0)[0]
        assert (m21[i, :n2] == inv).all()
    assert nm[0] > 100 and nm[4] == 0 and rounds.max() >= 2


def test_bow_batch_equals_reference_fixture(frame):
    import torch
    g = np.load(GOLDEN)
    s = bow_scene(*frame, int(g["seed"]))
    n = len(s["k1"])
    A, VA, keepA = _side([s], "k1", "d1", "fv1", n, 256)
    B, VB, keepB = _side([s], "k2", "d2", "fv2", n, 256)
    v1 = _t((s["valid1"] & (1 - s["bad1"]))[None])
    v2 = _t((s["valid2"] & (1 - s["bad2"]))[None])
    for k, (nnratio, check_ori, kf_kf) in enumerate(CASES):
        m12 = torch.zeros((1, n), dtype=torch.int32, device="cuda")
        nm = torch.zeros(1, dtype=torch.int32, device="cuda")
        ob.search_by_bow_batch(A, VA, v1, B, VB, v2 if kf_kf else None, kf_kf, nnratio, check_ori, m12, nm)
        torch.cuda.synchronize()
        assert int(nm[0]) == int(g[f"nm_{k}"]) and (m12[0].cpu().numpy() == g[f"match_{k}"]).all()


@pytest.mark.parametrize("nnratio,check_ori,kf_kf", CASES[:3])
def test_bow_single_problem_host_arrays(frame, nnratio, check_ori, kf_kf):
    """orbm_search_by_bow (what the C++ forwarder calls): host arrays in, same result as the oracle."""
    s = bow_scene(*frame, 8, n2=1700)
    B = (0.0, 1241.0, 0.0, 376.0)
    v1 = s["valid1"] & (1 - s["bad1"])
    v2 = (s["valid2"] & (1 - s["bad2"])) if kf_kf else None
    nm, m = ob.search_by_bow(ob.FrameView(s["k1"], s["d1"], B), s["fv1"], v1, ob.FrameView(s["k2"], s["d2"], B), s["fv2"], v2, kf_kf, nnratio, check_ori)
    want = search_by_bow("oracle", s, nnratio, check_ori, kf_kf)
    assert nm == want[0] and nm > 100 and (m == want[1]).all()
    # empty feature vector on one side: nothing matches
    empty = (np.zeros(0, np.int32), np.zeros(1, np.int32), np.zeros(0, np.int32))
    nm, m = ob.search_by_bow(ob.FrameView(s["k1"], s["d1"], B), s["fv1"], v1, ob.FrameView(s["k2"], s["d2"], B), empty, v2, kf_kf, nnratio, check_ori)
    assert nm == 0 and (m == -1).all()
