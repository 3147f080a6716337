"""ORBmatcher::SearchByBoW (src/ORBmatcher.cc:552-832): the plain-C oracle against the reference's unmodified
ORBmatcher.cc (where oracle/_ref was built) and against committed fixtures generated from it."""
import os

import numpy as np
import pytest

from bow_lib import bow_scene, search_by_bow
from matcher_lib import extract_frame
from oracle_lib import ref

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_bow.npz")
CASES = [(0.75, True, True), (0.75, True, False), (0.9, False, True), (0.6, True, False), (0.95, False, False)]


@pytest.fixture(scope="module")
def frame():
    kps, desc, _ = extract_frame(1241, 376, 2000, 2)
    return kps, desc


@pytest.mark.skipif(ref() is None, reason="oracle/_ref not built (no /root/reference here)")
@pytest.mark.parametrize("nnratio,check_ori,kf_kf", CASES)
def test_search_by_bow_vs_reference(frame, nnratio, check_ori, kf_kf):
    for seed, n2 in ((1, None), (2, 1500)):
        s = bow_scene(*frame, seed, n2=n2)
        a = search_by_bow("ref", s, nnratio, check_ori, kf_kf)
        b = search_by_bow("oracle", s, nnratio, check_ori, kf_kf)
        assert a[0] == b[0] and a[0] > 100, (a[0], b[0])
        assert (a[1] == b[1]).all()


def test_oracle_vs_bow_fixtures(frame):
    g = np.load(GOLDEN)
    s = bow_scene(*frame, int(g["seed"]))
    for k, (nnratio, check_ori, kf_kf) in enumerate(CASES):
        nm, m = search_by_bow("oracle", s, nnratio, check_ori, kf_kf)
        assert nm == int(g[f"nm_{k}"]) and (m == g[f"match_{k}"]).all()
