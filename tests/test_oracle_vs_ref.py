"""T1: the plain-C oracle (oracle/orb_oracle.c) against the reference's own unmodified
ORBextractor.cc (oracle/_ref/liborbref.so) and against the committed fixtures generated from it.
"""
import os

import numpy as np
import pytest

from oracle_lib import OracleExtractor, RefExtractor, have_ref, oracle_distribute, ref
from synth import synth_frame

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
needs_ref = pytest.mark.skipif(ref() is None, reason="oracle/_ref/liborbref.so not built (needs /root/reference)")


def same_kps(a, b):
    return len(a) == len(b) and all((a[f] == b[f]).all() for f in a.dtype.names)


@pytest.mark.parametrize("name", ["tum", "kitti", "small"])
def test_oracle_vs_golden(name):
    """Always runs (no reference needed): fixtures were produced by the reference sources."""
    g = np.load(os.path.join(GOLD, f"ref_extract_{name}.npz"))
    O = OracleExtractor(int(g["nfeatures"]))
    n, kps, desc = O.extract(synth_frame(int(g["w"]), int(g["h"]), int(g["seed"])))
    assert n == len(g["kps"])
    assert same_kps(kps, g["kps"]) and (desc == g["desc"]).all()
    t = O.tables()
    for k, v in t.items():
        assert (v == g["tab_" + k]).all(), k
    O.close()


def test_constructor_tables_known_answers():
    """SURVEY.md App. C (computed independently with float32 emulation)."""
    O = OracleExtractor(2000)
    t = O.tables()
    assert t["per_level"].tolist() == [434, 362, 302, 251, 209, 175, 145, 122]
    assert t["umax"].tolist() == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    assert np.float32(t["scale"][7]) == np.float32(3.583181619644165)
    assert np.float32(t["inv_scale"][3]) == np.float32(0.5787036418914795)
    assert OracleExtractor(1000).tables()["per_level"].tolist() == [217, 181, 151, 126, 105, 87, 73, 60]
    assert OracleExtractor(4000).tables()["per_level"].tolist() == [869, 724, 603, 503, 419, 349, 291, 242]


@needs_ref
@pytest.mark.parametrize("w,h,nf,seed", [(640, 480, 1000, 21), (1241, 376, 2000, 22), (752, 480, 1200, 23),
                                          (320, 240, 500, 24), (160, 200, 200, 25), (500, 150, 300, 26)])
def test_oracle_vs_reference_end_to_end(w, h, nf, seed):
    img = synth_frame(w, h, seed)
    R, O = RefExtractor(nf), OracleExtractor(nf)
    n1, k1, d1 = R.extract(img)
    n2, k2, d2 = O.extract(img)
    assert n1 == n2 and same_kps(k1, k2) and (d1 == d2).all()
    for l in range(8):
        assert (R.pyramid(l, True) == O.pyramid(l, True)).all()
    tr, to = R.tables(), O.tables()
    assert all((tr[k] == to[k]).all() for k in tr)
    R.close(); O.close()


@needs_ref
def test_adversarial_images():
    """flat, saturated, checkerboard, gradient: empty levels, ties, retries."""
    h, w = 240, 320
    yy, xx = np.mgrid[0:h, 0:w]
    imgs = {
        "flat": np.full((h, w), 127, np.uint8),
        "white": np.full((h, w), 255, np.uint8),
        "checker8": (((yy // 8 + xx // 8) % 2) * 255).astype(np.uint8),
        "checker3": (((yy // 3 + xx // 3) % 2) * 200 + 20).astype(np.uint8),
        "gradient": ((xx * 255) // w).astype(np.uint8),
        "noise": np.random.default_rng(9).integers(0, 256, (h, w), dtype=np.uint8),
    }
    R, O = RefExtractor(500), OracleExtractor(500)
    for name, img in imgs.items():
        n1, k1, d1 = R.extract(img)
        n2, k2, d2 = O.extract(img)
        assert n1 == n2, name
        if n1 > 0:
            assert same_kps(k1, k2) and (d1 == d2).all(), name
    assert R.extract(np.zeros((0, 0), np.uint8))[0] == -1 and O.extract(np.zeros((0, 0), np.uint8))[0] == -1
    R.close(); O.close()


@needs_ref
def test_reference_is_pure_under_the_arena():
    """Same input, different arena offsets -> identical output."""
    img = synth_frame(640, 480, 31)
    R = RefExtractor(1000)
    n1, k1, d1 = R.extract(img)
    for skew in (16, 4096, 1234560):
        ref().orbref_arena_skew(skew)
        n2, k2, d2 = R.extract(img)
        assert n1 == n2 and same_kps(k1, k2) and (d1 == d2).all()
    R.close()


@needs_ref
@pytest.mark.parametrize("W,H,N", [(1209, 344, 434), (608, 448, 217), (720, 448, 261), (314, 73, 122), (100, 180, 60)])
def test_distribute_octtree_pure_rule(W, H, N):
    """DistributeOctTree restatement == reference (ordered), over candidate densities."""
    R = RefExtractor(N)
    rng = np.random.default_rng(W + H)
    from oracle_lib import CAND_DTYPE
    for mult in (12, 3, 1, 0.5):
        for rep in range(3):
            n = int(N * mult)
            pts = set()
            while len(pts) < n:
                pts.add((int(rng.integers(3, W - 3)), int(rng.integers(3, H - 3))))
            pts = sorted(pts, key=lambda p: (p[1] // 30, p[0] // 30, p[1], p[0]))
            cand = np.zeros(n, CAND_DTYPE)
            cand["x"] = [p[0] for p in pts]; cand["y"] = [p[1] for p in pts]
            cand["score"] = rng.integers(7, 60, n)   # few distinct values: plenty of response ties
            out_ref = R.distribute(cand, 16, 16 + W, 16, 16 + H, N)
            idx = oracle_distribute(cand, 16, 16 + W, 16, 16 + H, N)
            assert len(idx) == len(out_ref)
            assert (cand["x"][idx] == out_ref["x"]).all() and (cand["y"][idx] == out_ref["y"]).all()
    R.close()
