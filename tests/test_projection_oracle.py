"""Map-point side of the path: the plain-C oracle against the reference's own MapPoint.cc /
Frame.cc (where oracle/_ref was built) and against the committed fixtures generated from them."""
import ctypes as C
import os

import numpy as np
import pytest

from mappoint_lib import descriptor_groups, distinctive, frustum_scene, is_in_frustum, mappoint_ref, predict_scale
from oracle_lib import oracle, ref

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_mappoint.npz")
needs_ref = pytest.mark.skipif(ref() is None or mappoint_ref() is None, reason="oracle/_ref not built (no /root/reference here)")


def test_logf_vs_libm():
    """cvp_logf restates glibc's logf: bit-equal on a dense sample (all 2^31 positive floats were compared once when
    the table was written; this keeps a guard in the suite)."""
    O = oracle()
    O.cvp_logf.restype = C.c_float
    O.cvp_logf.argtypes = [C.c_float]
    m = C.CDLL("libm.so.6")
    m.logf.restype = C.c_float
    m.logf.argtypes = [C.c_float]
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.integers(1, 0x7f800000, 20000).astype(np.uint32).view(np.float32),
                         (1.0 + rng.normal(0, 1e-3, 5000)).astype(np.float32), np.float32([1.0, 1.2, 0.5, 2.0, 1e-40, 3e38])])
    for x in xs:
        a, b = np.float32(O.cvp_logf(float(x))), np.float32(m.logf(float(x)))
        assert a.view(np.uint32) == b.view(np.uint32), x


@needs_ref
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_distinctive_descriptor_vs_reference(seed):
    for d, bad in descriptor_groups(seed):
        for b in (None, bad):
            want, _, _ = distinctive("ref", d, b)
            got, idx, med = distinctive("oracle", d, b)
            if want is None:
                assert got is None
            else:
                assert got is not None and (got == want).all(), (len(d), idx)


@needs_ref
def test_predict_scale_vs_reference():
    rng = np.random.default_rng(5)
    for max_d in (10.0, 3.7, 55.5):
        cur = np.concatenate([max_d / np.float32(1.2) ** np.arange(-3, 12), rng.uniform(0.05, 200, 3000)]).astype(np.float32)
        cur = np.concatenate([cur, np.nextafter(cur[:15], np.float32(0)), np.nextafter(cur[:15], np.float32(1e9))])
        for sf, nl in ((1.2, 8), (1.5, 5), (1.1, 12)):
            assert (predict_scale("ref", max_d, cur, sf, nl) == predict_scale("oracle", max_d, cur, sf, nl)).all()


@needs_ref
@pytest.mark.parametrize("seed,cos_limit", [(1, 0.5), (2, 0.5), (3, 0.0), (4, 0.8)])
def test_is_in_frustum_vs_reference(seed, cos_limit):
    s = frustum_scene(seed)
    a = is_in_frustum("ref", s, cos_limit=cos_limit)
    b = is_in_frustum("oracle", s, cos_limit=cos_limit)
    assert 300 < a[0].sum() < len(a[0]) - 300
    assert (a[0] == b[0]).all() and (a[2] == b[2]).all()
    assert (a[1].view(np.uint32) == b[1].view(np.uint32)).all() and (a[3].view(np.uint32) == b[3].view(np.uint32)).all()


def test_oracle_vs_mappoint_fixtures():
    """Always on (also where /root/reference does not exist): fixtures written by tests/golden/make_golden.py from the
    reference's unmodified MapPoint.cc / Frame.cc."""
    g = np.load(GOLDEN)
    for k, (d, bad) in enumerate(descriptor_groups(int(g["desc_seed"]))):
        for tag, b in (("all", None), ("bad", bad)):
            got, _, _ = distinctive("oracle", d, b)
            want = g[f"distinctive_{tag}_{k}"]
            assert (got is None and want.size == 0) or (got == want).all()
    cur = g["predict_cur"]
    assert (predict_scale("oracle", float(g["predict_max"]), cur) == g["predict_level"]).all()
    s = frustum_scene(int(g["frustum_seed"]))
    iv, proj, lv, vc = is_in_frustum("oracle", s)
    assert (iv == g["frustum_in_view"]).all() and (lv == g["frustum_level"]).all()
    assert (proj.view(np.uint32) == g["frustum_proj"].view(np.uint32)).all()
    assert (vc.view(np.uint32) == g["frustum_view_cos"].view(np.uint32)).all()
