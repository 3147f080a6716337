"""T0: the oracle's OpenCV / libm primitive restatements (oracle/cv_prims.c) are pinned
bit-for-bit to OpenCV 4.13 and glibc: live against Python cv2 when it is importable, and always
against the fixtures in tests/golden/ that tests/golden/make_golden.py generated from cv2 4.13.0.
"""
import ctypes as C
import os

import numpy as np
import pytest

from oracle_lib import CAND_DTYPE, oracle
from synth import synth_frame

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

try:
    import cv2
    HAVE_CV2 = cv2.__version__.startswith("4.13")
except Exception:  # pragma: no cover
    HAVE_CV2 = False
needs_cv2 = pytest.mark.skipif(not HAVE_CV2, reason="cv2 4.13 not importable")


def o_resize(img, dw, dh):
    out = np.zeros((dh, dw), np.uint8)
    oracle().cvp_resize_linear_8u(img.ctypes.data, img.shape[1], img.shape[0], img.strides[0], out.ctypes.data, dw, dh, dw)
    return out


def o_blur(img):
    out = np.zeros_like(img)
    oracle().cvp_gaussian7x7_s2(img.ctypes.data, img.shape[1], img.shape[0], img.strides[0], out.ctypes.data, out.strides[0])
    return out


def o_border(img, b=19):
    out = np.zeros((img.shape[0] + 2 * b, img.shape[1] + 2 * b), np.uint8)
    oracle().cvp_border_reflect101(img.ctypes.data, img.shape[1], img.shape[0], img.strides[0], out.ctypes.data, out.strides[0], b)
    return out


def o_fast(img, t):
    buf = np.zeros(max(16, img.size // 4 + 16), CAND_DTYPE)
    n = oracle().cvp_fast9_nms(img.ctypes.data, img.shape[1], img.shape[0], img.strides[0], t, buf.ctypes.data, len(buf))
    return [(int(b["x"]), int(b["y"]), int(b["score"])) for b in buf[:n]]


def cv_fast(img, t):
    det = cv2.FastFeatureDetector_create(threshold=t, nonmaxSuppression=True, type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    return [(int(k.pt[0]), int(k.pt[1]), int(k.response)) for k in det.detect(img)]


def level_sizes(w, h, n=8):
    out = []
    for l in range(n):
        inv = np.float32(1) / np.float32(np.float32(1.2) ** l) if l else np.float32(1)
        out.append((int(np.rint(np.float32(w) * inv)), int(np.rint(np.float32(h) * inv))))
    return out


@needs_cv2
@pytest.mark.parametrize("w,h", [(640, 480), (1241, 376), (752, 480), (97, 61)])
def test_resize_vs_cv2(w, h):
    prev = synth_frame(w, h, 11)
    for l in range(1, 8):
        dw, dh = max(1, int(round(w / 1.2 ** l))), max(1, int(round(h / 1.2 ** l)))
        ref = cv2.resize(prev, (dw, dh), interpolation=cv2.INTER_LINEAR)
        assert (ref == o_resize(prev, dw, dh)).all(), (w, h, l)
        prev = ref


@needs_cv2
@pytest.mark.parametrize("w,h", [(640, 480), (8, 8), (9, 31), (346, 105), (7, 200)])
def test_blur_vs_cv2(w, h):
    rng = np.random.default_rng(w * 1000 + h)
    for kind in range(3):
        img = rng.integers(0, 256, (h, w), dtype=np.uint8) if kind == 0 else (
            synth_frame(max(w, 64), max(h, 64), 5)[:h, :w].copy() if kind == 1 else np.full((h, w), 255, np.uint8))
        ref = cv2.GaussianBlur(img, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
        assert (ref == o_blur(img)).all()
        inplace = img.copy()
        oracle().cvp_gaussian7x7_s2(inplace.ctypes.data, w, h, w, inplace.ctypes.data, w)
        assert (ref == inplace).all()


@needs_cv2
def test_border_vs_cv2():
    rng = np.random.default_rng(3)
    for (w, h) in [(70, 50), (20, 25), (346, 105)]:
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        assert (cv2.copyMakeBorder(img, 19, 19, 19, 19, cv2.BORDER_REFLECT_101) == o_border(img)).all()


@needs_cv2
def test_fast_vs_cv2():
    rng = np.random.default_rng(4)
    total = 0
    for i in range(80):
        w, h = int(rng.integers(7, 50)), int(rng.integers(7, 50))
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        if i % 3 == 0:
            img = cv2.GaussianBlur(img, (0, 0), 1.5)
        if i % 7 == 0:
            img[h // 3:, w // 3:] = 200   # flat regions with hard edges: score ties
        t = int(rng.choice([7, 20, 40]))
        ref = cv_fast(img, t)
        total += len(ref)
        assert ref == o_fast(img, t), (w, h, t)
    big = synth_frame(1241, 376, 2)
    for t in (20, 7):
        assert cv_fast(big, t) == o_fast(big, t)
    view = big[100:140, 200:243]   # non-contiguous cell-sized view, as ORBextractor.cc:853 passes
    assert cv_fast(view, 20) == o_fast(view, 20)
    assert total > 1000


@needs_cv2
def test_fast_atan2_vs_cv2():
    rng = np.random.default_rng(5)
    L = oracle()
    ys = rng.integers(-2 ** 23, 2 ** 23, 20000).astype(np.float32)
    xs = rng.integers(-2 ** 23, 2 ** 23, 20000).astype(np.float32)
    for y, x in zip(ys, xs):
        assert np.float32(cv2.fastAtan2(float(y), float(x))) == np.float32(L.cvp_fast_atan2(float(y), float(x)))
    for y, x in [(0, 0), (0, 1), (1, 0), (0, -1), (-1, 0), (5, 5), (-5, 5), (5, -5)]:
        assert np.float32(cv2.fastAtan2(float(y), float(x))) == np.float32(L.cvp_fast_atan2(float(y), float(x)))


def test_sincosf_vs_libm():
    """glibc restatement == the libm this box runs (what the reference's cos()/sin() call)."""
    libm = C.CDLL("libm.so.6")
    L = oracle()
    for f in ("sinf", "cosf"):
        getattr(libm, f).restype = C.c_float
        getattr(libm, f).argtypes = [C.c_float]
    rng = np.random.default_rng(6)
    ang = np.concatenate([(rng.random(100000) * 2 * np.pi).astype(np.float32),
                          np.float32([0, 1e-5, 2 ** -12, 0.7853981, 0.7853982, 1.5707963, 3.1415927, 6.2831850])])
    for a in ang:
        a = float(a)
        assert libm.sinf(a) == L.cvp_sinf(a) and libm.cosf(a) == L.cvp_cosf(a), a


# ---- committed fixtures (generated from cv2 4.13.0 by tests/golden/make_golden.py) -----------------
def test_golden_primitives():
    g = np.load(os.path.join(GOLD, "cv2_primitives.npz"))
    img = g["img"]
    assert (o_resize(img, g["resized"].shape[1], g["resized"].shape[0]) == g["resized"]).all()
    assert (o_blur(img) == g["blurred"]).all()
    assert (o_border(img) == g["bordered"]).all()
    for t, key in ((20, "fast20"), (7, "fast7")):
        assert np.array(o_fast(img, t), np.int32).reshape(-1, 3).tolist() == g[key].tolist()
    L = oracle()
    got = np.float32([L.cvp_fast_atan2(float(y), float(x)) for y, x in g["atan_in"]])
    assert (got == g["atan_out"]).all()
    assert (np.float32([L.cvp_sinf(float(a)) for a in g["angles"]]) == g["sin"]).all()
    assert (np.float32([L.cvp_cosf(float(a)) for a in g["angles"]]) == g["cos"]).all()
