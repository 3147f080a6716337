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
