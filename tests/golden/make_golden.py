#!/usr/bin/env python3
"""Generates the committed fixtures in tests/golden/ (run in the authoring container).

cv2_primitives.npz   outputs of Python cv2 4.13.0 / libm for one small image: pins oracle/cv_prims.c
ref_extract_*.npz    outputs of the reference's UNMODIFIED ORBextractor.cc (oracle/_ref/liborbref.so,
                     built from /root/reference against oracle/cvshim + bump arena): pins the oracle
                     restatement and, through it, the CUDA path
ref_match_*.npz      outputs of the reference's UNMODIFIED ORBmatcher.cc on mock frames
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle_lib import RefExtractor  # noqa: E402
from synth import synth_frame  # noqa: E402


def primitives():
    import cv2
    assert cv2.__version__.startswith("4.13"), cv2.__version__
    img = synth_frame(160, 120, 77)
    det = lambda t: cv2.FastFeatureDetector_create(threshold=t, nonmaxSuppression=True, type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    fast = lambda t: np.array([(int(k.pt[0]), int(k.pt[1]), int(k.response)) for k in det(t).detect(img)], np.int32).reshape(-1, 3)
    rng = np.random.default_rng(1)
    atan_in = rng.integers(-2 ** 20, 2 ** 20, (2000, 2)).astype(np.float32)
    atan_out = np.float32([cv2.fastAtan2(float(y), float(x)) for y, x in atan_in])
    libm = C.CDLL("libm.so.6")
    for f in ("sinf", "cosf"):
        getattr(libm, f).restype = C.c_float
        getattr(libm, f).argtypes = [C.c_float]
    angles = (rng.random(5000) * 2 * np.pi).astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "cv2_primitives.npz"), img=img,
                        resized=cv2.resize(img, (133, 100), interpolation=cv2.INTER_LINEAR),
                        blurred=cv2.GaussianBlur(img, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101),
                        bordered=cv2.copyMakeBorder(img, 19, 19, 19, 19, cv2.BORDER_REFLECT_101),
                        fast20=fast(20), fast7=fast(7), atan_in=atan_in, atan_out=atan_out, angles=angles,
                        sin=np.float32([libm.sinf(float(a)) for a in angles]),
                        cos=np.float32([libm.cosf(float(a)) for a in angles]))


def extractor():
    for name, (w, h, nf, seed) in {"tum": (640, 480, 1000, 1), "kitti": (1241, 376, 2000, 2), "small": (200, 150, 300, 5)}.items():
        R = RefExtractor(nf)
        n, kps, desc = R.extract(synth_frame(w, h, seed))
        t = R.tables()
        np.savez_compressed(os.path.join(HERE, f"ref_extract_{name}.npz"), w=w, h=h, nfeatures=nf, seed=seed,
                            kps=kps, desc=desc, **{"tab_" + k: v for k, v in t.items()})
        R.close()


if __name__ == "__main__":
    primitives()
    extractor()
    print("golden fixtures written to", HERE)
