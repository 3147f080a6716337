"""ctypes bindings of the CHECKERS (test infrastructure):

  oracle/_build/liborb_oracle.so  plain-C restatement, built from committed sources
  oracle/_ref/liborbref.so        the reference's unmodified sources + shim + arena
                                  (built only where /root/reference exists; travels
                                  to the GPU box as a prebuilt file)
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "_build", "liborb_oracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "liborbref.so")
REFERENCE_ROOT = "/root/reference"

KP_DTYPE = np.dtype([("x", "f4"), ("y", "f4"), ("size", "f4"), ("angle", "f4"), ("response", "f4"),
                     ("octave", "i4"), ("class_id", "i4")])
CAND_DTYPE = np.dtype([("x", "i4"), ("y", "i4"), ("score", "i4")])


def build_oracle():
    subprocess.run(["make", "-C", ORACLE_DIR, "oracle"], check=True, capture_output=True)


def build_ref():
    if os.path.isdir(REFERENCE_ROOT):
        subprocess.run(["make", "-C", ORACLE_DIR, "ref"], check=True, capture_output=True)


def have_ref():
    return os.path.exists(REF_SO)


_oracle = None
_ref = None


def oracle():
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_SO) or any(
                os.path.getmtime(os.path.join(ORACLE_DIR, f)) > os.path.getmtime(ORACLE_SO)
                for f in ("cv_prims.c", "cv_prims.h", "orb_oracle.c", "orb_oracle.h", "orb_pattern.inc")):
            build_oracle()
        L = C.CDLL(ORACLE_SO)
        L.orbo_create.restype = C.c_void_p
        L.orbo_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.orbo_destroy.argtypes = [C.c_void_p]
        L.orbo_levels.argtypes = [C.c_void_p]
        L.orbo_tables.argtypes = [C.c_void_p] + [C.c_void_p] * 7
        L.orbo_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int]
        L.orbo_stage_level_size.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orbo_stage_pyramid.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
        L.orbo_stage_blurred.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
        L.orbo_stage_candidates.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.orbo_stage_level_keypoints.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.orbo_distribute.argtypes = [C.c_void_p, C.c_int] + [C.c_int] * 5 + [C.c_void_p, C.c_int]
        L.orbo_extract_bench.restype = C.c_double
        L.orbo_extract_bench.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                         C.c_int, C.c_int, C.POINTER(C.c_longlong)]
        L.orbo_descriptor_distance.argtypes = [C.c_void_p, C.c_void_p]
        L.orbo_hamming_bf.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.cvp_resize_linear_8u.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_int, C.c_int, C.c_size_t]
        L.cvp_border_reflect101.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
        L.cvp_gaussian7x7_s2.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_size_t]
        L.cvp_fast9_nms.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_void_p, C.c_int]
        L.cvp_fast_atan2.restype = C.c_float
        L.cvp_fast_atan2.argtypes = [C.c_float, C.c_float]
        for f in ("cvp_sinf", "cvp_cosf"):
            getattr(L, f).restype = C.c_float
            getattr(L, f).argtypes = [C.c_float]
        _oracle = L
    return _oracle


def ref():
    global _ref
    if _ref is None:
        if os.path.isdir(REFERENCE_ROOT):
            build_ref()
        if not os.path.exists(REF_SO):
            return None
        L = C.CDLL(REF_SO)
        L.orbref_extractor_create.restype = C.c_void_p
        L.orbref_extractor_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.orbref_extractor_destroy.argtypes = [C.c_void_p]
        L.orbref_extractor_levels.argtypes = [C.c_void_p]
        L.orbref_extractor_tables.argtypes = [C.c_void_p] + [C.c_void_p] * 7
        L.orbref_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int]
        L.orbref_pyramid_level.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orbref_keypoints_octtree.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_int, C.c_void_p]
        L.orbref_distribute.argtypes = [C.c_void_p, C.c_void_p, C.c_int] + [C.c_int] * 6 + [C.c_void_p, C.c_int]
        L.orbref_arena_skew.argtypes = [C.c_size_t]
        L.orbref_extract_bench.restype = C.c_double
        L.orbref_extract_bench.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                           C.c_int, C.c_int, C.POINTER(C.c_longlong)]
        if hasattr(L, "orbref_stereo_bench"):
            L.orbref_stereo_bench.restype = C.c_double
            L.orbref_stereo_bench.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                              C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.POINTER(C.c_longlong), C.c_void_p, C.c_int]
        _ref = L
    return _ref


class _Extractor:
    """Common Python face of the two checkers."""
    prefix = None

    def __init__(self, lib, nfeatures, scale_factor, nlevels, ini_th, min_th):
        self.L = lib
        self.nfeatures, self.nlevels = nfeatures, nlevels
        self.h = getattr(lib, self.prefix + "create")(nfeatures, scale_factor, nlevels, ini_th, min_th)
        assert self.h

    def close(self):
        if self.h:
            getattr(self.L, self.prefix + "destroy")(self.h)
            self.h = None

    def tables(self):
        n = self.nlevels
        sc, inv, s2, is2 = (np.zeros(n, np.float32) for _ in range(4))
        per = np.zeros(n, np.int32)
        umax = np.zeros(16, np.int32)
        pat = np.zeros(1024, np.int32)
        getattr(self.L, self.prefix + "tables")(self.h, *(a.ctypes.data for a in (sc, inv, s2, is2, per, umax, pat)))
        return dict(scale=sc, inv_scale=inv, sigma2=s2, inv_sigma2=is2, per_level=per, umax=umax, pattern=pat)

    def extract(self, img, cap=None):
        img = np.ascontiguousarray(img)
        h, w = img.shape if img.ndim == 2 else (0, 0)
        cap = cap or (2 * self.nfeatures + 256)
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = self._extract(self.h, img.ctypes.data if img.size else None, w, h, img.strides[0] if img.size else 0,
                          kps.ctypes.data, desc.ctypes.data, cap)
        if n < 0:
            return n, None, None
        assert n <= cap
        return n, kps[:n].copy(), desc[:n].copy()


class OracleExtractor(_Extractor):
    prefix = "orbo_"

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        super().__init__(oracle(), nfeatures, scale_factor, nlevels, ini_th, min_th)
        self._extract = self.L.orbo_extract

    def level_size(self, level):
        w, h = C.c_int(), C.c_int()
        assert self.L.orbo_stage_level_size(self.h, level, C.byref(w), C.byref(h)) == 0
        return w.value, h.value

    def pyramid(self, level, with_border=False):
        w, h = self.level_size(level)
        b = 19 if with_border else 0
        out = np.zeros((h + 2 * b, w + 2 * b), np.uint8)
        assert self.L.orbo_stage_pyramid(self.h, level, int(with_border), out.ctypes.data, out.strides[0]) == 0
        return out

    def blurred(self, level):
        w, h = self.level_size(level)
        out = np.zeros((h, w), np.uint8)
        if self.L.orbo_stage_blurred(self.h, level, out.ctypes.data, out.strides[0]) != 0:
            return None
        return out

    def candidates(self, level):
        n = self.L.orbo_stage_candidates(self.h, level, None, 0)
        out = np.zeros(max(n, 1), CAND_DTYPE)
        self.L.orbo_stage_candidates(self.h, level, out.ctypes.data, n)
        return out[:n]

    def level_keypoints(self, level):
        n = self.L.orbo_stage_level_keypoints(self.h, level, None, 0)
        out = np.zeros(max(n, 1), KP_DTYPE)
        self.L.orbo_stage_level_keypoints(self.h, level, out.ctypes.data, n)
        return out[:n]


class RefExtractor(_Extractor):
    prefix = "orbref_extractor_"

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        lib = ref()
        assert lib is not None, "oracle/_ref/liborbref.so not built"
        super().__init__(lib, nfeatures, scale_factor, nlevels, ini_th, min_th)
        self._extract = self.L.orbref_extract

    def pyramid(self, level, with_border=False):
        w, h = C.c_int(), C.c_int()
        assert self.L.orbref_pyramid_level(self.h, level, int(with_border), None, 0, C.byref(w), C.byref(h)) == 0
        out = np.zeros((h.value, w.value), np.uint8)
        self.L.orbref_pyramid_level(self.h, level, int(with_border), out.ctypes.data, out.strides[0], None, None)
        return out

    def keypoints_octtree(self, img):
        img = np.ascontiguousarray(img)
        cap = 4 * self.nfeatures + 256
        kps = np.zeros(cap, KP_DTYPE)
        per = np.zeros(self.nlevels, np.int32)
        n = self.L.orbref_keypoints_octtree(self.h, img.ctypes.data, img.shape[1], img.shape[0], img.strides[0],
                                            kps.ctypes.data, cap, per.ctypes.data)
        return kps[:n].copy(), per

    def distribute(self, cand, min_x, max_x, min_y, max_y, n_want, level=0):
        kin = np.zeros(len(cand), KP_DTYPE)
        kin["x"], kin["y"], kin["response"] = cand["x"], cand["y"], cand["score"]
        kin["size"], kin["angle"], kin["class_id"] = 7, -1, -1
        out = np.zeros(len(cand) + 8, KP_DTYPE)
        n = self.L.orbref_distribute(self.h, kin.ctypes.data, len(kin), min_x, max_x, min_y, max_y, n_want, level,
                                     out.ctypes.data, len(out))
        return out[:n].copy()


def oracle_distribute(cand, min_x, max_x, min_y, max_y, n_want):
    cand = np.ascontiguousarray(cand)
    idx = np.zeros(len(cand) + 8, np.int32)
    n = oracle().orbo_distribute(cand.ctypes.data, len(cand), min_x, max_x, min_y, max_y, n_want, idx.ctypes.data, len(idx))
    return idx[:n].copy()
