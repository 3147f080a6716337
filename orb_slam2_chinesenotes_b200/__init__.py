"""B200-native ORB front end: Python face of liborb_b200.so (ctypes over the C ABI).

The product is the CUDA library; this module only loads it and mirrors the reference's
class interface (ORB_SLAM2::ORBextractor, include/ORBextractor.h:47-111, and the Hamming part
of ORB_SLAM2::ORBmatcher, include/ORBmatcher.h:43-100) so tests and benchmarks read like calls
into the reference.  There is no CPU fallback: if the library or a CUDA device is missing,
every compute call raises.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ORB_B200_LIB") or os.path.join(_HERE, "lib", "liborb_b200.so")   # override: a differently tuned build
CSRC_DIR = os.path.join(_HERE, "csrc")

KP_DTYPE = np.dtype([("x", "f4"), ("y", "f4"), ("size", "f4"), ("angle", "f4"), ("response", "f4"),
                     ("octave", "i4"), ("class_id", "i4")])
assert KP_DTYPE.itemsize == 28

ORBX_OK, ORBX_E_ARG, ORBX_E_SHAPE, ORBX_E_CAPACITY, ORBX_E_CUDA, ORBX_E_EMPTY = range(6)
STAGES = ("pyramid", "fast", "blur", "octree", "describe", "stereo")


class OrbmFrame(C.Structure):
    """orbm_frame of include/orb_b200.h: the Frame fields the window searches read."""
    _fields_ = [("n", C.c_int), ("kps", C.c_void_p), ("desc", C.c_void_p), ("u_right", C.c_void_p),
                ("min_x", C.c_float), ("max_x", C.c_float), ("min_y", C.c_float), ("max_y", C.c_float)]


class OrbmFrames(C.Structure):
    """orbm_frames: a batch of device-resident frames in the layout orbx_extract_batch writes."""
    _fields_ = [("nprob", C.c_int), ("kps", C.c_void_p), ("desc", C.c_void_p), ("u_right", C.c_void_p), ("n", C.c_void_p),
                ("kp_stride", C.c_int), ("min_x", C.c_float), ("max_x", C.c_float), ("min_y", C.c_float), ("max_y", C.c_float),
                ("max_n", C.c_int)]


class OrbmPoints(C.Structure):
    """orbm_points: device arrays [nprob][nq_stride] of the map points' tracking fields."""
    _fields_ = [("nq", C.c_void_p), ("nq_stride", C.c_int), ("proj_xyxr", C.c_void_p), ("level", C.c_void_p), ("view_cos", C.c_void_p),
                ("in_view", C.c_void_p), ("bad", C.c_void_p), ("observations", C.c_void_p), ("qdesc", C.c_void_p)]


class OrbmFreeWindows(C.Structure):
    """orbm_free_windows: device arrays [nprob][nq_stride] of the projected points of Fuse / SearchBySim3."""
    _fields_ = [("nq", C.c_void_p), ("nq_stride", C.c_int), ("uvr", C.c_void_p), ("level", C.c_void_p), ("ur", C.c_void_p),
                ("valid", C.c_void_p), ("qdesc", C.c_void_p)]


class OrbmWindows(C.Structure):
    """orbm_windows: device arrays [nprob][nq_stride] of already projected best-only queries."""
    _fields_ = [("nq", C.c_void_p), ("nq_stride", C.c_int), ("uvr", C.c_void_p), ("min_level", C.c_void_p), ("max_level", C.c_void_p),
                ("ur", C.c_void_p), ("er_max", C.c_void_p), ("valid", C.c_void_p), ("qdesc", C.c_void_p), ("q_angle", C.c_void_p),
                ("q_obs", C.c_void_p)]


class OrbmFeatVec(C.Structure):
    """orbm_featvec: DBoW2::FeatureVector per frame in CSR form (device arrays)."""
    _fields_ = [("node_id", C.c_void_p), ("node_off", C.c_void_p), ("n_nodes", C.c_void_p), ("feat", C.c_void_p), ("node_stride", C.c_int)]


class _FrameOut(C.Structure):
    """orbx_frame_out (include/orb_b200.h)."""
    _fields_ = [("n", C.c_int), ("kps", C.c_void_p), ("desc", C.c_void_p), ("nlevels", C.c_int), ("level", C.c_void_p * 16),
                ("level_w", C.c_int * 16), ("level_h", C.c_int * 16), ("level_pitch", C.c_size_t * 16)]


class OrbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"orb_b200 status {code}: {msg}")
        self.code = code


def build(verbose=False):
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", CSRC_DIR], capture_output=True, text=True)
    if verbose or r.returncode:
        print(r.stdout[-4000:], r.stderr[-4000:])
    if r.returncode:
        raise RuntimeError("building liborb_b200.so failed")
    return LIB_PATH


_lib = None


def lib():
    """The loaded C-ABI library.  Fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OSError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        vp, i32, f32, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
        pi = C.POINTER(C.c_int)
        L.orbx_create.argtypes = [C.POINTER(vp), i32, f32, i32, i32, i32, i32]
        L.orbx_destroy.argtypes = [vp]
        L.orbx_destroy.restype = None
        L.orbx_levels.argtypes = [vp]
        L.orbx_tables.argtypes = [vp] * 6
        L.orbx_shape_supported.argtypes = [vp, i32, i32]
        L.orbx_extract.argtypes = [vp, vp, i32, i32, sz, vp, vp, i32, vp]
        L.orbx_extract_batch.argtypes = [vp, vp, sz, i32, i32, i32, sz, vp, vp, i32, vp]
        L.orbx_extract_batch_async.argtypes = [vp, vp, sz, i32, i32, i32, sz, vp, vp, i32, vp]
        L.orbx_sync.argtypes = [vp]
        L.orbx_extract_stereo_batch.argtypes = [vp, vp, sz, i32, i32, i32, sz, vp, vp, i32, vp, f32, f32, vp, vp, vp]
        L.orbx_extract_stereo_batch_async.argtypes = [vp, vp, sz, i32, i32, i32, sz, vp, vp, i32, vp, f32, f32, vp, vp, vp]
        L.orbx_pyramid_level.argtypes = [vp, i32, i32, i32, vp, sz, pi, pi]
        L.orbx_set_stream.argtypes = [vp, vp]
        L.orbx_stream.argtypes = [vp]
        L.orbx_stream.restype = vp
        L.orbx_set_chunk.argtypes = [vp, i32]
        L.orbx_last_error.argtypes = [vp]
        L.orbx_last_error.restype = C.c_char_p
        L.orbx_debug_blurred.argtypes = [vp, i32, i32, vp, sz]
        L.orbx_debug_candidates.argtypes = [vp, i32, i32, vp, i32, pi]
        L.orbx_debug_level_keypoints.argtypes = [vp, i32, i32, vp, i32, pi]
        L.orbx_profile.argtypes = [vp, i32]
        L.orbx_stage_ms.argtypes = [vp, vp, vp, i32]
        L.orbx_plan_describe.argtypes = [i32, f32, i32, i32, i32, i32, i32] + [vp] * 6
        L.orbx_max_keypoints.argtypes = [i32, f32, i32, i32, i32, i32, i32]
        L.orbx_launches_per_chunk.argtypes = [vp, i32]
        L.orbx_extract_frame.argtypes = [vp, vp, i32, i32, sz, i32, vp]
        L.orbx_debug_last_call_us.argtypes = [vp, vp]
        L.orbm_hamming_bf.argtypes = [vp, i32, vp, i32, i32, vp, vp, vp, i32]
        L.orbm_hamming_bf_async.argtypes = [vp, i32, vp, i32, i32, vp, vp, vp, vp]
        fp = C.POINTER(OrbmFrame)
        L.orbm_search_by_projection_points.argtypes = [fp, vp, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, f32, f32, pi, i32]
        L.orbm_search_by_projection_frame.argtypes = [fp, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, f32, vp, i32, vp, vp, f32, i32, i32, pi, i32]
        L.orbm_search_for_initialization.argtypes = [fp, fp, vp, vp, i32, f32, i32, pi, i32]
        L.orbm_window_search_best.argtypes = [fp, i32] + [vp] * 11 + [i32, i32, pi, i32]
        L.orbm_window_best_free.argtypes = [fp, i32] + [vp] * 6 + [i32, i32, vp, vp, pi, i32]
        L.orbm_fuse_project_batch.argtypes = [i32, i32, vp, vp, f32, f32, f32, f32, f32, f32, vp, i32, f32, vp, i32, i32] + [vp] * 9 + [vp]
        L.orbx_undistort_keypoints_batch.argtypes = [vp, vp, vp, i32, i32, vp, vp, i32, vp]
        L.orbx_stereo_from_rgbd_batch.argtypes = [vp, vp, vp, i32, i32, vp, C.c_size_t, C.c_size_t, i32, i32, f32, vp, vp, vp]
        L.orbm_window_best_free_batch.argtypes = [C.POINTER(OrbmFrames), C.POINTER(OrbmFreeWindows), vp, i32, i32, vp, vp, vp, vp]
        L.orbm_search_by_projection_points_batch.argtypes = [C.POINTER(OrbmFrames), vp, i32, C.POINTER(OrbmPoints), vp, vp, f32, f32, vp, vp, vp]
        L.orbm_window_search_best_batch.argtypes = [C.POINTER(OrbmFrames), C.POINTER(OrbmWindows), vp, vp, i32, i32, vp, vp, vp]
        L.orbm_search_by_projection_frame_batch.argtypes = [C.POINTER(OrbmFrames), vp, vp, vp, f32, vp, i32, vp, i32] + [vp] * 8 + [f32, i32, i32, vp, vp]
        L.orbm_search_for_initialization_batch.argtypes = [C.POINTER(OrbmFrames), C.POINTER(OrbmFrames), vp, vp, i32, f32, i32, vp, vp, vp]
        L.orbm_project_points_batch.argtypes = [i32, vp, vp, f32, f32, f32, f32, f32, f32, i32, f32, vp, i32, i32] + [vp] * 9 + [vp]
        L.orbm_distinctive_descriptors.argtypes = [vp, vp, i32, vp, vp, vp, vp]
        L.orbm_search_by_bow_batch.argtypes = [C.POINTER(OrbmFrames), C.POINTER(OrbmFeatVec), vp, C.POINTER(OrbmFrames), C.POINTER(OrbmFeatVec), vp,
                                               i32, f32, i32, vp, vp, vp, vp, vp]
        L.orbm_search_by_bow.argtypes = [fp, vp, i32, vp, vp, vp, fp, vp, i32, vp, vp, vp, i32, f32, i32, vp, pi, i32]
        L.orbm_search_for_triangulation.argtypes = [fp, vp, i32, vp, vp, vp, fp, vp, i32, vp, vp, vp, vp, vp, vp, vp, i32, i32, vp, pi, i32]
        L.orbm_search_for_triangulation_batch.argtypes = [C.POINTER(OrbmFrames), C.POINTER(OrbmFeatVec), vp, C.POINTER(OrbmFrames), C.POINTER(OrbmFeatVec), vp,
                                                          vp, vp, vp, vp, i32, i32, vp, vp, vp]
        L.orbm_project_points.argtypes = [vp, vp, f32, f32, f32, f32, f32, f32, i32, f32, i32] + [vp] * 8 + [i32]
        L.orbm_distinctive_descriptor.argtypes = [vp, i32, vp, pi, pi, i32]
        L.orbm_stereo_matches.argtypes = [vp, i32, vp, i32, i32, vp, vp, i32, vp, vp, f32, f32, vp, vp, pi]
        _lib = L
    return _lib


def _ptr(a):
    """Raw address of a numpy array or a torch tensor (host or CUDA)."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()


def plan_describe(nfeatures, scale_factor, nlevels, ini_th, min_th, w, h):
    """Host-only geometry of the extractor for one image shape (no GPU needed)."""
    arrs = [np.zeros(nlevels, np.int32) for _ in range(6)]
    rc = lib().orbx_plan_describe(nfeatures, scale_factor, nlevels, ini_th, min_th, w, h, *(a.ctypes.data for a in arrs))
    if rc:
        raise OrbError(rc, "unsupported shape" if rc == ORBX_E_SHAPE else "bad argument")
    return dict(zip(("level_w", "level_h", "cells", "quota", "n_ini", "cand_cap"), arrs))


class ORBextractor:
    """ORB_SLAM2::ORBextractor (include/ORBextractor.h:47-111) on one B200.

    ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST) and
    extractor(image, mask) -> (keypoints, descriptors) follow the reference; extract_batch
    runs many equally sized frames per launch.
    """

    HARRIS_SCORE, FAST_SCORE = 0, 1

    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, device=0):
        self._L = lib()
        self._h = C.c_void_p()
        self._profiling = False
        self.nfeatures, self.nlevels = int(nfeatures), int(nlevels)
        rc = self._L.orbx_create(C.byref(self._h), nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, device)
        if rc:
            self._h = C.c_void_p()
            raise OrbError(rc, "orbx_create failed (no CUDA device?)" if rc == ORBX_E_CUDA else "bad extractor parameters")
        self._scale_factor = float(np.float32(scaleFactor))
        self.device = device

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.orbx_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def _check(self, rc, allow=()):
        if rc and rc not in allow:
            raise OrbError(rc, self._L.orbx_last_error(self._h).decode())
        return rc

    # ---- accessors, include/ORBextractor.h:64-84
    def _tables(self):
        n = self.nlevels
        out = [np.zeros(n, np.float32) for _ in range(4)] + [np.zeros(n, np.int32)]
        self._check(self._L.orbx_tables(self._h, *(a.ctypes.data for a in out)))
        return out

    def GetLevels(self):
        return self._L.orbx_levels(self._h)

    def GetScaleFactor(self):
        return self._scale_factor

    def GetScaleFactors(self):
        return self._tables()[0]

    def GetInverseScaleFactors(self):
        return self._tables()[1]

    def GetScaleSigmaSquares(self):
        return self._tables()[2]

    def GetInverseScaleSigmaSquares(self):
        return self._tables()[3]

    def features_per_level(self):
        return self._tables()[4]

    def default_capacity(self):
        # DistributeOctTree returns >= N per level, at most N+2 (or 4*nIni): leave generous room
        return self.nfeatures + 4 * self.nlevels + 64

    # ---- operator(), include/ORBextractor.h:60
    def __call__(self, image, mask=None, capacity=None):
        """Returns (keypoints [n] structured as cv::KeyPoint, descriptors [n,32] uint8).
        An empty image leaves "outputs untouched" in the reference; here it returns (None, None)."""
        image = np.asarray(image)
        if image.size == 0:
            return None, None
        assert image.dtype == np.uint8 and image.ndim == 2, "CV_8UC1 expected (src/ORBextractor.cc:1091)"
        if image.strides[1] != 1 or image.strides[0] < image.shape[1]:   # also views with a negative / overlapping row stride
            image = np.ascontiguousarray(image)
        h, w = image.shape
        if capacity is None and not self._profiling:
            # orbx_extract_frame: the single-call CUDA graph; the results are copied out of the context's pinned memory
            out = _FrameOut()
            self._check(self._L.orbx_extract_frame(self._h, image.ctypes.data, w, h, image.strides[0], 0, C.byref(out)))
            k = out.n
            kps = np.empty(k, KP_DTYPE)
            desc = np.empty((k, 32), np.uint8)
            if k:
                C.memmove(kps.ctypes.data, out.kps, k * 28)
                C.memmove(desc.ctypes.data, out.desc, k * 32)
            return kps, desc
        cap = capacity or self.default_capacity()
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = np.zeros(1, np.int32)
        self._check(self._L.orbx_extract(self._h, image.ctypes.data, w, h, image.strides[0], kps.ctypes.data,
                                         desc.ctypes.data, cap, n.ctypes.data))
        k = int(n[0])
        return kps[:k].copy(), desc[:k].copy()

    def extract_batch(self, frames, capacity=None, out=None):
        """frames: [B,H,W] uint8 numpy array (host) -> (kps [B,cap], desc [B,cap,32], n [B])."""
        frames = np.ascontiguousarray(frames)
        assert frames.dtype == np.uint8 and frames.ndim == 3
        b, h, w = frames.shape
        cap = capacity or self.default_capacity()
        if out is None:
            out = (np.zeros((b, cap), KP_DTYPE), np.zeros((b, cap, 32), np.uint8), np.zeros(b, np.int32))
        kps, desc, n = out
        self._check(self._L.orbx_extract_batch(self._h, frames.ctypes.data, h * w, b, w, h, w, kps.ctypes.data,
                                               desc.ctypes.data, cap, n.ctypes.data))
        return kps, desc, n

    def extract_batch_raw(self, imgs, frame_stride, batch, w, h, pitch, kps, desc, cap, n_out, asynchronous=False):
        """Thin pass-through of orbx_extract_batch[_async]: every argument may be numpy or torch."""
        fn = self._L.orbx_extract_batch_async if asynchronous else self._L.orbx_extract_batch
        return self._check(fn(self._h, _ptr(imgs), frame_stride, batch, w, h, pitch, _ptr(kps), _ptr(desc), cap, _ptr(n_out)))

    def extract_stereo_batch(self, frames, bf, fx, capacity=None):
        """The stereo Frame constructor's front end (src/Frame.cc:61-115) for rectified pairs stored as
        L0,R0,L1,R1,...: frames [2P,H,W] uint8 -> (kps [2P,cap], desc [2P,cap,32], n [2P], u_right [P,cap],
        depth [P,cap], n_stereo [P]); u_right/depth are mvuRight/mvDepth of the left keypoints."""
        frames = np.ascontiguousarray(frames)
        assert frames.dtype == np.uint8 and frames.ndim == 3 and frames.shape[0] % 2 == 0
        b, h, w = frames.shape
        cap = capacity or self.default_capacity()
        kps, desc, n = np.zeros((b, cap), KP_DTYPE), np.zeros((b, cap, 32), np.uint8), np.zeros(b, np.int32)
        ur, dep, ns = np.full((b // 2, cap), -1, np.float32), np.full((b // 2, cap), -1, np.float32), np.zeros(b // 2, np.int32)
        self._check(self._L.orbx_extract_stereo_batch(self._h, frames.ctypes.data, h * w, b // 2, w, h, w, kps.ctypes.data,
                                                      desc.ctypes.data, cap, n.ctypes.data, bf, fx, ur.ctypes.data,
                                                      dep.ctypes.data, ns.ctypes.data))
        return kps, desc, n, ur, dep, ns

    def extract_stereo_batch_raw(self, imgs, frame_stride, pairs, w, h, pitch, kps, desc, cap, n_out, bf, fx,
                                 u_right, depth, n_stereo, asynchronous=False):
        """Thin pass-through of orbx_extract_stereo_batch[_async]: every argument may be numpy or torch."""
        fn = self._L.orbx_extract_stereo_batch_async if asynchronous else self._L.orbx_extract_stereo_batch
        return self._check(fn(self._h, _ptr(imgs), frame_stride, pairs, w, h, pitch, _ptr(kps), _ptr(desc), cap, _ptr(n_out),
                              bf, fx, _ptr(u_right), _ptr(depth), _ptr(n_stereo)))

    def sync(self):
        self._check(self._L.orbx_sync(self._h))

    def set_stream(self, cuda_stream_handle):
        self._check(self._L.orbx_set_stream(self._h, cuda_stream_handle))

    def set_chunk(self, frames):
        self._check(self._L.orbx_set_chunk(self._h, frames))

    # ---- mvImagePyramid, include/ORBextractor.h:86
    def pyramid(self, level, frame=0, with_border=False):
        w, h = C.c_int(), C.c_int()
        self._check(self._L.orbx_pyramid_level(self._h, frame, level, int(with_border), None, 0, C.byref(w), C.byref(h)))
        out = np.zeros((h.value, w.value), np.uint8)
        self._check(self._L.orbx_pyramid_level(self._h, frame, level, int(with_border), out.ctypes.data, out.strides[0], None, None))
        return out

    @property
    def mvImagePyramid(self):
        return [self.pyramid(l) for l in range(self.nlevels)]

    # ---- stage taps
    def blurred(self, level, frame=0):
        ref = self.pyramid(level, frame)
        out = np.zeros_like(ref)
        self._check(self._L.orbx_debug_blurred(self._h, frame, level, out.ctypes.data, out.strides[0]))
        return out

    def _packed(self, fn, level, frame):
        n = C.c_int()
        self._check(fn(self._h, frame, level, None, 0, C.byref(n)))
        out = np.zeros((max(n.value, 1), 3), np.int32)
        self._check(fn(self._h, frame, level, out.ctypes.data, n.value, C.byref(n)))
        return out[:n.value]

    def candidates(self, level, frame=0):
        return self._packed(self._L.orbx_debug_candidates, level, frame)

    def level_keypoints(self, level, frame=0):
        return self._packed(self._L.orbx_debug_level_keypoints, level, frame)

    def launches_per_chunk(self, stereo=False):
        """Kernels one chunk of frames launches for the shape last seen (orbx_launches_per_chunk)."""
        r = self._L.orbx_launches_per_chunk(self._h, int(stereo))
        if r < 0:
            raise OrbError(-r, "orbx_launches_per_chunk")
        return r

    def profile(self, enable=True):
        self._profiling = bool(enable)
        self._check(self._L.orbx_profile(self._h, int(enable)))

    def stage_ms(self, reset=True):
        ms = np.zeros(len(STAGES), np.float32)
        cnt = np.zeros(len(STAGES), np.int32)
        self._check(self._L.orbx_stage_ms(self._h, ms.ctypes.data, cnt.ctypes.data, int(reset)))
        return dict(zip(STAGES, ms.tolist())), dict(zip(STAGES, cnt.tolist()))


def _np(a, dtype=None):
    return None if a is None else np.ascontiguousarray(a, dtype)


class FrameView:
    """Host-side stand-in for the ORB_SLAM2::Frame members the matchers read: mvKeysUn, mDescriptors,
    mvuRight, mnMinX..mnMaxY (include/Frame.h:120-175)."""

    def __init__(self, kps, desc, bounds, u_right=None):
        self.kps = _np(kps)
        self.desc = _np(desc, np.uint8)
        self.u_right = _np(u_right, np.float32)
        self.bounds = tuple(float(b) for b in bounds)
        assert self.kps.dtype == KP_DTYPE and self.desc.shape == (len(self.kps), 32)

    def struct(self):
        return OrbmFrame(len(self.kps), self.kps.ctypes.data, self.desc.ctypes.data,
                         None if self.u_right is None else self.u_right.ctypes.data, *self.bounds)


class ResidentFrame(FrameView):
    """A FrameView whose keypoints / descriptors / mvuRight were uploaded once (orbm_frame_upload): every search that takes
    a FrameView takes this instead and skips the upload -- Tracking runs two to four searches per Frame."""

    def __init__(self, kps, desc, bounds, u_right=None, device=0):
        super().__init__(kps, desc, bounds, u_right)
        self._h = C.c_void_p()
        L = lib()
        L.orbm_frame_upload.argtypes = [C.POINTER(OrbmFrame), C.c_int, C.POINTER(C.c_void_p)]
        L.orbm_frame_view.argtypes = [C.c_void_p, C.POINTER(OrbmFrame)]
        L.orbm_frame_release.argtypes = [C.c_void_p]
        L.orbm_frame_release.restype = None
        host = FrameView.struct(self)
        rc = L.orbm_frame_upload(C.byref(host), device, C.byref(self._h))
        if rc:
            raise OrbError(rc, "orbm_frame_upload failed")

    def struct(self, with_right=True):
        v = OrbmFrame()
        rc = lib().orbm_frame_view(self._h, C.byref(v))
        if rc:
            raise OrbError(rc, "orbm_frame_view failed")
        if not with_right:
            v.u_right = None
        return v

    def close(self):
        if self._h:
            lib().orbm_frame_release(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ORBmatcher:
    """ORB_SLAM2::ORBmatcher (include/ORBmatcher.h:43-100): the Hamming searches of the hot path."""
    TH_LOW, TH_HIGH, HISTO_LENGTH = 50, 100, 30   # src/ORBmatcher.cc:37-39

    def __init__(self, nnratio=0.6, checkOri=True, device=0):
        self.mfNNratio, self.mbCheckOrientation, self.device = float(np.float32(nnratio)), bool(checkOri), device

    @staticmethod
    def DescriptorDistance(a, b, device=0):
        """src/ORBmatcher.cc:46-63 for one pair (runs the brute-force kernel on a 1x1 problem)."""
        _, d, _ = hamming_bf(np.asarray(a, np.uint8).reshape(1, 32), np.asarray(b, np.uint8).reshape(1, 32), device)
        return int(d[0])

    def _rc(self, rc, what):
        if rc:
            raise OrbError(rc, what + " failed")

    def SearchByProjection(self, F, scale, q, th=3.0, init_assign=None):
        """SearchByProjection(Frame&, const vector<MapPoint*>&, th), src/ORBmatcher.cc:73-157.
        q: dict of per-map-point arrays proj [nq,3], level, view_cos, in_view, bad, obs, desc [nq,32].
        Returns (nmatches, assign [F.n])."""
        scale = _np(scale, np.float32)
        proj, level, vc = _np(q["proj"], np.float32), _np(q["level"], np.int32), _np(q["view_cos"], np.float32)
        iv, bad, obs, qd = _np(q["in_view"], np.uint8), _np(q["bad"], np.uint8), _np(q["obs"], np.int32), _np(q["desc"], np.uint8)
        ia = _np(init_assign, np.int32)
        out = np.zeros(len(F.kps), np.int32)
        nm = C.c_int()
        fs = F.struct()
        self._rc(lib().orbm_search_by_projection_points(C.byref(fs), scale.ctypes.data, len(scale), len(level), proj.ctypes.data,
                 level.ctypes.data, vc.ctypes.data, iv.ctypes.data, bad.ctypes.data, obs.ctypes.data, qd.ctypes.data,
                 None if ia is None else ia.ctypes.data, out.ctypes.data, th, self.mfNNratio, C.byref(nm), self.device),
                 "orbm_search_by_projection_points")
        return nm.value, out

    def SearchByProjectionFrame(self, cur, last, Tcw_cur, Tcw_last, K, bf, scale, th, bMono, cur_init_obs=None):
        """SearchByProjection(Frame& cur, const Frame& last, th, bMono), src/ORBmatcher.cc:160-300.
        last: dict kps, has_mp, outlier, xyz, mp_desc, mp_obs.  Returns (nmatches, assign [cur.n])."""
        scale, K = _np(scale, np.float32), _np(K, np.float32)
        Tc, Tl = _np(Tcw_cur, np.float32), _np(Tcw_last, np.float32)
        lk, hm, ol = _np(last["kps"]), _np(last["has_mp"], np.uint8), _np(last["outlier"], np.uint8)
        xyz, md, mo = _np(last["xyz"], np.float32), _np(last["mp_desc"], np.uint8), _np(last["mp_obs"], np.int32)
        io = _np(cur_init_obs, np.int32)
        out = np.zeros(len(cur.kps), np.int32)
        nm = C.c_int()
        fs = cur.struct()
        self._rc(lib().orbm_search_by_projection_frame(C.byref(fs), len(lk), lk.ctypes.data, hm.ctypes.data, ol.ctypes.data, xyz.ctypes.data,
                 md.ctypes.data, mo.ctypes.data, Tc.ctypes.data, Tl.ctypes.data, K.ctypes.data, bf, scale.ctypes.data, len(scale),
                 None if io is None else io.ctypes.data, out.ctypes.data, th, int(bMono), int(self.mbCheckOrientation), C.byref(nm), self.device),
                 "orbm_search_by_projection_frame")
        return nm.value, out

    def SearchForInitialization(self, F1, F2, vbPrevMatched, windowSize=10):
        """src/ORBmatcher.cc:1055-1180.  Returns (nmatches, vnMatches12, updated vbPrevMatched)."""
        prev = np.ascontiguousarray(vbPrevMatched, np.float32).copy()
        m12 = np.zeros(len(F1.kps), np.int32)
        nm = C.c_int()
        f1, f2 = F1.struct(), F2.struct()
        self._rc(lib().orbm_search_for_initialization(C.byref(f1), C.byref(f2), prev.ctypes.data, m12.ctypes.data, windowSize,
                 self.mfNNratio, int(self.mbCheckOrientation), C.byref(nm), self.device), "orbm_search_for_initialization")
        return nm.value, m12, prev


def search_by_bow(A, fv_a, a_valid, B, fv_b, b_valid, kf_kf, nnratio, check_ori, device=0):
    """orbm_search_by_bow: ORBmatcher::SearchByBoW (src/ORBmatcher.cc:552-832) for one pair with host arrays.  A, B:
    FrameView; fv_*: (node_id, node_off, feat) int32 arrays; *_valid: uint8 per feature (b_valid None for the
    KeyFrame/Frame overload).  Returns (nmatches, match12 [A.n])."""
    (ia, oa, fa), (ib, ob, fb) = [tuple(_np(x, np.int32) for x in fv) for fv in (fv_a, fv_b)]
    av, bv = _np(a_valid, np.uint8), _np(b_valid, np.uint8)
    out = np.zeros(max(len(A.kps), 1), np.int32)
    nm = C.c_int()
    sa, sb = A.struct(), B.struct()
    p = lambda a: None if a is None else a.ctypes.data
    rc = lib().orbm_search_by_bow(C.byref(sa), p(av), len(ia), p(ia), p(oa), p(fa), C.byref(sb), p(bv), len(ib), p(ib), p(ob), p(fb),
                                  int(kf_kf), float(np.float32(nnratio)), int(check_ori), p(out), C.byref(nm), device)
    if rc:
        raise OrbError(rc, "orbm_search_by_bow failed")
    return nm.value, out[:len(A.kps)]


def search_for_triangulation(A, fv_a, a_valid, B, fv_b, b_valid, F12, epipole, scale, sigma2, check_ori, device=0):
    """orbm_search_for_triangulation: ORBmatcher::SearchForTriangulation (src/ORBmatcher.cc:1183-1361) for one key-frame pair
    with host arrays.  A, B: FrameView (u_right = mvuRight or None); *_valid: the feature has no map point yet (and is a stereo
    one when bOnlyStereo); F12 3x3, epipole (ex, ey), scale / sigma2 = pKF2's tables.  Returns (nmatches, match12 [A.n])."""
    (ia, oa, fa), (ib, ob, fb) = [tuple(_np(x, np.int32) for x in fv) for fv in (fv_a, fv_b)]
    av, bv = _np(a_valid, np.uint8), _np(b_valid, np.uint8)
    F12, epipole, scale, sigma2 = _np(F12, np.float32), _np(epipole, np.float32), _np(scale, np.float32), _np(sigma2, np.float32)
    out = np.zeros(max(len(A.kps), 1), np.int32)
    nm = C.c_int()
    sa, sb = A.struct(), B.struct()
    p = lambda a: None if a is None else a.ctypes.data
    rc = lib().orbm_search_for_triangulation(C.byref(sa), p(av), len(ia), p(ia), p(oa), p(fa), C.byref(sb), p(bv), len(ib), p(ib), p(ob), p(fb),
                                             p(F12), p(epipole), p(scale), p(sigma2), len(scale), int(check_ori), p(out), C.byref(nm), device)
    if rc:
        raise OrbError(rc, "orbm_search_for_triangulation failed")
    return nm.value, out[:len(A.kps)]


def project_points(Tcw, K, bf, bounds, scale_factor, nlevels, xyz, normal, max_distance, min_distance, cos_limit=0.5, device=0):
    """orbm_project_points: Frame::isInFrustum (src/Frame.cc:288-345) for n map points of one frame, host arrays.
    Returns (in_view u8 [n], proj [n,3], level [n], view_cos [n]); entries of points not in view are zero."""
    Tcw, K = _np(Tcw, np.float32), _np(K, np.float32)
    xyz, normal, mx, mn = _np(xyz, np.float32), _np(normal, np.float32), _np(max_distance, np.float32), _np(min_distance, np.float32)
    n = len(mx)
    iv, proj, lv, vc = np.zeros(n, np.uint8), np.zeros((n, 3), np.float32), np.zeros(n, np.int32), np.zeros(n, np.float32)
    rc = lib().orbm_project_points(Tcw.ctypes.data, K.ctypes.data, float(bf), *(float(b) for b in bounds), float(np.float32(scale_factor)),
                                   nlevels, cos_limit, n, xyz.ctypes.data, normal.ctypes.data, mx.ctypes.data, mn.ctypes.data,
                                   iv.ctypes.data, proj.ctypes.data, lv.ctypes.data, vc.ctypes.data, device)
    if rc:
        raise OrbError(rc, "orbm_project_points failed")
    return iv, proj, lv, vc


def distinctive_descriptor(desc, bad=None, device=0):
    """orbm_distinctive_descriptor: MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:275-340) for one map
    point's observations desc [n,32].  Returns (index or -1, median or -1)."""
    desc, bad = _np(desc, np.uint8), _np(bad, np.uint8)
    bi, bm = C.c_int(), C.c_int()
    rc = lib().orbm_distinctive_descriptor(desc.ctypes.data, len(desc), None if bad is None else bad.ctypes.data, C.byref(bi), C.byref(bm), device)
    if rc:
        raise OrbError(rc, "orbm_distinctive_descriptor failed")
    return bi.value, bm.value


def window_search_best(F, uvr, min_level, max_level, qdesc, th_accept, check_ori=False, q_angle=None, q_obs=None,
                       init_obs=None, ur=None, er_max=None, valid=None, device=0):
    """orbm_window_search_best: the best-candidate-only window search shared by the projection overloads of
    ORBmatcher (src/ORBmatcher.cc:160-300, 303-431, 434-549) once the caller has projected its points.
    Returns (nmatches, assign [F.n])."""
    uvr, qdesc = _np(uvr, np.float32), _np(qdesc, np.uint8)
    minl, maxl = _np(min_level, np.int32), _np(max_level, np.int32)
    qa, qo, io = _np(q_angle, np.float32), _np(q_obs, np.int32), _np(init_obs, np.int32)
    ur, em, va = _np(ur, np.float32), _np(er_max, np.float32), _np(valid, np.uint8)
    out = np.zeros(len(F.kps), np.int32)
    nm = C.c_int()
    fs = F.struct()
    p = lambda a: None if a is None else a.ctypes.data
    L = lib()
    rc = L.orbm_window_search_best(C.byref(fs), len(minl), p(uvr), p(minl), p(maxl), p(ur), p(em), p(va), p(qdesc), p(qa), p(qo), p(io),
                                   p(out), int(th_accept), int(check_ori), C.byref(nm), device)
    if rc:
        raise OrbError(rc, "orbm_window_search_best failed")
    return nm.value, out


def window_best_free(F, uvr, level, qdesc, th_accept, ur=None, valid=None, inv_sigma2=None, device=0):
    """orbm_window_best_free: the search ORBmatcher::Fuse (src/ORBmatcher.cc:1364-1513, :1516-1633) and SearchBySim3
    (:836-1052) run for every projected map point -- closest descriptor in the window at the predicted level or the one
    below, optionally behind Fuse's chi-square test (inv_sigma2 = mvInvLevelSigma2, ur = projected right coordinates).
    F: the KEY FRAME's keypoints with its (integer) image bounds.  Returns (nfound, best_idx [nq], best_dist [nq])."""
    uvr, qdesc, level = _np(uvr, np.float32), _np(qdesc, np.uint8), _np(level, np.int32)
    ur, va, s2 = _np(ur, np.float32), _np(valid, np.uint8), _np(inv_sigma2, np.float32)
    nq = len(level)
    bi, bd = np.zeros(nq, np.int32), np.zeros(nq, np.int32)
    nf = C.c_int()
    fs = F.struct()
    p = lambda a: None if a is None else a.ctypes.data
    rc = lib().orbm_window_best_free(C.byref(fs), nq, p(uvr), p(level), p(ur), p(va), p(qdesc), p(s2), 0 if s2 is None else len(s2),
                                     int(th_accept), p(bi), p(bd), C.byref(nf), device)
    if rc:
        raise OrbError(rc, "orbm_window_best_free failed")
    return nf.value, bi, bd


def window_best_free_batch(frames, q, nq, nq_stride, best_idx, best_dist, nfound, th_accept, inv_sigma2=None, stream=None):
    """orbm_window_best_free_batch.  q: dict of CUDA tensors uvr, level, desc and optionally ur, valid, each
    [P,nq_stride(,3|32)]; best_idx / best_dist [P,nq_stride], nfound [P] int32 CUDA tensors.  Only enqueues."""
    w = OrbmFreeWindows(_ptr(nq), nq_stride, _ptr(q["uvr"]), _ptr(q["level"]), _ptr(q.get("ur")), _ptr(q.get("valid")), _ptr(q["desc"]))
    s2 = None if inv_sigma2 is None else np.ascontiguousarray(inv_sigma2, np.float32)
    rc = lib().orbm_window_best_free_batch(C.byref(frames), C.byref(w), None if s2 is None else s2.ctypes.data, 0 if s2 is None else len(s2),
                                           int(th_accept), _ptr(best_idx), _ptr(best_dist), _ptr(nfound), stream)
    if rc:
        raise OrbError(rc, "orbm_window_best_free_batch failed")


def undistort_keypoints_batch(kps, kps_un, n, K, dist_coef, stream=None):
    """orbx_undistort_keypoints_batch: Frame::UndistortKeyPoints (src/Frame.cc:436-468) for every frame of a batch.  kps,
    kps_un: CUDA tensors [P,cap] of 28-byte records (kps_un may be kps), n [P] int32; K = fx, fy, cx, cy and dist_coef host."""
    K = np.ascontiguousarray(K, np.float32)
    d = np.ascontiguousarray(dist_coef, np.float32).ravel()
    rc = lib().orbx_undistort_keypoints_batch(_ptr(kps), _ptr(kps_un), _ptr(n), int(kps.shape[1]), int(kps.shape[0]), K.ctypes.data,
                                              d.ctypes.data if len(d) else None, len(d), stream)
    if rc:
        raise OrbError(rc, "orbx_undistort_keypoints_batch failed")


def stereo_from_rgbd_batch(kps, kps_un, n, depth, bf, u_right, depth_out, stream=None):
    """orbx_stereo_from_rgbd_batch: Frame::ComputeStereoFromRGBD (src/Frame.cc:702-727).  depth: CUDA float32 tensor [P,h,w]
    (any row / frame stride); kps_un may be None (= kps); u_right, depth_out: CUDA float32 [P,cap]."""
    P, h, w = (int(x) for x in depth.shape)
    es = depth.element_size()
    rc = lib().orbx_stereo_from_rgbd_batch(_ptr(kps), _ptr(kps_un), _ptr(n), int(kps.shape[1]), P, _ptr(depth), depth.stride(1) * es,
                                           depth.stride(0) * es, w, h, float(np.float32(bf)), _ptr(u_right), _ptr(depth_out), stream)
    if rc:
        raise OrbError(rc, "orbx_stereo_from_rgbd_batch failed")


def fuse_project_batch(pose, K, bf, bounds, scale_factor, scale, th, nq, nq_stride, pts, out, sim3=False, points_shared=False, stream=None):
    """orbm_fuse_project_batch: the projection / gates / PredictScale / radius prologue of ORBmatcher::Fuse (or, sim3, one direction
    of SearchBySim3) for every (key frame, map point).  pose [P,24], nq [P] and pts = dict(xyz, normal (None with sim3), max_d,
    min_d, skip (or absent)) are CUDA tensors; out = dict(uvr, level, ur (or absent), valid) CUDA tensors [P,nq_stride(,3)] -- the
    arrays window_best_free_batch takes.  Only enqueues."""
    K, scale = np.ascontiguousarray(K, np.float32), np.ascontiguousarray(scale, np.float32)
    rc = lib().orbm_fuse_project_batch(int(pose.shape[0]), int(sim3), _ptr(pose), K.ctypes.data, float(np.float32(bf)), *(float(b) for b in bounds),
                                       float(np.float32(scale_factor)), scale.ctypes.data, len(scale), float(np.float32(th)), _ptr(nq), nq_stride,
                                       int(points_shared), _ptr(pts["xyz"]), _ptr(pts.get("normal")), _ptr(pts["max_d"]), _ptr(pts["min_d"]),
                                       _ptr(pts.get("skip")), _ptr(out["uvr"]), _ptr(out["level"]), _ptr(out.get("ur")), _ptr(out["valid"]), stream)
    if rc:
        raise OrbError(rc, "orbm_fuse_project_batch failed")


def frames_batch(kps, desc, n, bounds, u_right=None, max_n=0):
    """orbm_frames over CUDA tensors laid out as orbx_extract_batch leaves them: kps [P,cap] (28-byte records, any
    dtype), desc [P,cap,32] uint8, n [P] int32, u_right [P,cap] float32 or None; max_n: a guaranteed bound on n
    (max_keypoints(...)), 0 = cap."""
    nprob, cap = int(desc.shape[0]), int(desc.shape[1])
    return OrbmFrames(nprob, _ptr(kps), _ptr(desc), _ptr(u_right), _ptr(n), cap, *(float(b) for b in bounds), int(max_n))


def max_keypoints(nfeatures, scale_factor, nlevels, ini_th, min_th, w, h):
    """orbx_max_keypoints: the largest keypoint count operator() can return for one w x h image (host only)."""
    r = lib().orbx_max_keypoints(nfeatures, scale_factor, nlevels, ini_th, min_th, w, h)
    if r < 0:
        raise OrbError(-r, "unsupported shape" if -r == ORBX_E_SHAPE else "bad argument")
    return r


def search_by_projection_points_batch(frames, scale, q, nq, nq_stride, assign_out, nmatches, th, nnratio, init_assign=None,
                                      rounds=None, stream=None):
    """orbm_search_by_projection_points_batch (src/ORBmatcher.cc:73-157 for every problem).  q: dict of CUDA tensors
    [P,nq_stride(,3|32)] proj, level, view_cos, in_view, bad, obs, desc; nq [P] int32.  Only enqueues."""
    scale = np.ascontiguousarray(scale, np.float32)
    pts = OrbmPoints(_ptr(nq), nq_stride, _ptr(q["proj"]), _ptr(q["level"]), _ptr(q["view_cos"]), _ptr(q["in_view"]), _ptr(q["bad"]),
                     _ptr(q["obs"]), _ptr(q["desc"]))
    rc = lib().orbm_search_by_projection_points_batch(C.byref(frames), scale.ctypes.data, len(scale), C.byref(pts), _ptr(init_assign),
                                                      _ptr(assign_out), th, float(np.float32(nnratio)), _ptr(nmatches), _ptr(rounds), stream)
    if rc:
        raise OrbError(rc, "orbm_search_by_projection_points_batch failed")


def window_search_best_batch(frames, q, nq, nq_stride, assign_out, nmatches, th_accept, check_ori, init_obs=None, rounds=None,
                             stream=None):
    """orbm_window_search_best_batch.  q: dict of CUDA tensors uvr, min_level, max_level, desc and optionally ur, er_max,
    valid, q_angle, q_obs, each [P,nq_stride(,3|32)].  Only enqueues."""
    w = OrbmWindows(_ptr(nq), nq_stride, _ptr(q["uvr"]), _ptr(q["min_level"]), _ptr(q["max_level"]), _ptr(q.get("ur")),
                    _ptr(q.get("er_max")), _ptr(q.get("valid")), _ptr(q["desc"]), _ptr(q.get("q_angle")), _ptr(q.get("q_obs")))
    rc = lib().orbm_window_search_best_batch(C.byref(frames), C.byref(w), _ptr(init_obs), _ptr(assign_out), int(th_accept),
                                             int(check_ori), _ptr(nmatches), _ptr(rounds), stream)
    if rc:
        raise OrbError(rc, "orbm_window_search_best_batch failed")


def search_by_projection_frame_batch(cur, Tcw_cur, Tcw_last, K, bf, scale, n_last, last_stride, last, assign_out, nmatches, th, bMono,
                                     check_ori, cur_init_obs=None, stream=None):
    """orbm_search_by_projection_frame_batch (src/ORBmatcher.cc:160-300 for every (current, last) frame pair).  cur:
    frames_batch(...); last: dict of CUDA tensors [P,last_stride(,..)] kps, has_mp, outlier (or None), xyz, mp_desc, mp_obs (or None)."""
    K, scale = np.ascontiguousarray(K, np.float32), np.ascontiguousarray(scale, np.float32)
    rc = lib().orbm_search_by_projection_frame_batch(C.byref(cur), _ptr(Tcw_cur), _ptr(Tcw_last), K.ctypes.data, float(bf), scale.ctypes.data,
                                                     len(scale), _ptr(n_last), last_stride, _ptr(last["kps"]), _ptr(last["has_mp"]),
                                                     _ptr(last.get("outlier")), _ptr(last["xyz"]), _ptr(last["mp_desc"]), _ptr(last.get("mp_obs")),
                                                     _ptr(cur_init_obs), _ptr(assign_out), th, int(bMono), int(check_ori), _ptr(nmatches), stream)
    if rc:
        raise OrbError(rc, "orbm_search_by_projection_frame_batch failed")


def search_for_initialization_batch(F1, F2, prev_matched, matches12, nmatches, window, nnratio, check_ori, rounds=None, stream=None):
    """orbm_search_for_initialization_batch (src/ORBmatcher.cc:1055-1180 for every (F1, F2) pair): F1, F2 frames_batch(...);
    prev_matched [P,cap1,2] float32 (updated in place), matches12 [P,cap1] int32, nmatches [P] int32 CUDA tensors."""
    rc = lib().orbm_search_for_initialization_batch(C.byref(F1), C.byref(F2), _ptr(prev_matched), _ptr(matches12), int(window),
                                                    float(np.float32(nnratio)), int(check_ori), _ptr(nmatches), _ptr(rounds), stream)
    if rc:
        raise OrbError(rc, "orbm_search_for_initialization_batch failed")


def project_points_batch(Tcw, K, bf, bounds, scale_factor, nlevels, nq, nq_stride, pts, out, cos_limit=0.5, points_shared=False,
                         n_in_view=None, stream=None):
    """orbm_project_points_batch: Frame::isInFrustum + MapPoint::PredictScale for every (frame, map point).
    Tcw [P,16], nq [P] and pts = dict(xyz, normal, max_d, min_d) are CUDA tensors; out = dict(in_view, proj, level,
    view_cos) CUDA tensors [P,nq_stride(,3)] -- the same arrays search_by_projection_points_batch takes.  Only enqueues."""
    K = np.ascontiguousarray(K, np.float32)
    rc = lib().orbm_project_points_batch(int(Tcw.shape[0]), _ptr(Tcw), K.ctypes.data, float(bf), *(float(b) for b in bounds),
                                         float(np.float32(scale_factor)), nlevels, cos_limit, _ptr(nq), nq_stride, int(points_shared),
                                         _ptr(pts["xyz"]), _ptr(pts["normal"]), _ptr(pts["max_d"]), _ptr(pts["min_d"]),
                                         _ptr(out["in_view"]), _ptr(out["proj"]), _ptr(out["level"]), _ptr(out["view_cos"]),
                                         _ptr(n_in_view), stream)
    if rc:
        raise OrbError(rc, "orbm_project_points_batch failed")


def distinctive_descriptors(desc, offsets, best_idx, best_median=None, bad=None, stream=None):
    """orbm_distinctive_descriptors: MapPoint::ComputeDistinctiveDescriptors for len(offsets)-1 map points whose
    observations are desc[offsets[p]:offsets[p+1]] (CUDA tensors).  Only enqueues."""
    rc = lib().orbm_distinctive_descriptors(_ptr(desc), _ptr(offsets), int(offsets.shape[0]) - 1, _ptr(bad), _ptr(best_idx),
                                            _ptr(best_median), stream)
    if rc:
        raise OrbError(rc, "orbm_distinctive_descriptors failed")


def search_by_bow_batch(A, VA, a_valid, B, VB, b_valid, kf_kf, nnratio, check_ori, match12, nmatches, match21=None, rounds=None, stream=None):
    """orbm_search_by_bow_batch: ORBmatcher::SearchByBoW (src/ORBmatcher.cc:552-832) for every (A, B) problem.  A, B:
    frames_batch(...); VA, VB: (node_id [P,S], node_off [P,S+1], n_nodes [P], feat [P,cap]) CUDA int32 tensors."""
    fa = OrbmFeatVec(_ptr(VA[0]), _ptr(VA[1]), _ptr(VA[2]), _ptr(VA[3]), int(VA[0].shape[1]))
    fb = OrbmFeatVec(_ptr(VB[0]), _ptr(VB[1]), _ptr(VB[2]), _ptr(VB[3]), int(VB[0].shape[1]))
    rc = lib().orbm_search_by_bow_batch(C.byref(A), C.byref(fa), _ptr(a_valid), C.byref(B), C.byref(fb), _ptr(b_valid), int(kf_kf),
                                        float(np.float32(nnratio)), int(check_ori), _ptr(match12), _ptr(match21), _ptr(nmatches),
                                        _ptr(rounds), stream)
    if rc:
        raise OrbError(rc, "orbm_search_by_bow_batch failed")


def search_for_triangulation_batch(A, VA, a_valid, B, VB, b_valid, F12, epipole, scale, sigma2, check_ori, match12, nmatches, stream=None):
    """orbm_search_for_triangulation_batch (src/ORBmatcher.cc:1183-1361 for every key-frame pair).  A, B: frames_batch(...) with
    u_right = mvuRight (or None); VA, VB as in search_by_bow_batch; F12 [P,9], epipole [P,2] CUDA float32 tensors; scale, sigma2 host."""
    fa = OrbmFeatVec(_ptr(VA[0]), _ptr(VA[1]), _ptr(VA[2]), _ptr(VA[3]), int(VA[0].shape[1]))
    fb = OrbmFeatVec(_ptr(VB[0]), _ptr(VB[1]), _ptr(VB[2]), _ptr(VB[3]), int(VB[0].shape[1]))
    scale, sigma2 = np.ascontiguousarray(scale, np.float32), np.ascontiguousarray(sigma2, np.float32)
    rc = lib().orbm_search_for_triangulation_batch(C.byref(A), C.byref(fa), _ptr(a_valid), C.byref(B), C.byref(fb), _ptr(b_valid), _ptr(F12),
                                                   _ptr(epipole), scale.ctypes.data, sigma2.ctypes.data, len(scale), int(check_ori),
                                                   _ptr(match12), _ptr(nmatches), stream)
    if rc:
        raise OrbError(rc, "orbm_search_for_triangulation_batch failed")


def stereo_matches(ex_left, ex_right, kps_l, desc_l, kps_r, desc_r, bf, fx, frame_l=0, frame_r=0):
    """Frame::ComputeStereoMatches (src/Frame.cc:513-699) on the pyramids still held by the two
    extractors.  Returns (mvuRight, mvDepth, number of matches before the median cut)."""
    kps_l, kps_r = _np(kps_l), _np(kps_r)
    desc_l, desc_r = _np(desc_l, np.uint8), _np(desc_r, np.uint8)
    ur, dep = np.zeros(len(kps_l), np.float32), np.zeros(len(kps_l), np.float32)
    n = C.c_int()
    rc = lib().orbm_stereo_matches(ex_left._h, frame_l, ex_right._h, frame_r, len(kps_l), kps_l.ctypes.data, desc_l.ctypes.data,
                                   len(kps_r), kps_r.ctypes.data, desc_r.ctypes.data, bf, fx, ur.ctypes.data, dep.ctypes.data, C.byref(n))
    if rc:
        raise OrbError(rc, "orbm_stereo_matches failed")
    return ur, dep, n.value


def hamming_bf(queries, train, device=0, nprob=1):
    """All-pairs Hamming search: per query (best index, best distance, second distance)."""
    L = lib()
    if isinstance(queries, np.ndarray):
        queries = np.ascontiguousarray(queries, np.uint8)
        train = np.ascontiguousarray(train, np.uint8)
        nq, nt = queries.shape[-2], train.shape[-2]
        outs = [np.zeros(nprob * nq, np.int32) for _ in range(3)]
    else:
        import torch
        nq, nt = queries.shape[-2], train.shape[-2]
        outs = [torch.zeros(nprob * nq, dtype=torch.int32, device=queries.device) for _ in range(3)]
    rc = L.orbm_hamming_bf(_ptr(queries), nq, _ptr(train), nt, nprob, _ptr(outs[0]), _ptr(outs[1]), _ptr(outs[2]), device)
    if rc:
        raise OrbError(rc, "orbm_hamming_bf failed")
    return tuple(outs)


def hamming_bf_async(queries, train, outs, nprob=1, stream=None):
    """orbm_hamming_bf_async: CUDA tensors in, results into the three int32 CUDA tensors `outs`; only enqueues."""
    nq, nt = queries.shape[-2], train.shape[-2]
    rc = lib().orbm_hamming_bf_async(_ptr(queries), nq, _ptr(train), nt, nprob, _ptr(outs[0]), _ptr(outs[1]), _ptr(outs[2]), stream)
    if rc:
        raise OrbError(rc, "orbm_hamming_bf_async failed")
