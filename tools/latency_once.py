"""Latency of the drop-in calls for ONE frame / ONE matching problem with host buffers (what a real-time SLAM front end
sees per frame), against the CPU checker on one core."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import orb_slam2_chinesenotes_b200 as ob
from matcher_lib import Matcher, projected_queries
from oracle_lib import OracleExtractor
from synth import synth_frame, stereo_pair

w, h, nf = 1241, 376, 2000
img = synth_frame(w, h, 2)
ex = ob.ORBextractor(nf, 1.2, 8, 20, 7)
for _ in range(5):
    kps, desc = ex(img)
t0 = time.perf_counter()
for _ in range(50):
    kps, desc = ex(img)
t_ex = (time.perf_counter() - t0) / 50
import ctypes as C
us = (C.c_double * 3)()
ex._L.orbx_debug_last_call_us.argtypes = [C.c_void_p, C.c_void_p]
ex._L.orbx_debug_last_call_us(ex._h, us)
print({"single_call_host_us": {"stage_copy": us[0], "graph_launch": us[1], "wait": us[2]}})
L, R = stereo_pair(w, h, 3)
pair = np.stack([L, R])
for _ in range(3):
    ex.extract_stereo_batch(pair, 386.1448, 718.856)
t0 = time.perf_counter()
for _ in range(30):
    ex.extract_stereo_batch(pair, 386.1448, 718.856)
t_st = (time.perf_counter() - t0) / 30
bounds = (0.0, float(w), 0.0, float(h))
q = projected_queries(kps, desc, 2000, 5)
F = ob.FrameView(kps, desc, bounds)
M = ob.ORBmatcher(0.9, True)
scale = ex.GetScaleFactors()
for _ in range(5):
    M.SearchByProjection(F, scale, q, 3.0)
t0 = time.perf_counter()
for _ in range(50):
    nm, a = M.SearchByProjection(F, scale, q, 3.0)
t_m = (time.perf_counter() - t0) / 50
R = ob.ResidentFrame(kps, desc, bounds)
for _ in range(5):
    M.SearchByProjection(R, scale, q, 3.0)
t0 = time.perf_counter()
for _ in range(50):
    nm_r, a_r = M.SearchByProjection(R, scale, q, 3.0)
t_mr = (time.perf_counter() - t0) / 50
assert nm_r == nm and (a_r == a).all()
O = OracleExtractor(nf)
t0 = time.perf_counter(); O.extract(img); t_cpu_ex = time.perf_counter() - t0
Mo = Matcher("oracle")
t0 = time.perf_counter()
for _ in range(5):
    Mo.search_by_projection_points(kps, desc, None, scale, bounds, q, 3.0, 0.9, None)
t_cpu_m = (time.perf_counter() - t0) / 5
# the C++ drop-in class itself (tests/cpp/dropin_harness.cc, timed inside C++): Frame::ExtractORB's call, with and without mvImagePyramid
import ctypes as C, subprocess
cpp = os.path.join(ROOT, "tests", "cpp")
subprocess.run(["make", "-C", cpp, "all"], capture_output=True)
D = C.CDLL(os.path.join(cpp, "_build", "libdropin.so"))
D.dropin_create.restype = C.c_void_p
D.dropin_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
D.dropin_time_ms.restype = C.c_double
D.dropin_time_ms.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int]
hd = D.dropin_create(nf, 1.2, 8, 20, 7)
t_drop_pyr = D.dropin_time_ms(hd, img.ctypes.data, w, h, img.strides[0], 50, 1)
t_drop = D.dropin_time_ms(hd, img.ctypes.data, w, h, img.strides[0], 50, 0)
print({"dropin_operator_ms": t_drop_pyr, "dropin_operator_no_pyramid_ms": t_drop})
print({"extract_1_frame_ms": t_ex * 1e3, "stereo_pair_ms": t_st * 1e3, "search_by_projection_ms": t_m * 1e3, "search_by_projection_resident_frame_ms": t_mr * 1e3,
       "cpu_extract_1_frame_ms": t_cpu_ex * 1e3, "cpu_search_by_projection_ms": t_cpu_m * 1e3, "nmatches": nm})
