"""One launch of every matcher / map-point kernel on freshly extracted frames (for `ncu`, see tools/profile_round.sh);
with ORB_BENCH_PROFILE=1 the warm-up is a single launch."""
import sys, os, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, orb_slam2_chinesenotes_b200 as ob
dev = torch.device("cuda", 0)
w, h, nf = 1241, 376, 2000
ex = ob.ORBextractor(nf, 1.2, 8, 20, 7, device=0)
cap = ex.default_capacity()
batch = 448 * 2
frames = bench.workload_frames("kitti_1241x376_nf2000", 0, batch, dev)
d_kps = torch.zeros((batch, cap, 7), dtype=torch.float32, device=dev)
d_desc = torch.zeros((batch, cap, 32), dtype=torch.uint8, device=dev)
d_n = torch.zeros(batch, dtype=torch.int32, device=dev)
ex.extract_batch_raw(frames, h * w, batch, w, h, w, d_kps, d_desc, cap, d_n, asynchronous=True)
torch.cuda.synchronize()
r = bench.bench_window_match(dev, d_kps, d_desc, d_n, cap, w, h, ex.GetScaleFactors(), nprob=444, cpu=False, reps=1 if os.environ.get("ORB_BENCH_PROFILE") else 40)
print({k: r[k] for k in ("value", "ms_per_launch", "rounds_max", "rounds_mean")})
m = bench.bench_mappoint_side(dev, d_kps, d_desc, d_n, cap, w, h, 1.2, reps=1, cpu=False)
print({k: (v["value"], v["ms_per_launch"]) for k, v in m.items()})
# SearchForInitialization: left frame p against a shifted, bit-flipped copy of itself, 148 pairs per launch
np_ = 148
k1 = d_kps[0:2 * np_:2].contiguous(); d1 = d_desc[0:2 * np_:2].contiguous(); n1 = d_n[0:2 * np_:2].contiguous()
g = torch.Generator(device=dev).manual_seed(3)
k2 = k1.clone(); k2[..., 0] += 6.0; k2[..., 1] -= 4.0
d2 = d1.clone(); d2[:, :, 8:] ^= (torch.rand(d2[:, :, 8:].shape, generator=g, device=dev) < 0.03).to(torch.uint8) * 4
bounds = (0.0, float(w), 0.0, float(h))
F1, F2 = ob.frames_batch(k1, d1, n1, bounds, None, 2016), ob.frames_batch(k2, d2, n1, bounds, None, 2016)
prev = k1[..., :2].contiguous().clone()
m12 = torch.zeros((np_, cap), dtype=torch.int32, device=dev); nm = torch.zeros(np_, dtype=torch.int32, device=dev)
rounds = torch.zeros(np_, dtype=torch.int32, device=dev)
for _ in range(2):
    p = prev.clone()
    ob.search_for_initialization_batch(F1, F2, p, m12, nm, 100, 0.9, True, rounds)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
p = prev.clone(); e0.record(); ob.search_for_initialization_batch(F1, F2, p, m12, nm, 100, 0.9, True, rounds); e1.record(); torch.cuda.synchronize()
print({"init_ms_per_launch": e0.elapsed_time(e1), "pairs": np_, "matches": float(nm.float().mean()), "rounds_max": int(rounds.max())})
h = bench.bench_hamming(dev, 0, reps=1)
print({k: h[k] for k in ("value", "ms_per_launch")})
os._exit(0)
