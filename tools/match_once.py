"""One launch of every matcher / map-point kernel on freshly extracted frames (for `ncu`, see tools/profile_round.sh);
with ORB_BENCH_PROFILE=1 the warm-up is a single launch."""
import sys, os, numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import bench, orb_slam2_chinesenotes_b200 as ob
dev = torch.device("cuda", 0)
w, h, nf = 1241, 376, 2000
ex = ob.ORBextractor(nf, 1.2, 8, 20, 7, device=0)
cap = ex.default_capacity()
batch = 296 * 2
frames = bench.synth_batch_torch(batch, w, h, 2000, dev)
d_kps = torch.zeros((batch, cap, 7), dtype=torch.float32, device=dev)
d_desc = torch.zeros((batch, cap, 32), dtype=torch.uint8, device=dev)
d_n = torch.zeros(batch, dtype=torch.int32, device=dev)
ex.extract_batch_raw(frames, h * w, batch, w, h, w, d_kps, d_desc, cap, d_n, asynchronous=True)
torch.cuda.synchronize()
r = bench.bench_window_match(dev, d_kps, d_desc, d_n, cap, w, h, ex.GetScaleFactors(), nprob=296, cpu=False, reps=1)
print({k: r[k] for k in ("value", "ms_per_launch")})
m = bench.bench_mappoint_side(dev, d_kps, d_desc, d_n, cap, w, h, 1.2, reps=1, cpu=False)
print({k: (v["value"], v["ms_per_launch"]) for k, v in m.items()})
h = bench.bench_hamming(dev, 0, reps=1)
print({k: h[k] for k in ("value", "ms_per_launch")})
os._exit(0)
