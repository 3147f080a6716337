import sys, os, ctypes as C, numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import bench, orb_slam2_chinesenotes_b200 as ob
dev = torch.device("cuda", 0)
w, h, nf = 1241, 376, 2000
ex = ob.ORBextractor(nf, 1.2, 8, 20, 7, device=0)
cap = ex.default_capacity()
batch = 1024
frames = bench.synth_batch_torch(batch, w, h, 2000, dev)
d_kps = torch.zeros((batch, cap, 7), dtype=torch.float32, device=dev)
d_desc = torch.zeros((batch, cap, 32), dtype=torch.uint8, device=dev)
d_n = torch.zeros(batch, dtype=torch.int32, device=dev)
ex.extract_batch_raw(frames, h * w, batch, w, h, w, d_kps, d_desc, cap, d_n, asynchronous=True)
torch.cuda.synchronize()
for nprob in (148, 296, 444, 512):
    for rep in range(2):
        r = bench.bench_window_match(dev, d_kps, d_desc, d_n, cap, w, h, ex.GetScaleFactors(), nprob=nprob, cpu=False)
        print(nprob, {k: r[k] for k in ("problems_per_launch", "value", "ms_per_launch", "rounds_max", "rounds_mean")})
