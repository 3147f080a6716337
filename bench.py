#!/usr/bin/env python3
"""Benchmark of the ORB front-end hot path (BASELINE.json metric: ORB frames/s @1241x376, nfeatures=2000).

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA, liborb_b200.so)
  python bench.py --impl reference ...                      the reference's own CPU code (oracle/_ref)

One step = ORBextractor::operator() over one batch of synthetic KITTI-shape frames
(BASELINE.json configs[1]); batches are larger than the 126 MB L2 so every step starts cold.
Multi-GPU (torchrun): frames are independent, so each rank extracts its own batch (weak scaling);
NCCL is used only to gather timings and counts.  Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (w, h, nfeatures, frames per step per GPU)
    # kitti = BASELINE.json configs[1]: rectified stereo pairs (frames L0,R0,L1,R1,...), left/right extraction +
    # Frame::ComputeStereoMatches per pair; the others are extraction only
    "kitti_1241x376_nf2000": (1241, 376, 2000, 1024),
    "tum_640x480_nf1000": (640, 480, 1000, 1024),
    "euroc_752x480_nf1200": (752, 480, 1200, 256),
    "hd_1920x1080_nf4000": (1920, 1080, 4000, 128),
}
LEVELS, SCALE, INI_TH, MIN_TH = 8, 1.2, 20, 7
STEREO = {"kitti_1241x376_nf2000"}
# KITTI 00-02 calibration (synthetic-test constants): fx, fy, cx, cy, bf, ThDepth * bf / fx
CAM = dict(fx=718.856, fy=718.856, cx=607.1928, cy=185.2157, bf=386.1448, th_depth=35.0 * 386.1448 / 718.856)


_STDOUT_FD = None


def guard_stdout():
    """The contract is ONE JSON line on stdout.  Libraries print there too (NCCL's version banner does, whatever
    NCCL_DEBUG_FILE says), so file descriptor 1 is pointed at stderr for the whole run and the JSON line is written to
    the saved descriptor at the end."""
    global _STDOUT_FD
    if _STDOUT_FD is None:
        sys.stdout.flush()
        _STDOUT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    sys.stdout.flush()
    if _STDOUT_FD is None:
        os.write(1, line)
    else:
        os.write(_STDOUT_FD, line)


def level_pixels(w, h):
    """Sum of level pixels P, last level p7."""
    import orb_slam2_chinesenotes_b200 as ob
    d = ob.plan_describe(1000, SCALE, LEVELS, INI_TH, MIN_TH, w, h)
    px = (d["level_w"].astype(np.int64) * d["level_h"].astype(np.int64))
    return int(px.sum()), int(px[-1]), int(px[0])


def algorithmic_bytes(w, h):
    """Per-frame algorithmic bytes of each dense stage."""
    P, p7, WH = level_pixels(w, h)
    return {"pyramid": 2 * P - p7 - WH, "fast": P, "blur": 2 * P}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic(kernel, frames_per_launch):
    """DRAM bytes per launch of `kernel` from the committed `ncu --set full` capture (profiles/ncu_traffic.json,
    written by tools/ncu_traffic.py): dram__bytes_read.sum + dram__bytes_write.sum per frame x frames per launch."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            t = json.load(f)
        return float(t["dram_bytes_per_frame"][kernel]) * frames_per_launch
    except Exception:
        return None


def synth_batch_torch(batch, w, h, seed, device):
    """Synthetic frames on the GPU: smoothed noise, mean 128 / std 48, low-contrast bottom band,
    hard-edged rectangles (the recipe of tests/synth.py, generated with torch for speed)."""
    import torch
    import torch.nn.functional as F
    g = torch.Generator(device=device).manual_seed(seed)
    out = torch.empty((batch, h, w), dtype=torch.uint8, device=device)
    sigma, r = 2.5, 8
    x = torch.arange(-r, r + 1, device=device, dtype=torch.float32)
    k = torch.exp(-0.5 * (x / sigma) ** 2)
    k = k / k.sum()
    rng = np.random.default_rng(seed)
    step = 32
    for b0 in range(0, batch, step):
        nb = min(step, batch - b0)
        n = torch.randint(0, 256, (nb, 1, h, w), generator=g, device=device, dtype=torch.uint8).float()
        n = F.conv2d(F.pad(n, (r, r, 0, 0), mode="reflect"), k.view(1, 1, 1, -1))
        n = F.conv2d(F.pad(n, (0, 0, r, r), mode="reflect"), k.view(1, 1, -1, 1))
        n = n[:, 0]
        n = (n - n.mean(dim=(1, 2), keepdim=True)) / n.std(dim=(1, 2), keepdim=True) * 48 + 128
        y0 = int(h * 0.75)
        n[:, y0:] = (n[:, y0:] - 128) * 0.12 + 128
        img = n.round().clamp(0, 255).to(torch.uint8)
        for i in range(nb):
            for _ in range(20):
                rx = int(rng.integers(0, w - 40)); ry = int(rng.integers(0, h - 40))
                img[i, ry:ry + int(rng.integers(8, 40)), rx:rx + int(rng.integers(8, 40))] = int(rng.integers(0, 256))
        out[b0:b0 + nb] = img
    return out


def stereo_right_torch(left, seed):
    """Right images for a batch of left images (tests/synth.py stereo_pair): every block of 16..63 rows is the left
    block shifted by its own disparity of 5..60 px, plus +-2 grey levels of noise."""
    import torch
    b, h, w = left.shape
    rng = np.random.default_rng(seed + 7919)
    right = torch.empty_like(left)
    for i in range(b):
        y = 0
        while y < h:
            bh, d = int(rng.integers(16, 64)), int(rng.integers(5, 61))
            right[i, y:y + bh, :w - d] = left[i, y:y + bh, d:]
            right[i, y:y + bh, w - d:] = left[i, y:y + bh, w - d - 1:w - d]
            y += bh
    g = torch.Generator(device=left.device).manual_seed(seed + 1)
    noise = torch.randint(-2, 3, right.shape, generator=g, device=left.device, dtype=torch.int16)
    return (right.to(torch.int16) + noise).clamp_(0, 255).to(torch.uint8)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ reference arm
def load_cpu_checker():
    """(library, kind): oracle/_ref (the reference's own sources) when built, else the oracle port.
    bench.py may execute oracle/ only here: as the timed CPU baseline, never as the product."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    r = oracle_lib.ref()
    if r is not None:
        return r.orbref_extract_bench, "reference"
    return oracle_lib.oracle().orbo_extract_bench, "port"


def cpu_frames_per_s(frames_np, nfeatures, threads, reps=1):
    fn, kind = load_cpu_checker()
    nfr, h, w = frames_np.shape
    best = None
    tot = C.c_longlong()
    for _ in range(reps):
        dt = fn(nfeatures, SCALE, LEVELS, INI_TH, MIN_TH, frames_np.ctypes.data, nfr, w, h, threads, C.byref(tot))
        best = dt if best is None else min(best, dt)
    return nfr / best, kind, int(tot.value)


def cpu_stereo_frames_per_s(frames_np, nfeatures, threads):
    """Frames/s of the reference's own stereo Frame constructor (two ExtractORB threads + ComputeStereoMatches,
    src/Frame.cc:61-115) over pairs L0,R0,L1,R1,...; None when oracle/_ref was not built."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    r = oracle_lib.ref()
    if r is None or not hasattr(r, "orbref_stereo_bench"):
        return None
    nfr, h, w = frames_np.shape
    tot = C.c_longlong()
    dt = r.orbref_stereo_bench(nfeatures, SCALE, LEVELS, INI_TH, MIN_TH, frames_np.ctypes.data, nfr // 2, w, h, max(threads // 2, 1),
                               CAM["fx"], CAM["fy"], CAM["cx"], CAM["cy"], CAM["bf"], CAM["th_depth"], C.byref(tot), None, 0)
    return nfr / dt, int(tot.value)


def cpu_workload_frames_per_s(workload, frames_np, nfeatures, cores):
    """(frames/s, kind, what ran) of the CPU implementation of one bench workload."""
    if workload in STEREO:
        r = cpu_stereo_frames_per_s(frames_np, nfeatures, cores)
        if r is not None:
            return r[0], "reference", f"stereo Frame constructor, {max(cores // 2, 1)} frames in flight x 2 extractor threads"
    v, kind, _ = cpu_frames_per_s(frames_np, nfeatures, cores)
    return v, kind, f"{cores} threads, one extractor per thread" + (" (extraction only)" if workload in STEREO else "")


def run_reference(args, rank, world):
    if rank != 0:
        return
    from synth import stereo_pair, synth_frame  # numpy generators of the parity tests (no torch / CUDA needed)
    w, h, nf, _ = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    nsample = int(min(256, max(16, 4 * cores)))
    if args.workload in STEREO:
        frames = np.stack([im for i in range(min(nsample, 16) // 2) for im in stereo_pair(w, h, 2000 + i)])
    else:
        frames = np.stack([synth_frame(w, h, 2000 + i) for i in range(min(nsample, 16))])
    frames = np.concatenate([frames] * ((nsample + len(frames) - 1) // len(frames)))[:nsample]
    for _ in range(args.warmup):
        cpu_workload_frames_per_s(args.workload, frames[:max(2, cores & ~1)], nf, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v, kind, what = cpu_workload_frames_per_s(args.workload, frames, nf, cores)
    dt = time.perf_counter() - t0
    value = nsample * args.steps / dt
    sample = f"{nsample} frames per step ({min(nsample, 16)} distinct), {what}"
    emit(({
        "impl": "reference", "metric": "orb_extract_frames_per_s", "value": value, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": args.workload, "w": w, "h": h, "nfeatures": nf, "levels": LEVELS},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def bench_hamming(dev, device_index, nq=2000, nt=2000, nprob=256, reps=5):
    """Second half of the BASELINE metric: brute-force Hamming matching in Gpairs/s (BASELINE.json configs[3]
    shape, 2000 query x 2000 frame descriptors, 256 independent problems per launch, device resident).
    Bound: integer pipe, 8 POPC per pair at 16 POPC/clk/SM (tools/ubench_pipes.cu measures that rate)."""
    import torch
    import orb_slam2_chinesenotes_b200 as ob
    g = torch.Generator(device=dev).manual_seed(7)
    q = torch.randint(0, 256, (nprob, nq, 32), generator=g, device=dev, dtype=torch.uint8)
    t = torch.randint(0, 256, (nprob, nt, 32), generator=g, device=dev, dtype=torch.uint8)
    outs = [torch.zeros(nprob * nq, dtype=torch.int32, device=dev) for _ in range(3)]
    stream = torch.cuda.current_stream()

    def run():
        ob.hamming_bf_async(q, t, outs, nprob, stream.cuda_stream)

    warm_up(run)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        run()
    e1.record(stream)
    torch.cuda.synchronize()
    dt = e0.elapsed_time(e1) * 1e-3 / reps
    props = torch.cuda.get_device_properties(dev)
    sms = props.multi_processor_count
    # The kernel trades POPC work for LOP3 work with three carry-save adders (csrc/orb_match.cu k_hamming_bf): 5 POPC +
    # 8 XOR + 6 LOP3 per pair instead of 8 POPC + 8 XOR, which balances the quarter-rate POPC pipe against the ALU pipe.
    peak = sms * 16 * 1.965e9 / 5 / 1e9
    gp = nprob * nq * nt / dt / 1e9
    return {"value": gp, "unit": "Gpairs/s", "nq": nq, "nt": nt, "problems_per_launch": nprob, "ms_per_launch": dt * 1e3,
            "bound": "integer pipes (POPC and ALU balanced)", "peak": peak, "frac": gp / peak,
            "peak_note": "SMs x 16 POPC/clk x 1965 MHz / 5 POPC per pair; the ALU pipe (21 instructions per pair at 1.95 per "
                         "clock and SM, profiles/r01_ubench_pipes.txt) allows 864",
            "plain_popc_peak": sms * 16 * 1.965e9 / 8 / 1e9,
            "plain_popc_peak_note": "what 8 POPC per pair allow (SMs x 16 POPC/clk x 1965 MHz / 8); 537 Gpairs/s were measured in that form"}


def warm_up(fn, seconds=0.3):
    """Short kernels after an idle GPU: keep launching until the clocks have ramped up (0.3 s), not a fixed count."""
    import torch
    if os.environ.get("ORB_BENCH_PROFILE"):          # under ncu: one launch
        fn()
        return
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        for _ in range(10):
            fn()
        torch.cuda.synchronize()


def bench_window_match(dev, d_kps, d_desc, d_n, cap, w, h, scale, nprob=512, nq=2000, reps=40, cpu=True):
    """BASELINE.json configs[3]: windowed SearchByProjection(Frame, MapPoints) matching (src/ORBmatcher.cc:73-157),
    2000 map points against the ~2000 keypoints of a frame, TH_HIGH = 100, nnratio = 0.9, th = 3, for `nprob` frames per
    launch.  The frames are the left images the extractor just produced and stay where it left them in HBM; the
    queries are built on the device as SURVEY.md section 8d describes (descriptor of a random keypoint with
    k ~ U[0,80] bit flips, projection = keypoint + N(0, 3 px), level = its octave)."""
    import torch
    import orb_slam2_chinesenotes_b200 as ob
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    nprob = min(nprob, d_kps.shape[0] // 2)
    nprob = (nprob // sms) * sms or nprob                           # whole waves: one block per SM
    kps, desc, n = d_kps[0:2 * nprob:2].contiguous(), d_desc[0:2 * nprob:2].contiguous(), d_n[0:2 * nprob:2].contiguous()
    g = torch.Generator(device=dev).manual_seed(11)
    tgt = (torch.rand((nprob, nq), generator=g, device=dev) * n[:, None]).long().clamp_(max=cap - 1)
    tgt[:, ::5] = tgt[:, ::5] % 40                                  # every 5th query competes for the first 40 keypoints
    gi = tgt[..., None].expand(-1, -1, 7)
    kq = torch.gather(kps, 1, gi)
    proj = torch.stack([kq[..., 0] + 3 * torch.randn((nprob, nq), generator=g, device=dev),
                        kq[..., 1] + 3 * torch.randn((nprob, nq), generator=g, device=dev),
                        kq[..., 0] - 30 * torch.rand((nprob, nq), generator=g, device=dev)], dim=2).contiguous()
    level = kq[..., 5].contiguous().view(torch.int32).clone()
    qd = torch.gather(desc, 1, tgt[..., None].expand(-1, -1, 32)).clone()
    pflip = torch.rand((nprob, nq, 1, 1), generator=g, device=dev) * (80.0 / 256.0)
    weights = (2 ** torch.arange(8, device=dev, dtype=torch.int32)).view(1, 1, 1, 8)
    for p0 in range(0, nprob, 32):
        bits = (torch.rand((min(32, nprob - p0), nq, 32, 8), generator=g, device=dev) < pflip[p0:p0 + 32]).to(torch.int32)
        qd[p0:p0 + 32] ^= (bits * weights).sum(-1).to(torch.uint8)
    q = dict(proj=proj, level=level, desc=qd,
             view_cos=torch.where(torch.rand((nprob, nq), generator=g, device=dev) < 0.5, 0.9995, 0.9).float().contiguous(),
             in_view=torch.ones((nprob, nq), dtype=torch.uint8, device=dev), bad=torch.zeros((nprob, nq), dtype=torch.uint8, device=dev),
             obs=torch.ones((nprob, nq), dtype=torch.int32, device=dev))
    d_nq = torch.full((nprob,), nq, dtype=torch.int32, device=dev)
    d_assign = torch.zeros((nprob, cap), dtype=torch.int32, device=dev)
    d_nm = torch.zeros(nprob, dtype=torch.int32, device=dev)
    d_rounds = torch.zeros(nprob, dtype=torch.int32, device=dev)
    bounds = (0.0, float(w), 0.0, float(h))
    max_n = ob.max_keypoints(WORKLOADS["kitti_1241x376_nf2000"][2], SCALE, LEVELS, INI_TH, MIN_TH, w, h)   # provable bound, no read-back
    F = ob.frames_batch(kps, desc, n, bounds, None, max_n)
    stream = torch.cuda.current_stream()
    th, nnratio = 3.0, 0.9

    def run():
        ob.search_by_projection_points_batch(F, scale, q, d_nq, nq, d_assign, d_nm, th, nnratio, None, d_rounds, stream.cuda_stream)

    warm_up(run)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        run()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    nm = d_nm.cpu().numpy()
    assert (nm > 0).all(), "window matcher found nothing"
    out = {"value": nprob / (ms * 1e-3), "unit": "problems/s", "what": "SearchByProjection(Frame, MapPoints), th=3, nnratio=0.9",
           "nq": nq, "nt": float(n.float().mean().item()), "problems_per_launch": nprob, "ms_per_launch": ms,
           "queries_per_s": nprob * nq / (ms * 1e-3), "matches_per_problem": float(nm.mean()),
           "rounds_max": int(d_rounds.max().item()), "rounds_mean": float(d_rounds.float().mean().item()), "launches": 1}
    # latency of ONE problem through the host-array entry point (what a real-time tracker calls once per frame)
    n0 = int(n[0].item())
    kp0 = np.ascontiguousarray(kps[0, :n0].cpu().numpy()).view(ob.KP_DTYPE).reshape(n0)
    F1 = ob.FrameView(kp0, np.ascontiguousarray(desc[0, :n0].cpu().numpy()), bounds)
    q1 = {k: np.ascontiguousarray(v[0].cpu().numpy()) for k, v in q.items()}
    M1 = ob.ORBmatcher(nnratio, True)
    nrep = 1 if os.environ.get("ORB_BENCH_PROFILE") else 30            # under ncu: one launch
    for _ in range(min(5, nrep)):
        M1.SearchByProjection(F1, scale, q1, th)
    t0 = time.perf_counter()
    for _ in range(nrep):
        M1.SearchByProjection(F1, scale, q1, th)
    out["single_call_ms"] = 1e3 * (time.perf_counter() - t0) / nrep
    if cpu:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib
        from matcher_lib import Matcher
        M = Matcher("ref" if oracle_lib.ref() is not None else "oracle")
        n0 = int(n[0].item())
        kp0 = np.ascontiguousarray(kps[0, :n0].cpu().numpy()).view(oracle_lib.KP_DTYPE).reshape(n0)
        de0 = np.ascontiguousarray(desc[0, :n0].cpu().numpy())
        q0 = {k: np.ascontiguousarray(v[0].cpu().numpy()) for k, v in q.items()}
        t0 = time.perf_counter()
        for _ in range(20):
            want = M.search_by_projection_points(kp0, de0, None, np.ascontiguousarray(scale, np.float32), bounds, q0, th, nnratio)
        dt = (time.perf_counter() - t0) / 20
        assert want[0] == int(nm[0]) and (want[1] == d_assign[0, :n0].cpu().numpy()).all(), "window matcher disagrees with the CPU checker"
        out["cpu_baseline"] = {"value": 1.0 / dt, "unit": "problems/s", "cores": 1, "kind": "reference" if M.impl == "ref" else "port",
                               "sample": "problem 0 of the launch, 20 repetitions, results compared"}
    return out


def bench_mappoint_side(dev, d_kps, d_desc, d_n, cap, w, h, scale_factor, reps=40, cpu=True):
    """SURVEY.md section 8f rows on the frames the extractor just produced (device resident, one launch each):
      projection   Frame::isInFrustum + MapPoint::PredictScale, 4000 map points x nprob frames (src/Frame.cc:288-345)
      bow          ORBmatcher::SearchByBoW(KeyFrame, KeyFrame): left frame p against left frame p+1 with a synthetic
                   256-node feature vector (src/ORBmatcher.cc:700-832)
      distinctive  MapPoint::ComputeDistinctiveDescriptors for 65536 map points with 2..33 observations each
    CPU figures: the oracle's C restatement of the same functions on one core (kind "port"), problem 0, results compared."""
    import torch
    import orb_slam2_chinesenotes_b200 as ob
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    stream = torch.cuda.current_stream()
    out = {}

    def timed(fn):
        warm_up(fn)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    g = torch.Generator(device=dev).manual_seed(5)
    # ---- projection
    nprob, npts = 2 * sms, 4000
    K = np.float32([CAM["fx"], CAM["fy"], CAM["cx"], CAM["cy"]])
    Tcw = torch.eye(4, device=dev).repeat(nprob, 1, 1)
    Tcw[:, :3, 3] = 0.3 * torch.randn((nprob, 3), generator=g, device=dev)
    Tcw = Tcw.reshape(nprob, 16).contiguous()
    z = 2 + 58 * torch.rand(npts, generator=g, device=dev)
    xyz = torch.stack([(torch.rand(npts, generator=g, device=dev) * 1.4 - 0.2) * w, (torch.rand(npts, generator=g, device=dev) * 1.4 - 0.2) * h, z], 1)
    xyz[:, 0] = (xyz[:, 0] - CAM["cx"]) / CAM["fx"] * z
    xyz[:, 1] = (xyz[:, 1] - CAM["cy"]) / CAM["fy"] * z
    xyz = xyz.contiguous()
    dist = xyz.norm(dim=1)
    normal = (xyz / dist[:, None] + 0.3 * torch.randn((npts, 3), generator=g, device=dev))
    normal = (normal / normal.norm(dim=1, keepdim=True)).contiguous()
    max_d = (dist * (0.7 + 2.0 * torch.rand(npts, generator=g, device=dev))).contiguous()
    min_d = (max_d / 1.2 ** 7).contiguous()
    pts = dict(xyz=xyz, normal=normal, max_d=max_d, min_d=min_d)
    po = dict(in_view=torch.zeros((nprob, npts), dtype=torch.uint8, device=dev), proj=torch.zeros((nprob, npts, 3), device=dev),
              level=torch.zeros((nprob, npts), dtype=torch.int32, device=dev), view_cos=torch.zeros((nprob, npts), device=dev))
    nq = torch.full((nprob,), npts, dtype=torch.int32, device=dev)
    cnt = torch.zeros(nprob, dtype=torch.int32, device=dev)
    bounds = (0.0, float(w), 0.0, float(h))
    ms = timed(lambda: ob.project_points_batch(Tcw, K, CAM["bf"], bounds, scale_factor, LEVELS, nq, npts, pts, po, 0.5, True, cnt, stream.cuda_stream))
    out["projection"] = {"value": nprob * npts / (ms * 1e-3), "unit": "points/s", "frames_per_launch": nprob, "points_per_frame": npts,
                         "ms_per_launch": ms, "in_view_fraction": float(cnt.float().mean().item()) / npts}
    # ---- bow
    nb = min((d_kps.shape[0] // 2 - 1) // sms * sms or 1, 3 * sms)
    kA, dA, nA = d_kps[0:2 * nb:2].contiguous(), d_desc[0:2 * nb:2].contiguous(), d_n[0:2 * nb:2].contiguous()
    kB, dB, nB = d_kps[2:2 * nb + 2:2].contiguous(), d_desc[2:2 * nb + 2:2].contiguous(), d_n[2:2 * nb + 2:2].contiguous()

    def featvec(desc, n):
        P = desc.shape[0]
        node = (desc[:, :, 0].to(torch.int32) >> 4) * 16 + (desc[:, :, 1].to(torch.int32) >> 4)
        node = torch.where(torch.arange(cap, device=dev)[None, :] < n[:, None], node, 1 << 20)   # padding sorts last
        order = torch.argsort(node, dim=1, stable=True).to(torch.int32).contiguous()
        counts = torch.zeros((P, 256), dtype=torch.int32, device=dev)
        counts.scatter_add_(1, node.clamp(max=255).long(), (node < 256).to(torch.int32))
        off = torch.zeros((P, 257), dtype=torch.int32, device=dev)
        off[:, 1:] = counts.cumsum(1)
        ids = torch.arange(256, device=dev, dtype=torch.int32).repeat(P, 1).contiguous()           # empty nodes keep empty runs
        return ids, off.contiguous(), torch.full((P,), 256, dtype=torch.int32, device=dev), order

    VA, VB = featvec(dA, nA), featvec(dB, nB)
    # the two left frames are different scenes: make side B a noisy copy of side A so that there is something to match
    dB = dA.clone()
    dB[:, :, 4:] ^= (torch.rand(dB[:, :, 4:].shape, generator=g, device=dev) < 0.02).to(torch.uint8) * 16
    kB, nB, VB = kA, nA, featvec(dB, nA)
    valid = torch.ones((nb, cap), dtype=torch.uint8, device=dev)
    max_n = ob.max_keypoints(WORKLOADS["kitti_1241x376_nf2000"][2], SCALE, LEVELS, INI_TH, MIN_TH, w, h)
    FA, FB = ob.frames_batch(kA, dA, nA, bounds, None, max_n), ob.frames_batch(kB, dB, nB, bounds, None, max_n)
    m12 = torch.zeros((nb, cap), dtype=torch.int32, device=dev)
    nm = torch.zeros(nb, dtype=torch.int32, device=dev)
    rounds = torch.zeros(nb, dtype=torch.int32, device=dev)
    ms = timed(lambda: ob.search_by_bow_batch(FA, VA, valid, FB, VB, valid, True, 0.75, True, m12, nm, None, rounds, stream.cuda_stream))
    out["bow"] = {"value": nb / (ms * 1e-3), "unit": "problems/s", "what": "SearchByBoW(KeyFrame, KeyFrame), nnratio 0.75, 256 nodes",
                  "problems_per_launch": nb, "ms_per_launch": ms, "matches_per_problem": float(nm.float().mean().item()),
                  "rounds_max": int(rounds.max().item())}
    # ---- distinctive descriptors
    npnt = 65536
    sizes = torch.randint(2, 34, (npnt,), generator=g, device=dev, dtype=torch.int32)
    off = torch.zeros(npnt + 1, dtype=torch.int32, device=dev)
    off[1:] = sizes.cumsum(0)
    total = int(off[-1].item())
    owner = torch.repeat_interleave(torch.arange(npnt, device=dev), sizes.long())
    base = torch.randint(0, 256, (npnt, 32), generator=g, device=dev, dtype=torch.uint8)
    obs = base[owner] ^ ((torch.rand((total, 32), generator=g, device=dev) < 0.25).to(torch.uint8) * torch.randint(1, 256, (total, 32), generator=g, device=dev, dtype=torch.uint8))
    obs = obs.contiguous()
    bi = torch.zeros(npnt, dtype=torch.int32, device=dev)
    bm = torch.zeros(npnt, dtype=torch.int32, device=dev)
    ms = timed(lambda: ob.distinctive_descriptors(obs, off, bi, bm, None, stream.cuda_stream))
    pairs = float((sizes.double() ** 2).sum().item())
    out["distinctive"] = {"value": npnt / (ms * 1e-3), "unit": "map points/s", "points_per_launch": npnt, "observations": total,
                          "ms_per_launch": ms, "Gpairs_per_s": pairs / (ms * 1e-3) / 1e9}
    if cpu:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from bow_lib import search_by_bow
        from mappoint_lib import distinctive, is_in_frustum
        import oracle_lib
        # projection, frame 0
        sc = dict(Tcw=Tcw[0].cpu().numpy().reshape(4, 4).copy(), K=K, bf=np.float32(CAM["bf"]), bounds=bounds, xyz=xyz.cpu().numpy(),
                  normal=normal.cpu().numpy(), max_d=max_d.cpu().numpy(), min_d=min_d.cpu().numpy())
        t0 = time.perf_counter()
        for _ in range(20):
            iv, pj, lv, vc = is_in_frustum("oracle", sc, scale_factor=scale_factor, nlevels=LEVELS)
        dt = (time.perf_counter() - t0) / 20
        m = iv.astype(bool)
        assert (po["in_view"][0].cpu().numpy() == iv).all() and (po["proj"][0].cpu().numpy()[m].view(np.uint32) == pj[m].view(np.uint32)).all() \
            and (po["level"][0].cpu().numpy()[m] == lv[m]).all(), "projection disagrees with the CPU checker"
        out["projection"]["cpu_baseline"] = {"value": npts / dt, "unit": "points/s", "cores": 1, "kind": "port", "sample": "frame 0, 20 repetitions, results compared"}
        # bow, problem 0
        n0 = int(nA[0].item())
        tonp = lambda t: np.ascontiguousarray(t.cpu().numpy())
        kp0 = tonp(kA[0, :n0]).view(oracle_lib.KP_DTYPE).reshape(n0)
        fv = lambda V: (tonp(V[0][0]), tonp(V[1][0]), tonp(V[3][0, :n0]))
        s = dict(k1=kp0, d1=tonp(dA[0, :n0]), k2=kp0, d2=tonp(dB[0, :n0]), valid1=np.ones(n0, np.uint8), bad1=np.zeros(n0, np.uint8),
                 valid2=np.ones(n0, np.uint8), bad2=np.zeros(n0, np.uint8), fv1=fv(VA), fv2=fv(VB))
        t0 = time.perf_counter()
        for _ in range(20):
            want = search_by_bow("oracle", s, 0.75, True, True)
        dt = (time.perf_counter() - t0) / 20
        assert want[0] == int(nm[0].item()) and (want[1] == m12[0, :n0].cpu().numpy()).all(), "SearchByBoW disagrees with the CPU checker"
        out["bow"]["cpu_baseline"] = {"value": 1.0 / dt, "unit": "problems/s", "cores": 1, "kind": "port", "sample": "problem 0, 20 repetitions, results compared"}
        # distinctive, first 2000 points
        offh, obsh = off.cpu().numpy(), obs.cpu().numpy()
        t0 = time.perf_counter()
        got = [distinctive("oracle", obsh[offh[p]:offh[p + 1]])[1:] for p in range(2000)]
        dt = time.perf_counter() - t0
        assert [g_[0] for g_ in got] == bi[:2000].cpu().tolist() and [g_[1] for g_ in got] == bm[:2000].cpu().tolist(), "distinctive descriptors disagree"
        out["distinctive"]["cpu_baseline"] = {"value": 2000 / dt, "unit": "map points/s", "cores": 1, "kind": "port",
                                              "sample": "first 2000 points through ctypes (call overhead included), results compared"}
    return out


# ------------------------------------------------------------------------------------------ our arm
def run_ours(args, rank, world, local_rank):
    import torch
    import orb_slam2_chinesenotes_b200 as ob
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # NCCL's banner must not share stdout with the JSON line
        dist.init_process_group("nccl", device_id=dev)
    w, h, nf, batch = WORKLOADS[args.workload]
    if args.batch:
        batch = args.batch
    ex = ob.ORBextractor(nf, SCALE, LEVELS, INI_TH, MIN_TH, device=local_rank)
    if args.chunk:
        ex.set_chunk(args.chunk)
    cap = ex.default_capacity()
    stereo = args.workload in STEREO
    if stereo:
        batch += batch % 2
        left = synth_batch_torch(batch // 2, w, h, 1000 * 2 + rank * 100003, dev)
        frames = torch.stack([left, stereo_right_torch(left, 1000 * 2 + rank * 100003)], dim=1).reshape(batch, h, w).contiguous()
        del left
    else:
        frames = synth_batch_torch(batch, w, h, 1000 * 2 + rank * 100003, dev)
    pairs = batch // 2
    d_ur = torch.full((pairs, cap), -1.0, dtype=torch.float32, device=dev)
    d_dep = torch.full((pairs, cap), -1.0, dtype=torch.float32, device=dev)
    d_ns = torch.zeros(pairs, dtype=torch.int32, device=dev)
    d_kps = torch.zeros((batch, cap, 7), dtype=torch.float32, device=dev)
    d_desc = torch.zeros((batch, cap, 32), dtype=torch.uint8, device=dev)
    d_n = torch.zeros(batch, dtype=torch.int32, device=dev)
    # a real (non-default) stream: handle 0 would mean "use the context's own stream", and
    # torch.cuda.Event only sees the stream it is recorded on
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ex.set_stream(stream.cuda_stream)

    def step_device():
        if stereo:
            ex.extract_stereo_batch_raw(frames, h * w, pairs, w, h, w, d_kps, d_desc, cap, d_n, CAM["bf"], CAM["fx"],
                                        d_ur, d_dep, d_ns, asynchronous=True)
        else:
            ex.extract_batch_raw(frames, h * w, batch, w, h, w, d_kps, d_desc, cap, d_n, asynchronous=True)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    # ---- device-resident throughput (`value`)
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    e1.record(stream)
    barrier()
    ms_dev = e0.elapsed_time(e1)
    # ---- per-stage kernel durations: a second timed pass of the same steps with CUDA events around every
    # launch; the library then keeps all kernels on the one stream, so an event pair brackets one stage
    ex.profile(True)
    ex.stage_ms(reset=True)
    for _ in range(args.steps):
        step_device()
    barrier()
    stage_ms, stage_cnt = ex.stage_ms(reset=True)
    ex.profile(False)
    nk = d_n.cpu().numpy()
    assert (nk > 0).all() and (nk <= cap).all(), "extraction produced no keypoints / overflowed"
    ns_dev = d_ns.cpu().numpy()
    if stereo:
        assert (ns_dev > 0).all(), "stereo matching found nothing"
        n_depth = int((d_dep > 0).sum().item())

    # ---- end to end through the C ABI with HOST buffers (pinned): H2D + kernels + D2H per step
    h_frames = torch.empty((batch, h, w), dtype=torch.uint8).pin_memory()
    h_frames.copy_(frames)
    h_kps = torch.empty((batch, cap, 7), dtype=torch.float32).pin_memory()
    h_desc = torch.empty((batch, cap, 32), dtype=torch.uint8).pin_memory()
    h_n = torch.empty(batch, dtype=torch.int32).pin_memory()
    h_ur = torch.empty((pairs, cap), dtype=torch.float32).pin_memory()
    h_dep = torch.empty((pairs, cap), dtype=torch.float32).pin_memory()
    h_ns = torch.empty(pairs, dtype=torch.int32).pin_memory()

    def step_host():
        if stereo:
            ex.extract_stereo_batch_raw(h_frames, h * w, pairs, w, h, w, h_kps, h_desc, cap, h_n, CAM["bf"], CAM["fx"],
                                        h_ur, h_dep, h_ns, asynchronous=False)
        else:
            ex.extract_batch_raw(h_frames, h * w, batch, w, h, w, h_kps, h_desc, cap, h_n, asynchronous=False)

    step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    barrier()
    ms_e2e = 1e3 * (time.perf_counter() - t0)
    clocks = sampler.stop()
    assert (h_n.numpy() == nk).all(), "host-buffer path and device-resident path disagree"
    if stereo:
        assert (h_ns.numpy() == ns_dev).all(), "stereo: host-buffer path and device-resident path disagree"

    # ---- max over ranks
    if dist is not None:
        t = torch.tensor([ms_dev, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_dev, ms_e2e = float(t[0]), float(t[1])
        kp = torch.tensor([float(nk.sum())], device=dev, dtype=torch.float64)
        dist.all_reduce(kp)
        total_kp = float(kp[0])
    else:
        total_kp = float(nk.sum())
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    total_frames = batch * world * args.steps
    value = total_frames / (ms_dev * 1e-3)
    e2e_value = total_frames / (ms_e2e * 1e-3)
    # ---- roofline of the dominant kernel (stage times measured above with CUDA events on the stream)
    bytes_per_frame = algorithmic_bytes(w, h)
    dom = max(("pyramid", "fast", "blur", "octree", "describe", "stereo"), key=lambda s: stage_ms[s])
    peak, peak_src = measured_peak()
    roof = {"bound": "hbm", "kernel": dom, "peak": peak, "unit": "GB/s", "peak_source": peak_src, "traffic": None,
            "stage_timing": "second pass of the same steps, kernels serialised on one stream, CUDA events around each launch",
            "stage_ms_per_step": {k: v / args.steps for k, v in stage_ms.items()}}
    if dom in bytes_per_frame:
        launches = max(stage_cnt[dom], 1)
        frames_per_launch = batch * args.steps / launches
        dur_s = stage_ms[dom] * 1e-3 / launches
        roof["achieved"] = bytes_per_frame[dom] * frames_per_launch / dur_s / 1e9
        roof["frac"] = roof["achieved"] / peak
        roof["algorithmic_bytes_per_frame"] = bytes_per_frame[dom]
        roof["algorithmic_bytes_per_launch"] = bytes_per_frame[dom] * frames_per_launch
        roof["traffic"] = measured_traffic(dom, frames_per_launch)
    else:
        roof["achieved"] = None
        roof["frac"] = None
    # what actually bounds the dominant kernel, from the committed ncu capture of this same command (not measured live)
    roof["limiter"] = {"fast": "instruction issue: 2.75 of 4 warp instructions per clock and SM, ALU pipe 62 %, 177 instructions per pixel pair; "
                               "DRAM traffic = algorithmic bytes (profiles/r01d_ncu_full_all_kernels.txt, r01d_ncu_full_extract_stalls.txt)"}.get(dom)
    roof["dense_stages"] = {s: {"GB/s": bytes_per_frame[s] * batch * args.steps / (stage_ms[s] * 1e-3) / 1e9,
                                "frac": bytes_per_frame[s] * batch * args.steps / (stage_ms[s] * 1e-3) / 1e9 / peak}
                            for s in bytes_per_frame if stage_ms[s] > 0}
    # ---- CPU baseline on a bounded sample of the same frames (rank 0, N=1 only)
    cpu = None
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        ns = int(min(batch, max(16, 4 * cores))) & ~1
        sample = frames[:ns].cpu().numpy()
        v, kind, what = cpu_workload_frames_per_s(args.workload, sample, nf, cores)
        cpu = {"value": v, "unit": "frames/s", "cores": cores, "kind": kind,
               "sample": f"first {ns} frames of the step's batch, {what}"}
    # ---- latency of ONE call with host buffers (what a real-time tracker sees per frame), rank 0 only
    latency = None
    if world == 1:
        one = h_frames[:2].numpy()
        for _ in range(5):
            ex(one[0])
        t0 = time.perf_counter()
        for _ in range(30):
            ex(one[0])
        latency = {"extract_1_frame_ms": 1e3 * (time.perf_counter() - t0) / 30, "note": "operator() on one frame, host buffers, ctypes overhead included"}
        if stereo:
            for _ in range(3):
                ex.extract_stereo_batch(one, CAM["bf"], CAM["fx"])
            t0 = time.perf_counter()
            for _ in range(30):
                ex.extract_stereo_batch(one, CAM["bf"], CAM["fx"])
            latency["stereo_pair_ms"] = 1e3 * (time.perf_counter() - t0) / 30
    chunk = args.chunk or 512                      # device-resident default of orb_capi.cu (the timed `value` path)
    chunks = (batch + chunk - 1) // chunk
    cfg_stereo = {}
    if stereo:
        cfg_stereo = {"stereo": "frames are rectified pairs L0,R0,L1,R1,...: left/right extraction + Frame::ComputeStereoMatches per pair",
                      "pairs_per_step_per_gpu": pairs, "depth_points_per_pair": n_depth / pairs}
    hamming = bench_hamming(dev, local_rank) if world == 1 else None
    window = mappoint = None
    if world == 1 and stereo:
        window = bench_window_match(dev, d_kps, d_desc, d_n, cap, w, h, ex.GetScaleFactors(), cpu=not args.no_cpu)
        mappoint = bench_mappoint_side(dev, d_kps, d_desc, d_n, cap, w, h, SCALE, cpu=not args.no_cpu)
    emit(({
        "metric": "orb_extract_frames_per_s", "value": value, "unit": "frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_dev / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": args.workload, "w": w, "h": h, "nfeatures": nf, "levels": LEVELS,
                   "frames_per_step_per_gpu": batch, "l2": "inputs larger than L2 (%.0f MB per step)" % (batch * w * h / 1e6),
                   "keypoints_per_frame": total_kp / (batch * world), **cfg_stereo},
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": batch * w * h,
                "d2h_bytes_per_step": batch * (cap * 60 + 4) + (pairs * (cap * 8 + 4) if stereo else 0),
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": (LEVELS - 1 + 4 + (3 if stereo else 0)) * chunks * args.steps,
        "roofline": roof, "cpu_baseline": cpu, "clocks": clocks, "hamming_bf": hamming,
        "window_match": window, "mappoint_side": mappoint, "latency": latency,
    }))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="kitti_1241x376_nf2000", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    guard_stdout()
    if args.impl == "reference":
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
