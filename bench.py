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


GEN_BLOCK = 32   # frames are generated in blocks of 32 whose random streams depend on the GLOBAL block index only


def synth_frames_torch(first, count, w, h, seed, device):
    """Synthetic frames [first, first+count) of the endless sequence `seed`: smoothed noise, mean 128 / std 48,
    low-contrast bottom band, hard-edged rectangles (the recipe of tests/synth.py, generated with torch for speed).
    A frame depends only on (seed, its global index), so any rank / any arm can produce any block of the sequence;
    `first` must be a multiple of GEN_BLOCK."""
    import torch
    import torch.nn.functional as F
    assert first % GEN_BLOCK == 0
    out = torch.empty((count, h, w), dtype=torch.uint8, device=device)
    sigma, r = 2.5, 8
    x = torch.arange(-r, r + 1, device=device, dtype=torch.float32)
    k = torch.exp(-0.5 * (x / sigma) ** 2)
    k = k / k.sum()
    for b0 in range(0, count, GEN_BLOCK):
        nb = min(GEN_BLOCK, count - b0)
        blk = (first + b0) // GEN_BLOCK
        g = torch.Generator(device=device).manual_seed(seed * 1000003 + blk)
        rng = np.random.default_rng([seed, blk])
        n = torch.randint(0, 256, (GEN_BLOCK, 1, h, w), generator=g, device=device, dtype=torch.uint8)[:nb].float()
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


def stereo_right_torch(left, first, seed):
    """Right images for left images [first, first+len) (tests/synth.py stereo_pair): every block of 16..63 rows is the
    left block shifted by its own disparity of 5..60 px, plus +-2 grey levels of noise; depends on (seed, index) only."""
    import torch
    b, h, w = left.shape
    right = torch.empty_like(left)
    for i in range(b):
        rng = np.random.default_rng([seed + 7919, first + i])
        y = 0
        while y < h:
            bh, d = int(rng.integers(16, 64)), int(rng.integers(5, 61))
            right[i, y:y + bh, :w - d] = left[i, y:y + bh, d:]
            right[i, y:y + bh, w - d:] = left[i, y:y + bh, w - d - 1:w - d]
            y += bh
        g = torch.Generator(device=left.device).manual_seed((seed + 1) * 1000003 + first + i)
        noise = torch.randint(-2, 3, (h, w), generator=g, device=left.device, dtype=torch.int16)
        right[i] = (right[i].to(torch.int16) + noise).clamp_(0, 255).to(torch.uint8)
    return right


def workload_frames(workload, first, count, device):
    """Frames [first, first+count) of a workload's input sequence (stereo workloads: L0,R0,L1,R1,...; first and count even)."""
    import torch
    w, h, _, _ = WORKLOADS[workload]
    seed = 2000 + sorted(WORKLOADS).index(workload)
    if workload in STEREO:
        assert first % 2 == 0 and count % 2 == 0 and (first // 2) % GEN_BLOCK == 0
        left = synth_frames_torch(first // 2, count // 2, w, h, seed, device)
        return torch.stack([left, stereo_right_torch(left, first // 2, seed)], dim=1).reshape(count, h, w).contiguous()
    return synth_frames_torch(first, count, w, h, seed, device)


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
    """(frames/s, kind, what ran, keypoints found) of the CPU implementation of one bench workload."""
    if workload in STEREO:
        r = cpu_stereo_frames_per_s(frames_np, nfeatures, cores)
        if r is not None:
            return r[0], "reference", f"stereo Frame constructor, {max(cores // 2, 1)} frames in flight x 2 extractor threads", r[1]
    v, kind, tot = cpu_frames_per_s(frames_np, nfeatures, cores)
    return v, kind, f"{cores} threads, one extractor per thread" + (" (extraction only)" if workload in STEREO else ""), tot


def opencv_dense_frames_per_s(frames_np, cores):
    """Informative only (never the reference arm): what REAL OpenCV (cv2, SIMD / IPP dispatch) needs for the dense stages of
    ORBextractor::operator() alone -- pyramid, FAST per level, GaussianBlur per level -- one process per core
    (tools/cv2_dense_baseline.py, run as a separate program: no fork of this CUDA process).  The reference arm runs the reference's
    sources over oracle/cvshim, a scalar stand-in for OpenCV, so its frames/s understate what the reference reaches with the real
    library; this line brackets it from the other side.  None when cv2 is missing or the helper fails."""
    import tempfile
    try:
        with tempfile.TemporaryDirectory() as d:
            path = os.path.join(d, "frames.npy")
            np.save(path, np.ascontiguousarray(frames_np))
            r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "cv2_dense_baseline.py"), path, str(cores), str(LEVELS), str(SCALE), str(INI_TH)],
                               capture_output=True, text=True, timeout=180)
        if r.returncode != 0:
            return None
        return json.loads(r.stdout.strip().splitlines()[-1])
    except Exception:
        return None


def workload_config(workload, batch, world, scaling):
    """The `config` object of a JSON line: what is measured, nothing that was measured (both arms print the same one)."""
    w, h, nf, _ = WORKLOADS[workload]
    cfg = {"workload": workload, "w": w, "h": h, "nfeatures": nf, "levels": LEVELS, "scale_factor": SCALE, "ini_th": INI_TH, "min_th": MIN_TH,
           "frames_per_step_per_gpu": batch if scaling == "weak" else batch // max(world, 1), "frames_per_step": batch * world if scaling == "weak" else batch,
           "l2": "inputs larger than L2 (%.0f MB per step and GPU)" % ((batch if scaling == "weak" else batch // max(world, 1)) * w * h / 1e6)}
    if workload in STEREO:
        cfg["stereo"] = "frames are rectified pairs L0,R0,L1,R1,...: left/right extraction + Frame::ComputeStereoMatches per pair"
    return cfg


def run_reference(args, rank, world):
    """The reference's own CPU code (oracle/_ref: the unmodified sources) on a bounded sample of the workload: the FIRST
    frames of the very sequence the CUDA arm extracts (same generator, same seeds), all host cores."""
    if rank != 0:
        return
    import torch
    w, h, nf, batch = WORKLOADS[args.workload]
    if args.batch:
        batch = args.batch
    cores = os.cpu_count() or 1
    nsample = 64 if args.workload in STEREO else int(min(256, max(16, 4 * cores))) // GEN_BLOCK * GEN_BLOCK or GEN_BLOCK
    nsample = min(nsample, batch)
    dev = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")   # the generator's random streams are per device type
    frames = np.ascontiguousarray(workload_frames(args.workload, 0, nsample, dev).cpu().numpy())
    for _ in range(args.warmup):
        cpu_workload_frames_per_s(args.workload, frames[:max(2, cores & ~1)], nf, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v, kind, what, total_kp = cpu_workload_frames_per_s(args.workload, frames, nf, cores)
    dt = time.perf_counter() - t0
    value = nsample * args.steps / dt
    sample = f"first {nsample} frames of the workload's sequence per step (generated on {dev.type}), {what}"
    base = {"value": value, "unit": "frames/s", "cores": cores, "kind": kind, "sample": sample}
    ocv = opencv_dense_frames_per_s(frames, cores)          # after the timed region: informative bracket, see the function
    if ocv is not None:
        base["opencv_dense_stages_only"] = {"value": ocv["frames_per_s"], "unit": "frames/s", "cores": cores,
                                            "ms_per_frame_per_core": ocv["ms_per_frame_per_core"],
                                            "what": "cv2 (real OpenCV, SIMD) pyramid + FAST per level + GaussianBlur only, one process per core, same "
                                                    "frames; informative: this arm runs the reference's sources over a scalar OpenCV stand-in"}
    emit(({
        "impl": "reference", "metric": "orb_extract_frames_per_s", "value": value, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(args.workload, batch, max(args.gpus, 1), "weak"),
        "stats": ({"depth_points_per_pair": total_kp / (nsample // 2)} if args.workload in STEREO and kind == "reference"
                  else {"keypoints_per_frame": total_kp / nsample}),
        "cpu_baseline": base,
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
    # the same with the Frame already on the device (orbm_frame_upload): the 2nd .. 4th search Tracking runs on a Frame
    R1 = ob.ResidentFrame(F1.kps, F1.desc, bounds)
    for _ in range(min(5, nrep)):
        res_r = M1.SearchByProjection(R1, scale, q1, th)
    t0 = time.perf_counter()
    for _ in range(nrep):
        res_r = M1.SearchByProjection(R1, scale, q1, th)
    out["single_call_resident_frame_ms"] = 1e3 * (time.perf_counter() - t0) / nrep
    res_h = M1.SearchByProjection(F1, scale, q1, th)
    assert res_r[0] == res_h[0] and (res_r[1] == res_h[1]).all(), "resident-frame search differs from the host-frame search"
    R1.close()
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
    # ---- triangulation: SearchForTriangulation on the same key-frame pairs (side B = side A's features seen again: same keypoints,
    # noisy descriptors, so the epipolar line of a feature passes through its twin), half of the features stereo ones
    F12 = torch.tensor([0, 0, 0, 0, 0, -1, 0, 1, 0], dtype=torch.float32, device=dev).repeat(nb, 1).contiguous()
    epi = torch.tensor([0.55 * w, 0.45 * h], dtype=torch.float32, device=dev).repeat(nb, 1).contiguous()
    urA = torch.where(torch.rand((nb, cap), generator=g, device=dev) < 0.5, kA[..., 0] - 20.0, torch.full((nb, cap), -1.0, device=dev)).contiguous()
    urB = torch.where(torch.rand((nb, cap), generator=g, device=dev) < 0.5, kA[..., 0] - 20.0, torch.full((nb, cap), -1.0, device=dev)).contiguous()
    tvA = (torch.rand((nb, cap), generator=g, device=dev) < 0.7).to(torch.uint8).contiguous()      # features without a map point yet
    tvB = (torch.rand((nb, cap), generator=g, device=dev) < 0.7).to(torch.uint8).contiguous()
    t_scale = np.float32([np.float32(scale_factor) ** i for i in range(LEVELS)])
    t_sigma2 = (t_scale * t_scale).astype(np.float32)
    TA, TB = ob.frames_batch(kA, dA, nA, bounds, urA, max_n), ob.frames_batch(kB, dB, nB, bounds, urB, max_n)
    t12 = torch.zeros((nb, cap), dtype=torch.int32, device=dev)
    tnm = torch.zeros(nb, dtype=torch.int32, device=dev)
    ms = timed(lambda: ob.search_for_triangulation_batch(TA, VA, tvA, TB, VB, tvB, F12, epi, t_scale, t_sigma2, True, t12, tnm, stream.cuda_stream))
    out["triangulation"] = {"value": nb / (ms * 1e-3), "unit": "problems/s", "what": "SearchForTriangulation(KeyFrame, KeyFrame), 256 nodes, "
                            "rotation check on", "problems_per_launch": nb, "ms_per_launch": ms, "matches_per_problem": float(tnm.float().mean().item())}
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
    # ---- fuse: the order-free search of ORBmatcher::Fuse (chi-square gate on) for nf key frames x 2000 projected points
    nfz = min((d_kps.shape[0] // 2) // sms * sms or 1, 3 * sms)
    kF, dF, nF = d_kps[0:2 * nfz:2].contiguous(), d_desc[0:2 * nfz:2].contiguous(), d_n[0:2 * nfz:2].contiguous()
    nqf = 2000
    tgt = (torch.rand((nfz, nqf), generator=g, device=dev) * nF[:, None]).long().clamp_(max=cap - 1)
    kq = torch.gather(kF, 1, tgt[..., None].expand(-1, -1, 7))
    lvl = kq[..., 5].contiguous().view(torch.int32).clone()
    scale_t = torch.tensor([scale_factor ** i for i in range(LEVELS)], device=dev, dtype=torch.float32)
    fq = dict(uvr=torch.stack([kq[..., 0] + 1.5 * torch.randn((nfz, nqf), generator=g, device=dev),
                               kq[..., 1] + 1.5 * torch.randn((nfz, nqf), generator=g, device=dev),
                               3.0 * scale_t[lvl.long().clamp_(0, LEVELS - 1)]], dim=2).contiguous(),
              level=lvl, desc=torch.gather(dF, 1, tgt[..., None].expand(-1, -1, 32)).clone())
    fq["desc"][:, :, :3] ^= torch.randint(0, 256, (nfz, nqf, 3), generator=g, device=dev, dtype=torch.uint8)
    fq["ur"] = (fq["uvr"][..., 0] - 30 * torch.rand((nfz, nqf), generator=g, device=dev)).contiguous()
    urF = torch.where(torch.rand((nfz, cap), generator=g, device=dev) < 0.5, kF[..., 0] - 30 * torch.rand((nfz, cap), generator=g, device=dev),
                      torch.full((nfz, cap), -1.0, device=dev)).contiguous()
    inv_sigma2 = (np.float32(1.0) / np.float32([np.float32(scale_factor) ** i for i in range(LEVELS)]) ** 2).astype(np.float32)
    f_nq = torch.full((nfz,), nqf, dtype=torch.int32, device=dev)
    f_bi = torch.zeros((nfz, nqf), dtype=torch.int32, device=dev)
    f_bd = torch.zeros((nfz, nqf), dtype=torch.int32, device=dev)
    f_nf = torch.zeros(nfz, dtype=torch.int32, device=dev)
    max_n = ob.max_keypoints(WORKLOADS["kitti_1241x376_nf2000"][2], SCALE, LEVELS, INI_TH, MIN_TH, w, h)
    FF = ob.frames_batch(kF, dF, nF, bounds, urF, max_n)
    ms = timed(lambda: ob.window_best_free_batch(FF, fq, f_nq, nqf, f_bi, f_bd, f_nf, 50, inv_sigma2, stream.cuda_stream))
    out["fuse"] = {"value": nfz / (ms * 1e-3), "unit": "problems/s", "what": "the search of ORBmatcher::Fuse(KeyFrame, MapPoints, th=3): 2000 projected "
                   "points per key frame, chi-square gate, TH_LOW", "problems_per_launch": nfz, "points_per_problem": nqf, "ms_per_launch": ms,
                   "points_per_s": nfz * nqf / (ms * 1e-3), "fused_per_problem": float(f_nf.float().mean().item())}
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
        # triangulation, problem 0
        O = oracle_lib.oracle()
        vpp, cii = C.c_void_p, C.c_int
        O.orbo_search_for_triangulation.argtypes = [cii, vpp, vpp, vpp, vpp, cii, vpp, vpp, vpp] * 2 + [vpp, vpp, vpp, vpp, cii, cii, vpp]
        tA = dict(kp=kp0, d=tonp(dA[0, :n0]), has=(1 - tonp(tvA[0, :n0])).astype(np.uint8), ur=tonp(urA[0, :n0]), fv=fv(VA))
        tB = dict(kp=kp0, d=tonp(dB[0, :n0]), has=(1 - tonp(tvB[0, :n0])).astype(np.uint8), ur=tonp(urB[0, :n0]), fv=fv(VB))
        f0, e0 = tonp(F12[0]), tonp(epi[0])
        w12 = np.zeros(n0, np.int32)
        pp = lambda a: a.ctypes.data
        t0 = time.perf_counter()
        for _ in range(20):
            w_nm = O.orbo_search_for_triangulation(n0, pp(tA["kp"]), pp(tA["d"]), pp(tA["has"]), pp(tA["ur"]), len(tA["fv"][0]), pp(tA["fv"][0]), pp(tA["fv"][1]), pp(tA["fv"][2]),
                                                   n0, pp(tB["kp"]), pp(tB["d"]), pp(tB["has"]), pp(tB["ur"]), len(tB["fv"][0]), pp(tB["fv"][0]), pp(tB["fv"][1]), pp(tB["fv"][2]),
                                                   pp(f0), pp(e0), pp(t_scale), pp(t_sigma2), 0, 1, pp(w12))
        dt = (time.perf_counter() - t0) / 20
        assert w_nm == int(tnm[0].item()) and w_nm > 100 and (w12 == t12[0, :n0].cpu().numpy()).all(), "SearchForTriangulation disagrees with the CPU checker"
        out["triangulation"]["cpu_baseline"] = {"value": 1.0 / dt, "unit": "problems/s", "cores": 1, "kind": "port", "sample": "problem 0, 20 repetitions, results compared"}
        # fuse, problem 0
        from fuse_lib import window_best_free_oracle
        n0 = int(nF[0].item())
        kp0 = tonp(kF[0, :n0]).view(oracle_lib.KP_DTYPE).reshape(n0)
        q0 = dict(uvr=tonp(fq["uvr"][0]), level=tonp(fq["level"][0]), ur=tonp(fq["ur"][0]))
        t0 = time.perf_counter()
        for _ in range(5):
            w_n, w_bi, w_bd = window_best_free_oracle(kp0, tonp(dF[0, :n0]), tonp(urF[0, :n0]), bounds, q0, tonp(fq["desc"][0]), inv_sigma2, 50)
        dt = (time.perf_counter() - t0) / 5
        assert w_n == int(f_nf[0].item()) and w_n > 100 and (w_bi == f_bi[0].cpu().numpy()).all() and (w_bd == f_bd[0].cpu().numpy()).all(), \
            "the Fuse search disagrees with the CPU checker"
        out["fuse"]["cpu_baseline"] = {"value": 1.0 / dt, "unit": "problems/s", "cores": 1, "kind": "port", "sample": "problem 0, 5 repetitions, results compared"}
        # one Fuse search through the host-array entry point (what LocalMapping::SearchInNeighbors calls per neighbour key frame)
        F1 = ob.FrameView(kp0.view(ob.KP_DTYPE), tonp(dF[0, :n0]), bounds, tonp(urF[0, :n0]))
        nrep = 1 if os.environ.get("ORB_BENCH_PROFILE") else 30
        for _ in range(min(5, nrep)):
            one = ob.window_best_free(F1, q0["uvr"], q0["level"], tonp(fq["desc"][0]), 50, ur=q0["ur"], inv_sigma2=inv_sigma2)
        t0 = time.perf_counter()
        for _ in range(nrep):
            one = ob.window_best_free(F1, q0["uvr"], q0["level"], tonp(fq["desc"][0]), 50, ur=q0["ur"], inv_sigma2=inv_sigma2)
        out["fuse"]["single_call_ms"] = 1e3 * (time.perf_counter() - t0) / nrep
        assert one[0] == w_n and (one[1] == w_bi).all(), "single Fuse search differs from the batch"
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
def kernel_limits():
    """What bounds each kernel according to the committed ncu capture of this round (profiles/r02_kernel_limits.json,
    written by tools/ncu_limits.py from an `ncu --set full` report): issue rate, pipe utilisation, instructions per unit."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_kernel_limits.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def oracle_parity(workload, frames_np, first, got, stereo_got=None):
    """Compare extraction results of frames `frames_np` with the CPU checker (oracle/, the C restatement pinned to the
    reference): every keypoint field, every descriptor byte, and for stereo pairs uRight / depth bit patterns.
    `got` = (kps [f,cap] structured, desc [f,cap,32], n [f]); stereo_got = (ur [p,cap], depth [p,cap], ns [p]).
    Raises SystemExit on the first difference: a number from a wrong result is not a number."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    _, _, nf, _ = WORKLOADS[workload]
    kps, desc, n = got
    O = [oracle_lib.OracleExtractor(nf, SCALE, LEVELS, INI_TH, MIN_TH) for _ in range(2)]
    L = oracle_lib.oracle()
    checked = 0
    for i in range(len(frames_np)):
        m, k_o, d_o = O[i & 1].extract(frames_np[i])
        if m != int(n[i]):
            raise SystemExit(f"bench.py: PARITY FAILURE frame {first + i}: {int(n[i])} keypoints, the checker has {m}")
        k_g = kps[i, :m]
        for f in k_o.dtype.names:
            if not (k_o[f].view(np.uint32) == k_g[f].view(np.uint32)).all():
                raise SystemExit(f"bench.py: PARITY FAILURE frame {first + i}: keypoint field {f}")
        if not (d_o == desc[i, :m]).all():
            raise SystemExit(f"bench.py: PARITY FAILURE frame {first + i}: descriptors")
        if stereo_got is not None and (i & 1):
            ur, dep, ns = stereo_got
            pl = i // 2
            nl = int(n[i - 1])
            kl, dl = np.ascontiguousarray(kps[i - 1, :nl]), np.ascontiguousarray(desc[i - 1, :nl])
            our, odep = np.zeros(nl, np.float32), np.zeros(nl, np.float32)
            L.orbo_stereo_matches.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                              C.c_float, C.c_float, C.c_void_p, C.c_void_p]
            kr, dr = np.ascontiguousarray(k_o), np.ascontiguousarray(d_o)
            nd = L.orbo_stereo_matches(O[0].h, O[1].h, nl, kl.ctypes.data, dl.ctypes.data, m, kr.ctypes.data, dr.ctypes.data,
                                       CAM["bf"], CAM["fx"], our.ctypes.data, odep.ctypes.data)
            if nd != int(ns[pl]) or not (ur[pl, :nl].view(np.uint32) == our.view(np.uint32)).all() or \
               not (dep[pl, :nl].view(np.uint32) == odep.view(np.uint32)).all():
                raise SystemExit(f"bench.py: PARITY FAILURE pair {(first + i) // 2}: mvuRight / mvDepth")
        checked += 1
    for o in O:
        o.close()
    return checked


def frame_crcs(kps, desc, n):
    """CRC32 of every frame's result (the valid keypoint records followed by their descriptors)."""
    import zlib
    out = np.zeros(len(n), np.int64)
    for i in range(len(n)):
        m = int(n[i])
        out[i] = zlib.crc32(desc[i, :m].tobytes(), zlib.crc32(kps[i, :m].tobytes()))
    return out


def measure_workload(workload, batch_total, scaling, rank, world, local_rank, dist, args, cpu, chunk=0):
    """One workload on this rank's GPU: device-resident throughput, per-stage times, end to end with pinned host
    buffers, parity against the CPU checker, roofline of the dominant stage.  scaling "weak": every rank extracts
    `batch_total` frames of its own; "strong": ONE batch of `batch_total` frames is split over the ranks
    (shard.frame_block) and the per-frame CRCs of all ranks are checked against rank 0's single-GPU result;
    "single": rank 0 alone.  Returns the result dict on rank 0, None elsewhere."""
    import torch
    import orb_slam2_chinesenotes_b200 as ob
    from orb_slam2_chinesenotes_b200.shard import frame_block
    dev = torch.device("cuda", local_rank)
    w, h, nf, _ = WORKLOADS[workload]
    stereo = workload in STEREO
    if scaling == "single" and rank != 0:
        return None
    nranks = 1 if scaling == "single" else world
    if scaling == "strong":
        first, stop = frame_block(batch_total, rank, nranks)
        batch = stop - first
    else:
        first, batch = rank * batch_total, batch_total
    sync_ranks = dist is not None and nranks > 1
    ex = ob.ORBextractor(nf, SCALE, LEVELS, INI_TH, MIN_TH, device=local_rank)
    if chunk:
        ex.set_chunk(chunk)
    cap = ex.default_capacity()
    frames = workload_frames(workload, first, batch, dev)
    pairs = batch // 2
    d_ur = torch.full((max(pairs, 1), cap), -1.0, dtype=torch.float32, device=dev)
    d_dep = torch.full((max(pairs, 1), cap), -1.0, dtype=torch.float32, device=dev)
    d_ns = torch.zeros(max(pairs, 1), dtype=torch.int32, device=dev)
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
        if sync_ranks:
            dist.barrier()
            torch.cuda.synchronize()

    steps = args.steps
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    # ---- device-resident throughput (`value`)
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(steps):
        step_device()
    e1.record(stream)
    barrier()
    ms_dev = e0.elapsed_time(e1)
    # ---- per-stage kernel durations: a second timed pass of the same steps with CUDA events around every
    # launch; the library then keeps all kernels on the one stream, so an event pair brackets one stage
    ex.profile(True)
    ex.stage_ms(reset=True)
    for _ in range(steps):
        step_device()
    barrier()
    stage_ms, stage_cnt = ex.stage_ms(reset=True)
    ex.profile(False)
    nk = d_n.cpu().numpy()
    if not ((nk > 0).all() and (nk <= cap).all()):
        raise SystemExit("bench.py: extraction produced no keypoints / overflowed")
    ns_dev = d_ns.cpu().numpy()
    n_depth = 0
    if stereo:
        if not (ns_dev > 0).all():
            raise SystemExit("bench.py: stereo matching found nothing")
        n_depth = int((d_dep > 0).sum().item())

    # ---- end to end through the C ABI with HOST buffers (pinned): H2D + kernels + D2H per step
    h_frames = torch.empty((batch, h, w), dtype=torch.uint8).pin_memory()
    h_frames.copy_(frames)
    h_kps = torch.empty((batch, cap, 7), dtype=torch.float32).pin_memory()
    h_desc = torch.empty((batch, cap, 32), dtype=torch.uint8).pin_memory()
    h_n = torch.empty(batch, dtype=torch.int32).pin_memory()
    h_ur = torch.empty((max(pairs, 1), cap), dtype=torch.float32).pin_memory()
    h_dep = torch.empty((max(pairs, 1), cap), dtype=torch.float32).pin_memory()
    h_ns = torch.empty(max(pairs, 1), dtype=torch.int32).pin_memory()
    h2d_bytes = batch * w * h
    d2h_bytes = batch * (cap * 60 + 4) + (pairs * (cap * 8 + 4) if stereo else 0)

    # what the copies alone cost on this box: the step's host->device and device->host bytes on two streams, all ranks
    # at once, nothing else running (the ceiling of any end-to-end number)
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    stage_in = torch.empty_like(frames)

    def copies():
        with torch.cuda.stream(s_in):
            stage_in.copy_(h_frames, non_blocking=True)
        with torch.cuda.stream(s_out):
            h_kps.copy_(d_kps, non_blocking=True)
            h_desc.copy_(d_desc, non_blocking=True)
            if stereo:
                h_ur.copy_(d_ur, non_blocking=True)
                h_dep.copy_(d_dep, non_blocking=True)

    copies()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        copies()
    barrier()
    ms_copy = 1e3 * (time.perf_counter() - t0)
    del stage_in

    def step_host():
        if stereo:
            ex.extract_stereo_batch_raw(h_frames, h * w, pairs, w, h, w, h_kps, h_desc, cap, h_n, CAM["bf"], CAM["fx"],
                                        h_ur, h_dep, h_ns, asynchronous=False)
        else:
            ex.extract_batch_raw(h_frames, h * w, batch, w, h, w, h_kps, h_desc, cap, h_n, asynchronous=False)

    step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        step_host()
    barrier()
    ms_e2e = 1e3 * (time.perf_counter() - t0)
    # the same steps SUBMITTED back to back (the _async entry points take pinned host buffers and only enqueue), one
    # synchronisation at the end: what a stream of batches sees -- the next step's uploads run beside this step's last kernels
    # and downloads.  Two sets of result buffers alternate, as a consumer of step k's results would need while step k+1 runs.
    h2 = [torch.empty_like(t).pin_memory() for t in (h_kps, h_desc, h_n, h_ur, h_dep, h_ns)]
    sets = [(h_kps, h_desc, h_n, h_ur, h_dep, h_ns), tuple(h2)]

    def submit(i):
        k_, d_, n_, u_, p_, s_ = sets[i % 2]
        if stereo:
            ex.extract_stereo_batch_raw(h_frames, h * w, pairs, w, h, w, k_, d_, cap, n_, CAM["bf"], CAM["fx"], u_, p_, s_, asynchronous=True)
        else:
            ex.extract_batch_raw(h_frames, h * w, batch, w, h, w, k_, d_, cap, n_, asynchronous=True)

    keep_sync = [t.clone() for t in (h_n, h_ns)] + [h_desc[:4].clone(), h_kps[:4].clone()]
    submit(0); submit(1)
    ex.sync()
    barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        submit(i)
    ex.sync()
    barrier()
    ms_pipe = 1e3 * (time.perf_counter() - t0)
    m_ = int(keep_sync[0][:4].min())                    # entries behind n_out are unspecified: compare the valid ones
    for k_, d_, n_, u_, p_, s_ in sets:                 # both buffer sets hold the synchronous call's results
        if not ((n_ == keep_sync[0]).all() and (not stereo or (s_ == keep_sync[1]).all()) and (d_[:4, :m_] == keep_sync[2][:, :m_]).all()
                and (k_[:4, :m_].contiguous().view(torch.uint8) == keep_sync[3][:, :m_].contiguous().view(torch.uint8)).all()):
            raise SystemExit("bench.py: pipelined submissions and the synchronous call disagree")
    del h2
    clocks = sampler.stop()

    # ---- the results that were timed are the reference's results: first and last frames of the batch, both paths
    kp_dt = ob.KP_DTYPE
    hk = h_kps.numpy().view(np.uint8).reshape(batch, cap, 28).view(kp_dt).reshape(batch, cap)
    hd, hn = h_desc.numpy(), h_n.numpy()
    if not (hn == nk).all() or (stereo and not (h_ns.numpy()[:pairs] == ns_dev[:pairs]).all()):
        raise SystemExit("bench.py: host-buffer path and device-resident path disagree")
    ncheck = min(batch, 4 if not stereo else 4)
    idx = list(range(ncheck)) + [i for i in range(batch - ncheck, batch) if i >= ncheck]
    blocks = [(0, ncheck)] + ([(batch - ncheck, batch)] if batch - ncheck >= ncheck else [])
    checked = 0
    for lo, hi in blocks:
        fr = np.ascontiguousarray(frames[lo:hi].cpu().numpy())
        dk = d_kps[lo:hi].cpu().numpy().view(np.uint8).reshape(hi - lo, cap, 28).view(kp_dt).reshape(hi - lo, cap)
        sg_d = sg_h = None
        if stereo:
            sg_d = (d_ur[lo // 2:hi // 2].cpu().numpy(), d_dep[lo // 2:hi // 2].cpu().numpy(), ns_dev[lo // 2:hi // 2])
            sg_h = (h_ur.numpy()[lo // 2:hi // 2], h_dep.numpy()[lo // 2:hi // 2], h_ns.numpy()[lo // 2:hi // 2])
        checked += oracle_parity(workload, fr, first + lo, (dk, d_desc[lo:hi].cpu().numpy(), nk[lo:hi]), sg_d)
        oracle_parity(workload, fr, first + lo, (hk[lo:hi], hd[lo:hi], hn[lo:hi]), sg_h)
    parity = {"frames_checked_against_cpu_checker": checked, "paths": ["device resident", "host buffers"],
              "fields": "all keypoint fields (bit patterns), descriptors" + (", mvuRight / mvDepth bits, match counts" if stereo else ""),
              "which": f"first and last {ncheck} frames of each rank's batch", "result": "identical"}

    # ---- strong scaling: the sharded batch equals the single-GPU batch (CRC32 per frame, gathered over NCCL)
    crc_check = None
    if scaling == "strong":
        mine = torch.from_numpy(frame_crcs(hk, hd, hn)).to(dev)
        if sync_ranks:
            sizes = [frame_block(batch_total, r, nranks) for r in range(nranks)]
            maxb = max(b - a for a, b in sizes)
            pad = torch.zeros(maxb, dtype=torch.int64, device=dev)
            pad[:batch] = mine
            gathered = [torch.zeros(maxb, dtype=torch.int64, device=dev) for _ in range(nranks)]
            dist.all_gather(gathered, pad)
            allcrc = torch.cat([g[:b - a] for g, (a, b) in zip(gathered, sizes)]).cpu().numpy()
        else:
            allcrc = mine.cpu().numpy()
        if rank == 0:
            # rank 0 extracts the WHOLE batch on its own GPU (device resident, in pieces) and compares
            want = np.zeros(batch_total, np.int64)
            piece = max(batch, 1)
            for p0 in range(0, batch_total, piece):
                pn = min(piece, batch_total - p0)
                if p0 == first and pn == batch:
                    fr = frames
                else:
                    fr = workload_frames(workload, p0, pn, dev)
                ex.extract_batch_raw(fr, h * w, pn, w, h, w, d_kps, d_desc, cap, d_n, asynchronous=True)
                torch.cuda.synchronize()
                pk = d_kps[:pn].cpu().numpy().view(np.uint8).reshape(pn, cap, 28).view(kp_dt).reshape(pn, cap)
                want[p0:p0 + pn] = frame_crcs(pk, d_desc[:pn].cpu().numpy(), d_n[:pn].cpu().numpy())
                del fr
            if not (want == allcrc).all():
                raise SystemExit(f"bench.py: SHARDING FAILURE: {int((want != allcrc).sum())} of {batch_total} frames differ from the single-GPU result")
            crc_check = {"frames": batch_total, "ranks": nranks, "result": "per-frame CRC32 (keypoints + descriptors) of the sharded batch = single-GPU batch"}

    # ---- max over ranks
    total_kp = float(nk.sum())
    if sync_ranks:
        t = torch.tensor([ms_dev, ms_e2e, ms_copy, ms_pipe], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_dev, ms_e2e, ms_copy, ms_pipe = float(t[0]), float(t[1]), float(t[2]), float(t[3])
        kp = torch.tensor([total_kp, float(n_depth)], device=dev, dtype=torch.float64)
        dist.all_reduce(kp)
        total_kp, n_depth = float(kp[0]), float(kp[1])
    torch.cuda.set_stream(torch.cuda.default_stream(dev))
    if rank != 0:
        ex.close()
        return None
    frames_all = batch_total * nranks if scaling == "weak" else batch_total
    value = frames_all * steps / (ms_dev * 1e-3)
    e2e_value = frames_all * steps / (ms_e2e * 1e-3)
    # ---- roofline of the dominant kernel (stage times measured above with CUDA events on the stream)
    bytes_per_frame = algorithmic_bytes(w, h)
    dom = max(("pyramid", "fast", "blur", "octree", "describe", "stereo"), key=lambda s_: stage_ms[s_])
    peak, peak_src = measured_peak()
    roof = {"bound": "hbm", "kernel": dom, "peak": peak, "unit": "GB/s", "peak_source": peak_src, "traffic": None,
            "stage_timing": "second pass of the same steps, kernels serialised on one stream, CUDA events around each launch",
            "stage_ms_per_step": {k: v / steps for k, v in stage_ms.items()}}
    if dom in bytes_per_frame:
        launches = max(stage_cnt[dom], 1)
        frames_per_launch = batch * steps / launches
        dur_s = stage_ms[dom] * 1e-3 / launches
        roof["achieved"] = bytes_per_frame[dom] * frames_per_launch / dur_s / 1e9
        roof["frac"] = roof["achieved"] / peak
        roof["algorithmic_bytes_per_frame"] = bytes_per_frame[dom]
        roof["algorithmic_bytes_per_launch"] = bytes_per_frame[dom] * frames_per_launch
        roof["traffic"] = measured_traffic(dom, frames_per_launch) if workload == "kitti_1241x376_nf2000" else None
    else:
        roof["achieved"] = None
        roof["frac"] = None
    # what actually bounds the dominant kernel: from the committed ncu capture of this round (not measured live)
    lim = kernel_limits().get(dom)
    if lim:
        roof["limiter"] = lim
        roof["issue_frac"] = lim.get("issue_frac")
        roof["binding"] = lim.get("binding")
    roof["dense_stages"] = {s_: {"GB/s": bytes_per_frame[s_] * batch * steps / (stage_ms[s_] * 1e-3) / 1e9,
                                 "frac": bytes_per_frame[s_] * batch * steps / (stage_ms[s_] * 1e-3) / 1e9 / peak}
                            for s_ in bytes_per_frame if stage_ms[s_] > 0}
    # ---- CPU baseline on a bounded sample of the same frames (rank 0, N=1 only)
    cpu_base = None
    if cpu:
        cores = os.cpu_count() or 1
        ns = int(min(batch, max(16, 4 * cores))) & ~1 or batch
        sample = frames[:ns].cpu().numpy()
        v, kind, what, _ = cpu_workload_frames_per_s(workload, sample, nf, cores)
        cpu_base = {"value": v, "unit": "frames/s", "cores": cores, "kind": kind,
                    "sample": f"first {ns} frames of the step's batch, {what}"}
        ocv = opencv_dense_frames_per_s(sample, cores)
        if ocv is not None:
            cpu_base["opencv_dense_stages_only"] = {"value": ocv["frames_per_s"], "unit": "frames/s", "cores": cores,
                                                    "ms_per_frame_per_core": ocv["ms_per_frame_per_core"],
                                                    "what": "cv2 (real OpenCV, SIMD) pyramid + FAST per level + GaussianBlur only, one thread per core, same "
                                                            "frames; informative: brackets the reference arm, which runs over a scalar OpenCV stand-in"}
    chunk_dev = chunk or 512                      # device-resident default of orb_capi.cu (the timed `value` path)
    chunks = (batch + chunk_dev - 1) // chunk_dev
    stats = {"keypoints_per_frame": total_kp / frames_all}
    if stereo:
        stats["depth_points_per_pair"] = n_depth / (frames_all / 2)
    res = {
        "value": value, "unit": "frames/s", "ms_per_step": ms_dev / steps, "scaling": "weak" if scaling == "single" else scaling,
        "n_gpus": nranks, "config": workload_config(workload, batch_total, nranks, "weak" if scaling != "strong" else "strong"), "stats": stats,
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                "ms_per_step": ms_e2e / steps, "copy_ceiling_frames_per_s": frames_all * steps / (ms_copy * 1e-3),
                "copy_ceiling_note": "the step's pinned host->device and device->host copies alone, two streams, all ranks at once (max over ranks)",
                "note": "value: one synchronous call per step (the call returns when the step's results are on the host)",
                "pipelined_submissions_frames_per_s": frames_all * steps / (ms_pipe * 1e-3),
                "pipelined_note": "the same steps through the _async entry points with the same pinned host buffers (two alternating result sets), "
                                  "one orbx_sync at the end: consecutive steps overlap; results compared with the synchronous call's"},
        "gpu_launches": ex.launches_per_chunk(stereo) * chunks * steps,
        "roofline": roof, "cpu_baseline": cpu_base, "clocks": clocks, "parity": parity,
    }
    if crc_check:
        res["sharding_check"] = crc_check
    res["_keep"] = (ex, d_kps, d_desc, d_n, cap, h_frames, frames)
    return res


def dropin_latency(image, nf):
    """Milliseconds per call of the C++ drop-in class ORB_SLAM2::ORBextractor (host/ORBextractor.{h,cc}) timed INSIDE C++
    by tests/cpp/dropin_harness.cc, i.e. what Frame::ExtractORB (src/Frame.cc:262-268) costs, mvImagePyramid included.
    The class needs OpenCV's cv::Mat / cv::KeyPoint types to compile; this image has no OpenCV C++, so the harness is
    built against the container-only stand-in the tests use (oracle/cvshim: types, no arithmetic on this path)."""
    cpp = os.path.join(ROOT, "tests", "cpp")
    try:
        if subprocess.run(["make", "-C", cpp, "all"], capture_output=True).returncode != 0:
            return {}
        D = C.CDLL(os.path.join(cpp, "_build", "libdropin.so"))
        D.dropin_create.restype = C.c_void_p
        D.dropin_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        D.dropin_destroy.argtypes = [C.c_void_p]
        D.dropin_time_ms.restype = C.c_double
        D.dropin_time_ms.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int]
        hd = D.dropin_create(nf, SCALE, LEVELS, INI_TH, MIN_TH)
        img = np.array(image)                          # pageable memory, as a cv::Mat from an image decoder would be
        h, w = img.shape
        out = {"dropin_operator_ms": D.dropin_time_ms(hd, img.ctypes.data, w, h, img.strides[0], 50, 1),
               "dropin_operator_no_pyramid_ms": D.dropin_time_ms(hd, img.ctypes.data, w, h, img.strides[0], 50, 0),
               "dropin_note": "ORB_SLAM2::ORBextractor::operator() of host/ORBextractor.cc timed inside C++ (tests/cpp/dropin_harness.cc), pageable image, "
                              "keypoint vector + descriptor Mat + all 8 REFLECT_101-padded mvImagePyramid levels filled"}
        D.dropin_destroy(hd)
        return out
    except Exception as e:                             # no compiler on the box: the ctypes numbers above still stand
        return {"dropin_note": f"not measured: {e}"}


def pin_rank_to_cores(local_rank, world):
    """One rank per GPU on one host: give every rank its own slice of the host cores so the staging threads of
    eight ranks do not migrate over each other (all GPUs of these boxes hang off NUMA node 0)."""
    try:
        cores = sorted(os.sched_getaffinity(0))
        per = len(cores) // world
        if world > 1 and per >= 2:
            os.sched_setaffinity(0, set(cores[local_rank * per:(local_rank + 1) * per]))
            return per
    except Exception:
        pass
    return None


def run_ours(args, rank, world, local_rank):
    import torch
    import orb_slam2_chinesenotes_b200 as ob
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    pinned = pin_rank_to_cores(local_rank, world)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # NCCL's banner must not share stdout with the JSON line
        dist.init_process_group("nccl", device_id=dev)
    w, h, nf, batch = WORKLOADS[args.workload]
    if args.batch:
        batch = args.batch
    stereo = args.workload in STEREO
    if stereo:
        batch += batch % 2
    cpu = world == 1 and not args.no_cpu
    main = measure_workload(args.workload, batch, args.scaling, rank, world, local_rank, dist, args, cpu, args.chunk)

    # ---- the other BASELINE.json configs in the same line (default run only): configs[0] one 640x480 frame, configs[2]
    # EuRoC x256 on one GPU, configs[4] 4096 HD frames sharded over all ranks (strong scaling)
    others = []
    if not args.only_main and args.workload == "kitti_1241x376_nf2000" and args.scaling == "weak" and not args.batch:
        for name, b, sc in (("tum_640x480_nf1000", GEN_BLOCK, "single"), ("euroc_752x480_nf1200", 256, "single"),
                            ("hd_1920x1080_nf4000", args.hd_batch, "strong")):
            sub = argparse.Namespace(**vars(args))
            sub.steps, sub.warmup = min(args.steps, 3), 3
            r = measure_workload(name, b, sc, rank, world, local_rank, dist, sub, cpu and rank == 0)
            if r is not None:
                ex1, _, _, _, _, h_fr, _ = r.pop("_keep")
                if name == "tum_640x480_nf1000":      # configs[0]: ONE frame through operator(), host buffers
                    one = h_fr[0].numpy()
                    for _ in range(5):
                        ex1(one)
                    t0 = time.perf_counter()
                    for _ in range(30):
                        ex1(one)
                    r["latency"] = {"extract_1_frame_ms": 1e3 * (time.perf_counter() - t0) / 30,
                                    "note": "configs[0]: operator() on one 640x480 frame, nfeatures 1000, host buffers, ctypes overhead included"}
                ex1.close()
                others.append(r)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    ex, d_kps, d_desc, d_n, cap, h_frames, frames = main.pop("_keep")
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
        latency.update(dropin_latency(one[0], nf))
    hamming = bench_hamming(dev, local_rank) if world == 1 else None
    window = mappoint = None
    if world == 1 and stereo:
        window = bench_window_match(dev, d_kps, d_desc, d_n, cap, w, h, ex.GetScaleFactors(), cpu=not args.no_cpu)
        mappoint = bench_mappoint_side(dev, d_kps, d_desc, d_n, cap, w, h, SCALE, cpu=not args.no_cpu)
    line = {"metric": "orb_extract_frames_per_s", "value": main.pop("value"), "unit": main.pop("unit"), "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": main.pop("ms_per_step"),
            "higher_is_better": True, "scaling": main.pop("scaling"), "vs_baseline": None, "dtype": "u8", "data": "synthetic"}
    main.pop("n_gpus")
    line.update(main)
    line.update({"hamming_bf": hamming, "window_match": window, "mappoint_side": mappoint, "latency": latency,
                 "other_workloads": others, "host_cores_per_rank": pinned})
    emit(line)
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
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: every rank extracts --batch frames; strong: ONE batch of --batch frames is sharded over the ranks")
    ap.add_argument("--only-main", action="store_true", help="skip the other BASELINE configs (other_workloads)")
    ap.add_argument("--hd-batch", type=int, default=4096, help="frames of the configs[4] strong-scaling workload")
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
