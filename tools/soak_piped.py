"""Ad-hoc soak of the host-buffer pipeline (chunks rotating through three buffer sets and two compute streams, uploads / kernels /
downloads overlapped): random shapes, feature counts, batch sizes and chunk sizes; the results of the piped call (host buffers,
stereo pairs included) must equal those of the device-resident call (one stream, everything in place) frame by frame, bit for bit.
Usage (GPU box): python tools/soak_piped.py [cases] [seed]"""
import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import torch
import orb_slam2_chinesenotes_b200 as ob

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 10
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
bad = 0
for c in range(cases):
    w, h = int(rng.integers(200, 1300)), int(rng.integers(150, 700))
    nf = int(rng.integers(200, 2500))
    pairs = int(rng.integers(9, 160))
    b = 2 * pairs
    G = ob.ORBextractor(nf, 1.2, 8, 20, 7)
    if not G._L.orbx_shape_supported(G._h, w, h):
        G.close(); continue
    g = torch.Generator(device="cuda").manual_seed(int(rng.integers(1 << 30)))
    base = torch.randint(0, 256, (b, h // 4 + 1, w // 4 + 1), generator=g, device="cuda", dtype=torch.uint8)
    frames = torch.nn.functional.interpolate(base[:, None].float(), size=(h, w), mode="bilinear")[:, 0].round().clamp(0, 255).to(torch.uint8).contiguous()
    frames[1::2, :, 7:] = frames[0::2, :, :-7]                                   # right = left shifted: stereo matches exist
    cap = G.default_capacity()
    mk = lambda *s, dt=torch.float32, fill=0: torch.full(s, fill, dtype=dt, device="cuda")
    d = dict(kps=mk(b, cap, 7), desc=mk(b, cap, 32, dt=torch.uint8), n=mk(b, dt=torch.int32), ur=mk(pairs, cap, fill=-1), dep=mk(pairs, cap, fill=-1),
             ns=mk(pairs, dt=torch.int32))
    G.extract_stereo_batch_raw(frames, h * w, pairs, w, h, w, d["kps"], d["desc"], cap, d["n"], 386.1448, 718.856, d["ur"], d["dep"], d["ns"])
    torch.cuda.synchronize()
    chunk = int(rng.choice([0, 0, 16, 32, 48]))
    if chunk:
        G.set_chunk(chunk)
    hk, hd, hn, hur, hdep, hns = G.extract_stereo_batch(frames.cpu().numpy(), 386.1448, 718.856)
    n_d = d["n"].cpu().numpy()
    ok = (hn == n_d).all() and (hns == d["ns"].cpu().numpy()).all()
    kd, dd = d["kps"].cpu().numpy().view(np.uint32), d["desc"].cpu().numpy()
    for f in range(b):
        k = int(hn[f])
        ok = ok and hk[f, :k].tobytes() == kd[f, :k].tobytes() and (hd[f, :k] == dd[f, :k]).all()
    for p_ in range(pairs):
        k = int(hn[2 * p_])
        ok = ok and (hur[p_, :k].view(np.uint32) == d["ur"][p_, :k].cpu().numpy().view(np.uint32)).all() \
            and (hdep[p_, :k].view(np.uint32) == d["dep"][p_, :k].cpu().numpy().view(np.uint32)).all()
    print(c, (w, h, nf, pairs, chunk), int(hn.sum()), int(hns.sum()), "ok" if ok else "MISMATCH", flush=True)
    bad += 0 if ok else 1
    G.close()
print("mismatches:", bad)
