"""T2/T3 (GPU): the CUDA extractor through the C ABI against the oracle, stage by stage and end
to end, on the BASELINE.json shapes, adversarial images and batches.  Bit-exact everywhere:
pyramid / blur bytes, FAST candidates, quadtree output order, angles (float bits), descriptors."""
import os

import numpy as np
import pytest

import orb_slam2_chinesenotes_b200 as ob
from oracle_lib import OracleExtractor
from synth import synth_frame

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def same_kps(a, b):
    return len(a) == len(b) and all((a[f] == b[f]).all() for f in a.dtype.names)


def assert_frame_equal(G, O, img, stages=True):
    n, k_o, d_o = O.extract(img)
    k_g, d_g = G(img)
    assert n == len(k_g)
    assert same_kps(k_o, k_g), "keypoint fields differ"
    assert (k_o["angle"].view(np.uint32) == k_g["angle"].view(np.uint32)).all(), "angle bits differ"
    assert (d_o == d_g).all(), "descriptors differ"
    if stages:
        for l in range(O.nlevels):
            assert (O.pyramid(l) == G.pyramid(l)).all(), f"pyramid level {l}"
            assert (O.pyramid(l, True) == G.pyramid(l, with_border=True)).all(), f"border level {l}"
            bo = O.blurred(l)
            if bo is not None:
                assert (bo == G.blurred(l)).all(), f"blur level {l}"
            co = O.candidates(l)
            so = set(zip(co["x"].tolist(), co["y"].tolist(), co["score"].tolist()))
            assert so == set(map(tuple, G.candidates(l).tolist())), f"FAST candidates level {l}"
            ko = O.level_keypoints(l)
            lo = [(int(a["x"]) - 16, int(a["y"]) - 16, int(a["response"])) for a in ko]
            assert lo == [tuple(r) for r in G.level_keypoints(l).tolist()], f"quadtree order level {l}"


@pytest.mark.parametrize("w,h,nf,seed", [(640, 480, 1000, 1), (1241, 376, 2000, 2), (752, 480, 1200, 3),
                                          (1920, 1080, 4000, 4), (200, 150, 300, 5), (160, 200, 200, 25),
                                          (500, 150, 300, 26), (333, 257, 700, 27)])
def test_stages_and_outputs_bit_exact(w, h, nf, seed):
    G, O = ob.ORBextractor(nf, 1.2, 8, 20, 7), OracleExtractor(nf)
    assert_frame_equal(G, O, synth_frame(w, h, seed))
    G.close(); O.close()


@pytest.mark.parametrize("name", ["tum", "kitti", "small"])
def test_against_reference_fixtures(name):
    """Fixtures produced by the reference's unmodified ORBextractor.cc (tests/golden/make_golden.py)."""
    g = np.load(os.path.join(GOLD, f"ref_extract_{name}.npz"))
    G = ob.ORBextractor(int(g["nfeatures"]), 1.2, 8, 20, 7)
    kps, desc = G(synth_frame(int(g["w"]), int(g["h"]), int(g["seed"])))
    assert same_kps(kps, g["kps"]) and (desc == g["desc"]).all()
    t = dict(scale=G.GetScaleFactors(), inv_scale=G.GetInverseScaleFactors(), sigma2=G.GetScaleSigmaSquares(),
             inv_sigma2=G.GetInverseScaleSigmaSquares(), per_level=G.features_per_level())
    for k, v in t.items():
        assert (v == g["tab_" + k]).all(), k
    assert G.GetLevels() == 8 and np.float32(G.GetScaleFactor()) == np.float32(1.2)
    G.close()


def test_other_parameters():
    for (nf, sf, nl, ini, mn) in [(500, 1.2, 4, 20, 7), (1500, 1.5, 5, 30, 10), (800, 1.1, 8, 12, 5), (3000, 1.2, 8, 20, 7)]:
        G = ob.ORBextractor(nf, sf, nl, ini, mn)
        O = OracleExtractor(nf, sf, nl, ini, mn)
        assert_frame_equal(G, O, synth_frame(640, 480, nf), stages=False)
        G.close(); O.close()


def test_adversarial_images():
    h, w = 240, 320
    yy, xx = np.mgrid[0:h, 0:w]
    imgs = {
        "flat": np.full((h, w), 127, np.uint8),
        "white": np.full((h, w), 255, np.uint8),
        "black": np.zeros((h, w), np.uint8),
        "checker8": (((yy // 8 + xx // 8) % 2) * 255).astype(np.uint8),
        "checker3": (((yy // 3 + xx // 3) % 2) * 200 + 20).astype(np.uint8),
        "gradient": ((xx * 255) // w).astype(np.uint8),
        "noise": np.random.default_rng(9).integers(0, 256, (h, w), dtype=np.uint8),
        "dots": ((((yy % 7) == 0) & ((xx % 7) == 0)) * 255).astype(np.uint8),
    }
    G, O = ob.ORBextractor(500, 1.2, 8, 20, 7), OracleExtractor(500)
    for name, img in imgs.items():
        n, k_o, d_o = O.extract(img)
        k_g, d_g = G(img)
        assert n == len(k_g), name
        assert same_kps(k_o, k_g) and (d_o == d_g).all(), name
    assert G(np.zeros((0, 0), np.uint8)) == (None, None)      # empty image: outputs untouched
    G.close(); O.close()


def test_non_contiguous_pitch():
    """operator() accepts a cv::Mat view with step > cols."""
    big = synth_frame(700, 500, 41)
    view = big[10:490, 30:670]
    G, O = ob.ORBextractor(1000, 1.2, 8, 20, 7), OracleExtractor(1000)
    n, k_o, d_o = O.extract(np.ascontiguousarray(view))
    k_g, d_g = G(view)
    assert same_kps(k_o, k_g) and (d_o == d_g).all()
    G.close(); O.close()


def test_batch_equals_single_and_chunking():
    """Chunk sizes pick different launch shapes (FAST blocks take 1 / 4 / all cell rows for < 16 / < 64 / >= 64 frames) and
    chunks alternate between compute streams; every frame must still equal its single-frame result."""
    w, h, nf, B = 416, 240, 600, 20
    frames = np.stack([synth_frame(w, h, 100 + i) for i in range(B)])
    G, O = ob.ORBextractor(nf, 1.2, 8, 20, 7), OracleExtractor(nf)
    want = [O.extract(frames[i]) for i in range(B)]
    for chunk in (32, 16, 4, 1):
        G.set_chunk(chunk)
        kps, desc, n = G.extract_batch(frames)
        for i in range(B):
            no, k_o, d_o = want[i]
            assert no == n[i]
            assert same_kps(k_o, kps[i, :no]) and (d_o == desc[i, :no]).all(), (chunk, i)
    G.close(); O.close()


def test_capacity_overflow_is_reported():
    G = ob.ORBextractor(1000, 1.2, 8, 20, 7)
    with pytest.raises(ob.OrbError) as e:
        G(synth_frame(640, 480, 1), capacity=100)
    assert e.value.code == ob.ORBX_E_CAPACITY
    with pytest.raises(ob.OrbError) as e:
        G(synth_frame(60, 60, 1))
    assert e.value.code == ob.ORBX_E_SHAPE
    G.close()


def test_device_resident_async_path():
    import torch
    w, h, nf, B = 640, 480, 1000, 6
    frames = np.stack([synth_frame(w, h, 200 + i) for i in range(B)])
    G, O = ob.ORBextractor(nf, 1.2, 8, 20, 7), OracleExtractor(nf)
    cap = G.default_capacity()
    d_f = torch.from_numpy(frames).cuda()
    d_k = torch.zeros((B, cap, 7), dtype=torch.float32, device="cuda")
    d_d = torch.zeros((B, cap, 32), dtype=torch.uint8, device="cuda")
    d_n = torch.zeros(B, dtype=torch.int32, device="cuda")
    G.set_stream(torch.cuda.current_stream().cuda_stream)
    G.extract_batch_raw(d_f, h * w, B, w, h, w, d_k, d_d, cap, d_n, asynchronous=True)
    G.sync()
    n = d_n.cpu().numpy()
    kps = d_k.cpu().numpy().view(np.uint8).reshape(B, cap, 28).view(ob.KP_DTYPE).reshape(B, cap)
    desc = d_d.cpu().numpy()
    for i in range(B):
        no, k_o, d_o = O.extract(frames[i])
        assert no == n[i] and same_kps(k_o, kps[i, :no]) and (d_o == desc[i, :no]).all()
    G.close(); O.close()


def test_full_size_batch_properties():
    """BASELINE size (KITTI shape, 64 frames): size-independent properties + spot checks."""
    w, h, nf, B = 1241, 376, 2000, 64
    base = [synth_frame(w, h, 300 + i) for i in range(4)]
    frames = np.stack([base[i % 4] for i in range(B)])
    G, O = ob.ORBextractor(nf, 1.2, 8, 20, 7), OracleExtractor(nf)
    kps, desc, n = G.extract_batch(frames)
    for i in range(B):                                         # identical frames -> identical results (determinism)
        j = i % 4
        assert n[i] == n[j] and (desc[i, :n[i]] == desc[j, :n[j]]).all() and same_kps(kps[i, :n[i]], kps[j, :n[j]])
    for j in range(4):
        no, k_o, d_o = O.extract(base[j])
        assert no == n[j] and same_kps(k_o, kps[j, :no]) and (d_o == desc[j, :no]).all()
        k = kps[j, :no]
        assert (np.diff(k["octave"]) >= 0).all()               # appended level by level
        assert no >= nf and (k["response"] >= 7).all() and (k["angle"] >= 0).all() and (k["angle"] < 360).all()
    G.close(); O.close()


def test_two_extractors_on_two_threads_with_different_shapes():
    """Frame::ExtractORB runs the left and right extractor on two std::threads (src/Frame.cc:82-85); a SLAM process may
    also hold extractors for different cameras.  Two contexts, two image shapes (different shared-memory footprints of
    the same kernels), concurrent calls: same results as alone."""
    import threading
    jobs = [(640, 480, 1000, 1), (1241, 376, 2000, 2)]
    want, errors = [], []
    for w, h, nf, seed in jobs:
        G = ob.ORBextractor(nf, 1.2, 8, 20, 7)
        want.append(G(synth_frame(w, h, seed)))
        G.close()

    def work(i):
        try:
            w, h, nf, seed = jobs[i]
            G = ob.ORBextractor(nf, 1.2, 8, 20, 7)
            img = synth_frame(w, h, seed)
            for _ in range(15):
                k, d = G(img)
                if not (same_kps(k, want[i][0]) and (d == want[i][1]).all()):
                    errors.append("mismatch")
            G.close()
        except Exception as e:
            errors.append(repr(e))

    threads = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:3]


@pytest.mark.gpu
def test_async_submissions_with_host_buffers_overlap_and_agree():
    """The _async entry points take (pinned) host buffers too: three batches submitted back to back without waiting, one
    orbx_sync at the end; every batch's results equal the synchronous call's (frames differ per batch, so a mix-up of
    staging slots between calls in flight would show)."""
    import torch
    w, h, nf, per = 640, 480, 1000, 40
    ex = ob.ORBextractor(nf, 1.2, 8, 20, 7)
    ex.set_chunk(16)                                                       # several chunks per call: the slots rotate across calls
    cap = ex.default_capacity()
    batches = [np.stack([synth_frame(w, h, 300 + 50 * b + i) for i in range(per)]) for b in range(3)]
    want = [ex.extract_batch(fr) for fr in batches]
    pin = lambda a: torch.from_numpy(a).pin_memory()
    h_in = [pin(fr) for fr in batches]
    outs = [(torch.zeros((per, cap, 7), dtype=torch.float32).pin_memory(), torch.zeros((per, cap, 32), dtype=torch.uint8).pin_memory(),
             torch.zeros(per, dtype=torch.int32).pin_memory()) for _ in range(3)]
    for b in range(3):
        ex.extract_batch_raw(h_in[b], h * w, per, w, h, w, outs[b][0], outs[b][1], cap, outs[b][2], asynchronous=True)
    ex.sync()
    for b in range(3):
        kps, desc, n = want[b]
        gk = outs[b][0].numpy().view(np.uint8).reshape(per, cap, 28).view(ob.KP_DTYPE).reshape(per, cap)
        gd, gn = outs[b][1].numpy(), outs[b][2].numpy()
        assert (gn == n).all()
        for f in range(per):
            assert (gd[f, :n[f]] == desc[f, :n[f]]).all()
            assert gk[f, :n[f]].tobytes() == kps[f, :n[f]].tobytes()
    ex.close()


@pytest.mark.gpu
@pytest.mark.parametrize("x0,pitch_pad", [(31, 0), (8, 5), (3, 1)])
def test_device_resident_view_with_odd_offset_and_pitch(x0, pitch_pad):
    """Frames that are a window of larger device images: base address, row pitch and frame stride with no alignment at all
    (the pyramid and blur rings fetch aligned words around such rows; FAST falls back from its tensor-map copies)."""
    import torch
    w, h, nf, B = 640, 480, 1000, 3
    Wb, Hb = w + 60 + pitch_pad, h + 20
    big = np.stack([synth_frame(Wb, Hb, 700 + i) for i in range(B)])
    y0 = 10
    roi = np.ascontiguousarray(big[:, y0:y0 + h, x0:x0 + w])
    O = OracleExtractor(nf)
    want = [O.extract(roi[i]) for i in range(B)]
    G = ob.ORBextractor(nf, 1.2, 8, 20, 7)
    cap = G.default_capacity()
    d_big = torch.from_numpy(big).cuda()
    view = d_big[:, y0:y0 + h, x0:x0 + w]                                   # data_ptr() = first pixel of the window
    d_k = torch.zeros((B, cap, 7), dtype=torch.float32, device="cuda")
    d_d = torch.zeros((B, cap, 32), dtype=torch.uint8, device="cuda")
    d_n = torch.zeros(B, dtype=torch.int32, device="cuda")
    G.extract_batch_raw(view, Hb * Wb, B, w, h, Wb, d_k, d_d, cap, d_n, asynchronous=True)
    G.sync()
    n = d_n.cpu().numpy()
    kps = d_k.cpu().numpy().view(np.uint8).reshape(B, cap, 28).view(ob.KP_DTYPE).reshape(B, cap)
    desc = d_d.cpu().numpy()
    for i in range(B):
        no, k_o, d_o = want[i]
        assert no == n[i]
        assert same_kps(k_o, kps[i, :no]) and (d_o == desc[i, :no]).all(), i
    G.close(); O.close()
