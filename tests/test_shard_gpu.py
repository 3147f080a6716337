"""T4 on hardware: a batch sharded over ranks equals the single-GPU batch, bit for bit and in
order.  Two processes share the one GPU of the test box (the ranks of a real run own one GPU each); each extracts
its shard.frame_block through the C ABI, the per-rank statistics travel over gloo (NCCL refuses two ranks on one
device; bench.py --scaling strong does the same check over NCCL on N GPUs with per-frame CRCs)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
W, H, NF, NFRAMES = 752, 480, 1200, 7


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import torch.distributed as dist
    import orb_slam2_chinesenotes_b200 as ob
    from orb_slam2_chinesenotes_b200.shard import extract_sharded
    from synth import synth_frame
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    frames = np.stack([synth_frame(W, H, 900 + i) for i in range(NFRAMES)])
    G = ob.ORBextractor(NF, 1.2, 8, 20, 7, device=0)
    start, stop, res, stats = extract_sharded(frames, lambda block: G.extract_batch(block), rank, world)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), start=start, stop=stop, kps=res[0], desc=res[1], n=res[2],
             frames=[s["frames"] for s in stats], keypoints=[s["keypoints"] for s in stats])
    G.close()
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_gpu_batch_equals_single_gpu_batch(tmp_path):
    import torch.multiprocessing as mp
    import orb_slam2_chinesenotes_b200 as ob
    from synth import synth_frame
    world, port = 2, 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    assert [int(p["start"]) for p in parts] == [0, 3] and [int(p["stop"]) for p in parts] == [3, 7]
    n = np.concatenate([p["n"] for p in parts])
    kps = np.concatenate([p["kps"] for p in parts])
    desc = np.concatenate([p["desc"] for p in parts])
    frames = np.stack([synth_frame(W, H, 900 + i) for i in range(NFRAMES)])
    G = ob.ORBextractor(NF, 1.2, 8, 20, 7, device=0)
    k1, d1, n1 = G.extract_batch(frames)          # the whole batch on one GPU, one process
    G.close()
    assert (n == n1).all() and (n > 0).all()
    for i in range(NFRAMES):
        m = int(n[i])
        assert (desc[i, :m] == d1[i, :m]).all()
        assert all((kps[i, :m][f].view(np.uint32) == k1[i, :m][f].view(np.uint32)).all() for f in k1.dtype.names)
    for p in parts:                                # every rank saw the same gathered statistics
        assert p["frames"].tolist() == [3, 4] and p["keypoints"].tolist() == [int(n[:3].sum()), int(n[3:].sum())]
