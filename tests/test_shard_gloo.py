"""Multi-process sharding logic on CPU (gloo, world_size 2): block partition, per-rank extraction of the
rank's block, all-gather of counts; the concatenation in rank order equals the single-process result.
The per-frame worker here is the ORACLE (there is no GPU in this container); on GPUs the worker is
ORBextractor.extract_batch and the backend is NCCL (bench.py)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def test_frame_block_partition():
    from orb_slam2_chinesenotes_b200.shard import frame_block
    for n in (0, 1, 7, 8, 64, 4096, 4097):
        for world in (1, 2, 3, 4, 8):
            blocks = [frame_block(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import torch.distributed as dist
    from oracle_lib import KP_DTYPE, OracleExtractor
    from orb_slam2_chinesenotes_b200.shard import extract_sharded
    from synth import synth_frame
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    frames = np.stack([synth_frame(320, 240, 500 + i) for i in range(5)])
    cap = 700

    def oracle_worker(block):
        O = OracleExtractor(500)
        kps, desc, n = np.zeros((len(block), cap), KP_DTYPE), np.zeros((len(block), cap, 32), np.uint8), np.zeros(len(block), np.int32)
        for i, f in enumerate(block):
            m, k, d = O.extract(f)
            n[i] = m; kps[i, :m] = k; desc[i, :m] = d
        O.close()
        return kps, desc, n

    start, stop, res, stats = extract_sharded(frames, oracle_worker, rank, world)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), start=start, stop=stop, kps=res[0], desc=res[1], n=res[2],
             frames=[s["frames"] for s in stats], keypoints=[s["keypoints"] for s in stats])
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_equals_single_process(tmp_path):
    import torch.multiprocessing as mp
    world, port = 2, 29000 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, HERE)
    from oracle_lib import OracleExtractor
    from synth import synth_frame
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    assert [int(p["start"]) for p in parts] == [0, 2] and [int(p["stop"]) for p in parts] == [2, 5]
    n = np.concatenate([p["n"] for p in parts])
    kps = np.concatenate([p["kps"] for p in parts])
    desc = np.concatenate([p["desc"] for p in parts])
    O = OracleExtractor(500)
    for i in range(5):
        m, k, d = O.extract(synth_frame(320, 240, 500 + i))
        assert m == n[i] and (d == desc[i, :m]).all() and all((k[f] == kps[i, :m][f]).all() for f in k.dtype.names)
    # every rank saw the same gathered statistics
    for p in parts:
        assert p["frames"].tolist() == [2, 3] and p["keypoints"].tolist() == [int(n[:2].sum()), int(n[2:].sum())]
    O.close()
