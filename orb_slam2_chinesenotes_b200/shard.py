"""Frame-level sharding across GPUs.

Frames (and stereo pairs as a unit) are independent, so a batch is cut into contiguous blocks
[r*F/G, (r+1)*F/G), one per rank / GPU; there is NO collective on the per-frame path.  The only
communication is one all-gather of per-rank {frames, keypoints, elapsed} at the end (NCCL on GPUs,
gloo in the CPU tests).  One process per GPU, launched with torchrun.
"""
import time

import numpy as np


def frame_block(n_frames, rank, world):
    """Contiguous block of frame indices owned by `rank`: [start, stop)."""
    return (rank * n_frames) // world, ((rank + 1) * n_frames) // world


def gather_stats(stats, device=None):
    """All-gather one small dict of numbers per rank; returns the list ordered by rank (every rank gets it)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [dict(stats)]
    keys = sorted(stats)
    t = torch.tensor([float(stats[k]) for k in keys], dtype=torch.float64, device=device or "cpu")
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [dict(zip(keys, o.tolist())) for o in out]


def extract_sharded(frames, worker, rank=0, world=1, device=None):
    """Run `worker(block_of_frames) -> (kps [b,cap], desc [b,cap,32], n [b])` on this rank's block.

    Returns (start, stop, local results, per-rank stats).  Concatenating the per-rank results in rank
    order restores the frame order of the batch."""
    start, stop = frame_block(len(frames), rank, world)
    t0 = time.perf_counter()
    res = worker(frames[start:stop]) if stop > start else None
    dt = time.perf_counter() - t0
    nk = int(np.asarray(res[2]).sum()) if res is not None else 0
    stats = gather_stats({"rank": rank, "frames": stop - start, "keypoints": nk, "elapsed_s": dt}, device)
    return start, stop, res, stats
