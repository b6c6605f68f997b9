"""Read sharding across the GPUs of one box: reads are independent, so a batch is cut into contiguous,
anchor-balanced shards of whole reads, one per rank, and nothing is exchanged on the data path.  torch.distributed is used
for the measurement only (sum of the work, max of the time over ranks).

The reference has no multi-GPU path at all (one stream, one worker thread: README.md:46-47, gpu/plchain.cu:299); in the
drop-in the same partition happens per driver thread (thread_id % n_gpus, csrc/plchain_dropin.cpp)."""
from __future__ import annotations

import numpy as np


def shard_bounds(off: np.ndarray, world: int) -> np.ndarray:
    """Read boundaries r[0..world] of `world` contiguous shards balanced by ANCHORS (not reads): shard k = reads
    [r[k], r[k+1]).  Every read lands in exactly one shard; shards may be empty when there are fewer reads than ranks."""
    off = np.asarray(off, np.int64)
    n_reads = len(off) - 1
    if world < 1:
        raise ValueError("world must be positive")
    total = int(off[-1]) - int(off[0])
    targets = int(off[0]) + (total * np.arange(1, world, dtype=np.int64)) // world
    # cut at the read boundary closest to each target
    hi = np.searchsorted(off, targets, side="left").clip(0, n_reads)
    lo = (hi - 1).clip(0, n_reads)
    cut = np.where(np.abs(off[hi] - targets) <= np.abs(off[lo] - targets), hi, lo)
    cut = np.maximum.accumulate(cut)
    return np.concatenate([[0], cut, [n_reads]]).astype(np.int64)


def shard(a: np.ndarray, off: np.ndarray, world: int, rank: int):
    """This rank's reads: (anchors, offsets rebased to 0, first read, one past the last read)."""
    b = shard_bounds(off, world)
    r0, r1 = int(b[rank]), int(b[rank + 1])
    o = np.asarray(off[r0:r1 + 1], np.int64)
    return a[int(o[0]):int(o[-1])], o - o[0], r0, r1


def reduce_job(dist, sums, maxes, device="cpu"):
    """Whole-job figures: element-wise SUM of `sums` and MAX of `maxes` over the ranks (identity without a process group).
    Returns two lists of floats."""
    import torch
    s = torch.tensor([float(x) for x in sums], dtype=torch.float64, device=device)
    m = torch.tensor([float(x) for x in maxes], dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
        dist.all_reduce(m, op=dist.ReduceOp.MAX)
    return [float(x) for x in s.tolist()], [float(x) for x in m.tolist()]
