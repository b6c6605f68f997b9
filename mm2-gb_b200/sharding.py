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


# ---- host placement of a rank: the CPUs next to its GPU ----------------------------------------------------------------
# The end-to-end path of a rank is bound by host memory passes (gather into pinned staging, compact_a's gather).
# On a box with more than one NUMA node a rank whose worker threads or staging pages sit on the far node pays the socket
# interconnect for every one of those passes, so a rank is confined to the CPUs of its GPU's node BEFORE it allocates
# anything (first touch then places its buffers there too).  Where the box exposes a single node nothing is changed.

def plan_rank_cpus(gpu_cpus, local_rank: int, all_cpus):
    """CPUs this rank should run on and the number of ranks sharing them: (cpus, sharers), or (None, local_world) when there
    is nothing to gain -- no topology (a GPU's set is empty or covers every CPU), or a GPU whose node has no usable CPU.
    `gpu_cpus[r]` = CPUs NVML reports as local to the GPU of local rank r; `all_cpus` = the CPUs the process may use."""
    allc = frozenset(all_cpus)
    sets = [frozenset(c) & allc for c in gpu_cpus]
    world = len(sets)
    if not 0 <= local_rank < world:
        raise ValueError("local_rank outside the box")
    if any(len(s) == 0 for s in sets) or all(s == allc for s in sets):
        return None, world
    mine = sets[local_rank]
    sharers = sum(1 for s in sets if s == mine)
    return sorted(mine), sharers


def gpu_local_cpus(n_gpus: int):
    """NVML's ideal CPU set of each of the first `n_gpus` GPUs (honouring a numeric CUDA_VISIBLE_DEVICES); [] for a GPU NVML
    cannot describe."""
    import os
    out = []
    try:
        import pynvml
        pynvml.nvmlInit()
    except Exception:
        return [[] for _ in range(n_gpus)]
    vis = [v.strip() for v in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if v.strip()]
    n_words = (max(os.cpu_count() or 1, 1) + 63) // 64
    for r in range(n_gpus):
        try:
            idx = int(vis[r]) if r < len(vis) and vis[r].isdigit() else r
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            words = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
            out.append([64 * i + b for i, wd in enumerate(words) for b in range(64) if (int(wd) >> b) & 1])
        except Exception:
            out.append([])
    return out


def place_rank(local_rank: int, local_world: int):
    """Confine this process to the CPUs of its GPU's NUMA node (see above).  Returns a small report for the bench line."""
    import os
    try:
        allc = sorted(os.sched_getaffinity(0))
    except Exception:
        return {"pinned": False, "why": "no sched_getaffinity"}
    if local_world <= 1:
        return {"pinned": False, "why": "one rank", "cpus": len(allc)}
    cpus, sharers = plan_rank_cpus(gpu_local_cpus(local_world), local_rank, allc)
    if cpus is None:
        return {"pinned": False, "why": "no NUMA topology exposed for the GPUs", "cpus": len(allc), "ranks_sharing": local_world}
    try:
        os.sched_setaffinity(0, cpus)
    except Exception as e:                                   # keep running unpinned
        return {"pinned": False, "why": "sched_setaffinity: %s" % e, "cpus": len(allc), "ranks_sharing": local_world}
    return {"pinned": True, "cpus": len(cpus), "first_cpu": cpus[0], "last_cpu": cpus[-1], "ranks_sharing": sharers}
