"""Data-parallel sharding of a batch of independent problems across GPUs.

The path has no data-path collective: rank r owns a contiguous block of problems and solves it
alone.  The only exchange is the reduction of convergence statistics and of the timing (max over
ranks) after the solve, over NCCL on the GPU box (gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np


def weak_shard(batch_per_gpu: int, rank: int):
    """Problems [first, first + count) of the global stream owned by `rank` (fixed work per GPU)."""
    return rank * batch_per_gpu, batch_per_gpu


def strong_shard(total: int, rank: int, world: int):
    """Contiguous split of `total` problems over `world` ranks (sizes differ by at most one)."""
    base, rem = divmod(total, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def summarize_results(results) -> np.ndarray:
    """[converged, sum backward passes, line-search failures, max-iters exits, sum rollouts, n]."""
    st = results["status"]
    return np.array([np.sum((st == 1) | (st == 2)), results["backward_passes"].sum(), np.sum(st == 4),
                     np.sum(st == 3), results["rollouts"].sum(), st.size], dtype=np.float64)


def reduce_stats(dist, vec):
    """(max over ranks, sum over ranks) of a 1-D float64 torch tensor; identity without a process group."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return vec.clone(), vec.clone()
    mx, sm = vec.clone(), vec.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    return mx, sm
