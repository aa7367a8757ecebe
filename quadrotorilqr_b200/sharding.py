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
    return np.array([np.sum((st == 1) | (st == 2)), results["backward_passes"].sum(), np.sum(st >= 4),
                     np.sum(st == 3), results["rollouts"].sum(), st.size], dtype=np.float64)


def reduce_stats(dist, vec):
    """(max over ranks, sum over ranks) of a 1-D float64 torch tensor; identity without a process group."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return vec.clone(), vec.clone()
    mx, sm = vec.clone(), vec.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    return mx, sm


def gather_results(dist, results):
    """All ranks' per-problem result records, in rank order (the one collective of the path, after the solve).

    ``results``: a structured numpy array (``RESULT_DTYPE``) or a torch tensor of any dtype whose first
    dimension is this rank's problem count; shards must have equal sizes (weak scaling).  NCCL when the
    tensor lives on a GPU, gloo on the CPU.  Without a process group the input comes back unchanged."""
    import torch

    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return results
    if isinstance(results, np.ndarray):
        raw = torch.from_numpy(np.ascontiguousarray(results).view(np.uint8).reshape(results.shape[0], -1))
        out = torch.empty((dist.get_world_size() * raw.shape[0], raw.shape[1]), dtype=torch.uint8)
        dist.all_gather_into_tensor(out, raw)
        return out.numpy().reshape(-1).view(results.dtype)
    out = torch.empty((dist.get_world_size() * results.shape[0],) + tuple(results.shape[1:]), dtype=results.dtype,
                      device=results.device)
    dist.all_gather_into_tensor(out, results.contiguous())
    return out


def gather_trajectories(dist, traj_soa):
    """All ranks' device-resident trajectories ``[N, 17, B]`` -> ``[N, 17, world * B]`` in rank order (only on
    request: 377 MB at B = 65536, N = 40).  The batch is the last axis of the SoA layout, so each rank's block is
    gathered along a leading axis and folded back."""
    import torch

    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return traj_soa
    w = dist.get_world_size()
    out = torch.empty((w,) + tuple(traj_soa.shape), dtype=traj_soa.dtype, device=traj_soa.device)
    dist.all_gather_into_tensor(out, traj_soa.contiguous().unsqueeze(0))
    return out.permute(1, 2, 0, 3).reshape(traj_soa.shape[0], traj_soa.shape[1], w * traj_soa.shape[2])
