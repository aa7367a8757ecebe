"""world_size-2 gloo test of the multi-GPU host logic: each rank generates and solves its
own shard of the Philox problem stream (with the CPU oracle standing in for the GPU), statistics are
reduced over the process group, and per-problem results must equal the unsharded run bit for bit."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B_PER_RANK, N, SEED = 6, 12, 77


def _setup():
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import oracle as O
    from quadrotorilqr_b200 import problems

    m, opts = problems.hover_model(), problems.default_options(False)
    cfg = O.make_config(mass_kg=m["mass_kg"], inertia=m["inertia"], arm_length_m=m["arm_length_m"],
                        torque_to_thrust_ratio_m=m["torque_to_thrust_ratio_m"], g_mpss=m["g_mpss"], Q=m["Q"],
                        R=m["R"], dt_s=m["dt_s"])
    desired = problems.hover_desired_trajectory(N, m["dt_s"], m["mass_kg"], m["g_mpss"])
    return O, problems, m, cfg, desired


def _solve_range(first, count):
    O, problems, m, cfg, desired = _setup()
    x0 = problems.hover_initial_states(count, seed=SEED, first=first)
    init = np.stack([O.forward_sim(cfg, desired, problems.constant_state_trajectory(x, N, m["dt_s"], desired[0, 14:18])[0],
                                   np.zeros((N, 4)), np.zeros((N, 4, 12))) for x in x0])
    return O.solve_batch(cfg, desired, init, nthreads=1)


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from quadrotorilqr_b200 import sharding

    first, count = sharding.weak_shard(B_PER_RANK, rank)
    r = _solve_range(first, count)
    res = np.zeros(count, dtype=[("status", "<i4"), ("backward_passes", "<i4"), ("rollouts", "<i4")])
    res["status"], res["backward_passes"], res["rollouts"] = r["status"], r["backward_passes"], r["rollouts"]
    vec = torch.from_numpy(sharding.summarize_results(res))
    mx, sm = sharding.reduce_stats(dist, vec)
    np.save(os.path.join(out_dir, f"traj{rank}.npy"), r["traj"])
    allres = sharding.gather_results(dist, res)  # per-problem records of every rank, in rank order
    soa = torch.from_numpy(np.ascontiguousarray(r["traj"][:, :, 1:].transpose(1, 2, 0)))  # [N, 17, B]
    alltraj = sharding.gather_trajectories(dist, soa)
    if rank == 1:
        np.save(os.path.join(out_dir, "gathered_res.npy"), allres)
        np.save(os.path.join(out_dir, "gathered_traj.npy"), alltraj.numpy())
    if rank == 0:
        np.save(os.path.join(out_dir, "sum.npy"), sm.numpy())
        np.save(os.path.join(out_dir, "max.npy"), mx.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shards_match_unsharded(tmp_path):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    full = _solve_range(0, world * B_PER_RANK)
    got = np.concatenate([np.load(tmp_path / f"traj{r}.npy") for r in range(world)])
    assert np.array_equal(got, full["traj"])  # bit-identical per problem, whatever the shard
    sm = np.load(tmp_path / "sum.npy")
    conv = np.sum((full["status"] == 1) | (full["status"] == 2))
    assert sm[0] == conv and sm[1] == full["backward_passes"].sum() and sm[5] == world * B_PER_RANK
    assert sm[4] == full["rollouts"].sum()
    # the gather of per-problem results and (on request) trajectories, as seen by rank 1
    g = np.load(tmp_path / "gathered_res.npy")
    assert np.array_equal(g["status"], full["status"]) and np.array_equal(g["backward_passes"], full["backward_passes"])
    gt = np.load(tmp_path / "gathered_traj.npy")  # [N, 17, world * B]
    assert np.array_equal(gt.transpose(2, 0, 1), full["traj"][:, :, 1:])


def test_strong_shard_partition():
    from quadrotorilqr_b200 import sharding

    for total in (1, 7, 64, 65536, 65537):
        for world in (1, 2, 3, 8):
            spans = [sharding.strong_shard(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (f0, c0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
