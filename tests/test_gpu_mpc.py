"""Receding-horizon MPC (BASELINE config 5) on the device vs the same loop built from oracle calls:
solve from the warm start, apply u_0 to the plant, shift the solution one knot (last knot duplicated),
put the plant state in knot 0."""
import numpy as np
import pytest

from conftest import make_solver, oracle_config

pytestmark = pytest.mark.gpu


def oracle_mpc(O, cfg, desired, traj0, plant0, steps, disturbance=None):
    traj, plant = traj0.copy(), plant0.copy()
    states, controls, iters = [], [], []
    for _ in range(steps):
        r = O.solve(cfg, desired, traj)
        sol = r["traj"]
        u0 = sol[0, 14:18].copy()
        plant = O.discrete_dynamics(cfg, plant, u0)
        if disturbance is not None:
            plant[7:13] += disturbance
        nxt = sol.copy()
        nxt[:-1, 1:] = sol[1:, 1:]  # time_s column stays
        nxt[0, 1:14] = plant
        traj = nxt
        states.append(plant.copy())
        controls.append(u0)
        iters.append(r["backward_passes"])
    return np.array(states), np.array(controls), np.array(iters), traj


def test_mpc_loop_matches_oracle(O):
    import torch
    from quadrotorilqr_b200 import problems

    model, opts = problems.hover_model(), problems.default_options(False)
    s = make_solver(model, opts)
    cfg = oracle_config(O, model, opts)
    B, N, T = 5, 20, 6
    desired = problems.hover_desired_trajectory(N, model["dt_s"], model["mass_kg"], model["g_mpss"])
    x0 = problems.hover_initial_states(B, seed=31)
    seedtraj = problems.constant_state_trajectory(x0, N, model["dt_s"], desired[0, 14:18])
    initial = s.forward_sim(seedtraj, np.zeros((B, N, 4)), np.zeros((B, N, 48)))
    dist = np.random.default_rng(1).uniform(-0.02, 0.02, (B, 6))

    dev = torch.device("cuda:0")
    aos = torch.from_numpy(initial).to(dev)
    soa = torch.empty((N, 17, B), dtype=torch.float64, device=dev)
    des = torch.empty((N, 17, 1), dtype=torch.float64, device=dev)
    des_aos = torch.from_numpy(desired[None].copy()).to(dev)
    plant = torch.from_numpy(np.ascontiguousarray(x0.T)).to(dev)
    dist_d = torch.from_numpy(np.ascontiguousarray(dist.T)).to(dev)
    slog = torch.zeros((T, 13, B), dtype=torch.float64, device=dev)
    ulog = torch.zeros((T, 4, B), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    s.pack_trajectory_device(aos, soa)
    s.pack_trajectory_device(des_aos, des)
    tot = s.mpc_run_device(T, soa, des, plant, disturbance=dist_d, state_log=slog, control_log=ulog)
    torch.cuda.synchronize()
    slog, ulog = slog.cpu().numpy(), ulog.cpu().numpy()
    total_iters = 0
    for b in range(B):
        st, ct, it, _ = oracle_mpc(O, cfg, desired, initial[b], x0[b], T, dist[b])
        total_iters += int(it.sum())
        assert np.max(np.abs(slog[:, :, b] - st)) <= 1e-9 * max(1.0, np.max(np.abs(st)))
        assert np.max(np.abs(ulog[:, :, b] - ct)) <= 1e-9 * max(1.0, np.max(np.abs(ct)))
    assert tot["backward_passes"] == total_iters and tot["resolves"] == B * T
    # warm starts pay off: later re-solves need fewer iterations than the first
    assert tot["backward_passes"] < B * T * 17
