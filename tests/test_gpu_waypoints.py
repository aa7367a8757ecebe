"""Waypoint variant of BASELINE configs 2/3: every problem has its OWN desired trajectory
(desired_count = batch), solved by the default quad/split backward path (block-diagonal Q)."""
import numpy as np
import pytest

from conftest import make_solver, oracle_config

pytestmark = pytest.mark.gpu


def test_per_problem_waypoint_desired_matches_oracle(O):
    from quadrotorilqr_b200 import problems

    model, opts = problems.hover_model(), problems.default_options(False)
    s = make_solver(model, opts)
    cfg = oracle_config(O, model, opts)
    B, N = 19, 40
    rng = np.random.default_rng(21)
    base = problems.hover_desired_trajectory(N)
    desired = np.repeat(base[None], B, axis=0)
    way = rng.uniform(-1, 1, (B, 4, 3))
    for seg in range(4):
        desired[:, seg * N // 4:(seg + 1) * N // 4, 1:4] = way[:, seg][:, None, :]
    x0 = problems.hover_initial_states(B, seed=9)
    seedtraj = problems.constant_state_trajectory(x0, N, model["dt_s"], base[0, 14:18])
    initial = s.forward_sim(seedtraj, np.zeros((B, N, 4)), np.zeros((B, N, 48)))
    r = s.solve(initial, desired, want_gains=True, hist_cap=100)
    o = O.solve_batch(cfg, desired, initial, want_gains=True, hist_cap=100)
    assert np.array_equal(r["results"]["status"], o["status"])
    assert np.array_equal(r["results"]["backward_passes"], o["backward_passes"])
    assert np.array_equal(r["results"]["rollouts"], o["rollouts"])
    for b in range(B):
        for key, rtol in (("traj", 1e-9), ("K", 1e-9), ("cost_history", 1e-9)):
            a, e = np.asarray(r[key][b], dtype=float).ravel(), np.asarray(o[key][b], dtype=float).ravel()
            assert np.max(np.abs(a - e)) <= rtol * max(1.0, np.max(np.abs(e))), (key, b)
    # the shared-desired and per-problem-desired code paths agree bit for bit when the desired is the same
    same = np.repeat(base[None], B, axis=0)
    r1 = s.solve(initial, base, want_gains=True)
    r2 = s.solve(initial, same, want_gains=True)
    assert np.array_equal(r1["traj"], r2["traj"]) and np.array_equal(r1["K"], r2["K"])
    assert np.array_equal(r1["results"], r2["results"])
