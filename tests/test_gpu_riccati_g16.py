"""k_riccati_g16 (16 lanes per problem, used for latency-bound launches) against k_riccati_g4 (4 lanes per problem):
bit-identical gains, cost-reduction terms and whole solves -- which is what allows picking the kernel by launch size
without a problem's result depending on the batch it is solved in -- and parity with the oracle."""
import dataclasses

import numpy as np
import pytest

from conftest import make_solver, oracle_config

pytestmark = pytest.mark.gpu


def pair(monkeypatch, model, opts):
    monkeypatch.setenv("QILQR_G16_THRESHOLD", "1000000")
    g16 = make_solver(model, opts)
    monkeypatch.setenv("QILQR_G16_THRESHOLD", "0")
    g4 = make_solver(model, opts)
    monkeypatch.delenv("QILQR_G16_THRESHOLD")
    return g16, g4


def hover(s, B, N, seed, dt=0.1):
    from quadrotorilqr_b200 import problems

    m = problems.hover_model()
    desired = problems.hover_desired_trajectory(N, dt, m["mass_kg"], m["g_mpss"])
    x0 = problems.hover_initial_states(B, seed=seed)
    initial = s.forward_sim(problems.constant_state_trajectory(x0, N, dt, desired[0, 14:18]),
                            np.zeros((B, N, 4)), np.zeros((B, N, 48)))
    return desired, initial


@pytest.mark.parametrize("B,N", [(1, 40), (7, 40), (8, 3), (77, 40), (1030, 25)])
def test_backward_pass_is_bit_identical(monkeypatch, B, N):
    from quadrotorilqr_b200 import problems

    rng = np.random.default_rng(3)
    A = rng.uniform(-1, 1, (3, 3))
    base = problems.hover_model()
    rnd = dict(base, inertia=A @ A.T + 3 * np.eye(3), mass_kg=1.3, arm_length_m=0.4,
               Q=np.diag(rng.uniform(0.5, 50, 12)), R=np.diag([1.0, 1.5, 2.0, 0.5]) + 0.1)
    for model in (base, rnd):
        for sym in (False, True):
            opts = dataclasses.replace(problems.default_options(False), symmetrize_vxx=sym, quu_regularization=0.25 * sym)
            g16, g4 = pair(monkeypatch, model, opts)
            desired, traj = hover(g16, B, N, seed=31 + B)
            per_problem = np.repeat(desired[None], B, axis=0)
            per_problem[:, :, 1:4] += rng.uniform(-0.3, 0.3, (B, 1, 3))
            for des in (desired, per_problem):
                a, b = g16.backwards_pass(traj, des), g4.backwards_pass(traj, des)
                for x, y, name in zip(a, b, ("k", "K", "QuTk", "kTQuuk")):
                    assert np.array_equal(x, y), (name, B, N, sym)


def test_solves_are_bit_identical_and_match_the_oracle(O, monkeypatch):
    from quadrotorilqr_b200 import problems

    model, opts = problems.hover_model(), problems.default_options(False)
    g16, g4 = pair(monkeypatch, model, opts)
    desired, initial = hover(g16, 333, 40, seed=77)
    a = g16.solve(initial, desired, want_gains=True, hist_cap=100)
    b = g4.solve(initial, desired, want_gains=True, hist_cap=100)
    assert np.array_equal(a["results"], b["results"])
    for key in ("traj", "k", "K", "cost_history"):
        assert np.array_equal(a[key], b[key]), key
    cfg = oracle_config(O, model, opts)
    o = O.solve_batch(cfg, desired, initial, hist_cap=100)
    assert np.array_equal(a["results"]["status"], o["status"])
    assert np.array_equal(a["results"]["backward_passes"], o["backward_passes"])
    assert np.max(np.abs(a["traj"] - o["traj"])) <= 1e-9 * max(1.0, np.max(np.abs(o["traj"])))


def test_long_horizon_with_parallel_step_sizes(monkeypatch):
    """BASELINE config 4 in small: N = 1000, symmetrised V_xx, 8 parallel step sizes."""
    from quadrotorilqr_b200 import problems

    N, dt = 1000, 0.02
    model = dict(problems.hover_model(), dt_s=dt)
    opts = dataclasses.replace(problems.default_options(False), symmetrize_vxx=True, num_parallel_alphas=8)
    g16, g4 = pair(monkeypatch, model, opts)
    d = problems.figure_eight_desired(N, dt)
    x0 = problems.figure_eight_initial_states(5, d)
    init = g16.forward_sim(problems.constant_state_trajectory(x0, N, dt, d[0, 14:18]), np.zeros((5, N, 4)),
                           np.zeros((5, N, 48)))
    a, b = g16.solve(init, d), g4.solve(init, d)
    assert np.array_equal(a["results"], b["results"]) and np.all(np.isin(a["results"]["status"], [1, 2]))
    assert np.array_equal(a["traj"], b["traj"])
