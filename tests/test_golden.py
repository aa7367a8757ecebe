"""Committed golden vectors (tests/golden/*.npz, written by tests/golden/make_golden.py from the oracle).
CPU: the oracle still reproduces them bit for bit.  GPU: the CUDA path matches them to RTOL."""
import os

import numpy as np
import pytest

from conftest import make_solver, oracle_config

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL = 1e-9


def close(a, b, rtol=RTOL):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.max(np.abs(a - b)) <= rtol * max(1.0, np.max(np.abs(b)))


def test_oracle_reproduces_golden_default_problem(O):
    from quadrotorilqr_b200 import problems

    g = np.load(os.path.join(GOLD, "c1_default_problem.npz"))
    cfg = oracle_config(O, problems.default_model(), problems.default_options(True))
    r = O.solve(cfg, g["desired"], g["desired"])
    assert np.array_equal(g["desired"], problems.default_desired_trajectory())
    assert r["backward_passes"] == int(g["backward_passes"]) == 77 and r["status"] == int(g["status"]) == 1
    assert np.array_equal(r["traj"], g["traj"]) and np.array_equal(r["cost_history"], g["cost_history"])
    assert np.array_equal(r["step_history"], g["step_history"]) and g["step_history"][4] == 0.5
    assert float(g["final_cost"]) == 22556.502591980552


def test_oracle_reproduces_golden_hover(O):
    from quadrotorilqr_b200 import problems

    g = np.load(os.path.join(GOLD, "c2_hover_prefix16.npz"))
    assert np.array_equal(g["x0"], problems.hover_initial_states(16, seed=2026))
    cfg = oracle_config(O, problems.hover_model(), problems.default_options(False))
    b = O.solve_batch(cfg, g["desired"], g["initial"], want_gains=True, hist_cap=100)
    for key in ("traj", "k", "K", "cost_history", "status", "backward_passes", "rollouts", "final_cost"):
        assert np.array_equal(b[key], g[key]), key


@pytest.mark.gpu
def test_gpu_matches_golden_default_problem():
    from quadrotorilqr_b200 import problems

    g = np.load(os.path.join(GOLD, "c1_default_problem.npz"))
    s = make_solver(problems.default_model(), problems.default_options(True))
    r = s.solve(g["desired"][None], g["desired"], want_gains=True, hist_cap=100, want_debug=True)
    res = r["results"][0]
    assert res["status"] == int(g["status"]) and res["backward_passes"] == int(g["backward_passes"])
    assert res["rollouts"] == int(g["rollouts"])
    assert close(r["traj"][0], g["traj"]) and close(r["k"][0], g["k"]) and close(r["K"][0], g["K"])
    assert close(r["cost_history"][0][:76], g["cost_history"])
    assert close(r["debug"][0][0], g["debug_first"]) and close(r["debug"][0][75], g["debug_last"])


@pytest.mark.gpu
def test_gpu_matches_golden_hover_and_single_iteration():
    from quadrotorilqr_b200 import problems

    s = make_solver(problems.hover_model(), problems.default_options(False))
    g = np.load(os.path.join(GOLD, "c2_hover_prefix16.npz"))
    r = s.solve(g["initial"], g["desired"], want_gains=True, hist_cap=100)
    assert np.array_equal(r["results"]["status"], g["status"])
    assert np.array_equal(r["results"]["backward_passes"], g["backward_passes"])
    assert np.array_equal(r["results"]["rollouts"], g["rollouts"])
    assert close(r["traj"], g["traj"]) and close(r["K"], g["K"]) and close(r["cost_history"], g["cost_history"])
    h = np.load(os.path.join(GOLD, "c2_single_iteration.npz"))
    k, K, a, c = s.backwards_pass(h["traj"], h["desired"])
    assert close(k, h["k"]) and close(K, h["K"]) and close(a, h["QuTk"]) and close(c, h["kTQuuk"])
    assert close(s.forward_sim(h["traj"], h["k"], h["K"], 1.0), h["rolled"], 1e-8)
    assert close(s.cost_trajectory(h["traj"], h["desired"]), h["cost0"])
    assert close(s.cost_trajectory(h["rolled"], h["desired"]), h["cost1"])
