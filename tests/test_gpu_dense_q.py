"""General 12x12 cost weights (pose/velocity coupling, even non-symmetric -- the reference accepts any Q and
uses it as written, cost.hh:47-52) go through the same quad/TMA backward path as block-diagonal ones."""
import numpy as np
import pytest

from conftest import make_solver, oracle_config

pytestmark = pytest.mark.gpu


def batch(s, model, B, N, seed):
    from quadrotorilqr_b200 import problems

    d = problems.hover_desired_trajectory(N, model["dt_s"], model["mass_kg"], model["g_mpss"])
    x0 = problems.hover_initial_states(B, seed=seed)
    init = s.forward_sim(problems.constant_state_trajectory(x0, N, model["dt_s"], d[0, 14:18]), np.zeros((B, N, 4)),
                         np.zeros((B, N, 48)))
    return d, init


@pytest.mark.parametrize("symmetric", [True, False])
def test_dense_q_backward_pass_and_solve_match_oracle(O, symmetric, monkeypatch):
    from quadrotorilqr_b200 import problems

    rng = np.random.default_rng(17)
    A = rng.uniform(-1, 1, (12, 12))
    Q = np.diag([100.0] * 6 + [1.0] * 6) + 2.0 * (A @ A.T)
    if not symmetric:
        Q = Q + rng.uniform(-0.3, 0.3, (12, 12))
    model = dict(problems.hover_model(), Q=Q, R=np.diag([1.0, 1.5, 2.0, 0.5]))
    opts = problems.default_options(False)
    s = make_solver(model, opts)
    cfg = oracle_config(O, model, opts)
    B, N = 23, 30
    d, init = batch(s, model, B, N, seed=5)
    k, K, a, c = s.backwards_pass(init, d)
    for b in range(B):
        ko, Ko, ao, co = O.backwards_pass(cfg, d, init[b])
        for got, exp in ((k[b], ko), (K[b], Ko), (a[b], ao), (c[b], co)):
            got, exp = np.asarray(got, dtype=float), np.asarray(exp, dtype=float)
            assert np.max(np.abs(got - exp)) <= 1e-9 * max(1.0, np.max(np.abs(exp)))
    # the thread-per-problem kernel (no structure assumptions on Q) agrees too
    monkeypatch.setenv("QILQR_BACKWARD", "t1")
    s1 = make_solver(model, opts)
    monkeypatch.delenv("QILQR_BACKWARD")
    k1, K1, a1, c1 = s1.backwards_pass(init, d)
    assert np.max(np.abs(K1 - K)) <= 1e-11 * max(1.0, np.max(np.abs(K)))
    if symmetric:  # full solves (a non-symmetric Q makes the reference's own gradient inconsistent; not iterated)
        r = s.solve(init, d, hist_cap=100)
        o = O.solve_batch(cfg, d, init, hist_cap=100)
        assert np.array_equal(r["results"]["status"], o["status"])
        assert np.array_equal(r["results"]["backward_passes"], o["backward_passes"])
        assert np.max(np.abs(r["traj"] - o["traj"])) <= 1e-9 * max(1.0, np.max(np.abs(o["traj"])))


def test_diagonal_q_cost_shortcut_is_bit_neutral(monkeypatch):
    """A diagonal Q lets the running cost skip the off-diagonal products of dx^T Q (DeviceParams::q_diagonal): every
    skipped term is (+-0) + (+-0), so costs, decisions and trajectories are the same doubles as with the dense chain
    (QILQR_Q_DIAGONAL=0 forces the dense chain).  The role-specialised rollout and the cost API go through it too."""
    from quadrotorilqr_b200 import problems

    model, opts = problems.hover_model(), problems.default_options(False)
    fast = make_solver(model, opts)
    monkeypatch.setenv("QILQR_Q_DIAGONAL", "0")
    dense = make_solver(model, opts)
    monkeypatch.delenv("QILQR_Q_DIAGONAL")
    B, N = 5000, 40      # above the role-specialised rollout's launch size at first, below it later
    d, init = batch(fast, model, B, N, seed=11)
    assert np.array_equal(fast.cost_trajectory(init, d), dense.cost_trajectory(init, d))
    a, b = fast.solve(init, d, hist_cap=100), dense.solve(init, d, hist_cap=100)
    assert np.array_equal(a["results"], b["results"])
    assert np.array_equal(a["traj"], b["traj"]) and np.array_equal(a["cost_history"], b["cost_history"])
