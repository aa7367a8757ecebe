"""The model-agnostic CUDA path (dense J_x, J_u; include/qilqr.h qilqr_set_model_variant) -- SURVEY.md 8(f)-4.

(1) For the reference's QuadrotorModel the generic kernels must reproduce the quadrotor-specific ones
    (and the oracle): same iteration counts and flags, values within the FP64 parity tolerance 1e-9.
(2) For a second model (RK4 integrator and/or Coriolis term; oracle: QuadrotorModelVariant) the CUDA
    path must match the oracle the same way, piece by piece and over whole solves.
"""
import numpy as np
import pytest

from conftest import make_solver, oracle_config
from test_gpu_parity import assert_close, check_solve_against_oracle

pytestmark = pytest.mark.gpu


def hover_problem(s, O, cfg, B, N=40, seed=0, small=False):
    from quadrotorilqr_b200 import problems

    m = problems.hover_model()
    desired = problems.hover_desired_trajectory(N, m["dt_s"], m["mass_kg"], m["g_mpss"])
    # `small`: SURVEY.md App. C "small perturbation" family (6-8 iterations)
    x0 = (problems.hover_initial_states(B, seed=seed, pos=0.5, theta_max=0.2, vel=0.1) if small
          else problems.hover_initial_states(B, seed=seed))
    seedtraj = problems.constant_state_trajectory(x0, N, m["dt_s"], desired[0, 14:18])
    initial = s.forward_sim(seedtraj, np.zeros((B, N, 4)), np.zeros((B, N, 48)))
    return desired, initial, seedtraj


@pytest.mark.parametrize("flags", [1, 2, 3])
def test_variant_dynamics_match_oracle(O, flags):
    from quadrotorilqr_b200 import problems
    from test_gpu_parity import random_states

    model, opts = problems.hover_model(), problems.default_options(False)
    model = dict(model, inertia=np.array([[1.0, 0.1, 0.0], [0.1, 2.0, 0.2], [0.0, 0.2, 1.5]]))
    s = make_solver(model, opts, model_flags=flags)
    cfg = oracle_config(O, model, opts, model_kind=flags)
    x = random_states(O, 33, seed=flags)
    u = np.random.default_rng(flags).normal(size=(33, 4)) + 2.5
    xn, A, B = s.discrete_dynamics(x, u, diffs=True)
    xd, Jx, Ju = s.continuous_dynamics(x, u, diffs=True)
    for b in range(33):
        xo, Ao, Bo = O.discrete_dynamics(cfg, x[b], u[b], diffs=True)
        assert_close(xn[b], xo, what="x_next")
        assert_close(A[b].reshape(12, 12), Ao, what="J_x")
        assert_close(B[b].reshape(12, 4), Bo, what="J_u")
        xdo, Jxo, Juo = O.continuous_dynamics(cfg, x[b], u[b], diffs=True)
        assert_close(xd[b], xdo, what="xdot")
        assert_close(Jx[b].reshape(12, 12), Jxo, what="Jc_x")
        assert_close(Ju[b].reshape(12, 4), Juo, what="Jc_u")


@pytest.mark.parametrize("flags", [4, 1, 2, 3])
def test_generic_path_pieces_and_solves_match_oracle(O, flags):
    from quadrotorilqr_b200 import problems

    model, opts = problems.hover_model(), problems.default_options(False)
    s = make_solver(model, opts, model_flags=flags)
    cfg = oracle_config(O, model, opts, model_kind=flags & 3)
    B, N = 37, 40
    # with the Coriolis term the default-size perturbations leave ~10 % of the problems wandering to max_iters
    # through hundreds of backtracks, where rounding decides individual Armijo tests: use the small family
    desired, initial, seedtraj = hover_problem(s, O, cfg, B, N, small=bool(flags & 2))
    cost = s.cost_trajectory(initial, desired)
    k, K, QuTk, kTQuuk = s.backwards_pass(initial, desired)
    new = s.forward_sim(initial, k, K, 1.0)
    for b in range(B):
        assert_close(initial[b], O.forward_sim(cfg, desired, seedtraj[b], np.zeros((N, 4)), np.zeros((N, 4, 12))),
                     what="open loop")
        assert_close(cost[b], O.cost_trajectory(cfg, desired, initial[b]), what="cost")
        ko, Ko, a, c = O.backwards_pass(cfg, desired, initial[b])
        assert_close(k[b], ko, what="k")
        assert_close(K[b], Ko, what="K")
        assert_close(QuTk[b], a, what="QuTk")
        assert_close(kTQuuk[b], c, what="kTQuuk")
        assert_close(new[b], O.forward_sim(cfg, desired, initial[b], ko, Ko, 1.0), rtol=1e-8, what="fwd")
    r, o = check_solve_against_oracle(O, s, cfg, desired, initial)
    assert np.all(np.isin(r["results"]["status"], [1, 2]))


def test_generic_path_equals_quadrotor_kernels_on_default_problem(O):
    """BASELINE config 1 through both kernel families: identical decisions, values to 1e-9."""
    from quadrotorilqr_b200 import problems

    model, opts = problems.default_model(), problems.default_options(False)
    desired = problems.default_desired_trajectory()
    fast = make_solver(model, opts).solve(desired[None], desired, want_gains=True, hist_cap=100)
    gen = make_solver(model, opts, model_flags=4).solve(desired[None], desired, want_gains=True, hist_cap=100)
    for f in ("status", "backward_passes", "rollouts", "num_debug"):
        assert np.array_equal(fast["results"][f], gen["results"][f]), f
    assert_close(gen["traj"], fast["traj"], what="traj")
    assert_close(gen["K"], fast["K"], rtol=1e-8, what="K")
    assert_close(gen["cost_history"], fast["cost_history"], what="cost history")


def test_generic_path_symmetrised_long_horizon(O):
    """The dense Riccati kernel's symmetrisation branch, ragged tile (B not a multiple of 8), N = 300."""
    from quadrotorilqr_b200 import problems
    from quadrotorilqr_b200.options import ILQROptions

    model = dict(problems.hover_model(), dt_s=0.02)
    opts = problems.default_options(False)
    opts.symmetrize_vxx = True
    N, B = 300, 5
    s = make_solver(model, opts, model_flags=1)
    cfg = oracle_config(O, model, opts, model_kind=1)
    desired = problems.figure_eight_desired(N, model["dt_s"])
    rng = np.random.default_rng(0)
    x0 = np.tile(desired[0, 1:14], (B, 1))
    x0[:, :3] += rng.uniform(-0.3, 0.3, (B, 3))
    seedtraj = problems.constant_state_trajectory(x0, N, model["dt_s"], desired[0, 14:18])
    initial = s.forward_sim(seedtraj, np.zeros((B, N, 4)), np.zeros((B, N, 48)))
    k, K, QuTk, kTQuuk = s.backwards_pass(initial, desired)
    for b in range(B):
        ko, Ko, a, c = O.backwards_pass(cfg, desired, initial[b])
        assert_close(k[b], ko, rtol=1e-7, what="k")
        assert_close(K[b], Ko, rtol=1e-7, what="K")
        assert_close(QuTk[b], a, rtol=1e-8, what="QuTk")
