"""The oracle's second model (QuadrotorModelVariant: RK4 integrator and/or Coriolis term) -- CPU.

The variant is NOT in the reference (SURVEY.md section 8(f)-4 asks for a second model behind the
ModelT concept of ilqr.hh:25-44), so there is nothing to pin it against: these tests check that it
is self-consistent, the way quadrotor_model_test.cc:145-346 checks the reference's model -- analytic
Jacobians against central finite differences on the manifold -- and that kind 0 is untouched.
"""
import numpy as np
import pytest

from conftest import random_spd_inertia


def rand_state(O, rng, scale=0.5):
    return np.concatenate([O.se3_exp(rng.normal(size=6) * scale), rng.normal(size=6) * scale])


def fd_discrete(O, cfg, x, u, eps=1e-6):
    A, B = np.zeros((12, 12)), np.zeros((12, 4))
    for j in range(12):
        d = np.zeros(12)
        d[j] = eps
        A[:, j] = O.state_minus(O.discrete_dynamics(cfg, O.state_add(x, d), u),
                                O.discrete_dynamics(cfg, O.state_add(x, -d), u)) / (2 * eps)
    for j in range(4):
        d = np.zeros(4)
        d[j] = eps
        B[:, j] = O.state_minus(O.discrete_dynamics(cfg, x, u + d), O.discrete_dynamics(cfg, x, u - d)) / (2 * eps)
    return A, B


@pytest.mark.parametrize("kind", [1, 2, 3])
def test_variant_jacobians_match_finite_differences(O, kind):
    rng = np.random.default_rng(kind)
    cfg = O.make_config(mass_kg=1.3, inertia=random_spd_inertia(3), arm_length_m=0.4,
                        torque_to_thrust_ratio_m=0.1, dt_s=0.1, model_kind=kind)
    for _ in range(5):
        x, u = rand_state(O, rng), rng.normal(size=4) + 3.0
        _, A, B = O.discrete_dynamics(cfg, x, u, diffs=True)
        Afd, Bfd = fd_discrete(O, cfg, x, u)
        np.testing.assert_allclose(A, Afd, atol=2e-8)
        np.testing.assert_allclose(B, Bfd, atol=2e-8)
        _, Jx, _ = O.continuous_dynamics(cfg, x, u, diffs=True)
        Jfd = np.zeros((12, 12))
        for j in range(12):
            d = np.zeros(12)
            d[j] = 1e-6
            Jfd[:, j] = (O.continuous_dynamics(cfg, O.state_add(x, d), u)
                         - O.continuous_dynamics(cfg, O.state_add(x, -d), u)) / 2e-6
        np.testing.assert_allclose(Jx, Jfd, atol=2e-8)


def test_coriolis_term_is_the_only_difference(O):
    rng = np.random.default_rng(0)
    base = O.make_config(inertia=random_spd_inertia(1), torque_to_thrust_ratio_m=0.1)
    cor = O.make_config(inertia=random_spd_inertia(1), torque_to_thrust_ratio_m=0.1, model_kind=2)
    x, u = rand_state(O, rng), rng.normal(size=4) + 2.0
    a, b = O.continuous_dynamics(base, x, u), O.continuous_dynamics(cor, x, u)
    v, w = x[7:10], x[10:13]
    np.testing.assert_allclose(b[6:9] - a[6:9], -np.cross(w, v), atol=1e-15)
    assert np.array_equal(np.delete(a, [6, 7, 8]), np.delete(b, [6, 7, 8]))


def test_rk4_is_far_more_accurate_than_euler(O):
    """Against a fine-step integration of the same ODE; on SE(3) the scheme of quadrotor_model.cc:51-63
    (no dexp^-1 correction of the stage increments) converges at second order."""
    rng = np.random.default_rng(5)
    x0, u0 = rand_state(O, rng), np.array([2.5, 2.4, 2.6, 2.45])
    inertia = random_spd_inertia(2)

    def integrate(kind, dt, T=1.0):
        cfg = O.make_config(inertia=inertia, torque_to_thrust_ratio_m=0.1, dt_s=dt, model_kind=kind)
        x = x0.copy()
        for _ in range(int(round(T / dt))):
            x = O.discrete_dynamics(cfg, x, u0)
        return x

    ref = integrate(1, 1e-3)
    err = {(k, dt): np.abs(O.state_minus(integrate(k, dt), ref)).max() for k in (0, 1) for dt in (0.1, 0.05)}
    assert err[(1, 0.1)] < err[(0, 0.1)] / 20
    assert 3.0 < err[(1, 0.1)] / err[(1, 0.05)] < 5.0   # second order
    assert 1.7 < err[(0, 0.1)] / err[(0, 0.05)] < 2.3   # first order


def test_variant_solve_converges_on_hover(O):
    from quadrotorilqr_b200 import problems

    m = problems.hover_model()
    desired = problems.hover_desired_trajectory(40, m["dt_s"], m["mass_kg"], m["g_mpss"])
    x0 = problems.hover_initial_states(4, seed=1)
    for kind in (1, 2, 3):
        cfg = O.make_config(mass_kg=m["mass_kg"], inertia=m["inertia"], arm_length_m=m["arm_length_m"],
                            torque_to_thrust_ratio_m=m["torque_to_thrust_ratio_m"], g_mpss=m["g_mpss"],
                            Q=m["Q"], R=m["R"], dt_s=m["dt_s"], model_kind=kind)
        seed = problems.constant_state_trajectory(x0, 40, m["dt_s"], desired[0, 14:18])
        init = np.stack([O.forward_sim(cfg, desired, seed[b], np.zeros((40, 4)), np.zeros((40, 4, 12)))
                         for b in range(4)])
        r = O.solve_batch(cfg, desired, init)
        assert np.all(np.isin(r["status"], [1, 2]))
        assert np.all(r["final_cost"] < [O.cost_trajectory(cfg, desired, init[b]) for b in range(4)])
