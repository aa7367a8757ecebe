"""The reference's own gtest assertions, ported 1:1 and run against the CPU oracle.

This is what pins the oracle: ilqr_test.cc:102-190,
quadrotor_model_test.cc:94-447, cost_test.cc:27-151.  Everything here runs on the CPU.
"""
import numpy as np
import pytest

from conftest import identity_traj, random_spd_inertia

DT = 0.1


def ilqr_fixture(O):
    """ILQRFixture (ilqr_test.cc:68-100): N=3, dt=.1, m=1, I=eye, arm=1, ratio=1, g=0, Q=I, R=I."""
    cfg = O.make_config(mass_kg=1.0, inertia=np.eye(3), arm_length_m=1.0, torque_to_thrust_ratio_m=1.0,
                        g_mpss=0.0, Q=np.eye(12), R=np.eye(4), dt_s=DT, step_update=0.5,
                        desired_reduction_frac=0.5, ls_max_iters=10, rtol=1e-12, atol=1e-12, max_iters=100)
    N = 3
    cur = identity_traj(N, DT)
    k = np.ones((N, 4))
    K = np.zeros((N, 4, 12))
    return cfg, cur, k, K


def approx_state_eq(O, lhs, rhs, tol):  # ilqr_test.cc:38-48
    rel = O.se3_compose(O.se3_inverse(lhs[:7]), rhs[:7])
    return np.linalg.norm(O.se3_log(rel)) < tol and np.allclose(lhs[7:], rhs[7:], rtol=1e-12, atol=tol)


def check_approx_traj_eq(O, a, b, tol):  # ilqr_test.cc:55-63
    assert a.shape == b.shape
    for i in range(a.shape[0]):
        assert approx_state_eq(O, a[i, 1:14], b[i, 1:14], tol)
        assert np.allclose(a[i, 14:], b[i, 14:], rtol=tol, atol=tol)


def test_forward_sim_generates_correct_trajectory(O):  # ilqr_test.cc:102-126
    cfg, cur, k, K = ilqr_fixture(O)
    accel = 4.0
    exp = identity_traj(3, DT)
    exp[:, 14:] = 1.0
    exp[1, 10] = DT * accel
    exp[2, 1:8] = O.se3_compose(exp[2, 1:8], O.se3_exp([0, 0, DT * DT * accel, 0, 0, 0]))
    exp[2, 10] = 2.0 * DT * accel
    new = O.forward_sim(cfg, cur, cur, k, K)
    check_approx_traj_eq(O, new, exp, 1e-6)
    assert np.array_equal(new[:, 0], cur[:, 0])


def test_cost_trajectory_calculates_correct_cost(O):  # ilqr_test.cc:128-141
    cfg, cur, k, K = ilqr_fixture(O)
    new = O.forward_sim(cfg, cur, cur, k, K)
    cost = O.cost_trajectory(cfg, cur, new)
    accel = 4.0
    expected = (DT * accel) ** 2.0 + (DT * DT * accel) ** 2.0 + (2.0 * DT * accel) ** 2.0 + 3 * 4
    assert abs(cost - expected) <= 4 * np.spacing(expected)  # EXPECT_DOUBLE_EQ = 4 ULP


def test_backward_pass_returns_zero_update_if_zero_gradient(O):  # ilqr_test.cc:143-153
    cfg, cur, _, _ = ilqr_fixture(O)
    k, K, QuTk, kTQuuk = O.backwards_pass(cfg, cur, cur)
    assert k.shape[0] == 3
    assert QuTk == 0.0 and kTQuuk == 0.0
    assert np.all(k == 0.0)


def test_backward_pass_expected_reduction_negative(O):  # ilqr_test.cc:155-164
    cfg, cur, k, K = ilqr_fixture(O)
    new = O.forward_sim(cfg, cur, cur, k, K)
    _, _, QuTk, _ = O.backwards_pass(cfg, cur, new)
    assert QuTk < 0.0


def test_line_search_finds_step_size_that_reduces_cost(O):  # ilqr_test.cc:166-177
    cfg, cur, k, K = ilqr_fixture(O)
    traj = O.forward_sim(cfg, cur, cur, k, K)
    cost = O.cost_trajectory(cfg, cur, traj)
    k2, K2, QuTk, kTQuuk = O.backwards_pass(cfg, cur, traj)
    new_traj, new_cost, step = O.line_search(cfg, cur, traj, cost, k2, K2, QuTk, kTQuuk)
    assert new_cost - cost < 0.5 * (step * QuTk + step * step * kTQuuk / 2.0)


def test_solve_finds_optimal_trajectory(O):  # ilqr_test.cc:179-190
    cfg, cur, k, K = ilqr_fixture(O)
    k = k.copy()
    k[:, 0] *= 100
    k[:, 2] *= 100
    initial = O.forward_sim(cfg, cur, cur, k, K)
    r = O.solve(cfg, cur, initial)
    check_approx_traj_eq(O, cur, r["traj"], 1e-6)


def test_line_search_exhaustion_raises(O):  # ilqr.hh:191-193
    cfg, cur, k, K = ilqr_fixture(O)
    traj = O.forward_sim(cfg, cur, cur, k, K)
    cost = O.cost_trajectory(cfg, cur, traj)
    k2, K2, QuTk, kTQuuk = O.backwards_pass(cfg, cur, traj)
    with pytest.raises(RuntimeError):
        O.line_search(cfg, cur, traj, -1e30, k2, K2, QuTk, kTQuuk)  # no step can beat -1e30


def test_cost_trajectory_longer_than_desired_raises(O):  # cost.hh:39-40
    cfg, cur, _, _ = ilqr_fixture(O)
    with pytest.raises(IndexError):
        O.cost_trajectory(cfg, cur[:2], cur)


# ---- quadrotor_model_test.cc -------------------------------------------------------------------
def model_cfg(O, inertia=None, ratio=1.0, g=9.81):
    return O.make_config(mass_kg=1.0, inertia=np.eye(3) if inertia is None else inertia, arm_length_m=1.0,
                         torque_to_thrust_ratio_m=ratio, g_mpss=g)


def x_identity():
    x = np.zeros(13)
    x[6] = 1.0
    return x


def test_discrete_dynamics_translational(O):  # quadrotor_model_test.cc:94-116
    cfg = model_cfg(O)
    x = x_identity()
    x[7:10] = [1.0, 2.0, 3.0]
    xn = O.discrete_dynamics(cfg, x, np.ones(4), DT)
    assert np.allclose(xn[0:3], [0.1, 0.2, 0.3], rtol=1e-6)
    assert np.allclose(xn[7:13], [1.0, 2.0, 3.0 + (4.0 - 9.81) * DT, 0, 0, 0], rtol=1e-6)


def test_discrete_dynamics_rotational(O):  # quadrotor_model_test.cc:118-143
    cfg = model_cfg(O)
    x = x_identity()
    x[10:13] = [1.2, 0.0, 0.0]
    xn = O.discrete_dynamics(cfg, x, [0.0, -1.0, 0.0, 1.0], DT)
    expected_pose = O.se3_compose(x[:7], O.se3_exp([0, 0, 0, 1.2 * DT, 0, 0]))
    rel = O.se3_log(O.se3_compose(O.se3_inverse(expected_pose), xn[:7]))
    assert np.linalg.norm(rel[3:]) < 1e-6
    assert np.allclose(xn[10:13], [1.2 + 2.0 * DT, 0, 0], rtol=1e-6)


X_INIT = None


def x_init(O):
    x = np.zeros(13)
    x[:7] = O.se3_exp([1.0, 2.0, 3.0, 4.0, 5.0, 6.0])
    x[7:] = [2.0, 3.0, 4.0, 5.0, 6.0, 7.0]
    return x


def check_state_jacobian(O, fun, analytic, minus=None):  # quadrotor_model_test.cc:30-55
    EPS = 1e-6
    for i in range(12):
        d = np.zeros(12)
        d[i] = EPS
        yp, ym = fun(d), fun(-d)
        fd = (minus(yp, ym) if minus else (yp - ym)) / (2 * EPS)
        col = analytic[:, i]
        err_abs = np.linalg.norm(col - fd)
        err_rel = err_abs / np.linalg.norm(col) if np.linalg.norm(col) > 0 else np.inf
        assert err_rel < 0.01 or err_abs < 1e-12, (i, err_rel, err_abs)


def check_control_jacobian(O, fun, analytic, minus=None):  # quadrotor_model_test.cc:57-79
    EPS = 1e-6
    for i in range(4):
        d = np.zeros(4)
        d[i] = EPS
        yp, ym = fun(d), fun(-d)
        fd = (minus(yp, ym) if minus else (yp - ym)) / (2 * EPS)
        col = analytic[:, i]
        err_abs = np.linalg.norm(col - fd)
        err_rel = err_abs / np.linalg.norm(col) if np.linalg.norm(col) > 0 else np.inf
        assert err_rel < 0.01 or err_abs < 1e-12, (i, err_rel, err_abs)


def test_discrete_dynamics_jacobians_fd(O):  # quadrotor_model_test.cc:145-197
    cfg = model_cfg(O, inertia=random_spd_inertia())
    x = x_init(O)
    _, Jx, _ = O.discrete_dynamics(cfg, x, np.zeros(4), DT, diffs=True)
    check_state_jacobian(O, lambda d: O.discrete_dynamics(cfg, O.state_add(x, d), np.zeros(4), DT), Jx,
                         minus=O.state_minus)
    u = np.array([1.0, 2.0, 3.0, 4.0])
    _, _, Ju = O.discrete_dynamics(cfg, x, u, DT, diffs=True)
    check_control_jacobian(O, lambda d: O.discrete_dynamics(cfg, x, u + d, DT), Ju, minus=O.state_minus)


def test_continuous_dynamics_jacobians_fd(O):  # quadrotor_model_test.cc:199-249
    cfg = model_cfg(O, inertia=random_spd_inertia())
    x = x_init(O)
    _, Jx, _ = O.continuous_dynamics(cfg, x, np.zeros(4), diffs=True)
    check_state_jacobian(O, lambda d: O.continuous_dynamics(cfg, O.state_add(x, d), np.zeros(4)), Jx)
    u = np.array([1.0, 2.0, 3.0, 4.0])
    _, _, Ju = O.continuous_dynamics(cfg, x, u, diffs=True)
    check_control_jacobian(O, lambda d: O.continuous_dynamics(cfg, x, u + d), Ju)


def test_state_add_jacobians_fd(O):  # quadrotor_model_test.cc:251-296
    x = x_init(O)
    tangent = np.array([3.0, 4, 5, 6, 7, 8, 4, 5, 6, 7, 8, 9])
    _, Jl, Jr = O.state_add(x, tangent, diffs=True)
    check_state_jacobian(O, lambda d: O.state_add(O.state_add(x, d), tangent), Jl, minus=O.state_minus)
    check_state_jacobian(O, lambda d: O.state_add(x, tangent + d), Jr, minus=O.state_minus)


def test_state_minus_jacobians_fd(O):  # quadrotor_model_test.cc:298-346
    lhs = x_init(O)
    rhs = np.zeros(13)
    rhs[:7] = O.se3_exp(2.0 * np.array([1.0, 2, 3, 4, 5, 6]))
    rhs[7:] = 2.0 * np.array([2.0, 3, 4, 5, 6, 7])
    _, Jl, Jr = O.state_minus(lhs, rhs, diffs=True)
    check_state_jacobian(O, lambda d: O.state_minus(O.state_add(lhs, d), rhs), Jl)
    check_state_jacobian(O, lambda d: O.state_minus(lhs, O.state_add(rhs, d)), Jr)


def test_euler_step_jacobians_fd(O):  # quadrotor_model_test.cc:399-447
    x = x_init(O)
    xdot = np.array([3.0, 4, 5, 6, 7, 8, 4, 5, 6, 7, 8, 9])
    _, Jl, Jr = O.euler_step(x, xdot, DT, diffs=True)
    check_state_jacobian(O, lambda d: O.euler_step(O.state_add(x, d), xdot, DT), Jl, minus=O.state_minus)
    check_state_jacobian(O, lambda d: O.euler_step(x, xdot + d, DT), Jr, minus=O.state_minus)


def test_state_tangent_ordering(O):  # quadrotor_model_test.cc:348-369: velocity 0-5, acceleration 6-11
    cfg = model_cfg(O)
    x = x_identity()
    x[7:] = [1, 2, 3, 4, 5, 6]
    xdot = O.continuous_dynamics(cfg, x, np.zeros(4))
    assert np.array_equal(xdot[:6], x[7:])


def test_inertia_not_positive_definite_rejected(O):  # quadrotor_model.cc:21-24
    assert O.check_model(model_cfg(O)) == 0
    assert O.check_model(model_cfg(O, inertia=np.diag([1.0, -1.0, 1.0]))) == 1
    asym = np.eye(3)
    asym[0, 1] = 0.1
    assert O.check_model(model_cfg(O, inertia=asym)) == 1


# ---- cost_test.cc -------------------------------------------------------------------------------
def random_point(seed):
    rng = np.random.default_rng(seed)
    return rng.uniform(-1, 1, 6), rng.uniform(-1, 1, 6), rng.uniform(-1, 1, 4)


def test_cost_zero_when_zero_error(O):  # cost_test.cc:27-39
    tau, vel, u = random_point(1)
    x = np.concatenate([O.se3_exp(tau), vel])
    cfg = O.make_config(Q=np.eye(12), R=np.eye(4))
    assert O.cost(cfg, x, u, x, u) == 0.0


def test_cost_differentials_fd(O):  # cost_test.cc:66-151
    # desired = the random point itself, evaluated AT the desired point as the reference does
    tau, vel, u = random_point(2)
    x = np.concatenate([O.se3_exp(tau), vel])
    cfg = O.make_config(Q=np.eye(12), R=np.eye(4))
    c, Cx, Cu, Cxx, Cuu, Cxu = O.cost(cfg, x, u, x, u, diffs=True)
    EPS = 1e-6
    f = lambda xx, uu: O.cost(cfg, xx, uu, x, u)
    Cx_fd = np.array([(f(O.state_add(x, e), u) - f(O.state_add(x, -e), u)) / (2 * EPS) for e in EPS * np.eye(12)])
    assert np.linalg.norm(Cx - Cx_fd) / 12 < EPS
    Cu_fd = np.array([(f(x, u + e) - f(x, u - e)) / (2 * EPS) for e in EPS * np.eye(4)])
    assert np.linalg.norm(Cu - Cu_fd) / 4 < EPS
    H_fd = np.zeros((12, 12))
    E = EPS * np.eye(12)
    for i in range(12):
        for j in range(12):
            H_fd[i, j] = (f(O.state_add(O.state_add(x, E[i]), E[j]), u) - f(O.state_add(x, E[i]), u)
                          - f(O.state_add(x, E[j]), u) + f(x, u)) / (EPS * EPS)
    assert np.linalg.norm(np.linalg.inv(Cxx) @ H_fd - np.eye(12)) < 11  # the reference's loose bound
    assert np.allclose(Cuu, 2 * np.eye(4)) and np.all(Cxu == 0)
    # and at a point with non-zero error the gradient is still right
    tau2, vel2, u2 = random_point(3)
    x2 = np.concatenate([O.se3_exp(tau2), vel2])
    _, Cx2, Cu2, _, _, _ = O.cost(cfg, x2, u2, x, u, diffs=True)
    Cx2_fd = np.array([(f(O.state_add(x2, e), u2) - f(O.state_add(x2, -e), u2)) / (2 * EPS) for e in E])
    assert np.linalg.norm(Cx2 - Cx2_fd) / 12 < 1e-5
