"""Cross-checks of the oracle's manif restatement against scipy (expm / logm / Rotation)
and central finite differences -- the guardrails of SURVEY.md section 8c(ii)."""
import numpy as np
import pytest
from scipy.linalg import expm, logm
from scipy.spatial.transform import Rotation


def hat6(tau):
    v, w = tau[:3], tau[3:]
    M = np.zeros((4, 4))
    M[:3, :3] = [[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]
    M[:3, 3] = v
    return M


def to_matrix(O, X):
    T = np.eye(4)
    T[:3, :3] = O.rotation_matrix(X[3:7])
    T[:3, 3] = X[:3]
    return T


TAUS = [np.array([1.0, 2, 3, 0.4, 0.5, 0.6]), np.array([0.1, -0.2, 0.3, 1e-9, -2e-9, 1e-9]),
        np.array([-1.0, 0.5, 2.0, 2.0, -1.5, 1.0]), np.zeros(6), np.array([0.3, 0.1, -0.2, 0, 0, 3.0])]


@pytest.mark.parametrize("tau", TAUS)
def test_exp_log_match_scipy(O, tau):
    X = O.se3_exp(tau)
    assert np.allclose(to_matrix(O, X), expm(hat6(tau)), atol=1e-13)
    assert np.allclose(O.rotation_matrix(X[3:7]), Rotation.from_rotvec(tau[3:]).as_matrix(), atol=1e-14)
    back = O.se3_log(X)
    assert np.allclose(back, tau, atol=1e-12)
    L = np.real(logm(to_matrix(O, X)))
    assert np.allclose(hat6(back), L, atol=1e-9)


def test_compose_inverse_adjoint(O):
    A, B = O.se3_exp(TAUS[0]), O.se3_exp(TAUS[2])
    TA, TB = to_matrix(O, A), to_matrix(O, B)
    assert np.allclose(to_matrix(O, O.se3_compose(A, B)), TA @ TB, atol=1e-13)
    assert np.allclose(to_matrix(O, O.se3_inverse(A)), np.linalg.inv(TA), atol=1e-13)
    # Ad(A) tau^ = A tau^ A^-1
    tau = TAUS[1] + 0.3
    lhs = hat6(O.se3_adj(A) @ tau)
    assert np.allclose(lhs, TA @ hat6(tau) @ np.linalg.inv(TA), atol=1e-12)


@pytest.mark.parametrize("tau", TAUS[:3] + [TAUS[4]])
def test_jacobians_fd_and_inverses(O, tau):
    h = 1e-6
    Jr, Jl = O.se3_rjac(tau), O.se3_ljac(tau)
    X = O.se3_exp(tau)
    Jr_fd, Jl_fd = np.zeros((6, 6)), np.zeros((6, 6))
    for i in range(6):
        e = np.zeros(6)
        e[i] = h
        Xp, Xm = O.se3_exp(tau + e), O.se3_exp(tau - e)
        # right: Log(X^-1 Exp(tau+e)) ; left: Log(Exp(tau+e) X^-1)
        Jr_fd[:, i] = (O.se3_log(O.se3_compose(O.se3_inverse(X), Xp)) - O.se3_log(O.se3_compose(O.se3_inverse(X), Xm))) / (2 * h)
        Jl_fd[:, i] = (O.se3_log(O.se3_compose(Xp, O.se3_inverse(X))) - O.se3_log(O.se3_compose(Xm, O.se3_inverse(X)))) / (2 * h)
    # manif's (1 - cos t)/t^2 loses ~1e-16/t^2 relative accuracy for tiny angles (its small-angle
    # branch only starts at t^2 <= 1e-14), which finite differences amplify by 1/h: loosen there.
    atol = 2e-8 if np.linalg.norm(tau[3:]) > 1e-3 else 1e-3
    assert np.allclose(Jr, Jr_fd, atol=atol)
    assert np.allclose(Jl, Jl_fd, atol=atol)
    assert np.allclose(O.se3_rjacinv(tau) @ Jr, np.eye(6), atol=1e-12)
    assert np.allclose(O.se3_ljacinv(tau) @ Jl, np.eye(6), atol=1e-12)
    assert np.allclose(Jl, O.se3_adj(X) @ Jr, atol=1e-12)


def test_plus_minus_jacobians(O):
    X, tau = O.se3_exp(TAUS[0]), TAUS[2] * 0.3
    Y, JX, Jt = O.se3_plus(X, tau)
    assert np.allclose(JX, np.linalg.inv(O.se3_adj(O.se3_exp(tau))), atol=1e-12)
    assert np.allclose(Jt, O.se3_rjac(tau), atol=0)
    t, JA, JB = O.se3_minus(Y, X)
    assert np.allclose(t, tau, atol=1e-12)
    assert np.allclose(JA, O.se3_rjacinv(t), atol=0)
    assert np.allclose(JB, -O.se3_ljacinv(t), atol=0)


def test_log_angle_branch_near_pi(O):
    # w < 0 branch of SO3::log keeps the angle in (-pi, pi]
    for ang in (3.0, 3.14, -3.1):
        q = np.array([np.sin(ang / 2), 0, 0, np.cos(ang / 2)])
        for sgn in (1.0, -1.0):
            w = O.se3_log(np.concatenate([[0, 0, 0], sgn * q]))[3:]
            assert np.allclose(w, [ang, 0, 0], atol=1e-12)


def test_ldlt4_matches_numpy(O):
    rng = np.random.default_rng(0)
    for _ in range(20):
        A = rng.uniform(-1, 1, (4, 4))
        A = A @ A.T + 0.1 * np.eye(4)
        b = rng.uniform(-1, 1, (4, 13))
        x = O.ldlt4_solve(A, b)
        assert np.allclose(x, np.linalg.solve(A, b), rtol=1e-10, atol=1e-12)
    # lower triangle only
    A2 = A.copy()
    A2[0, 3] = 99.0
    assert np.array_equal(O.ldlt4_solve(A2, b), O.ldlt4_solve(A, b))
