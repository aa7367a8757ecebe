"""The oracle against an INDEPENDENT 50-digit restatement (guardrail iii) -- CPU only.

No reference test pins an N > 3 result, so the oracle is the golden source for every BASELINE config.  This test
rebuilds one full iLQR iteration -- linearisation, Riccati sweep, closed-loop rollout, cost -- from the mathematics
alone, in mpmath at 50 digits, sharing no formula with the oracle:

  * SE(3) through 4x4 homogeneous matrices with the MATRIX exponential / logarithm (mpmath expm / logm) instead of
    the closed forms (Rodrigues, Jl, Jr^-1, Barfoot's Q block) that manif and the oracle use;
  * every Jacobian (dynamics A, B; the cost's d(x (-) x_d)/dx) by central differences at h = 1e-18 in 50-digit
    arithmetic (truncation ~1e-36) instead of the analytic chain rule of quadrotor_model.cc / cost.hh;
  * the Riccati recursion (ilqr.hh:118-140) and rollout (ilqr.hh:157-169) on mpmath matrices, Q_uu inverted explicitly.

Agreement to ~1e-12 therefore checks the oracle's formulas AND bounds its double-precision rounding.
"""
import numpy as np
import pytest

mp = pytest.importorskip("mpmath")
mp.mp.dps = 50
M, mpf = mp.matrix, mp.mpf


def hat3(w):
    return M([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])


def hat6(tau):  # manif SE3 tangent: linear part first, angular part second
    H = mp.zeros(4, 4)
    H[0:3, 0:3] = hat3(tau[3:6])
    for i in range(3):
        H[i, 3] = tau[i]
    return H


def vee6(H):
    return [H[0, 3], H[1, 3], H[2, 3], H[2, 1], H[0, 2], H[1, 0]]


def pose_from_tq(t, q):  # q = (x, y, z, w)
    x, y, z, w = [mpf(float(v)) for v in q]
    n = mp.sqrt(x * x + y * y + z * z + w * w)
    x, y, z, w = x / n, y / n, z / n, w / n
    R = M([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
           [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
           [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    T = mp.eye(4)
    T[0:3, 0:3] = R
    for i in range(3):
        T[i, 3] = mpf(float(t[i]))
    return T


class MpProblem:
    def __init__(self, model, dt, rk4=False, coriolis=False):
        self.rk4, self.coriolis = rk4, coriolis
        self.m, self.g, self.dt = mpf(model["mass_kg"]), mpf(model["g_mpss"]), mpf(dt)
        self.I = M(np.asarray(model["inertia"], dtype=float).tolist())
        a, r = mpf(model["arm_length_m"]), mpf(model["torque_to_thrust_ratio_m"])
        self.arms = M([[0, -a, 0, a], [a, 0, -a, 0], [-r, r, -r, r]])  # quadrotor_model.cc:15-18
        self.Q = M(np.asarray(model["Q"], dtype=float).tolist())
        self.R = M(np.asarray(model["R"], dtype=float).tolist())

    # state = (T 4x4, v list of 6)
    def plus(self, x, d):
        T, v = x
        return T * mp.expm(hat6(d[0:6])), [v[i] + d[6 + i] for i in range(6)]

    def minus(self, x, y):
        return vee6(mp.logm(mp.inverse(y[0]) * x[0])) + [x[1][i] - y[1][i] for i in range(6)]

    def xdot(self, x, u):  # quadrotor_model.cc:65-78
        T, v = x
        R = T[0:3, 0:3]
        ez = M([0, 0, 1])
        lin = -self.g * (R.T * ez) + (sum(u) / self.m) * ez
        w = M(v[3:6])
        if self.coriolis:  # QuadrotorModelVariant: transport term -omega x v
            lin = lin - hat3(w) * M(v[0:3])
        ang = mp.lu_solve(self.I, self.arms * M(u) - hat3(w) * (self.I * w))
        return list(v) + [lin[i] for i in range(3)] + [ang[i] for i in range(3)]

    def step(self, x, u):
        if not self.rk4:  # explicit Euler on the manifold, quadrotor_model.cc:33-49
            return self.plus(x, [self.dt * c for c in self.xdot(x, u)])
        k, acc = [mpf(0)] * 12, [mpf(0)] * 12  # the scheme of quadrotor_model.cc:51-63
        for ci, dti in ((mpf(1) / 6, mpf(0)), (mpf(2) / 6, self.dt / 2), (mpf(2) / 6, self.dt / 2), (mpf(1) / 6, self.dt)):
            k = self.xdot(self.plus(x, [dti * c for c in k]), u)
            acc = [a + ci * c for a, c in zip(acc, k)]
        return self.plus(x, [self.dt * c for c in acc])

    def jac(self, f, n, h=mpf(10) ** -18):
        cols = []
        for j in range(n):
            e = [mpf(0)] * n
            e[j] = h
            cols.append([(a - b) / (2 * h) for a, b in zip(f(e), f([-c for c in e]))])
        return M(cols).T

    def linearise(self, x, u, xd, ud):
        fx = self.step(x, u)
        A = self.jac(lambda d: self.minus(self.step(self.plus(x, d), u), fx), 12)
        B = self.jac(lambda d: self.minus(self.step(x, [u[i] + d[i] for i in range(4)]), fx), 4)
        dx, du = M(self.minus(x, xd)), M([u[i] - ud[i] for i in range(4)])
        J = self.jac(lambda d: self.minus(self.plus(x, d), xd), 12)
        Cx, Cxx = 2 * (J.T * (self.Q.T * dx)), 2 * (J.T * self.Q * J)  # cost.hh:49-52: 2 dx^T Q J, 2 J^T Q J
        Cu, Cuu = 2 * (self.R.T * du), 2 * self.R
        return A, B, Cx, Cu, Cxx, Cuu

    def cost(self, x, u, xd, ud):
        dx, du = M(self.minus(x, xd)), M([u[i] - ud[i] for i in range(4)])
        return (dx.T * self.Q * dx)[0] + (du.T * self.R * du)[0]


def to_mp_traj(traj):
    return [((pose_from_tq(p[1:4], p[4:8]), [mpf(float(v)) for v in p[8:14]]), [mpf(float(v)) for v in p[14:18]])
            for p in traj]


def test_one_ilqr_iteration_against_50_digit_arithmetic(O):
    from quadrotorilqr_b200 import problems

    model = dict(problems.hover_model(), inertia=np.array([[1.0, 0.05, 0.0], [0.05, 1.4, 0.1], [0.0, 0.1, 1.2]]),
                 torque_to_thrust_ratio_m=0.3)
    A_ = np.random.default_rng(1).uniform(-0.5, 0.5, (12, 12))
    model["Q"] = model["Q"] + A_ @ A_.T  # a symmetric Q with pose/velocity coupling
    N, dt = 8, model["dt_s"]
    cfg = O.make_config(mass_kg=model["mass_kg"], inertia=model["inertia"], arm_length_m=model["arm_length_m"],
                        torque_to_thrust_ratio_m=model["torque_to_thrust_ratio_m"], g_mpss=model["g_mpss"],
                        Q=model["Q"], R=model["R"], dt_s=dt)
    desired = problems.hover_desired_trajectory(N, dt, model["mass_kg"], model["g_mpss"])
    x0 = problems.hover_initial_states(1, seed=3)[0]
    seed = problems.constant_state_trajectory(x0[None], N, dt, desired[0, 14:18])[0]
    traj = O.forward_sim(cfg, desired, seed, np.zeros((N, 4)), np.zeros((N, 4, 12)))
    traj[:, 14:18] += np.random.default_rng(2).uniform(-0.3, 0.3, (N, 4))  # non-trivial control gradient

    P = MpProblem(model, dt)
    tr, des = to_mp_traj(traj), to_mp_traj(desired)

    # ---- backward pass (ilqr.hh:97-147) ----
    Vx, Vxx = mp.zeros(12, 1), mp.zeros(12, 12)
    k_mp, K_mp, QuTk, kTQuuk = [None] * N, [None] * N, mpf(0), mpf(0)
    for i in range(N - 1, -1, -1):
        (x, u), (xd, ud) = tr[i], des[i]
        A, B, Cx, Cu, Cxx, Cuu = P.linearise(x, u, xd, ud)
        Qx, Qu = Cx + A.T * Vx, Cu + B.T * Vx
        Qxx, Quu, Qxu = Cxx + A.T * Vxx * A, Cuu + B.T * Vxx * B, A.T * Vxx * B
        Quu_inv = mp.inverse(Quu)
        K = -(Quu_inv * Qxu.T)
        k = -(Quu_inv * Qu)
        Vx = Qx - K.T * Quu * k
        Vxx = Qxx - K.T * Quu * K
        QuTk += (Qu.T * k)[0]
        kTQuuk += (k.T * Quu * k)[0]
        k_mp[i], K_mp[i] = k, K
    ko, Ko, a, c = O.backwards_pass(cfg, desired, traj)
    k_ref = np.array([[float(v) for v in k] for k in k_mp])
    K_ref = np.array([[[float(K[r, s]) for s in range(12)] for r in range(4)] for K in K_mp])
    print(f"oracle vs 50-digit: k {np.max(np.abs(ko - k_ref)) / max(1.0, np.abs(k_ref).max()):.1e}, "
          f"K {np.max(np.abs(Ko - K_ref)) / max(1.0, np.abs(K_ref).max()):.1e}, "
          f"QuTk {abs(a - float(QuTk)) / abs(float(QuTk)):.1e}, kTQuuk {abs(c - float(kTQuuk)) / abs(float(kTQuuk)):.1e}")
    assert np.max(np.abs(ko - k_ref)) <= 1e-12 * max(1.0, np.abs(k_ref).max())
    assert np.max(np.abs(Ko - K_ref)) <= 1e-12 * max(1.0, np.abs(K_ref).max())
    assert abs(a - float(QuTk)) <= 1e-12 * abs(float(QuTk)) and abs(c - float(kTQuuk)) <= 1e-12 * abs(float(kTQuuk))

    # ---- rollout with those gains at alpha = 0.5 and its cost (ilqr.hh:149-172, 89-95) ----
    alpha = mpf("0.5")
    x, cost, states, controls = tr[0][0], mpf(0), [], []
    for i in range(N):
        d = M(P.minus(x, tr[i][0]))
        u = [tr[i][1][j] + alpha * k_mp[i][j] + (K_mp[i] * d)[j] for j in range(4)]
        states.append(x)
        controls.append(u)
        cost += P.cost(x, u, des[i][0], des[i][1])
        x = P.step(x, u)
    new = O.forward_sim(cfg, desired, traj, ko, Ko, 0.5)
    new_mp = to_mp_traj(new)
    for i in range(N):
        err = max(abs(v) for v in P.minus(new_mp[i][0], states[i]))
        assert err < mpf(10) ** -12, (i, err)
        assert max(abs(mpf(float(new[i, 14 + j])) - controls[i][j]) for j in range(4)) < mpf(10) ** -12
    print(f"rollout cost {float(abs(mpf(O.cost_trajectory(cfg, desired, new)) - cost) / abs(cost)):.1e}")
    assert abs(mpf(O.cost_trajectory(cfg, desired, new)) - cost) <= mpf(10) ** -13 * abs(cost)


@pytest.mark.parametrize("kind", [0, 1, 2, 3])
def test_discrete_dynamics_and_jacobians_against_50_digit_arithmetic(O, kind):
    """x+, J_x, J_u of the reference model (kind 0) and of the RK4 / Coriolis variants, against the matrix-exponential
    model differentiated numerically at 50 digits."""
    from quadrotorilqr_b200 import problems

    model = dict(problems.hover_model(), inertia=np.array([[1.0, 0.05, 0.0], [0.05, 1.4, 0.1], [0.0, 0.1, 1.2]]),
                 torque_to_thrust_ratio_m=0.3, mass_kg=1.3)
    cfg = O.make_config(mass_kg=model["mass_kg"], inertia=model["inertia"], arm_length_m=model["arm_length_m"],
                        torque_to_thrust_ratio_m=model["torque_to_thrust_ratio_m"], g_mpss=model["g_mpss"],
                        dt_s=model["dt_s"], model_kind=kind)
    P = MpProblem(model, model["dt_s"], rk4=bool(kind & 1), coriolis=bool(kind & 2))
    rng = np.random.default_rng(kind)
    xs = np.concatenate([O.se3_exp(rng.normal(size=6) * 0.6), rng.normal(size=6) * 0.8])
    us = rng.normal(size=4) + 3.0
    xn, A, B = O.discrete_dynamics(cfg, xs, us, diffs=True)
    x = (pose_from_tq(xs[0:3], xs[3:7]), [mpf(float(v)) for v in xs[7:13]])
    u = [mpf(float(v)) for v in us]
    fx = P.step(x, u)
    got = (pose_from_tq(xn[0:3], xn[3:7]), [mpf(float(v)) for v in xn[7:13]])
    assert max(abs(v) for v in P.minus(got, fx)) < mpf(10) ** -14
    A_mp = P.jac(lambda d: P.minus(P.step(P.plus(x, d), u), fx), 12)
    B_mp = P.jac(lambda d: P.minus(P.step(x, [u[i] + d[i] for i in range(4)]), fx), 4)
    A_ref = np.array([[float(A_mp[r, c]) for c in range(12)] for r in range(12)])
    B_ref = np.array([[float(B_mp[r, c]) for c in range(4)] for r in range(12)])
    assert np.max(np.abs(A - A_ref)) < 1e-13 and np.max(np.abs(B - B_ref)) < 1e-13
