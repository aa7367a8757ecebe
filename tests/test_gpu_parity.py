"""Parity of the CUDA path (through the C ABI) with the CPU oracle -- the first gate.

Tolerance (FP64 mode, BASELINE.json north_star): trajectories, gains and per-iteration
costs agree within RTOL = 1e-9 relative to the array's scale; iteration counts, rollout
counts and convergence flags must be IDENTICAL.
"""
import numpy as np
import pytest

from conftest import identity_traj, make_solver, oracle_config, random_spd_inertia

pytestmark = pytest.mark.gpu

RTOL = 1e-9


FLOOR = 1e-3  # entries below FLOOR * (the array's largest magnitude) are compared as if they had that magnitude


def assert_close(a, b, rtol=RTOL, what="", floor=FLOOR):
    """ELEMENT-WISE relative comparison: |a_ij - b_ij| <= rtol * max(|b_ij|, floor * max|b|, tiny).

    Small gain entries are thus held to rtol relative to themselves down to 1e-3 of the array's scale (below that
    an entry is the result of cancellation and only its absolute error is meaningful).  rtol = 0 demands equality."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if not b.size:
        return
    scale = max(1.0, float(np.max(np.abs(b))))
    tol = rtol * np.maximum(np.abs(b), floor * scale)
    err = np.abs(a - b)
    bad = err > tol
    if np.any(bad):
        i = np.unravel_index(np.argmax(err / np.maximum(tol, 1e-300)), err.shape) if err.ndim else ()
        raise AssertionError(f"{what}: element {i}: |{a[i]!r} - {b[i]!r}| = {err[i]:.3e} > {tol[i]:.3e} "
                             f"(rtol {rtol:g}, array scale {scale:.3e})")


def random_states(O, n, seed, big=False):
    rng = np.random.default_rng(seed)
    x = np.zeros((n, 13))
    for i in range(n):
        s = 3.0 if big else 0.7
        tau = rng.uniform(-s, s, 6)
        if i % 7 == 3:
            tau[3:] *= 1e-9  # small-angle branches
        if i % 11 == 5:
            tau[3:] = 0.0
        x[i, :7] = O.se3_exp(tau)
        x[i, 7:] = rng.uniform(-2, 2, 6)
    return x


@pytest.fixture(scope="module")
def models():
    from quadrotorilqr_b200 import problems

    dflt = problems.default_model()
    rnd = dict(dflt, inertia=random_spd_inertia(), torque_to_thrust_ratio_m=1.0, mass_kg=1.3,
               arm_length_m=0.4)
    rng = np.random.default_rng(5)
    A = rng.uniform(-1, 1, (12, 12))
    dense = dict(rnd, Q=A @ A.T + np.eye(12), R=np.diag([1.0, 2.0, 3.0, 4.0]) + 0.1)
    return dict(default=dflt, random_inertia=rnd, dense_Q=dense)


# ---------------------------------------------------------------------------------------------
# model / Lie library
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["default", "random_inertia"])
def test_dynamics_match_oracle(O, models, name):
    from quadrotorilqr_b200 import ILQROptions

    model = models[name]
    s = make_solver(model)
    cfg = oracle_config(O, model, ILQROptions())
    x = random_states(O, 96, 1)
    u = np.random.default_rng(2).uniform(-3, 6, (96, 4))
    xn, Jx, Ju = s.discrete_dynamics(x, u, diffs=True)
    xd, Jxc, Juc = s.continuous_dynamics(x, u, diffs=True)
    assert_close(s.discrete_dynamics(x, u), xn, 0, "with/without diffs")
    for i in range(x.shape[0]):
        r = O.discrete_dynamics(cfg, x[i], u[i], diffs=True)
        assert_close(xn[i], r[0], what="x_next")
        assert_close(Jx[i], r[1], what="J_x")
        assert_close(Ju[i], r[2], what="J_u")
        c = O.continuous_dynamics(cfg, x[i], u[i], diffs=True)
        assert_close(xd[i], c[0], what="xdot")
        assert_close(Jxc[i], c[1], what="J_x cont")
        assert_close(Juc[i], c[2], what="J_u cont")


def test_state_minus_add_match_oracle(O, models):
    s = make_solver(models["default"])
    a, b = random_states(O, 80, 3, big=True), random_states(O, 80, 4, big=True)
    tangent = np.random.default_rng(6).uniform(-2, 2, (80, 12))
    tangent[5, 3:6] = 0.0
    tangent[6, 3:6] *= 1e-9
    d, Jl, Jr = s.state_minus(a, b, diffs=True)
    y, Al, Ar = s.state_add(a, tangent, diffs=True)
    for i in range(80):
        r = O.state_minus(a[i], b[i], diffs=True)
        assert_close(d[i], r[0], what="minus")
        assert_close(Jl[i], r[1], what="minus J_lhs")
        assert_close(Jr[i], r[2], what="minus J_rhs")
        q = O.state_add(a[i], tangent[i], diffs=True)
        assert_close(y[i], q[0], what="add")
        assert_close(Al[i], q[1], what="add J_lhs")
        assert_close(Ar[i], q[2], what="add J_rhs")
    # x (-) x is exactly zero; x (+) 0 is x up to the renormalisation rule
    z = s.state_minus(a, a)
    assert np.all(z == 0.0)


@pytest.mark.parametrize("name", ["default", "dense_Q"])
def test_cost_matches_oracle(O, models, name):
    from quadrotorilqr_b200 import ILQROptions

    model = models[name]
    s = make_solver(model)
    cfg = oracle_config(O, model, ILQROptions())
    x, xd = random_states(O, 64, 7), random_states(O, 64, 8)
    rng = np.random.default_rng(9)
    u, ud = rng.uniform(-3, 6, (64, 4)), rng.uniform(-3, 6, (64, 4))
    c, Cx, Cu, Cxx, Cuu, Cxu = s.cost(x, u, xd, ud, diffs=True)
    assert_close(s.cost(x, u, xd, ud), c, 0)
    for i in range(64):
        r = O.cost(cfg, x[i], u[i], xd[i], ud[i], diffs=True)
        assert_close(c[i], r[0], what="cost")
        assert_close(Cx[i], r[1], what="C.x")
        assert_close(Cu[i], r[2], what="C.u")
        assert_close(Cxx[i], r[3], what="C.xx")
        assert_close(Cuu[i], r[4], what="C.uu")
        assert_close(Cxu[i], r[5], what="C.xu")
    assert np.all(s.cost(x, u, x, u) == 0.0)  # cost_test.cc:27-39


def test_reference_dynamics_known_answers(models):
    """quadrotor_model_test.cc:94-143 through the CUDA path."""
    s = make_solver(models["default"], torque_to_thrust_ratio_m=1.0)
    x = np.zeros(13)
    x[6] = 1.0
    x[7:10] = [1.0, 2.0, 3.0]
    xn = s.discrete_dynamics(x, np.ones(4))
    assert np.allclose(xn[0:3], [0.1, 0.2, 0.3], rtol=1e-6)
    assert np.allclose(xn[7:13], [1.0, 2.0, 3.0 + (4.0 - 9.81) * 0.1, 0, 0, 0], rtol=1e-6)
    x = np.zeros(13)
    x[6] = 1.0
    x[10] = 1.2
    xn = s.discrete_dynamics(x, [0.0, -1.0, 0.0, 1.0])
    assert np.allclose(xn[10:13], [1.4, 0, 0], rtol=1e-6)
    assert np.allclose(xn[3:7], [np.sin(0.06), 0, 0, np.cos(0.06)], atol=1e-9)


# ---------------------------------------------------------------------------------------------
# reference ILQRFixture known answers (ilqr_test.cc:68-190) through the CUDA path
# ---------------------------------------------------------------------------------------------
def ilqr_fixture():
    from quadrotorilqr_b200 import ConvergenceCriteria, ILQROptions, LineSearchParams, problems

    model = dict(problems.default_model(), torque_to_thrust_ratio_m=1.0, g_mpss=0.0, Q=np.eye(12))
    opts = ILQROptions(LineSearchParams(0.5, 0.5, 10), ConvergenceCriteria(1e-12, 1e-12, 100.0))
    s = make_solver(model, opts)
    cur = identity_traj(3, 0.1)
    return s, model, opts, cur, np.ones((3, 4)), np.zeros((3, 4, 12))


def test_ilqr_fixture_known_answers(O):
    s, model, opts, cur, k, K = ilqr_fixture()
    new = s.forward_sim(cur, k, K)
    assert np.allclose(new[:, 10], [0.0, 0.4, 0.8], atol=1e-6)  # v_z
    assert np.allclose(new[:, 3], [0.0, 0.0, 0.04], atol=1e-6)  # z
    assert np.all(new[:, 14:] == 1.0) and np.array_equal(new[:, 0], cur[:, 0])
    cost = s.cost_trajectory(new, cur)
    assert abs(cost - 12.8016) <= 4 * np.spacing(12.8016)  # EXPECT_DOUBLE_EQ (ilqr_test.cc:140)
    k0, K0, QuTk, kTQuuk = s.backwards_pass(cur, cur)
    assert QuTk == 0.0 and kTQuuk == 0.0 and np.all(k0 == 0.0)  # ilqr_test.cc:143-153
    k1, K1, QuTk, kTQuuk = s.backwards_pass(new, cur)
    assert QuTk < 0.0  # ilqr_test.cc:155-164
    t2, c2, step = s.line_search(new, cur, cost, k1, K1, QuTk, kTQuuk)
    assert c2 - cost < 0.5 * (step * QuTk + step * step * kTQuuk / 2.0)  # ilqr_test.cc:166-177
    kk = k.copy()
    kk[:, 0] *= 100
    kk[:, 2] *= 100
    initial = s.forward_sim(cur, kk, K)
    r = s.solve(initial, cur)
    assert np.max(np.abs(r["traj"][0] - cur)) < 1e-6  # ilqr_test.cc:179-190
    # and the same numbers as the oracle
    cfg = oracle_config(O, model, opts)
    ro = O.solve(cfg, cur, initial)
    assert r["results"]["status"][0] == ro["status"]
    assert r["results"]["backward_passes"][0] == ro["backward_passes"]
    assert_close(r["traj"][0], ro["traj"], what="solve traj")


def test_line_search_exhaustion_and_range_errors():
    from quadrotorilqr_b200 import QilqrError, _capi

    s, model, opts, cur, k, K = ilqr_fixture()
    new = s.forward_sim(cur, k, K)
    cost = s.cost_trajectory(new, cur)
    k1, K1, QuTk, kTQuuk = s.backwards_pass(new, cur)
    with pytest.raises(QilqrError) as e:  # ilqr.hh:191-193
        s.line_search(new, cur, -1e30, k1, K1, QuTk, kTQuuk)
    assert e.value.code == _capi.ERR_LINE_SEARCH
    with pytest.raises(QilqrError) as e:  # cost.hh:39-40
        s.cost_trajectory(new, cur[:2])
    assert e.value.code == _capi.ERR_OUT_OF_RANGE


# ---------------------------------------------------------------------------------------------
# solver pieces and full solves on the BASELINE configs
# ---------------------------------------------------------------------------------------------
def hover_batch(s, B, N=40, seed=0, first=0):
    from quadrotorilqr_b200 import problems

    m = problems.hover_model()
    desired = problems.hover_desired_trajectory(N, m["dt_s"], m["mass_kg"], m["g_mpss"])
    x0 = problems.hover_initial_states(B, seed=seed, first=first)
    seedtraj = problems.constant_state_trajectory(x0, N, m["dt_s"], desired[0, 14:18])
    initial = s.forward_sim(seedtraj, np.zeros((B, N, 4)), np.zeros((B, N, 48)))
    return desired, initial


def test_pieces_match_oracle_on_hover_batch(O):
    from quadrotorilqr_b200 import problems

    model, opts = problems.hover_model(), problems.default_options(False)
    s = make_solver(model, opts)
    cfg = oracle_config(O, model, opts)
    B, N = 37, 40  # ragged batch (not a multiple of the warp size)
    desired, initial = hover_batch(s, B, N)
    cost = s.cost_trajectory(initial, desired)
    k, K, QuTk, kTQuuk = s.backwards_pass(initial, desired)
    new = s.forward_sim(initial, k, K, 1.0)
    half = s.forward_sim(initial, k, K, 0.5)
    for b in range(B):
        seed = problems.constant_state_trajectory(initial[b, 0, 1:14], N, model["dt_s"], desired[0, 14:18])[0]
        assert_close(initial[b], O.forward_sim(cfg, desired, seed, np.zeros((N, 4)), np.zeros((N, 4, 12))),
                     what="open loop")
        assert_close(cost[b], O.cost_trajectory(cfg, desired, initial[b]), what="cost")
        ko, Ko, a, c = O.backwards_pass(cfg, desired, initial[b])
        assert_close(k[b], ko, what="k")
        assert_close(K[b], Ko, what="K")
        assert_close(QuTk[b], a, what="QuTk")
        assert_close(kTQuuk[b], c, what="kTQuuk")
        assert_close(new[b], O.forward_sim(cfg, desired, initial[b], ko, Ko, 1.0), what="fwd 1.0")
        assert_close(half[b], O.forward_sim(cfg, desired, initial[b], ko, Ko, 0.5), what="fwd 0.5")


def check_solve_against_oracle(O, s, cfg, desired, initial, hist_cap=100):
    B = initial.shape[0]
    r = s.solve(initial, desired, want_gains=True, hist_cap=hist_cap)
    o = O.solve_batch(cfg, desired, initial, want_gains=True, hist_cap=hist_cap)
    res = r["results"]
    assert np.array_equal(res["status"], o["status"]), "convergence flags differ"
    assert np.array_equal(res["backward_passes"], o["backward_passes"]), "iteration counts differ"
    assert np.array_equal(res["rollouts"], o["rollouts"]), "rollout counts differ"
    assert np.array_equal(res["num_debug"], o["num_debug"])
    for b in range(B):
        assert_close(r["traj"][b], o["traj"][b], what=f"traj[{b}]")
        assert_close(r["k"][b], o["k"][b], what=f"k[{b}]")
        assert_close(r["K"][b], o["K"][b].reshape(r["K"][b].shape), what=f"K[{b}]")
        assert_close(r["cost_history"][b], o["cost_history"][b], what=f"cost history[{b}]")
        assert_close(res["final_cost"][b], o["final_cost"][b], what="final cost")
    return r, o


def test_solve_matches_oracle_hover_batch(O):
    """BASELINE config 2 at a size the oracle finishes in seconds."""
    from quadrotorilqr_b200 import problems

    model, opts = problems.hover_model(), problems.default_options(False)
    s = make_solver(model, opts)
    cfg = oracle_config(O, model, opts)
    desired, initial = hover_batch(s, 200, 40)
    r, o = check_solve_against_oracle(O, s, cfg, desired, initial)
    assert np.all(np.isin(r["results"]["status"], [1, 2]))
    assert r["results"]["backward_passes"].min() >= 3


def test_solve_matches_oracle_default_problem(O):
    """BASELINE config 1: the reference's default problem (quadrotor_ilqr.py:256-306)."""
    from quadrotorilqr_b200 import problems

    model, opts = problems.default_model(), problems.default_options(True)
    s = make_solver(model, opts)
    cfg = oracle_config(O, model, opts)
    desired = problems.default_desired_trajectory()
    r = s.solve(desired[None], desired, want_gains=True, hist_cap=100, want_debug=True)
    o = O.solve(cfg, desired, desired)
    res = r["results"][0]
    assert res["status"] == o["status"] == 1
    assert res["backward_passes"] == o["backward_passes"] == 77
    assert res["rollouts"] == o["rollouts"] == 77
    assert res["num_debug"] == o["num_debug"] == 76
    assert_close(r["traj"][0], o["traj"], what="traj")
    assert_close(r["cost_history"][0][:76], o["cost_history"], what="cost history")
    assert_close(r["k"][0], o["k"], what="k")
    assert_close(r["K"][0], o["K"], what="K")
    assert_close(r["debug"][0][:76], o["debug"], what="ILQRDebug trajectories")
    assert abs(res["final_cost"] - 22556.502591980552) < 1e-9 * 22556.5


def test_solve_per_problem_desired_and_options(O):
    """Waypoint variant (per-problem desired), symmetrised V_xx, regularised Q_uu, dense Q."""
    from quadrotorilqr_b200 import ConvergenceCriteria, ILQROptions, LineSearchParams, problems

    model = problems.hover_model()
    rng = np.random.default_rng(3)
    A = rng.uniform(-1, 1, (12, 12))
    model["Q"] = model["Q"] + 0.5 * (A @ A.T)
    opts = ILQROptions(LineSearchParams(0.5, 0.5, 20), ConvergenceCriteria(1e-10, 1e-10, 30.0),
                       symmetrize_vxx=True, quu_regularization=1e-3)
    s = make_solver(model, opts)
    cfg = oracle_config(O, model, opts)
    B, N = 33, 24
    _, initial = hover_batch(s, B, N, seed=11)
    base = problems.hover_desired_trajectory(N)
    desired = np.repeat(base[None], B, axis=0)
    way = rng.uniform(-1, 1, (B, 4, 3))
    for seg in range(4):
        desired[:, seg * N // 4:(seg + 1) * N // 4, 1:4] = way[:, seg][:, None, :]
    check_solve_against_oracle(O, s, cfg, desired, initial, hist_cap=30)


def test_solve_edge_cases(O):
    """N = 1 and N = 2 horizons, batch 1, max_iters 1, and a forced line-search failure."""
    from quadrotorilqr_b200 import ConvergenceCriteria, ILQROptions, LineSearchParams, problems

    model = problems.hover_model()
    for N, B, max_it in ((1, 3, 100.0), (2, 1, 100.0), (5, 4, 1.0), (7, 2, 2.5)):
        opts = ILQROptions(LineSearchParams(0.5, 0.5, 100), ConvergenceCriteria(1e-12, 1e-12, max_it))
        s = make_solver(model, opts)
        cfg = oracle_config(O, model, opts)
        desired, initial = hover_batch(s, B, N, seed=N)
        check_solve_against_oracle(O, s, cfg, desired, initial)
    # line search failure: desired_reduction_frac > 1 can never be met near the optimum
    opts = ILQROptions(LineSearchParams(0.5, 4.0, 3), ConvergenceCriteria(1e-14, 0.0, 20.0))
    s = make_solver(model, opts)
    cfg = oracle_config(O, model, opts)
    desired, initial = hover_batch(s, 6, 10, seed=2)
    r, o = check_solve_against_oracle(O, s, cfg, desired, initial)
    assert np.any(r["results"]["status"] == 4)


def test_results_do_not_depend_on_batch_composition(O):
    """Multi-GPU invariant: a problem's outputs are bit-identical whichever
    shard / batch it is solved in."""
    from quadrotorilqr_b200 import problems

    model, opts = problems.hover_model(), problems.default_options(False)
    s = make_solver(model, opts)
    desired, initial = hover_batch(s, 96, 40, seed=4)
    full = s.solve(initial, desired, want_gains=True)
    for lo, hi in ((0, 48), (48, 96), (10, 11), (64, 96)):
        part = s.solve(initial[lo:hi], desired, want_gains=True)
        assert np.array_equal(part["traj"], full["traj"][lo:hi])
        assert np.array_equal(part["K"], full["K"][lo:hi])
        assert np.array_equal(part["results"], full["results"][lo:hi])


def test_device_resident_path_matches_host_path():
    import torch
    from quadrotorilqr_b200 import problems

    model, opts = problems.hover_model(), problems.default_options(False)
    s = make_solver(model, opts)
    B, N = 70, 40
    desired, initial = hover_batch(s, B, N, seed=6)
    host = s.solve(initial, desired)
    dev = torch.device("cuda:0")
    aos = torch.from_numpy(initial).to(dev)
    soa = torch.empty((N, 17, B), dtype=torch.float64, device=dev)
    des_aos = torch.from_numpy(desired[None].copy()).to(dev)
    des = torch.empty((N, 17, 1), dtype=torch.float64, device=dev)
    res = torch.zeros(B * 24, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    s.pack_trajectory_device(aos, soa)
    s.pack_trajectory_device(des_aos, des)
    s.solve_device(soa, des, results=res)
    out = torch.zeros_like(aos)
    s.unpack_trajectory_device(soa, out, time_src=aos)
    torch.cuda.current_stream().synchronize()
    import ctypes
    # the solver runs on its own stream; solve_device returns after completion
    from quadrotorilqr_b200 import RESULT_DTYPE
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), host["traj"])
    r = np.frombuffer(res.cpu().numpy().tobytes(), dtype=RESULT_DTYPE)
    assert np.array_equal(r, host["results"])
    stats = s.last_solve_stats()
    assert stats["problem_iterations"] == int(host["results"]["backward_passes"].sum())
    assert stats["problem_rollouts"] == int(host["results"]["rollouts"].sum())
    assert stats["kernel_launches"] > 0


# ---------------------------------------------------------------------------------------------
# parallel line search (north star item 4) and the long-horizon config (BASELINE config 4)
# ---------------------------------------------------------------------------------------------
def test_parallel_alphas_equal_sequential_search(O):
    """Evaluating P step sizes per round as parallel rollouts must accept exactly the step the
    sequential backtracking accepts: same trajectories (bit for bit vs the sequential GPU path),
    same counts as the oracle.  The default problem backtracks once (alpha = 0.5 at iteration 4)."""
    import dataclasses
    from quadrotorilqr_b200 import ConvergenceCriteria, ILQROptions, LineSearchParams, problems

    model = problems.default_model()
    desired = problems.default_desired_trajectory()
    seq = make_solver(model, problems.default_options(False)).solve(desired[None], desired, want_gains=True, hist_cap=100)
    for P in (2, 8):
        opts = dataclasses.replace(problems.default_options(False), num_parallel_alphas=P)
        par = make_solver(model, opts).solve(desired[None], desired, want_gains=True, hist_cap=100)
        assert np.array_equal(par["traj"], seq["traj"]) and np.array_equal(par["K"], seq["K"])
        assert np.array_equal(par["results"], seq["results"])
        assert np.array_equal(par["cost_history"], seq["cost_history"])
    # forced exhaustion and multi-round backtracking (line search limit 5, P = 2 -> 3 rounds)
    model = problems.hover_model()
    opts = ILQROptions(LineSearchParams(0.5, 4.0, 5), ConvergenceCriteria(1e-14, 0.0, 20.0), num_parallel_alphas=2)
    s = make_solver(model, opts)
    cfg = oracle_config(O, model, opts)
    desired, initial = hover_batch(s, 9, 10, seed=2)
    r, o = check_solve_against_oracle(O, s, cfg, desired, initial)
    assert np.any(r["results"]["status"] == 4)
    # hover batch with P = 4
    opts = dataclasses.replace(problems.default_options(False), num_parallel_alphas=4)
    s = make_solver(model, opts)
    cfg = oracle_config(O, model, opts)
    desired, initial = hover_batch(s, 40, 40, seed=8)
    check_solve_against_oracle(O, s, cfg, desired, initial)


def test_long_horizon_figure_eight(O):
    """BASELINE config 4 at a size the oracle finishes in seconds: N = 1000, dt = 0.02, figure-eight
    tracking, 8 parallel alphas, symmetrised V_xx in BOTH oracle and GPU (the unsymmetrised reference
    recursion diverges at this horizon -- SURVEY fact 5).  Tolerance: iteration counts / flags identical;
    trajectories within 1e-7 relative (rounding differences are amplified along 1000 knots)."""
    import dataclasses
    from quadrotorilqr_b200 import problems

    N, dt = 1000, 0.02
    model = dict(problems.hover_model(), dt_s=dt)
    opts = dataclasses.replace(problems.default_options(False), symmetrize_vxx=True, num_parallel_alphas=8)
    s = make_solver(model, opts)
    cfg = oracle_config(O, model, opts)
    desired = problems.figure_eight_desired(N, dt)
    rng = np.random.default_rng(4)
    B = 3
    x0 = np.tile(desired[0, 1:14], (B, 1))
    x0[:, 0:3] += rng.uniform(-0.3, 0.3, (B, 3))
    seedtraj = problems.constant_state_trajectory(x0, N, dt, desired[0, 14:18])
    initial = s.forward_sim(seedtraj, np.zeros((B, N, 4)), np.zeros((B, N, 48)))
    r = s.solve(initial, desired, hist_cap=100)
    o = O.solve_batch(cfg, desired, initial, hist_cap=100)
    assert np.array_equal(r["results"]["status"], o["status"])
    assert np.array_equal(r["results"]["backward_passes"], o["backward_passes"])
    assert np.array_equal(r["results"]["rollouts"], o["rollouts"])
    assert np.all(np.isin(o["status"], [1, 2])) and o["backward_passes"].max() < 30
    for b in range(B):
        assert_close(r["traj"][b], o["traj"][b], what="long-horizon traj")  # measured over 4096 problems: 7.5e-15
        assert_close(r["cost_history"][b], o["cost_history"][b], rtol=1e-9, what="long-horizon cost history")


def test_quad_and_thread_backward_kernels_agree(monkeypatch):
    """The 4-lanes-per-problem kernel and the one-thread-per-problem kernel implement the same
    recursion; they must agree to rounding on the same inputs."""
    from quadrotorilqr_b200 import problems

    model, opts = problems.hover_model(), problems.default_options(False)
    s_quad = make_solver(model, opts)
    desired, initial = hover_batch(s_quad, 21, 40, seed=12)
    monkeypatch.setenv("QILQR_BACKWARD", "t1")
    s_thread = make_solver(model, opts)
    monkeypatch.delenv("QILQR_BACKWARD")
    kq, Kq, aq, cq = s_quad.backwards_pass(initial, desired)
    kt, Kt, at, ct = s_thread.backwards_pass(initial, desired)
    assert_close(Kq, Kt, rtol=1e-11, what="K")
    assert_close(kq, kt, rtol=1e-11, what="k")
    assert_close(aq, at, rtol=1e-11, what="QuTk")
    assert_close(cq, ct, rtol=1e-11, what="kTQuuk")
    for kpp in ("1", "2"):
        monkeypatch.setenv("QILQR_KPP", kpp)
        s_k = make_solver(model, opts)
        monkeypatch.delenv("QILQR_KPP")
        k2, K2, a2, c2 = s_k.backwards_pass(initial, desired)
        assert np.array_equal(K2, Kq) and np.array_equal(k2, kq)  # same arithmetic, different staging


def test_rollout_kernels_agree(monkeypatch):
    """The role-specialised rollout (three warps per 32 problems: control / pose / cost) runs the same
    instruction sequences as the one-thread-per-problem rollout: bit-identical trajectories, costs and solves."""
    from quadrotorilqr_b200 import problems

    model, opts = problems.hover_model(), problems.default_options(False)
    s_ws = make_solver(model, opts)
    monkeypatch.setenv("QILQR_ROLLOUT", "thread")
    s_th = make_solver(model, opts)
    monkeypatch.delenv("QILQR_ROLLOUT")
    for B in (1, 37, 200):  # ragged: partial warps and partial CTAs
        desired, initial = hover_batch(s_ws, B, 40, seed=3)
        k, K, _, _ = s_ws.backwards_pass(initial, desired)
        for alpha in (1.0, 0.25):
            fa, fb = s_ws.forward_sim(initial, k, K, alpha), s_th.forward_sim(initial, k, K, alpha)
            assert np.array_equal(fa, fb), f"max abs diff {np.abs(fa - fb).max():.3e}"
        a, b = s_ws.solve(initial, desired, want_gains=True, hist_cap=100), s_th.solve(initial, desired, want_gains=True,
                                                                                      hist_cap=100)
        assert np.array_equal(a["results"], b["results"])
        assert np.array_equal(a["traj"], b["traj"]) and np.array_equal(a["K"], b["K"])
        assert np.array_equal(a["cost_history"], b["cost_history"])
    # the reference's default problem through line_search() as well
    m2, o2 = problems.default_model(), problems.default_options(False)
    d2 = problems.default_desired_trajectory()
    s2 = make_solver(m2, o2)
    monkeypatch.setenv("QILQR_ROLLOUT", "thread")
    s3 = make_solver(m2, o2)
    monkeypatch.delenv("QILQR_ROLLOUT")
    r2, r3 = s2.solve(d2[None], d2, hist_cap=100), s3.solve(d2[None], d2, hist_cap=100)
    assert np.array_equal(r2["results"], r3["results"]) and np.array_equal(r2["traj"], r3["traj"])
    assert np.array_equal(r2["cost_history"], r3["cost_history"])


def test_tail_compaction_is_bit_identical(O, monkeypatch):
    """When few problems are left they are moved into a dense mini-batch (k_tail_gather / k_tail_scatter).
    Same kernels on the same values: every output equals the uncompacted solve bit for bit, for a shared and
    for per-problem desired trajectories, with the sequential and the parallel line search; and the compacted
    solve still matches the oracle."""
    import dataclasses

    from quadrotorilqr_b200 import problems

    model = problems.hover_model()
    monkeypatch.setenv("QILQR_HI_THRESHOLD", "24")  # the production threshold (2048) needs a batch above it
    for opts in (problems.default_options(False),
                 dataclasses.replace(problems.default_options(False), num_parallel_alphas=4, symmetrize_vxx=True)):
        s_on = make_solver(model, opts)
        monkeypatch.setenv("QILQR_TAIL_COMPACTION", "0")
        s_off = make_solver(model, opts)
        monkeypatch.delenv("QILQR_TAIL_COMPACTION")
        B, N = 150, 40
        desired, initial = hover_batch(s_on, B, N, seed=11)
        per_problem = np.repeat(desired[None], B, axis=0)
        per_problem[:, 20:, 1:4] += np.random.default_rng(2).uniform(-0.5, 0.5, (B, 1, 3))
        for des in (desired, per_problem):
            a = s_on.solve(initial, des, want_gains=True, hist_cap=100)
            b = s_off.solve(initial, des, want_gains=True, hist_cap=100)
            assert np.array_equal(a["results"], b["results"])
            assert a["results"]["backward_passes"].max() > a["results"]["backward_passes"].min() + 3  # a real tail
            for key in ("traj", "k", "K", "cost_history"):
                assert np.array_equal(a[key], b[key]), key
    s = make_solver(model, problems.default_options(False))
    cfg = oracle_config(O, model, problems.default_options(False))
    desired, initial = hover_batch(s, 120, 40, seed=4)
    check_solve_against_oracle(O, s, cfg, desired, initial)


def test_persistent_tail_matches_host_loop(O, monkeypatch):
    """The persistent tail kernel (one launch runs the rest of solve() once few problems are alive; default on,
    QILQR_PERSISTENT_TAIL=0 keeps the host-driven loop to the end) executes the same device code as the host-driven
    loop, and the library is compiled with -fmad=false: the results are BIT-IDENTICAL.  Block-diagonal and coupled
    Q, shared and per-problem desired trajectories, ragged tiles, hand-over thresholds below / at / above the batch
    size, line-search exhaustion; and it matches the oracle."""
    import dataclasses

    from quadrotorilqr_b200 import problems

    def pair(model, opts, threshold):
        monkeypatch.setenv("QILQR_PERSISTENT_TAIL", "1")
        monkeypatch.setenv("QILQR_PERSIST_THRESHOLD", str(threshold))
        on = make_solver(model, opts)
        monkeypatch.setenv("QILQR_PERSISTENT_TAIL", "0")
        off = make_solver(model, opts)
        monkeypatch.delenv("QILQR_PERSISTENT_TAIL")
        monkeypatch.delenv("QILQR_PERSIST_THRESHOLD")
        return on, off

    monkeypatch.setenv("QILQR_HI_THRESHOLD", "40")  # tail compaction into a mini-batch before the hand-over
    base = problems.hover_model()
    rng = np.random.default_rng(5)
    A = rng.uniform(-0.3, 0.3, (12, 12))
    coupled = dict(base, Q=base["Q"] + A @ A.T)
    for model, threshold in ((base, 21), (coupled, 21), (base, 1000), (base, 3)):
        s_on, s_off = pair(model, problems.default_options(False), threshold)
        B, N = 150, 40
        desired, initial = hover_batch(s_on, B, N, seed=11)
        per_problem = np.repeat(desired[None], B, axis=0)
        per_problem[:, 20:, 1:4] += rng.uniform(-0.5, 0.5, (B, 1, 3))
        for des in (desired, per_problem):
            a = s_on.solve(initial, des, want_gains=True, hist_cap=100)
            b = s_off.solve(initial, des, want_gains=True, hist_cap=100)
            assert a["results"]["backward_passes"].max() > a["results"]["backward_passes"].min() + 3  # a real tail
            assert np.array_equal(a["results"], b["results"])
            for key in ("traj", "k", "K", "cost_history"):
                assert np.array_equal(a[key], b[key]), key
    s_on, _ = pair(base, problems.default_options(False), 64)
    cfg = oracle_config(O, base, problems.default_options(False))
    desired, initial = hover_batch(s_on, 120, 40, seed=4)
    check_solve_against_oracle(O, s_on, cfg, desired, initial)
    tight = dataclasses.replace(problems.default_options(False))
    tight.line_search_params = dataclasses.replace(tight.line_search_params, max_iters=1, desired_reduction_frac=0.999)
    s_on, s_off = pair(base, tight, 64)
    a, b = s_on.solve(initial, desired), s_off.solve(initial, desired)
    assert np.array_equal(a["results"], b["results"]) and (a["results"]["status"] == 4).any()
    assert np.array_equal(a["traj"], b["traj"])
