"""Numerical blow-ups: a NaN makes every Armijo comparison false, so the reference ends in its line-search
exception (ilqr.hh:182-193).  The batch solver reports that exit as QILQR_STATUS_NONFINITE (5) when the last
candidate cost was not finite, QILQR_STATUS_LINE_SEARCH_FAILED (4) otherwise; the other problems of the batch are
unaffected."""
import numpy as np
import pytest

from conftest import make_solver, oracle_config

pytestmark = pytest.mark.gpu


def test_nan_problem_is_flagged_and_does_not_disturb_its_neighbours(O):
    from quadrotorilqr_b200 import ConvergenceCriteria, ILQROptions, LineSearchParams, QilqrError, problems, protos
    from quadrotorilqr_b200.quadrotor_ilqr_binding import QuadrotorILQR

    model = problems.hover_model()
    opts = ILQROptions(LineSearchParams(0.5, 0.5, 6), ConvergenceCriteria(1e-12, 1e-12, 30.0))
    s = make_solver(model, opts)
    cfg = oracle_config(O, model, opts)
    B, N = 12, 20
    d = problems.hover_desired_trajectory(N)
    x0 = problems.hover_initial_states(B, seed=77)
    init = s.forward_sim(problems.constant_state_trajectory(x0, N, model["dt_s"], d[0, 14:18]), np.zeros((B, N, 4)),
                         np.zeros((B, N, 48)))
    clean = s.solve(init, d)
    bad = init.copy()
    bad[3, 5, 15] = np.nan      # one control of problem 3
    bad[8, 0, 2] = np.inf       # initial position of problem 8
    r = s.solve(bad, d)
    o = O.solve_batch(cfg, d, bad)
    assert np.array_equal(r["results"]["status"], o["status"])
    assert r["results"]["status"][3] == 5 and r["results"]["status"][8] == 5
    assert np.array_equal(r["results"]["backward_passes"], o["backward_passes"])
    assert np.array_equal(r["results"]["rollouts"], o["rollouts"])
    ok = np.ones(B, dtype=bool)
    ok[[3, 8]] = False
    assert np.array_equal(r["traj"][ok], clean["traj"][ok]) and np.array_equal(r["results"][ok], clean["results"][ok])
    # the single-problem drop-in raises like the reference (std::runtime_error -> RuntimeError)
    q = QuadrotorILQR(model["mass_kg"], model["inertia"], model["arm_length_m"], model["torque_to_thrust_ratio_m"],
                      model["g_mpss"], model["Q"], model["R"], protos.trajectory_to_proto(d), model["dt_s"],
                      protos.options_to_proto(opts))
    with pytest.raises(RuntimeError):
        q.solve(protos.trajectory_to_proto(bad[3]))
