"""Behaviour of the reference algorithm that SURVEY.md records (section 0 "facts", Appendix C) -- reproduced with the
oracle on the CPU.  The survey measured these with an independent numpy scratch implementation; they are sanity
anchors for the oracle's control flow (iteration 0, both convergence exits, Armijo, line-search failure), not
reference test vectors."""
import numpy as np


def cfg_for(O, m, **kw):
    return O.make_config(mass_kg=m["mass_kg"], inertia=m["inertia"], arm_length_m=m["arm_length_m"],
                         torque_to_thrust_ratio_m=m["torque_to_thrust_ratio_m"], g_mpss=m["g_mpss"], Q=m["Q"],
                         R=m["R"], dt_s=m["dt_s"], **kw)


def test_default_problem_convergence_history(O):
    """App. C row 2: cost 0 -> 304,906 after the unconditional first step -> 22,556.50259198; 76 completed
    iterations, exit A at i = 76 (77 backward passes, 77 rollouts), one backtrack (alpha = 0.5 at i = 4),
    linear convergence at ~0.73 per iteration."""
    from quadrotorilqr_b200 import problems

    m, d = problems.default_model(), problems.default_desired_trajectory()
    r = O.solve(cfg_for(O, m), d, d)
    assert O.cost_trajectory(cfg_for(O, m), d, d) == 0.0
    assert (r["status"], r["backward_passes"], r["rollouts"], r["num_debug"]) == (1, 77, 77, 76)
    assert abs(r["cost_history"][0] - 304906) < 1.0 and abs(r["final_cost"] - 22556.50259198) < 1e-6
    steps = r["step_history"]
    assert steps[4] == 0.5 and np.all(np.delete(steps, 4) == 1.0)
    excess = r["cost_history"] - r["final_cost"]
    ratios = excess[41:61] / excess[40:60]
    assert np.all((ratios > 0.6) & (ratios < 0.85))


def test_long_horizon_needs_symmetrisation(O):
    """Fact 5 / App. C last rows: at N = 1000, dt = 0.02 the reference recursion (V_xx never symmetrised) is unstable
    -- the first step multiplies the cost by 10^6 and the solve never converges (the survey's instance ended in
    the line-search throw; this one crawls until max_iters) -- while with V_xx <- (V_xx + V_xx^T)/2 the same
    problem converges in 7-8 iterations without backtracking."""
    from quadrotorilqr_b200 import problems

    N, dt = 1000, 0.02
    m = dict(problems.hover_model(), dt_s=dt)
    d = problems.figure_eight_desired(N, dt)
    x0 = d[0, 1:14].copy()
    x0[0:3] += [0.2, -0.1, 0.15]
    seed = problems.constant_state_trajectory(x0[None], N, dt, d[0, 14:18])[0]
    init = O.forward_sim(cfg_for(O, m), d, seed, np.zeros((N, 4)), np.zeros((N, 4, 12)))
    plain = O.solve(cfg_for(O, m), d, init)
    sym = O.solve(cfg_for(O, m, symmetrize_vxx=True), d, init)
    assert plain["status"] not in (1, 2)
    assert plain["cost_history"][0] > 1e5 * sym["cost_history"][0] and plain["final_cost"] > 1e3 * sym["final_cost"]
    assert sym["status"] in (1, 2) and 5 <= sym["backward_passes"] <= 9
    assert np.all(sym["step_history"] == 1.0)
    assert 1000.0 < sym["final_cost"] < 1700.0


def test_hover_families(O):
    """App. C rows 3-5: small perturbations converge in 6-8 iterations, the medium family (the BASELINE workload)
    always converges within ~5-22, large perturbations with the default weights mostly run into max_iters."""
    from quadrotorilqr_b200 import problems

    m = problems.hover_model()
    d = problems.hover_desired_trajectory(40)
    cfg = cfg_for(O, m)

    def solve_family(n, **kw):
        x0 = problems.hover_initial_states(n, seed=7, **kw)
        seed = problems.constant_state_trajectory(x0, 40, m["dt_s"], d[0, 14:18])
        init = np.stack([O.forward_sim(cfg, d, s, np.zeros((40, 4)), np.zeros((40, 4, 12))) for s in seed])
        return O.solve_batch(cfg, d, init)

    small = solve_family(16, pos=0.5, theta_max=0.2, vel=0.1)
    assert np.all(np.isin(small["status"], [1, 2])) and small["backward_passes"].min() >= 5 and small["backward_passes"].max() <= 10
    medium = solve_family(32)
    assert np.all(np.isin(medium["status"], [1, 2])) and 4 <= medium["backward_passes"].min() and medium["backward_passes"].max() <= 25
    large = solve_family(16, pos=2.0, theta_max=2.0, vel=1.0)
    assert np.mean(large["status"] == 3) >= 0.5
