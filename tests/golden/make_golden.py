#!/usr/bin/env python
"""Regenerates tests/golden/*.npz from the CPU oracle (run from the repo root:
`python tests/golden/make_golden.py`).  The literal reference cannot be run here (Eigen/manif
absent), so these vectors pin the ORACLE's outputs at the time it passed the
reference's own known-answer tests; they guard against drift of the oracle and give the GPU
tests a fixture that does not depend on the oracle binary of the day."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402
from quadrotorilqr_b200 import problems  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def cfg_for(m, opts):
    return O.make_config(mass_kg=m["mass_kg"], inertia=m["inertia"], arm_length_m=m["arm_length_m"],
                         torque_to_thrust_ratio_m=m["torque_to_thrust_ratio_m"], g_mpss=m["g_mpss"], Q=m["Q"],
                         R=m["R"], dt_s=m["dt_s"], populate_debug=opts.populate_debug,
                         symmetrize_vxx=opts.symmetrize_vxx)


def main():
    # C1: the reference's default problem (quadrotor_ilqr.py:256-306)
    m, opts = problems.default_model(), problems.default_options(True)
    desired = problems.default_desired_trajectory()
    r = O.solve(cfg_for(m, opts), desired, desired)
    np.savez_compressed(os.path.join(HERE, "c1_default_problem.npz"), desired=desired, traj=r["traj"],
                        cost_history=r["cost_history"], step_history=r["step_history"], k=r["k"], K=r["K"],
                        status=r["status"], backward_passes=r["backward_passes"], rollouts=r["rollouts"],
                        final_cost=r["final_cost"], debug_first=r["debug"][0], debug_last=r["debug"][-1])
    # C2 (prefix): 16 hover problems of the Philox stream, seed 2026
    m, opts = problems.hover_model(), problems.default_options(False)
    cfg = cfg_for(m, opts)
    N = 40
    des = problems.hover_desired_trajectory(N, m["dt_s"], m["mass_kg"], m["g_mpss"])
    x0 = problems.hover_initial_states(16, seed=2026)
    init = np.stack([O.forward_sim(cfg, des, problems.constant_state_trajectory(x, N, m["dt_s"], des[0, 14:18])[0],
                                   np.zeros((N, 4)), np.zeros((N, 4, 12))) for x in x0])
    b = O.solve_batch(cfg, des, init, want_gains=True, hist_cap=100)
    np.savez_compressed(os.path.join(HERE, "c2_hover_prefix16.npz"), x0=x0, desired=des, initial=init,
                        traj=b["traj"], k=b["k"], K=b["K"], cost_history=b["cost_history"], status=b["status"],
                        backward_passes=b["backward_passes"], rollouts=b["rollouts"], final_cost=b["final_cost"])
    # one backward pass + rollout on the first hover problem (gains are the most sensitive output)
    k, K, a, c = O.backwards_pass(cfg, des, init[0])
    new = O.forward_sim(cfg, des, init[0], k, K, 1.0)
    np.savez_compressed(os.path.join(HERE, "c2_single_iteration.npz"), desired=des, traj=init[0], k=k, K=K, QuTk=a,
                        kTQuuk=c, rolled=new, cost0=O.cost_trajectory(cfg, des, init[0]),
                        cost1=O.cost_trajectory(cfg, des, new))
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
