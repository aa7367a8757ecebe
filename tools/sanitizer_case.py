#!/usr/bin/env python
"""Small solves exercising every kernel, for compute-sanitizer (memcheck / racecheck / initcheck):
   compute-sanitizer --tool racecheck python tools/sanitizer_case.py"""
import dataclasses
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quadrotorilqr_b200 import BatchILQR, problems  # noqa: E402


def run(opts, B, N, env=None, model_flags=0):
    for k, v in (env or {}).items():
        os.environ[k] = v
    m = problems.hover_model()
    s = BatchILQR(m["mass_kg"], m["inertia"], m["arm_length_m"], m["torque_to_thrust_ratio_m"], m["g_mpss"], m["Q"],
                  m["R"], m["dt_s"], opts, model_flags=model_flags)
    for k in (env or {}):
        del os.environ[k]
    d = problems.hover_desired_trajectory(N)
    x0 = problems.hover_initial_states(B, seed=3)
    seed = problems.constant_state_trajectory(x0, N, m["dt_s"], d[0, 14:18])
    init = s.forward_sim(seed, np.zeros((B, N, 4)), np.zeros((B, N, 48)))
    r = s.solve(init, d, want_gains=True, hist_cap=20, want_debug=opts.populate_debug)
    return r["results"]


if __name__ == "__main__":
    base = dataclasses.replace(problems.default_options(False))
    base.convergence_criteria.max_iters = 6.0
    print(run(base, 21, 9)["backward_passes"])                                        # split backward (default)
    print(run(base, 13, 6, {"QILQR_BACKWARD": "fused"})["backward_passes"])           # fused quad kernel
    print(run(dataclasses.replace(base, num_parallel_alphas=3, symmetrize_vxx=True, populate_debug=True), 10, 7)["rollouts"])
    print(run(base, 5, 5, {"QILQR_BACKWARD": "t1"})["backward_passes"])               # thread-per-problem kernel
    print(run(base, 40, 7, {"QILQR_ROLLOUT": "thread"})["rollouts"])                  # thread-per-problem rollout only
    print(run(base, 70, 7)["rollouts"])                                               # role-specialised rollout (3 CTAs, ragged)
    print(run(dataclasses.replace(base, symmetrize_vxx=True), 11, 6, model_flags=3)["backward_passes"])  # dense kernels, RK4 + Coriolis
    print(run(base, 9, 5, model_flags=4)["backward_passes"])                          # dense kernels, reference model
    # round 2: the 16-lane Riccati kernel is what the small cases above use; force the 4-lane one as well, the
    # persistent tail kernel, tail compaction at a small threshold, sampled debug rings, controls-in entry point
    print(run(base, 21, 9, {"QILQR_G16_THRESHOLD": "0"})["backward_passes"])
    print(run(base, 37, 8, {"QILQR_PERSISTENT_TAIL": "1", "QILQR_PERSIST_THRESHOLD": "12", "QILQR_HI_THRESHOLD": "20"})["backward_passes"])
    m = problems.hover_model()
    s = BatchILQR(m["mass_kg"], m["inertia"], m["arm_length_m"], m["torque_to_thrust_ratio_m"], m["g_mpss"], m["Q"], m["R"],
                  m["dt_s"], dataclasses.replace(base, populate_debug=True))
    d = problems.hover_desired_trajectory(7)
    x0 = problems.hover_initial_states(19, seed=4)
    s.set_debug_sampling(np.array([3, 17, 0], dtype=np.int32), every=2, ring=2)
    r = s.solve_from_controls(x0, np.tile(d[0, 14:18], (7, 1)), d, want_controls=True)
    print(r["results"]["backward_passes"], s.read_debug_samples(7)["counts"], s.last_cost_history(count=19).shape)
