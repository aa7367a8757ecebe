#!/usr/bin/env python
"""All 65536 problems of the headline workload against the CPU oracle: how many discrete decisions
(status / iteration count / rollout count) differ, and by how much the results differ when they do."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O  # noqa: E402
from quadrotorilqr_b200 import BatchILQR, problems  # noqa: E402

B, N = int(os.environ.get("B", 65536)), 40
m, opts = problems.hover_model(), problems.default_options(False)
s = BatchILQR(m["mass_kg"], m["inertia"], m["arm_length_m"], m["torque_to_thrust_ratio_m"], m["g_mpss"], m["Q"], m["R"],
              m["dt_s"], opts)
d = problems.hover_desired_trajectory(N)
x0 = problems.hover_initial_states(B, seed=int(os.environ.get("SEED", 2026)))
init = s.forward_sim(problems.constant_state_trajectory(x0, N, m["dt_s"], d[0, 14:18]), np.zeros((B, N, 4)),
                     np.zeros((B, N, 48)))
r = s.solve(init, d, hist_cap=100)
cfg = O.make_config(mass_kg=m["mass_kg"], inertia=m["inertia"], arm_length_m=m["arm_length_m"],
                    torque_to_thrust_ratio_m=m["torque_to_thrust_ratio_m"], g_mpss=m["g_mpss"], Q=m["Q"], R=m["R"],
                    dt_s=m["dt_s"])
o = O.solve_batch(cfg, d, init, hist_cap=100)
res = r["results"]
same = (res["status"] == o["status"]) & (res["backward_passes"] == o["backward_passes"]) & (res["rollouts"] == o["rollouts"])
scale = np.maximum(1.0, np.max(np.abs(o["traj"]), axis=(1, 2)))
err = np.max(np.abs(r["traj"] - o["traj"]), axis=(1, 2)) / scale
cerr = np.abs(res["final_cost"] - o["final_cost"]) / np.maximum(1.0, np.abs(o["final_cost"]))
diff = np.where(~same)[0]
out = {
    "batch": B, "seed": int(os.environ.get("SEED", 2026)), "identical_decisions": int(same.sum()), "different_decisions": int(diff.size),
    "max_rel_traj_err_where_identical": float(err[same].max()), "max_rel_cost_err_where_identical": float(cerr[same].max()),
    "max_rel_traj_err_where_different": float(err[diff].max()) if diff.size else 0.0,
    "max_rel_cost_err_where_different": float(cerr[diff].max()) if diff.size else 0.0,
    "iteration_count_differences": np.unique(res["backward_passes"][diff].astype(int) - o["backward_passes"][diff].astype(int),
                                             return_counts=True)[0].tolist() if diff.size else [],
    "examples": [{"problem": int(b), "gpu": [int(res["status"][b]), int(res["backward_passes"][b]), int(res["rollouts"][b])],
                  "oracle": [int(o["status"][b]), int(o["backward_passes"][b]), int(o["rollouts"][b])],
                  "rel_traj_err": float(err[b]),
                  "last_rel_cost_step_gpu": float(abs(r["cost_history"][b][max(0, res["num_debug"][b] - 1)] -
                                                      r["cost_history"][b][max(0, res["num_debug"][b] - 2)]) /
                                                  abs(r["cost_history"][b][max(0, res["num_debug"][b] - 2)]))}
                 for b in diff[:12]],
}
print(json.dumps(out))
