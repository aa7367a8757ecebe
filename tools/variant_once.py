"""One solve of the hover batch with a given model variant (for ncu captures of the generic kernels)."""
import sys

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from quadrotorilqr_b200 import BatchILQR, problems  # noqa: E402

flags, B = int(sys.argv[1]), int(sys.argv[2])
m, opts = problems.hover_model(), problems.default_options(False)
N = 40
desired = problems.hover_desired_trajectory(N, m["dt_s"], m["mass_kg"], m["g_mpss"])
x0 = problems.hover_initial_states(B, seed=0)
seed = problems.constant_state_trajectory(x0, N, m["dt_s"], desired[0, 14:18])
s = BatchILQR(m["mass_kg"], m["inertia"], m["arm_length_m"], m["torque_to_thrust_ratio_m"], m["g_mpss"], m["Q"], m["R"],
              m["dt_s"], opts, model_flags=flags)
init = s.forward_sim(seed, np.zeros((B, N, 4)), np.zeros((B, N, 48)))
r = s.solve(init, desired)
print("converged", np.isin(r["results"]["status"], [1, 2]).mean(), "iters", r["results"]["backward_passes"].mean())
