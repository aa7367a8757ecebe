"""Time forward_sim-sized launches of the rollout kernels (thread-per-problem vs role-specialised) through
solve() statistics: serial step time and rollout_ms for B = 1 and B = 65536."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from quadrotorilqr_b200 import BatchILQR, problems  # noqa: E402

m, opts = problems.hover_model(), problems.default_options(False)
N = 40
desired = problems.hover_desired_trajectory(N, m["dt_s"], m["mass_kg"], m["g_mpss"])
for mode in ("thread", "ws"):
    os.environ["QILQR_ROLLOUT"] = mode
    s = BatchILQR(m["mass_kg"], m["inertia"], m["arm_length_m"], m["torque_to_thrust_ratio_m"], m["g_mpss"], m["Q"],
                  m["R"], m["dt_s"], opts)
    for B in (1, 2048, 65536):
        x0 = problems.hover_initial_states(B, seed=0)
        seed = problems.constant_state_trajectory(x0, N, m["dt_s"], desired[0, 14:18])
        init = s.forward_sim(seed, np.zeros((B, N, 4)), np.zeros((B, N, 48)))
        s.set_profiling(True)
        for _ in range(2):
            r = s.solve(init, desired)
        st = s.last_solve_stats()
        s.set_profiling(False)
        t = time.time()
        for _ in range(3):
            r = s.solve(init, desired)
        dt = (time.time() - t) / 3
        print(mode, "B", B, "rollout_ms %.3f backward_ms %.3f" % (st["rollout_ms"], st["backward_ms"]),
              "rollout launches ~", st["solver_iterations"], "us/rollout-launch %.1f" % (1e3 * st["rollout_ms"] / max(1, st["solver_iterations"])),
              "host-API solve %.1f ms" % (dt * 1e3), flush=True)
