#!/usr/bin/env python
"""One backward pass (k_linearise + Riccati sweep) through qilqr_backwards_pass_host for a given batch / horizon:
the launch that ncu captures when tuning the latency-bound Riccati kernels.  usage: riccati_latency.py B N [reps]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quadrotorilqr_b200 import BatchILQR, problems  # noqa: E402

B, N = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
m, opts = problems.hover_model(), problems.default_options(False)
dt = 0.1 if N <= 100 else 0.02
m = dict(m, dt_s=dt)
s = BatchILQR(m["mass_kg"], m["inertia"], m["arm_length_m"], m["torque_to_thrust_ratio_m"], m["g_mpss"], m["Q"], m["R"],
              m["dt_s"], opts)
d = problems.hover_desired_trajectory(N, dt)
x0 = problems.hover_initial_states(B, seed=1)
traj = s.forward_sim(problems.constant_state_trajectory(x0, N, dt, d[0, 14:18]), np.zeros((B, N, 4)), np.zeros((B, N, 48)))
for _ in range(reps):
    t0 = time.perf_counter()
    s.backwards_pass(traj, d)
    print("backwards_pass_host wall ms", 1e3 * (time.perf_counter() - t0))
