"""Drop-in for the reference's pybind module ``src.quadrotor_ilqr_binding``
(``src/quadrotor_ilqr_binding.cc:20-49``): the same class name, constructor arguments and
``solve`` signature, taking and returning the same protobuf messages -- backed by the CUDA library.

    from quadrotorilqr_b200.quadrotor_ilqr_binding import QuadrotorILQR
    ilqr = QuadrotorILQR(mass_kg, inertia, arm_length_m, torque_to_thrust_ratio_m, g_mpss, Q, R,
                         desired_traj, dt_s, options)          # protos as in quadrotor_ilqr.py:294-305
    opt_traj, debug = ilqr.solve(initial_traj)                   # (QuadrotorTrajectory, QuadrotorILQRDebug)
"""
from __future__ import annotations

import numpy as np

from . import _capi, protos
from .solver import BatchILQR


class QuadrotorILQR:
    def __init__(self, mass_kg, inertia, arm_length_m, torque_to_thrust_ratio_m, g_mpss, Q, R, desired_traj, dt_s,
                 options, device: int = 0):
        self._options = protos.options_from_proto(options)
        self._desired = protos.trajectory_from_proto(desired_traj)
        self._solver = BatchILQR(mass_kg, np.asarray(inertia, dtype=np.float64), arm_length_m,
                                 torque_to_thrust_ratio_m, g_mpss, np.asarray(Q, dtype=np.float64),
                                 np.asarray(R, dtype=np.float64), dt_s, self._options, device=device)

    def solve(self, initial_traj):
        init = protos.trajectory_from_proto(initial_traj)
        n = init.shape[0]
        if n > self._desired.shape[0]:
            raise IndexError("vector::_M_range_check")  # std::out_of_range, cost.hh:39-40
        if n == 0:
            raise ValueError("empty trajectory")  # undefined behaviour in the reference (ilqr.hh:156)
        # `for (int i = 0; i < max_iters; ++i)` with max_iters a double (ilqr.hh:58): a NaN or non-positive bound runs
        # no iteration; the history needs one entry per completed iteration
        mi = self._options.convergence_criteria.max_iters
        cap = int(min(np.ceil(mi), 1 << 20)) + 1 if (mi == mi and mi > 0) else 1
        r = self._solver.solve(init[None], self._desired[:n], hist_cap=cap,
                               want_debug=self._options.populate_debug and cap > 1)
        res = r["results"][0]
        if res["status"] in (_capi.STATUS_LINE_SEARCH_FAILED, _capi.STATUS_NONFINITE):  # std::runtime_error, ilqr.hh:191-193
            why = " (the candidate costs were not finite)" if res["status"] == _capi.STATUS_NONFINITE else ""
            raise RuntimeError("Reached maximum number of line search iterations, "
                               f"{self._options.line_search_params.max_iters}{why}\n")
        nd = int(res["num_debug"]) if (self._options.populate_debug and r["debug"] is not None) else 0
        debug = protos.debug_to_proto(r["debug"][0][:nd] if nd else [], r["cost_history"][0][:nd] if nd else [])
        return protos.trajectory_to_proto(r["traj"][0]), debug
