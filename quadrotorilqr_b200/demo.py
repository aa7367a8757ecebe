"""The reference demo (``src/quadrotor_ilqr.py:256-323``, ``bazel run //src:quadrotor_ilqr``) on the GPU path,
without the matplotlib/STL visualisation: builds the same protos, calls the drop-in ``QuadrotorILQR`` and
prints / returns what the reference script plots.

    python -m quadrotorilqr_b200.demo [--plot_iters]
"""
from __future__ import annotations

import argparse

import numpy as np

from . import problems, protos
from .quadrotor_ilqr_binding import QuadrotorILQR


def extract_traj_array(trajectory) -> np.ndarray:
    """``extract_traj_array`` of the reference script (quadrotor_ilqr.py:40-65): [N, 18] in IDX order."""
    return problems.to_idx_layout(protos.trajectory_from_proto(trajectory))


def main(plot_iters: bool = False, verbose: bool = True):
    traj, opts = protos.trajectory_pb2, protos.ilqr_options_pb2
    dt_s = 0.1
    desired_traj = protos.trajectory_to_proto(problems.default_desired_trajectory())  # quadrotor_ilqr.py:257-270
    options = opts.ILQROptions(  # quadrotor_ilqr.py:272-284
        line_search_params=opts.LineSearchParams(step_update=0.5, desired_reduction_frac=0.5, max_iters=100),
        convergence_criteria=opts.ConvergenceCriteria(rtol=1e-12, atol=1e-12, max_iters=100),
        populate_debug=True,
    )
    m = problems.default_model()  # quadrotor_ilqr.py:286-292
    ilqr = QuadrotorILQR(m["mass_kg"], m["inertia"], m["arm_length_m"], m["torque_to_thrust_ratio_m"], m["g_mpss"],
                         m["Q"], m["R"], desired_traj, dt_s, options)
    opt_traj, debug = ilqr.solve(desired_traj)  # quadrotor_ilqr.py:306
    traj_dict = {"desired": desired_traj, "optimized": opt_traj}
    if plot_iters:
        for i, iter_debug in enumerate(debug.iter_debugs):
            traj_dict[f"iter {i}"] = iter_debug.trajectory
    costs = [d.cost for d in debug.iter_debugs]
    if verbose:
        arr = extract_traj_array(opt_traj)
        print(f"iterations: {len(costs)}   final cost: {costs[-1]:.6f}")
        print("cost per iteration:", " ".join(f"{c:.4g}" for c in costs[:8]), "...")
        print(f"final position: {arr[-1, problems.IDX.translation_x_m:problems.IDX.translation_z_m + 1]}")
    return traj_dict, costs


if __name__ == "__main__":
    ap = argparse.ArgumentParser(description="Run the Quadrotor iLQR Trajectory Generator (GPU path).")
    ap.add_argument("--plot_iters", action="store_true", help="also return the intermediate trajectories")
    a = ap.parse_args()
    main(a.plot_iters)
