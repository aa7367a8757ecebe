#!/usr/bin/env python
"""Every problem of a BASELINE config through the CUDA path against the CPU oracle: how many discrete
decisions (status / backward passes / rollouts) differ, where, and by how much the results differ.

  CONFIG=c3        batch 65536 hover problems, N = 40 (bench.py's workload; SEED picks the Philox stream)
  CONFIG=c3w       the waypoint variant: per-problem desired trajectories
  CONFIG=c4        N = 1000 figure-eight, batch 4096, symmetrised V_xx, 8 parallel step sizes
  CONFIG=libmfree  hover problems whose attitude stays the identity (position / linear-velocity offsets
                   along z only, so that no sin / cos / atan2 is ever evaluated: every rotation angle
                   stays below manif's small-angle threshold).  The STRICT build must reproduce the
                   oracle BIT FOR BIT on these -- that is what shows that the CUDA math library is the
                   only rounding difference left in that build.
  B, SEED          batch size and seed overrides
  QILQR_LIB        which build of the library (production: default; strict: libqilqr_b200_strict.so)

One JSON line on stdout.  Differing problems are enumerated with the quantity that decided them:
  rel_step   |cost_{k-1} - cost_k| / |cost_{k-1}| of the last completed iteration, in units of rtol
             (exit B of solve(), ilqr.hh:82-84, tests rel_step < 1)
and their final costs on both sides."""
import dataclasses
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O  # noqa: E402
from quadrotorilqr_b200 import BatchILQR, _capi, problems  # noqa: E402

CONFIG = os.environ.get("CONFIG", "c3")
SEED = int(os.environ.get("SEED", 2026))


def setup():
    m, opts = problems.hover_model(), problems.default_options(False)
    okw = {}
    if CONFIG in ("c3", "c3w", "libmfree"):
        B, N = int(os.environ.get("B", 65536 if CONFIG != "libmfree" else 4096)), 40
        d = problems.hover_desired_trajectory(N)
        x0 = problems.hover_initial_states(B, seed=SEED)
        if CONFIG == "libmfree":
            x0[:, 0:2] = 0.0
            x0[:, 3:6] = 0.0
            x0[:, 6] = 1.0
            x0[:, 7:9] = 0.0
            x0[:, 10:13] = 0.0
        desired = problems.waypoint_desired_trajectories(B, N) if CONFIG == "c3w" else d
        u0 = d[0, 14:18]
    elif CONFIG == "c4":
        B, N, dt_s = int(os.environ.get("B", 4096)), 1000, 0.02
        m = dict(m, dt_s=dt_s)
        opts = dataclasses.replace(opts, symmetrize_vxx=True, num_parallel_alphas=8)
        okw = dict(symmetrize_vxx=True)
        desired = d = problems.figure_eight_desired(N, dt_s)
        x0 = problems.figure_eight_initial_states(B, d)
        u0 = d[0, 14:18]
    else:
        raise SystemExit(f"unknown CONFIG {CONFIG}")
    return m, opts, okw, B, N, d, desired, x0, u0


def main():
    m, opts, okw, B, N, d, desired, x0, u0 = setup()
    s = BatchILQR(m["mass_kg"], m["inertia"], m["arm_length_m"], m["torque_to_thrust_ratio_m"], m["g_mpss"], m["Q"],
                  m["R"], m["dt_s"], opts)
    init = s.forward_sim(problems.constant_state_trajectory(x0, N, m["dt_s"], u0), np.zeros((B, N, 4)),
                         np.zeros((B, N, 48)))
    r = s.solve(init, desired, hist_cap=100)
    cfg = O.make_config(mass_kg=m["mass_kg"], inertia=m["inertia"], arm_length_m=m["arm_length_m"],
                        torque_to_thrust_ratio_m=m["torque_to_thrust_ratio_m"], g_mpss=m["g_mpss"], Q=m["Q"], R=m["R"],
                        dt_s=m["dt_s"], **okw)
    t0 = time.perf_counter()
    if CONFIG == "libmfree":  # the open-loop initial rollout must be bit-identical too
        oi = np.stack([O.forward_sim(cfg, d, problems.constant_state_trajectory(x0[b], N, m["dt_s"], u0)[0],
                                     np.zeros((N, 4)), np.zeros((N, 4, 12))) for b in range(min(B, 64))])
        init_identical = bool(np.array_equal(oi, init[: len(oi)]))
    o = O.solve_batch(cfg, desired, init, hist_cap=100)
    oracle_s = time.perf_counter() - t0
    res = r["results"]
    same = ((res["status"] == o["status"]) & (res["backward_passes"] == o["backward_passes"])
            & (res["rollouts"] == o["rollouts"]))
    scale = np.maximum(1.0, np.max(np.abs(o["traj"]), axis=(1, 2)))
    err = np.max(np.abs(r["traj"] - o["traj"]), axis=(1, 2)) / scale
    cerr = np.abs(res["final_cost"] - o["final_cost"]) / np.maximum(1.0, np.abs(o["final_cost"]))
    herr = np.max(np.abs(r["cost_history"] - o["cost_history"]) / np.maximum(1.0, np.abs(o["cost_history"])), axis=1)
    diff = np.where(~same)[0]
    rtol = opts.convergence_criteria.rtol

    def last_rel_step(hist, nd):
        if nd < 2:
            return None
        return float(abs(hist[nd - 2] - hist[nd - 1]) / abs(hist[nd - 2]) / rtol)

    cases = []
    for b in diff[:64]:
        nd_g, nd_o = int(res["num_debug"][b]), int(np.count_nonzero(o["cost_history"][b]))
        kind = ("line search (rollout count differs at the same iteration count)"
                if res["backward_passes"][b] == o["backward_passes"][b] and res["status"][b] == o["status"][b]
                else "convergence test at its threshold")
        cases.append({"problem": int(b), "kind": kind,
                      "gpu": [int(res["status"][b]), int(res["backward_passes"][b]), int(res["rollouts"][b])],
                      "oracle": [int(o["status"][b]), int(o["backward_passes"][b]), int(o["rollouts"][b])],
                      "rel_step_over_rtol_gpu": last_rel_step(r["cost_history"][b], nd_g),
                      "rel_step_over_rtol_oracle": last_rel_step(o["cost_history"][b], nd_o),
                      "final_cost_gpu": float(res["final_cost"][b]), "final_cost_oracle": float(o["final_cost"][b]),
                      "rel_final_cost_diff": float(cerr[b]), "rel_traj_err": float(err[b])})
    out = {
        "config": CONFIG, "build": _capi.lib().qilqr_build_info().decode(), "batch": B, "knots": N, "seed": SEED,
        "identical_decisions": int(same.sum()), "different_decisions": int(diff.size),
        "bit_identical_trajectories": int(np.sum(np.all(r["traj"] == o["traj"], axis=(1, 2)))),
        "bit_identical_cost_histories": int(np.sum(np.all(r["cost_history"] == o["cost_history"], axis=1))),
        "max_rel_traj_err_where_identical": float(err[same].max()) if same.any() else None,
        "max_rel_cost_history_err_where_identical": float(herr[same].max()) if same.any() else None,
        "max_rel_final_cost_err_where_identical": float(cerr[same].max()) if same.any() else None,
        "max_rel_traj_err_where_different": float(err[diff].max()) if diff.size else 0.0,
        "max_rel_final_cost_err_where_different": float(cerr[diff].max()) if diff.size else 0.0,
        "iteration_count_differences": sorted(set((res["backward_passes"][diff].astype(int)
                                                   - o["backward_passes"][diff].astype(int)).tolist())),
        "status_histogram_gpu": {int(k): int(v) for k, v in zip(*np.unique(res["status"], return_counts=True))},
        "iterations_mean": float(res["backward_passes"].mean()), "iterations_max": int(res["backward_passes"].max()),
        "oracle_seconds": oracle_s, "oracle_threads": O.hardware_threads(),
        "differing_problems": cases,
    }
    if CONFIG == "libmfree":
        out["initial_rollout_bit_identical"] = init_identical
        out["max_abs_quaternion_vector_part"] = float(np.max(np.abs(r["traj"][:, :, 4:7])))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
