#!/usr/bin/env python
"""Timings of the other BASELINE.json configs (the driver's headline is bench.py = config 3):
  c1  the reference's default problem, batch 1 (latency)
  c2  batch 1024 random hover problems
  c3w batch 65536 WAYPOINT problems: every problem tracks its own 4-segment piecewise-constant position
      path (desired_count = batch), one batch at a time (bench.py measures the hover variant, pipelined)
  c4  N = 1000 figure-eight tracking, batch 4096, 8 parallel alphas, symmetrised V_xx
  c5  receding-horizon MPC, 16384 quadrotors x 500 closed-loop steps, warm-started re-solves
Prints one JSON line per config; run under gpurun and keep the output under profiles/."""
import argparse
import dataclasses
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quadrotorilqr_b200 import BatchILQR, RESULT_DTYPE, problems  # noqa: E402


def mk(model, opts):
    return BatchILQR(model["mass_kg"], model["inertia"], model["arm_length_m"], model["torque_to_thrust_ratio_m"],
                     model["g_mpss"], model["Q"], model["R"], model["dt_s"], opts)


def device_problem(s, x0, desired, N, dev):
    B = x0.shape[0]
    x0_soa = torch.from_numpy(np.ascontiguousarray(x0.T)).to(dev)
    init = torch.empty((N, 17, B), dtype=torch.float64, device=dev)
    des_aos = torch.from_numpy(desired[None].copy()).to(dev)
    des = torch.empty((N, 17, 1), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    s.pack_trajectory_device(des_aos, des)
    s.rollout_constant_control_device(x0_soa, desired[0, 14:18], init)
    torch.cuda.synchronize()
    return x0_soa, init, des


def time_solves(s, init, des, reps, warm=1):
    B = init.shape[2]
    work = torch.empty_like(init)
    res = torch.zeros(B * 24, dtype=torch.uint8, device=init.device)
    out = None
    for r in range(warm + reps):
        if r == warm:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
        work.copy_(init)
        torch.cuda.synchronize()
        s.solve_device(work, des, results=res)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    out = np.frombuffer(res.cpu().numpy().tobytes(), dtype=RESULT_DTYPE)
    return dt, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="c1,c2,c3w,c4,c5")
    ap.add_argument("--mpc-steps", type=int, default=500)
    ap.add_argument("--mpc-batch", type=int, default=16384)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    todo = args.configs.split(",")
    if "c1" in todo:
        m, o = problems.default_model(), problems.default_options(True)
        s = mk(m, o)
        d = problems.default_desired_trajectory()
        s.solve(d[None], d)
        t0 = time.perf_counter()
        for _ in range(5):
            r = s.solve(d[None], d, want_debug=True, hist_cap=100)
        dt = (time.perf_counter() - t0) / 5
        print(json.dumps({"config": "c1 reference default problem, B=1, N=40, populate_debug, host API",
                          "ms_per_solve": 1e3 * dt, "backward_passes": int(r["results"]["backward_passes"][0]),
                          "final_cost": float(r["results"]["final_cost"][0]),
                          "us_per_iteration": 1e6 * dt / int(r["results"]["backward_passes"][0])}), flush=True)
    if "c2" in todo:
        m, o = problems.hover_model(), problems.default_options(False)
        s = mk(m, o)
        N, B = 40, 1024
        d = problems.hover_desired_trajectory(N)
        _, init, des = device_problem(s, problems.hover_initial_states(B, seed=2026), d, N, dev)
        dt, res = time_solves(s, init, des, 5)
        conv = int(np.sum((res["status"] == 1) | (res["status"] == 2)))
        print(json.dumps({"config": "c2 batch 1024 hover, N=40, device-resident", "ms_per_batch": 1e3 * dt,
                          "solves_per_s": conv / dt, "converged": conv, "iterations_per_solve":
                          float(res["backward_passes"].mean()), "max_iterations": int(res["backward_passes"].max())}),
              flush=True)
    if "c3w" in todo:
        m, o = problems.hover_model(), problems.default_options(False)
        s = mk(m, o)
        N, B = 40, 65536
        base = problems.hover_desired_trajectory(N)
        desired = problems.waypoint_desired_trajectories(B, N)  # waypoints U[-1,1]^3
        _, init, _ = device_problem(s, problems.hover_initial_states(B, seed=2026), base, N, dev)
        des = torch.empty((N, 17, B), dtype=torch.float64, device=dev)
        s.pack_trajectory_device(torch.from_numpy(desired).to(dev), des)
        torch.cuda.synchronize()
        dt, res = time_solves(s, init, des, 3)
        conv = int(np.sum((res["status"] == 1) | (res["status"] == 2)))
        print(json.dumps({"config": "c3w batch 65536 waypoint problems (per-problem desired trajectories), N=40, "
                                    "device-resident, one batch at a time", "ms_per_batch": 1e3 * dt,
                          "solves_per_s": conv / dt, "converged": conv, "converged_fraction": conv / B,
                          "iterations_per_solve": float(res["backward_passes"].mean()),
                          "max_iters_hit": int(np.sum(res["status"] == 3)),
                          "line_search_failures": int(np.sum(res["status"] >= 4))}), flush=True)
    if "c4" in todo:
        N, dt_s, B = 1000, 0.02, 4096
        m = dict(problems.hover_model(), dt_s=dt_s)
        o = dataclasses.replace(problems.default_options(False), symmetrize_vxx=True, num_parallel_alphas=8)
        s = mk(m, o)
        d = problems.figure_eight_desired(N, dt_s)
        x0 = problems.figure_eight_initial_states(B, d)
        _, init, des = device_problem(s, x0, d, N, dev)
        dt, res = time_solves(s, init, des, 2)
        conv = int(np.sum((res["status"] == 1) | (res["status"] == 2)))
        its = int(res["backward_passes"].sum())
        print(json.dumps({"config": "c4 N=1000 figure-eight, B=4096, 8 parallel alphas, symmetrize_vxx",
                          "ms_per_batch": 1e3 * dt, "solves_per_s": conv / dt, "converged": conv,
                          "iterations_per_solve": its / B, "max_iterations": int(res["backward_passes"].max()),
                          "us_per_iteration_amortised": 1e6 * dt / its,
                          "problem_knot_iterations_per_s": its * N / dt}), flush=True)
    if "c5" in todo:
        m, o = problems.hover_model(), problems.default_options(False)
        s = mk(m, o)
        N, B, T = 40, args.mpc_batch, args.mpc_steps
        d = problems.hover_desired_trajectory(N)
        x0 = problems.hover_initial_states(B, seed=5)
        plant, traj, des = device_problem(s, x0, d, N, dev)
        rng = np.random.Generator(np.random.Philox(key=55))
        dist = torch.from_numpy(np.ascontiguousarray(rng.uniform(-0.01, 0.01, (6, B)))).to(dev)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        tot = s.mpc_run_device(T, traj, des, plant, disturbance=dist)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        final = plant.cpu().numpy()
        print(json.dumps({"config": f"c5 MPC {B} quadrotors x {T} closed-loop steps, N=40, warm-started, constant "
                                    "body-velocity disturbance U[-0.01,0.01]^6 per step",
                          "seconds": dt, "resolves_per_s": tot["resolves"] / dt, "iterations_per_resolve":
                          tot["backward_passes"] / tot["resolves"], "not_converged": tot["not_converged"],
                          "ms_per_closed_loop_step": 1e3 * dt / T,
                          "final_position_error_max_m": float(np.abs(final[0:3]).max())}), flush=True)


if __name__ == "__main__":
    main()
