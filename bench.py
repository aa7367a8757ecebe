#!/usr/bin/env python
"""bench.py -- throughput of the batched quadrotor iLQR hot path on B200.

Metric (BASELINE.json): converged iLQR solves/sec at batch 65536 (config
"batch 65536 hover problems", N=40 knots, FP64, reference default options), plus the
amortised microseconds per iLQR iteration.

One "step" = one full batched solve (ILQR::solve, ilqr.hh:53-87, for every problem of the
batch).  `value` times the device-resident entry point (inputs already in HBM, the timed
region includes restoring the initial trajectories with a device-to-device copy);
`e2e` times the reference-facing host entry point qilqr_solve_host with pinned host
buffers (H2D + AoS->SoA + solve + SoA->AoS + D2H inside the timed region).

N > 1 (torchrun): one process per GPU, each rank solves its own `batch` problems
(weak scaling, no data-path collective); NCCL only reduces the convergence statistics
and the max-over-ranks time.

`--impl reference` times the CPU oracle (the restated reference algorithm; the literal
reference cannot be built here) on all host cores over a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "converged iLQR solves/sec (batch 65536 per GPU, N=40 hover problems, FP64)"
UNIT = "solves/s"
N_KNOTS = 40


def workload_string(batch, seed):
    """config.workload -- the same string on the GPU arm and on the reference arm."""
    return (f"batch {batch} hover problems/GPU, N={N_KNOTS}, dt=0.1, reference default options (rtol=atol=1e-12, "
            "max_iters=100, line search 0.5/0.5/100), torque_to_thrust_ratio=0.1, x0: pos U[-1,1]^3, angle U[0,0.5] rad, "
            f"vel U[-0.25,0.25]^6 (Philox seed {seed})")
# dense as-written FLOPs per knot, counted by the oracle's FLOP-counting scalar
# (oracle.count_flops on a converged hover trajectory)
F_BWD, F_ROLL, F_COST = 30231.3, 721.2, 581.2
# the same counts for the model variants of --model-variant (dense chain rule through the RK4 stages as
# the oracle's QuadrotorModelVariant writes it); 4 = the reference model on the model-agnostic kernels
VARIANT_FLOPS = {1: (72487.0, 2118.1, 585.0), 2: (30266.0, 745.2, 585.0), 3: (72559.0, 2190.1, 585.0),
                 4: (F_BWD, F_ROLL, F_COST)}


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.stop_flag, self.thread = index, [], False, None

    def _run(self):
        # one streaming nvidia-smi (-lms) instead of a process per sample: ~10 samples per second
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                parts = [p.strip() for p in line.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append(parts)
                if self.stop_flag:
                    break
        except Exception:
            pass

    def start(self):
        self.proc = None
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()
        # nvidia-smi takes 0.2 - 2 s to come up (longer on a box with 8 GPUs): wait for its first line, so that it is
        # already streaming when the (sub-second) timed region starts; only what arrives after this point counts
        t0 = time.perf_counter()
        while not self.samples and time.perf_counter() - t0 < 10.0:
            time.sleep(0.02)
        self.mark = len(self.samples)

    def stop(self):
        self.stop_flag = True
        if getattr(self, "proc", None) is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass
        if self.thread:
            self.thread.join(timeout=6)
        samples = self.samples[getattr(self, "mark", 0):] or self.samples[-1:]
        sm = [float(s[0]) for s in samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in samples if s[1].replace(".", "").isdigit()]
        reasons = set()
        for s in samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "sm_min_mhz": min(sm) if sm else None, "reasons": sorted(reasons), "samples": len(samples)}


def oracle_config(O, m, opts):
    return O.make_config(mass_kg=m["mass_kg"], inertia=m["inertia"], arm_length_m=m["arm_length_m"],
                         torque_to_thrust_ratio_m=m["torque_to_thrust_ratio_m"], g_mpss=m["g_mpss"], Q=m["Q"],
                         R=m["R"], dt_s=m["dt_s"], step_update=opts.line_search_params.step_update,
                         desired_reduction_frac=opts.line_search_params.desired_reduction_frac,
                         ls_max_iters=opts.line_search_params.max_iters, rtol=opts.convergence_criteria.rtol,
                         atol=opts.convergence_criteria.atol, max_iters=opts.convergence_criteria.max_iters)


def cpu_initial_trajectories(O, cfg, problems, m, desired, x0):
    """Open-loop hover rollout of each x0 (the initial trajectory of configs C2/C3) on the CPU."""
    B = x0.shape[0]
    out = np.zeros((B, N_KNOTS, 18))
    zk, zK = np.zeros((N_KNOTS, 4)), np.zeros((N_KNOTS, 4, 12))
    for b in range(B):
        seed = problems.constant_state_trajectory(x0[b], N_KNOTS, m["dt_s"], desired[0, 14:18])[0]
        out[b] = O.forward_sim(cfg, desired, seed, zk, zK)
    return out


def run_cpu_sample(O, cfg, desired, initial, threads):
    t0 = time.perf_counter()
    r = O.solve_batch(cfg, desired, initial, nthreads=threads)
    dt = time.perf_counter() - t0
    conv = int(np.sum((r["status"] == 1) | (r["status"] == 2)))
    return dt, conv, int(r["backward_passes"].sum())


def reference_arm(args, rank, world):
    """The reference's CPU algorithm (oracle port) on all host cores, bounded sample per step."""
    if rank != 0:
        return 0
    import oracle as O
    from quadrotorilqr_b200 import problems

    O.build()
    m, opts = problems.hover_model(), problems.default_options(False)
    cfg = oracle_config(O, m, opts)
    cores = O.hardware_threads()
    sample = int(min(args.batch, max(64, args.cpu_sample_per_core * cores)))
    desired = problems.hover_desired_trajectory(N_KNOTS, m["dt_s"], m["mass_kg"], m["g_mpss"])
    x0 = problems.hover_initial_states(sample, seed=args.seed)
    initial = cpu_initial_trajectories(O, cfg, problems, m, desired, x0)
    for _ in range(args.warmup):
        run_cpu_sample(O, cfg, desired, initial[: max(8, sample // 8)], cores)
    tot_t, tot_conv, tot_it = 0.0, 0, 0
    for _ in range(args.steps):
        dt, conv, its = run_cpu_sample(O, cfg, desired, initial, cores)
        tot_t += dt
        tot_conv += conv
        tot_it += its
    value = tot_conv / tot_t
    sample_desc = (f"first {sample} problems of the same Philox stream (seed {args.seed}), {args.steps} step(s), "
                   f"{cores} threads, g++ -O2 -ffp-contract=off")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_string(args.batch, args.seed),
                   "sample": "the CPU arm solves a bounded sample of that workload per step (cpu_baseline.sample)"},
        "us_per_iteration": 1e6 * tot_t / max(1, tot_it),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample_desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def other_config(args, local_rank):
    """`--config c4|c5`: the long-horizon and the receding-horizon BASELINE configs on one GPU, one JSON line each with
    the roofline of the backward pass measured the same way as for the headline (CUDA events around every launch of a
    profiled repetition).  Not the driver's headline: that is `--config c3` (the default)."""
    import ctypes
    import dataclasses

    import torch

    from quadrotorilqr_b200 import BatchILQR, RESULT_DTYPE, _capi, problems

    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)

    def mk(model, opts):
        return BatchILQR(model["mass_kg"], model["inertia"], model["arm_length_m"], model["torque_to_thrust_ratio_m"],
                         model["g_mpss"], model["Q"], model["R"], model["dt_s"], opts, device=local_rank)

    def device_problem(s_, x0, desired, N):
        B = x0.shape[0]
        x0_soa = torch.from_numpy(np.ascontiguousarray(x0.T)).to(dev)
        init = torch.empty((N, 17, B), dtype=torch.float64, device=dev)
        des = torch.empty((N, 17, 1), dtype=torch.float64, device=dev)
        torch.cuda.synchronize()
        s_.pack_trajectory_device(torch.from_numpy(desired[None].copy()).to(dev), des)
        s_.rollout_constant_control_device(x0_soa, desired[0, 14:18], init)
        torch.cuda.synchronize()
        return x0_soa, init, des

    peak = ctypes.c_double(0.0)
    _capi.lib().qilqr_measure_fp64_peak(ctypes.c_int(local_rank), ctypes.byref(peak))
    sampler = ClockSampler(local_rank)
    if args.config == "c4":
        N, dt_s, B = 1000, 0.02, args.batch if args.batch != 65536 else 4096
        m = dict(problems.hover_model(), dt_s=dt_s)
        o = dataclasses.replace(problems.default_options(False), symmetrize_vxx=True, num_parallel_alphas=8)
        P = max(1, min(args.pipeline, 4))  # batches in flight (6 GB of workspace each)
        handles = [mk(m, o) for _ in range(P)]
        s_ = handles[0]
        d = problems.figure_eight_desired(N, dt_s)
        _, init, des = device_problem(s_, problems.figure_eight_initial_states(B, d), d, N)
        works = [torch.empty_like(init) for _ in range(P)]
        ress = [torch.zeros(B * 24, dtype=torch.uint8, device=dev) for _ in range(P)]
        streams = [torch.cuda.ExternalStream(h_.stream_handle, device=dev) for h_ in handles]
        res = ress[0]

        def step(j=0):
            with torch.cuda.stream(streams[j]):
                works[j].copy_(init, non_blocking=True)
            handles[j].solve_device(works[j], des, results=ress[j])

        def run(n_each):
            ths = [threading.Thread(target=lambda j=j: [step(j) for _ in range(n_each)]) for j in range(P)]
            for t_ in ths:
                t_.start()
            for t_ in ths:
                t_.join()
            torch.cuda.synchronize()

        run(max(1, args.warmup // 2))
        t0 = time.perf_counter()
        for _ in range(2):
            step(0)
        torch.cuda.synchronize()
        serial_dt = (time.perf_counter() - t0) / 2
        sampler.start()
        n_each = max(2, args.steps // 5)
        steps = n_each * P
        t0 = time.perf_counter()
        run(n_each)
        dt = (time.perf_counter() - t0) / steps
        clocks = sampler.stop()
        s_.set_profiling(True)
        step(0)
        torch.cuda.synchronize()
        st = s_.last_solve_stats()
        s_.set_profiling(False)
        r = np.frombuffer(res.cpu().numpy().tobytes(), dtype=RESULT_DTYPE)
        conv = int(np.sum((r["status"] == 1) | (r["status"] == 2)))
        its = int(r["backward_passes"].sum())
        achieved = st["backward_problem_knots"] * F_BWD / (st["backward_ms"] * 1e-3) / 1e12
        line = {"metric": "converged iLQR solves/sec (BASELINE config 4: N=1000 figure-eight tracking, batch 4096, 8 parallel "
                          "line-search step sizes, symmetrised V_xx, FP64)", "value": conv / dt, "unit": UNIT,
                "n_gpus": 1, "steps": steps, "warmup": max(1, args.warmup // 2), "ms_per_step": 1e3 * dt,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"c4: batch {B}, N={N}, dt={dt_s}, figure-eight p(t)=(2 sin wt, 2 sin wt cos wt, 1), "
                                       "x0 = p(0) + U[-0.3,0.3]^3, hover-thrust initial rollout",
                           "pipeline": f"{P} batches in flight (one solver handle + host thread + stream pair each)"},
                "serial_ms_per_step": 1e3 * serial_dt, "serial_value": conv / serial_dt,
                "iterations_per_solve": its / B, "max_iterations": int(r["backward_passes"].max()),
                "problem_knot_iterations_per_s": its * N / dt, "us_per_iteration": 1e6 * dt / its,
                "gpu_launches": int(st["kernel_launches"]), "clocks": clocks,
                "roofline": {"kernel": "backward pass = k_linearise + k_riccati_g4", "bound": "fp64", "achieved": achieved,
                             "peak": peak.value, "unit": "TFLOP/s", "frac": achieved / peak.value if peak.value else None,
                             "frac_definition": "dense as-written reference FLOPs / CUDA-event time of every backward launch "
                                                "of one profiled solve / measured DFMA peak (512 Riccati warps on 148 SMs: "
                                                "latency-bound)",
                             "share_of_profiled_step": st["backward_ms"] / (st["backward_ms"] + st["rollout_ms"]),
                             "rollout_ms": st["rollout_ms"], "backward_ms": st["backward_ms"], "traffic": None}}
    else:  # c5
        N, B, T = N_KNOTS, (args.batch if args.batch != 65536 else 16384), max(10, args.steps * 5)
        m, o = problems.hover_model(), problems.default_options(False)
        s_ = mk(m, o)
        d = problems.hover_desired_trajectory(N)
        plant, traj, des = device_problem(s_, problems.hover_initial_states(B, seed=5), d, N)
        rng = np.random.Generator(np.random.Philox(key=55))
        dist = torch.from_numpy(np.ascontiguousarray(rng.uniform(-0.01, 0.01, (6, B)))).to(dev)
        s_.mpc_run_device(max(2, args.warmup), traj, des, plant, disturbance=dist)
        torch.cuda.synchronize()
        sampler.start()
        t0 = time.perf_counter()
        tot = s_.mpc_run_device(T, traj, des, plant, disturbance=dist)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        clocks = sampler.stop()
        s_.set_profiling(True)
        s_.mpc_run_device(5, traj, des, plant, disturbance=dist)
        torch.cuda.synchronize()
        st = s_.last_solve_stats()
        s_.set_profiling(False)
        achieved = st["backward_problem_knots"] * F_BWD / (st["backward_ms"] * 1e-3) / 1e12 if st["backward_ms"] else None
        line = {"metric": "warm-started iLQR re-solves/sec (BASELINE config 5: receding-horizon MPC, 16384 quadrotors "
                          "closed loop, N=40, FP64)", "value": tot["resolves"] / dt, "unit": "re-solves/s",
                "n_gpus": 1, "steps": T, "warmup": max(2, args.warmup), "ms_per_step": 1e3 * dt / T,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"c5: {B} quadrotors x {T} closed-loop steps, N={N}, plant = discrete_dynamics + constant "
                                       "body-velocity disturbance U[-0.01,0.01]^6, warm start = previous solution shifted one knot"},
                "iterations_per_resolve": tot["backward_passes"] / tot["resolves"], "not_converged": tot["not_converged"],
                "gpu_launches": int(st["kernel_launches"]), "clocks": clocks,
                "roofline": {"kernel": "backward pass = k_linearise + k_riccati_g4 / k_riccati_g16", "bound": "fp64",
                             "achieved": achieved, "peak": peak.value, "unit": "TFLOP/s",
                             "frac": achieved / peak.value if (achieved and peak.value) else None,
                             "frac_definition": "dense as-written reference FLOPs / CUDA-event time of every backward launch of 5 "
                                                "profiled closed-loop steps / measured DFMA peak",
                             "share_of_profiled_step": st["backward_ms"] / max(1e-9, st["backward_ms"] + st["rollout_ms"]),
                             "traffic": None}}
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c3", choices=["c3", "c4", "c5"],
                    help="c3 (default) = the headline, BASELINE config 3; c4 / c5: long-horizon and MPC configs, one GPU")
    ap.add_argument("--batch", type=int, default=65536, help="problems per GPU")
    ap.add_argument("--seed", type=int, default=2026)
    ap.add_argument("--cpu-sample-per-core", type=int, default=1024)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--model-variant", type=int, default=0, choices=[0, 1, 2, 3, 4],
                    help="NOT the headline: 0 = the reference's QuadrotorModel (BASELINE config); 1 RK4, 2 Coriolis, "
                         "3 both, 4 = reference model on the model-agnostic kernels (include/qilqr.h QILQR_MODEL_*)")
    ap.add_argument("--stagger-ms", type=float, default=-1.0,
                    help="start offset between pipelined handles (default: one measured step time; 0 = start together)")
    ap.add_argument("--pipeline", type=int, default=8,
                    help="batches in flight per GPU: host threads, each driving its own solver handle(s) and stream(s)")
    ap.add_argument("--handles-per-thread", type=int, default=1, choices=[1, 2],
                    help="2 = a thread begins its next batch on a second handle before finishing the previous one "
                         "(useful with QILQR_PERSISTENT_TAIL=1; measured no faster on one B200, profiles/r2)")
    args = ap.parse_args()

    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        return reference_arm(args, rank, world)
    # Two streams per handle: with the default 8 hardware queues, streams that share a queue serialise (a tail kernel
    # waits for another handle's bulk kernel to be dispatched).  Must be set before the CUDA context exists.
    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

    import torch

    from quadrotorilqr_b200 import BatchILQR, RESULT_DTYPE, _capi, problems

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    if args.config != "c3":
        return other_config(args, local_rank)
    import ctypes

    dist = None
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world,
                                device_id=torch.device(f"cuda:{local_rank}"))
    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)

    m, opts = problems.hover_model(), problems.default_options(False)
    if os.environ.get("QILQR_BENCH_MAX_ITERS"):  # diagnosis only (cuts the max_iters tail); not the BASELINE config
        opts.convergence_criteria.max_iters = float(os.environ["QILQR_BENCH_MAX_ITERS"])
    B, N = args.batch, N_KNOTS
    P = max(1, args.pipeline)

    H = max(1, args.handles_per_thread)

    def make_solver():
        s_ = BatchILQR(m["mass_kg"], m["inertia"], m["arm_length_m"], m["torque_to_thrust_ratio_m"], m["g_mpss"],
                       m["Q"], m["R"], m["dt_s"], opts, device=local_rank, model_flags=args.model_variant)
        return s_

    # P host threads x H solver handles = a software pipeline of batches.  A thread begins a batch on one handle
    # (qilqr_solve_*_begin returns once the device finishes the batch on its own: the few problems that creep to
    # max_iters run in the persistent tail kernel), begins the next batch on its other handle, and only then
    # finishes the first.  Every step is still one full batch, solved to the reference's own termination rule.
    solvers = [[make_solver() for _ in range(H)] for _ in range(P)]
    streams = [[torch.cuda.ExternalStream(s_.stream_handle, device=dev) for s_ in row] for row in solvers]
    solver = solvers[0][0]

    # ---- synthetic inputs: this rank's problems [rank*B, (rank+1)*B) of the Philox stream ----
    desired = problems.hover_desired_trajectory(N, m["dt_s"], m["mass_kg"], m["g_mpss"])
    x0 = problems.hover_initial_states(B, seed=args.seed, first=rank * B)
    x0_soa = torch.from_numpy(np.ascontiguousarray(x0.T)).to(dev)  # [13][B]
    init_soa = torch.empty((N, 17, B), dtype=torch.float64, device=dev)
    work_soa = [[torch.empty_like(init_soa) for _ in range(H)] for _ in range(P)]
    des_aos = torch.from_numpy(desired[None].copy()).to(dev)
    des_soa = torch.empty((N, 17, 1), dtype=torch.float64, device=dev)
    res_dev = [[torch.zeros(B * 24, dtype=torch.uint8, device=dev) for _ in range(H)] for _ in range(P)]
    torch.cuda.synchronize()
    solver.pack_trajectory_device(des_aos, des_soa)
    solver.rollout_constant_control_device(x0_soa, desired[0, 14:18], init_soa)  # open-loop hover rollout
    torch.cuda.synchronize()

    STAT_KEYS = ("backward_ms", "rollout_ms", "backward_problem_knots", "rollout_problem_knots",
                 "problem_iterations", "problem_rollouts", "solver_iterations", "bulk_wall_ms", "tail_wall_ms",
                 "backward_ms_bulk", "rollout_ms_bulk", "backward_problem_knots_bulk", "backward_launches_bulk",
                 "kernel_launches")

    def device_begin(j, h):
        with torch.cuda.stream(streams[j][h]):
            work_soa[j][h].copy_(init_soa, non_blocking=True)
        solvers[j][h].solve_device_begin(work_soa[j][h], des_soa, results=res_dev[j][h])

    def device_finish(j, h, acc):
        solvers[j][h].solve_device_finish()
        st = solvers[j][h].last_solve_stats()
        for k_ in STAT_KEYS:
            acc[k_] += st[k_]

    def run_pipelined(begin_fn, finish_fn, nsteps, stagger_s=0.0):
        """`nsteps` batches round-robin over the P threads; thread j alternates between its H handles and finishes
        a batch only when it needs the handle again (or at the end).  `stagger_s`: thread j starts j * stagger_s
        late (threads that start together stay in lock step)."""
        accs = [{k_: 0 for k_ in STAT_KEYS} for _ in range(P)]
        counts = [nsteps // P + (1 if j < nsteps % P else 0) for j in range(P)]
        errors = []

        def worker(j):
            try:
                if stagger_s > 0 and j > 0 and counts[j] > 0:
                    time.sleep(j * stagger_s)
                pending = [False] * H
                for s_ in range(counts[j]):
                    h = s_ % H
                    if pending[h]:
                        finish_fn(j, h, accs[j])
                    begin_fn(j, h)
                    pending[h] = True
                for d_ in range(H):  # oldest first
                    h = (counts[j] + d_) % H
                    if pending[h]:
                        finish_fn(j, h, accs[j])
                        pending[h] = False
            except Exception as e:  # noqa: BLE001 -- re-raised on the main thread
                errors.append(e)

        if P == 1:
            worker(0)
        else:
            ths = [threading.Thread(target=worker, args=(j,)) for j in range(P)]
            for t_ in ths:
                t_.start()
            for t_ in ths:
                t_.join()
        if errors:
            raise errors[0]
        return {k_: sum(a_[k_] for a_ in accs) for k_ in STAT_KEYS}

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- `value`: device-resident ----------------------------------------------------------
    n_warm = max(args.warmup, P * H)
    run_pipelined(device_begin, device_finish, n_warm)  # first calls: allocation of the workspaces
    barrier()
    t_w0 = time.perf_counter()
    run_pipelined(device_begin, device_finish, n_warm)
    barrier()
    # start offset between threads inside the timed regions: one step time (measured on the warm-up), unless given
    stagger_s = (args.stagger_ms * 1e-3 if args.stagger_ms >= 0 else (time.perf_counter() - t_w0) / n_warm) if P > 1 else 0.0
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(P)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(P)]
    barrier()
    for j in range(P):
        ev0[j].record(streams[j][0])
    t_wall0 = time.perf_counter()
    tot = run_pipelined(device_begin, device_finish, args.steps, stagger_s)
    for j in range(P):
        ev1[j].record(streams[j][0])
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = tot["kernel_launches"]
    clocks = sampler.stop()
    dev_ms = max(ev0[j].elapsed_time(ev1[j]) for j in range(P))
    prob_iters, prob_rollouts, solver_iters = tot["problem_iterations"], tot["problem_rollouts"], tot["solver_iterations"]
    res = np.frombuffer(res_dev[0][0].cpu().numpy().tobytes(), dtype=RESULT_DTYPE)
    converged = int(np.sum((res["status"] == 1) | (res["status"] == 2)))
    ls_failed = int(np.sum(res["status"] >= 4))
    max_iter_hit = int(np.sum(res["status"] == 3))

    # One batch at a time on one handle (no pipelining): the latency of a single batch ...
    def serial_steps(n_, acc):
        t0_ = time.perf_counter()
        for _ in range(n_):
            device_begin(0, 0)
            device_finish(0, 0, acc)
        torch.cuda.synchronize()
        return 1e3 * (time.perf_counter() - t0_) / n_

    barrier()
    n_serial = 3
    serial_ms = serial_steps(n_serial, {k_: 0 for k_ in STAT_KEYS})
    # ... and the place where the kernels are timed ALONE with CUDA events on the solver's stream around every
    # launch (inside the pipelined region kernels of different handles overlap, so per-kernel durations are not
    # separable).  Profiling keeps the host-driven loop to the end, so the latency-bound tail launches (a handful
    # of problems each) are timed too and reported apart from the throughput-bound bulk launches.
    solver.set_profiling(True)
    ser = {k_: 0 for k_ in STAT_KEYS}
    prof_ms = serial_steps(n_serial, ser)
    solver.set_profiling(False)
    bwd_ms, roll_ms = ser["backward_ms"], ser["rollout_ms"]
    bwd_knots, roll_knots = ser["backward_problem_knots"], ser["rollout_problem_knots"]
    bwd_ms_bulk, bwd_knots_bulk = ser["backward_ms_bulk"], ser["backward_problem_knots_bulk"]

    # ---- `e2e`: host buffers through qilqr_solve_host_begin / _finish -----------------------------------
    e2e = None
    if not args.no_e2e:
        init_host = torch.empty((B, N, 18), dtype=torch.float64, pin_memory=True)
        out_host = [[torch.empty((B, N, 18), dtype=torch.float64, pin_memory=True) for _ in range(H)] for _ in range(P)]
        res_host = [[torch.zeros(B * 24, dtype=torch.uint8, pin_memory=True) for _ in range(H)] for _ in range(P)]
        aos = torch.empty((B, N, 18), dtype=torch.float64, device=dev)
        torch.cuda.synchronize()
        solver.unpack_trajectory_device(init_soa, aos, time_src=None)
        torch.cuda.synchronize()
        init_host.copy_(aos)
        init_host[:, :, 0] = torch.arange(N, dtype=torch.float64) * m["dt_s"]
        del aos
        desired_c = np.ascontiguousarray(desired)

        def host_begin(j, h):
            solvers[j][h].solve_host_begin(init_host, desired_c, out_host[j][h], res_host[j][h])

        def host_finish(j, h, acc):
            solvers[j][h].solve_host_finish()

        e2e_steps = max(P, args.steps)
        run_pipelined(host_begin, host_finish, P * H)
        barrier()
        t0 = time.perf_counter()
        run_pipelined(host_begin, host_finish, e2e_steps, stagger_s)
        barrier()
        e2e_t = time.perf_counter() - t0
        r2 = np.frombuffer(res_host[0][0].numpy().tobytes(), dtype=RESULT_DTYPE)
        conv2 = int(np.sum((r2["status"] == 1) | (r2["status"] == 2)))
        e2e = {"t": e2e_t, "steps": e2e_steps, "converged": conv2,
               "h2d": B * N * 18 * 8 + N * 18 * 8, "d2h": B * N * 18 * 8 + B * 24}

        # the same batches through qilqr_solve_from_controls_host: x0 and ONE nominal control sequence go up (the
        # initial trajectory of this workload is their open-loop rollout), full trajectories come back
        x0_host = torch.from_numpy(np.ascontiguousarray(x0)).pin_memory()
        u_nom = np.ascontiguousarray(np.tile(desired[0, 14:18], (N, 1)))

        def ctl_begin(j, h):
            solvers[j][h].solve_from_controls(x0_host, u_nom, desired_c, out_traj=out_host[j][h], results=res_host[j][h],
                                              begin_only=True)

        run_pipelined(ctl_begin, host_finish, P * H)
        barrier()
        t0 = time.perf_counter()
        run_pipelined(ctl_begin, host_finish, e2e_steps, stagger_s)
        barrier()
        e2e["ctl_t"] = time.perf_counter() - t0
        r3 = np.frombuffer(res_host[0][0].numpy().tobytes(), dtype=RESULT_DTYPE)
        e2e["ctl_converged"] = int(np.sum((r3["status"] == 1) | (r3["status"] == 2)))
        e2e["ctl_same_results"] = bool(np.array_equal(r3, r2))
        e2e["ctl_h2d"] = B * 13 * 8 + N * 4 * 8 + N * 18 * 8

        # ... and with the solved CONTROLS [B][N][4] as the only trajectory output (what a receding-horizon caller
        # consumes; the states are their rollout from x0): 22 % of the bytes of the full trajectories come back
        traj_ref = out_host[0][0][:, :, 14:18].clone()
        uout_host = [[torch.empty((B, N, 4), dtype=torch.float64, pin_memory=True) for _ in range(H)] for _ in range(P)]

        def ctl_out_begin(j, h):
            solvers[j][h].solve_from_controls(x0_host, u_nom, desired_c, out_controls=uout_host[j][h], want_traj=False,
                                              results=res_host[j][h], begin_only=True)

        run_pipelined(ctl_out_begin, host_finish, P * H)
        barrier()
        t0 = time.perf_counter()
        run_pipelined(ctl_out_begin, host_finish, e2e_steps, stagger_s)
        barrier()
        e2e["cc_t"] = time.perf_counter() - t0
        r4 = np.frombuffer(res_host[0][0].numpy().tobytes(), dtype=RESULT_DTYPE)
        e2e["cc_converged"] = int(np.sum((r4["status"] == 1) | (r4["status"] == 2)))
        e2e["cc_same_results"] = bool(np.array_equal(r4, r2) and torch.equal(uout_host[0][0], traj_ref))
        e2e["cc_d2h"] = B * N * 4 * 8 + B * 24
        del traj_ref

        # what the host's memory system gives this rank while every rank copies at once: the H2D + D2H traffic of one
        # step (pinned buffers, both directions concurrently) -- the denominator for the end-to-end scaling
        dma_dev = torch.empty((B, N, 18), dtype=torch.float64, device=dev)
        dma_dev2 = torch.empty((B, N, 18), dtype=torch.float64, device=dev)
        s_up, s_dn = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        barrier()
        t0 = time.perf_counter()
        for _ in range(4):
            with torch.cuda.stream(s_up):
                dma_dev.copy_(init_host, non_blocking=True)
            with torch.cuda.stream(s_dn):
                out_host[0][0].copy_(dma_dev2, non_blocking=True)
        torch.cuda.synchronize()
        e2e["dma_gbs_per_direction"] = 4 * B * N * 18 * 8 / (time.perf_counter() - t0) / 1e9
        del dma_dev, dma_dev2

    # ---- reduce over ranks (NCCL: stats and max time only) ---------------------------------------
    vec = torch.tensor([dev_ms, t_wall * 1e3, float(converged), float(prob_iters), float(ls_failed),
                        float(max_iter_hit), float(launches), e2e["t"] if e2e else 0.0,
                        float(e2e["converged"]) if e2e else 0.0, e2e["ctl_t"] if e2e else 0.0,
                        float(e2e["ctl_converged"]) if e2e else 0.0,
                        e2e["dma_gbs_per_direction"] if e2e else 0.0, e2e["cc_t"] if e2e else 0.0,
                        float(e2e["cc_converged"]) if e2e else 0.0], dtype=torch.float64, device=dev)
    gathered_converged = None
    if dist is not None:
        mx = vec.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = vec.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        mx, sm = mx.cpu().numpy(), sm.cpu().numpy()
        # the path's one collective: per-problem result records of every shard (24 B/problem) over NCCL
        from quadrotorilqr_b200 import sharding

        allres = sharding.gather_results(dist, res_dev[0][0].view(B, 24))
        st_all = np.frombuffer(allres.cpu().numpy().tobytes(), dtype=RESULT_DTYPE)["status"]
        gathered_converged = int(np.sum((st_all == 1) | (st_all == 2)))
    else:
        mx = sm = vec.cpu().numpy()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    step_ms = max(mx[0], 0.0) / args.steps
    wall_ms = mx[1] / args.steps
    # device events bracket the stream work; the solver also synchronises with the host every
    # super-step, so the honest per-step time is the larger of the two clocks
    ms_per_step = max(step_ms, wall_ms)
    total_converged_per_step = sm[2]
    value = total_converged_per_step / (ms_per_step * 1e-3)
    total_iters = sm[3]
    us_per_iter = (ms_per_step * 1e3 * world) / max(1.0, total_iters / args.steps) if total_iters else None

    # ---- roofline of the dominant kernel (backward pass), rank 0 -----------------------------------
    peak = ctypes.c_double(0.0)
    _capi.lib().qilqr_measure_fp64_peak(ctypes.c_int(local_rank), ctypes.byref(peak))
    f_bwd, f_roll, f_cost = VARIANT_FLOPS.get(args.model_variant, (F_BWD, F_ROLL, F_COST))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    counters = None
    try:
        if args.model_variant:
            raise LookupError("no ncu capture is wired in for the model variants")
        counters = json.load(open(os.path.join(ROOT, "profiles", "r2", "r2_kernel_counters.json")))["backward_pass"]
    except Exception:
        pass

    def tflops(knots, flops_per_knot, ms):
        return knots * flops_per_knot / (ms * 1e-3) / 1e12 if ms > 0 else None

    achieved = tflops(bwd_knots_bulk, f_bwd, bwd_ms_bulk)          # bulk launches: the throughput-bound part
    achieved_all = tflops(bwd_knots, f_bwd, bwd_ms)                # every launch, the latency-bound tail included
    exe = counters["executed_flops_per_problem_knot"] if counters else None
    executed = tflops(bwd_knots_bulk, exe, bwd_ms_bulk) if exe else None
    n_bulk_launches = max(1, ser["backward_launches_bulk"])  # backward passes (launch pairs) timed as "bulk"
    roofline = {
        "kernel": ("backward pass = k_linearise + k_riccati_g4 (ILQR::backwards_pass, ilqr.hh:97-147)"
                   if not args.model_variant else
                   "backward pass = k_linearise_dense + k_riccati_dense (model-agnostic ILQR::backwards_pass)"),
        "bound": "fp64", "achieved": achieved, "peak": peak.value, "unit": "TFLOP/s",
        "frac": (achieved / peak.value) if (achieved and peak.value) else None,
        "frac_definition": "DENSE as-written reference FLOPs (30231 per problem-knot, counted by the oracle's "
                           "FLOP-counting scalar) / CUDA-event time of the throughput-bound bulk launches / measured "
                           "DFMA peak; the kernels skip structural zeros, so this mixes 'fewer FLOPs executed' with "
                           "pipe utilisation -- executed_frac is the utilisation",
        "executed_tflops": executed,
        "executed_frac": (executed / peak.value) if (executed and peak.value) else None,
        "executed_flops_per_problem_knot": exe,
        "executed_definition": "2*DFMA + DMUL + DADD thread-instructions per problem-knot from the committed ncu capture "
                               "(profiles/r2/r2_kernel_counters.json: smsp__sass_thread_inst_executed_op_d*_pred_on.sum) x "
                               "the problem-knots of the bulk launches / their CUDA-event time",
        "frac_all_launches": (achieved_all / peak.value) if (achieved_all and peak.value) else None,
        "peak_source": "measured in this run: register-resident DFMA kernel (qilqr_measure_fp64_peak); "
                       "MEASURED_PEAKS.json has no FP64 figure",
        "flops_per_problem_knot": f_bwd,
        "traffic": (counters["dram_bytes_per_problem_knot"] * bwd_knots_bulk / n_bulk_launches) if counters else None,
        "traffic_note": ("per backward pass (k_linearise + k_riccati_g4 launch pair), averaged over the bulk launches: "
                         "dram__bytes_read+write per problem-knot from the committed ncu capture x the mean problem-knots "
                         "of a bulk launch; algorithmic: "
                         f"{counters['algorithmic_bytes_per_problem_knot']:.0f} B per problem-knot -- the difference is the "
                         "linearisation record written by k_linearise and read once by k_riccati_g4") if counters else None,
        "algorithmic_bytes": 552.0 * bwd_knots_bulk / n_bulk_launches,
        "bulk_backward_passes_per_step": n_bulk_launches / n_serial,
        "mean_bulk_launch_ms": bwd_ms_bulk / n_bulk_launches,
        "timed": (f"CUDA events on the solver's stream around every launch of {n_serial} un-pipelined steps run right "
                  "after the timed region (kernels of different handles overlap inside it)"),
        "bulk_ms_per_step": bwd_ms_bulk / n_serial, "all_launches_ms_per_step": bwd_ms / n_serial,
        "share_of_profiled_step": bwd_ms / (prof_ms * n_serial),
        "bulk_kernel_ms_vs_ms_per_step": [bwd_ms_bulk / n_serial, ms_per_step],
        "hbm": {"achieved": (bwd_knots_bulk * (17 + 52) * 8.0) / (bwd_ms_bulk * 1e-3) / 1e9 if bwd_ms_bulk > 0 else None,
                "peak": hbm_peak, "unit": "GB/s",
                "frac": ((bwd_knots_bulk * (17 + 52) * 8.0) / (bwd_ms_bulk * 1e-3) / 1e9 / hbm_peak) if bwd_ms_bulk > 0 else None,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "of fallback"},
        "rollout_kernel": {"achieved": tflops(roll_knots, f_roll + f_cost, roll_ms), "unit": "TFLOP/s",
                           "share_of_profiled_step": roll_ms / (prof_ms * n_serial)},
    }

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_string(B, args.seed),
                   "cache": "inputs larger than L2 (356 MB trajectories + 1.4 GB gains per step vs 126 MB L2)",
                   "parallelism": f"{world} independent shard(s), one process per GPU",
                   **({"model_variant": f"{args.model_variant} (NOT the BASELINE model: see --model-variant)"}
                      if args.model_variant else {}),
                   "pipeline": f"{P} batches in flight per GPU ({P} host threads x {H} solver handle(s), each on its own "
                               "stream pair), "
                               f"CUDA_DEVICE_MAX_CONNECTIONS={os.environ.get('CUDA_DEVICE_MAX_CONNECTIONS')}, "
                               f"threads started {1e3 * stagger_s:.1f} ms apart; each step is one full batch"},
        "serial_ms_per_step": serial_ms,
        "serial_value": (converged / (serial_ms * 1e-3)) if serial_ms else None,
        "us_per_iteration": us_per_iter,
        "converged_fraction": total_converged_per_step / (B * world),
        "line_search_failures": int(sm[4]), "max_iters_hit": int(sm[5]),
        "iterations_per_solve": total_iters / args.steps / (B * world),
        "solver_iterations_per_step": solver_iters / args.steps,
        # host wall time a batch spends in its throughput-bound bulk and in its latency-bound tail (rank 0, pipelined)
        "bulk_wall_ms_per_batch": tot["bulk_wall_ms"] / args.steps, "tail_wall_ms_per_batch": tot["tail_wall_ms"] / args.steps,
        "device_event_ms_per_step": step_ms, "wall_ms_per_step": wall_ms,
        "gpu_launches": int(sm[6] / world),
        **({"gathered_results": {"problems": B * world, "converged": gathered_converged,
                                 "note": "ncclAllGather of the per-problem result records after the timed region"}}
           if gathered_converged is not None else {}),
        "clocks": clocks,
        "roofline": roofline,
    }
    if e2e:
        e2e_value = sm[8] / (mx[7] / e2e["steps"])
        line["e2e"] = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": e2e["h2d"],
                       "d2h_bytes_per_step": e2e["d2h"], "ms_per_step": 1e3 * mx[7] / e2e["steps"],
                       "steps": e2e["steps"],
                       "api": "qilqr_solve_host_begin / _finish (pinned host AoS in/out, results struct per problem; "
                              "same results as qilqr_solve_host)"}
        line["e2e_from_controls"] = {
            "value": sm[10] / (mx[9] / e2e["steps"]), "unit": UNIT, "ms_per_step": 1e3 * mx[9] / e2e["steps"],
            "h2d_bytes_per_step": e2e["ctl_h2d"], "d2h_bytes_per_step": e2e["d2h"], "same_results_as_e2e": e2e["ctl_same_results"],
            "api": "qilqr_solve_from_controls_host_begin / qilqr_solve_host_finish: x0 [B][13] and one nominal control "
                   "sequence in (this workload's initial trajectories are their open-loop rollout, made on the device), "
                   "full trajectories [B][N][18] out"}
        line["e2e_controls_in_controls_out"] = {
            "value": sm[13] / (mx[12] / e2e["steps"]), "unit": UNIT, "ms_per_step": 1e3 * mx[12] / e2e["steps"],
            "h2d_bytes_per_step": e2e["ctl_h2d"], "d2h_bytes_per_step": e2e["cc_d2h"],
            "same_results_as_e2e": e2e["cc_same_results"],
            "api": "qilqr_solve_from_controls_host_begin (out_traj = NULL, out_controls [B][N][4]) / "
                   "qilqr_solve_host_finish: the solved control sequences and the per-problem results come back"}
        line["host_dma"] = {
            "gbs_per_direction_all_gpus": sm[11], "gbs_per_direction_per_gpu": sm[11] / world,
            "note": "pinned-memory H2D and D2H copies of one step's trajectories, both directions at once, on every rank at "
                    "the same time (no compute): what the host's memory system and PCIe fabric give the end-to-end path. "
                    f"A step moves {(e2e['h2d'] + e2e['d2h']) / 1e6:.0f} MB per GPU through it (e2e) or "
                    f"{(e2e['ctl_h2d'] + e2e['d2h']) / 1e6:.0f} MB (e2e_from_controls)"}
    if not args.no_cpu_baseline and world >= 1:
        import oracle as O

        O.build()
        cfg = oracle_config(O, m, opts)
        cores = O.hardware_threads()
        # one pass of ~10-15 s of wall time (the whole batch on a 16-core box); the reference arm's steps use a
        # quarter of that so that its W + K steps end within a few minutes
        sample = int(min(B, max(64, 4 * args.cpu_sample_per_core * cores)))
        init_cpu = cpu_initial_trajectories(O, cfg, problems, m, desired, x0[:sample])
        dt, conv, its = run_cpu_sample(O, cfg, desired, init_cpu, cores)
        line["cpu_baseline"] = {
            "value": conv / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"first {sample} problems of the same batch, {cores} threads, oracle (restated reference "
                      "algorithm, g++ -O2 -ffp-contract=off); the literal reference cannot be built here",
            "us_per_iteration": 1e6 * dt / max(1, its), "seconds": dt}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
