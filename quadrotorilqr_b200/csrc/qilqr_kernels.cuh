// =============================================================================
// qilqr_kernels.cuh -- CUDA kernels of the batched iLQR hot path (sm_100a).
//
// Data layout in HBM (structure of arrays, problem index fastest so that a warp
// of 32 problems reads 256 contiguous bytes per row):
//   trajectory  double[N][17][B]   rows 0-2 t, 3-6 q(x,y,z,w), 7-12 body velocity, 13-16 control
//   desired     double[N][17][Bd]  Bd = 1 (shared) or B
//   k           double[N][4][B]
//   K           double[N][48][B]   row = 12*control + state   (ILQR::FeedbackGains, ilqr.hh:40-41)
// Each problem has two trajectory buffers; `sel[b]` says which one is current, so
// accepting a line-search candidate is a bit flip, not a copy.
//
// Kernels
//   k_cost_trajectory   ILQR::cost_trajectory      (ilqr.hh:89-95)
//   k_backward_t1       ILQR::backwards_pass       (ilqr.hh:97-147) + exit A of solve (:61-68), one thread
//                       per problem, no structure assumptions: the independent cross-check of the production
//                       backward pass (k_linearise + k_riccati_g4 in qilqr_backward_split.cuh)
//   k_rollout           ILQR::forward_sim + cost   (ilqr.hh:149-172, 89-95) + Armijo test of
//                       line_search (:182-189) + exit B of solve (:78-84); MODE_WIDE: parallel step sizes
//   k_select_alpha      first step size that passes the Armijo test among the parallel candidates
//   k_compact*          ordered stream compaction of the active / searching problem lists
//   k_mpc_advance       plant step + warm-start shift of the receding-horizon loop
//   k_pack / k_unpack   array-of-structs <-> structure-of-arrays
// =============================================================================
#pragma once
#include "qilqr_device.cuh"
#include "qilqr_model_generic.cuh"
#include "../../include/qilqr.h"

namespace qilqr {

// ACTIVE: needs a backward pass; SEARCH: needs a rollout at alpha[b]; WIDE: needs a round of parallel
// step-size evaluations (num_parallel_alphas > 1); DONE: finished.
enum Phase : int { PHASE_ACTIVE = 0, PHASE_SEARCH = 1, PHASE_DONE = 2, PHASE_WIDE = 3 };
// FORWARD: plain forward_sim; SOLVE / LINE_SEARCH: with the bookkeeping of solve() / line_search();
// WIDE: cost only, for step sizes alpha[b] * step_update^j, j = 0..palpha-1 (one thread per (problem, j)).
enum RolloutMode : int { MODE_FORWARD = 0, MODE_SOLVE = 1, MODE_LINE_SEARCH = 2, MODE_WIDE = 3 };

struct Problem {
  int B;   // batch (also the pitch of every SoA row)
  int N;   // knots
  int Bd;  // desired_count: 1 or B
  double *buf0, *buf1;    // trajectory double-buffer
  const double *desired;  // [N][17][Bd]
  double *gk, *gK;        // gains
};

struct SolveState {  // one entry per problem
  double *cost;      // actual cost of the current trajectory ("cost"/"new_cost" in ilqr.hh:56-75)
  double *new_cost;  // cost of the last candidate
  double *qutk, *ktquuk;  // detail::CostReductionTerms (ilqr.hh:13-16)
  double *alpha;
  int *ls_iter, *status, *bwd, *rollouts, *ndebug, *sel, *phase, *accepted_iter;
  double *cost_hist;  // [hist_cap][B] or nullptr
  int hist_cap;
};

QD size_t row_index(int i, int c, int rows, int B, int b) { return (size_t(i) * rows + c) * size_t(B) + b; }

QD void load_point(const double *traj, int i, int B, int b, double *x /*13*/, double *u /*4*/) {
#pragma unroll
  for (int c = 0; c < 13; ++c) x[c] = traj[row_index(i, c, 17, B, b)];
#pragma unroll
  for (int c = 0; c < 4; ++c) u[c] = traj[row_index(i, 13 + c, 17, B, b)];
}
QD void store_point(double *traj, int i, int B, int b, const double *x, const double *u) {
#pragma unroll
  for (int c = 0; c < 13; ++c) traj[row_index(i, c, 17, B, b)] = x[c];
#pragma unroll
  for (int c = 0; c < 4; ++c) traj[row_index(i, 13 + c, 17, B, b)] = u[c];
}

QD void prefetch_l2(const void *ptr) { asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr)); }

// A global load the compiler may not sink towards its first use (volatile asm keeps program order):
// lets a kernel issue all the loads of a loop iteration up front, ahead of long dependent arithmetic.
// NC = true reads through the non-coherent path: only for data that no thread of the running kernel writes (the
// per-iteration kernels).  The persistent tail kernel rewrites gains and trajectories between its own
// iterations and must use NC = false.
template <bool NC = true>
QD double ldg_early(const double *ptr) {
  double v;
  if (NC) asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(ptr));
  else asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(ptr) : "memory");
  return v;
}

// ILQR::is_converged (ilqr.hh:196-205)
QD bool is_converged(const DeviceParams &p, double cost, double new_cost) {
  const double d = fabs(cost - new_cost);
  if (d / fabs(cost) < p.rtol) return true;
  if (d < p.atol) return true;
  return false;
}

// ---------------------------------------------------------------------------
// cost_trajectory
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_cost_trajectory(const __grid_constant__ DeviceParams p, const double *traj, const double *desired,
                  int B, int N, int Bd, double *cost_out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int bd = (Bd == 1) ? 0 : b;
  double cost = 0.0;
  for (int i = 0; i < N; ++i) {
    double x[13], u[4], xd[13], ud[4], dx[12], du[4];
    load_point(traj, i, B, b, x, u);
    load_point(desired, i, Bd, bd, xd, ud);
    state_minus(x, xd, dx, nullptr);
#pragma unroll
    for (int j = 0; j < 4; ++j) du[j] = u[j] - ud[j];
    cost = cost + quadratic_cost(p, dx, du);
  }
  cost_out[b] = cost;
}

// ---------------------------------------------------------------------------
// forward_sim + cost_trajectory (+ line-search bookkeeping), one thread per problem
// ---------------------------------------------------------------------------
struct RolloutArgs {
  Problem pr;
  SolveState st;
  const int *list;  // problems to roll out (nullptr: 0..n-1)
  int n;
  int iter;  // MODE_SOLVE: the solver's super-step number (marks which problems were accepted in this launch; a
             // problem's own iteration index i of solve() is its count of backward passes - 1);
             // MODE_LINE_SEARCH: the iteration index to assume (any value > 0)
  int mode;  // RolloutMode
  // MODE_FORWARD only: explicit buffers and step sizes
  const double *cur;
  double *out;
  const double *alpha_in;
  double *cost_out;  // may be nullptr (MODE_WIDE: [palpha][B])
  int palpha;        // MODE_WIDE: step sizes evaluated per problem
  int check_phase;   // MODE_SOLVE / MODE_WIDE: `list` is the list of ALIVE problems; skip those that are not in
                     // PHASE_SEARCH / PHASE_WIDE
  int wide_on_reject;  // MODE_SOLVE with parallel step sizes: a rejected candidate sends the problem to PHASE_WIDE
                       // (the next `num_parallel_alphas` step sizes are then evaluated concurrently)
};

// The candidate trajectory of problem b (in its other buffer, cost `cost`) is accepted: it becomes the current
// trajectory; ILQRDebug / cost history entry; exit B of solve() (ilqr.hh:72-84).
QD void accept_candidate(const DeviceParams &p, const SolveState &st, int B, int b, int epoch, int mode, int it,
                         double cur_cost, double cost) {
  st.sel[b] ^= 1;
  st.cost[b] = cost;
  st.accepted_iter[b] = epoch;
  const int nd = st.ndebug[b];
  if (st.cost_hist && nd < st.hist_cap) st.cost_hist[size_t(nd) * B + b] = cost;
  st.ndebug[b] = nd + 1;
  if (mode == MODE_SOLVE && it > 0 && is_converged(p, cur_cost, cost)) {
    st.status[b] = QILQR_STATUS_CONVERGED_ACTUAL;  // ilqr.hh:82-84
    st.phase[b] = PHASE_DONE;
  } else if (mode == MODE_LINE_SEARCH) {
    st.phase[b] = PHASE_DONE;
  } else {
    // the loop bound of ilqr.hh:58, `i < max_iters` with max_iters a double, for the next i = it + 1
    // (k_finalize turns "done without a status" into QILQR_STATUS_MAX_ITERS)
    st.phase[b] = (double(it + 1) < p.max_iters) ? PHASE_ACTIVE : PHASE_DONE;
  }
}

// line_search() / solve() bookkeeping after one candidate rollout of problem b (ilqr.hh:70-84, 182-193)
QD void rollout_finish(const DeviceParams &p, const RolloutArgs &a, int b, double alpha, double cost) {
  const int B = a.pr.B;
  const SolveState &st = a.st;
  st.rollouts[b] += 1;
  st.new_cost[b] = cost;
  const double cur_cost = st.cost[b];
  // i of solve()'s loop (ilqr.hh:58) for this problem: every problem counts its own iterations, so that problems of
  // one batch may be at different iterations (a rejected candidate costs only that problem a round)
  const int it = (a.mode == MODE_SOLVE) ? st.bwd[b] - 1 : a.iter;
  bool accept;
  if (a.mode == MODE_SOLVE && it == 0) {
    accept = true;  // ilqr.hh:70-73: no acceptance test on the first iteration
  } else {
    const double desired = p.desired_reduction_frac * (alpha * st.qutk[b] + alpha * alpha * st.ktquuk[b] / 2.0);
    accept = (cost - cur_cost < desired);  // ilqr.hh:186; NaN -> reject
  }
  if (accept) {
    accept_candidate(p, st, B, b, a.iter, a.mode, it, cur_cost, cost);
  } else {
    st.alpha[b] = alpha * p.step_update;  // ilqr.hh:189
    const int ls = st.ls_iter[b] + 1;
    st.ls_iter[b] = ls;
    if (ls >= p.ls_max_iters) {
      st.status[b] = isfinite(cost) ? QILQR_STATUS_LINE_SEARCH_FAILED : QILQR_STATUS_NONFINITE;  // ilqr.hh:191-193
      st.phase[b] = PHASE_DONE;
    } else if (a.wide_on_reject) {
      st.phase[b] = PHASE_WIDE;
    }
  }
}

#ifndef QILQR_ROLLOUT_MINB
#define QILQR_ROLLOUT_MINB 2
#endif
// GENERIC = false: the reference's QuadrotorModel, inlined; true: any model variant (qilqr_model_generic.cuh)
template <bool GENERIC>
__global__ void __launch_bounds__(128, QILQR_ROLLOUT_MINB) k_rollout(const __grid_constant__ DeviceParams p, const __grid_constant__ RolloutArgs a) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int wide_j = (a.mode == MODE_WIDE) ? tid / a.n : 0;  // j-major so that a warp covers consecutive problems
  const int t = (a.mode == MODE_WIDE) ? tid % a.n : tid;
  if (tid >= a.n * ((a.mode == MODE_WIDE) ? a.palpha : 1)) return;
  const int b = a.list ? a.list[t] : t;
  if (a.check_phase && a.st.phase[b] != ((a.mode == MODE_WIDE) ? PHASE_WIDE : PHASE_SEARCH)) return;
  const int B = a.pr.B, N = a.pr.N, Bd = a.pr.Bd;
  const int bd = (Bd == 1) ? 0 : b;
  const double *cur;
  double *cand;
  double alpha;
  if (a.mode == MODE_FORWARD) {
    cur = a.cur; cand = a.out; alpha = a.alpha_in[b];
  } else if (a.mode == MODE_WIDE) {
    const int s = a.st.sel[b];
    cur = s ? a.pr.buf1 : a.pr.buf0;
    // the first step size of a round also writes its trajectory: it is the one the search accepts nearly always,
    // and k_select_alpha can then accept it without another rollout
    cand = (wide_j == 0) ? (s ? a.pr.buf0 : a.pr.buf1) : nullptr;
    alpha = a.st.alpha[b];
    for (int j = 0; j < wide_j; ++j) alpha *= p.step_update;  // the same products as the sequential search
  } else {
    const int s = a.st.sel[b];
    cur = s ? a.pr.buf1 : a.pr.buf0;
    cand = s ? a.pr.buf0 : a.pr.buf1;
    alpha = a.st.alpha[b];
  }
  const bool want_cost = (a.mode != MODE_FORWARD) || (a.cost_out != nullptr);

  double x[13], ubar[4];
  load_point(cur, 0, B, b, x, ubar);  // state = current_traj.front().state (ilqr.hh:156)
  double cost = 0.0;
  for (int i = 0; i < N; ++i) {
    // every load of this knot is issued here, before the long dependent chain of state_minus: the
    // gains arrive in its shadow instead of stalling the feedback product
    double xbar[13], gk[4], gK[48];
#pragma unroll
    for (int c = 0; c < 13; ++c) xbar[c] = ldg_early(&cur[row_index(i, c, 17, B, b)]);
#pragma unroll
    for (int c = 0; c < 4; ++c) ubar[c] = ldg_early(&cur[row_index(i, 13 + c, 17, B, b)]);
#pragma unroll
    for (int e = 0; e < 4; ++e) gk[e] = ldg_early(&a.pr.gk[row_index(i, e, 4, B, b)]);
    // the state part of the cost needs nothing that was just requested: it runs while the loads are in flight
    double cx = 0.0, ud[4] = {0.0, 0.0, 0.0, 0.0};
    if (want_cost) {
      double xd[13], dx[12];
      load_point(a.pr.desired, i, Bd, bd, xd, ud);
      state_minus(x, xd, dx, nullptr);
      cx = quadratic_cost_state(p, dx);
    }
#pragma unroll
    for (int e = 0; e < 48; ++e) gK[e] = ldg_early(&a.pr.gK[row_index(i, e, 48, B, b)]);
    double d[12];
    state_minus(x, xbar, d, nullptr);  // (state - current_traj[i].state).coeffs()
    double u[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double Kd = gK[12 * j] * d[0];
#pragma unroll
      for (int s = 1; s < 12; ++s) Kd = QFMA(gK[12 * j + s], d[s], Kd);
      u[j] = QFMA(alpha, gk[j], ubar[j]) + Kd;
    }
    if (cand) store_point(cand, i, B, b, x, u);
    if (want_cost) {
      double du[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) du[j] = u[j] - ud[j];
      cost = cost + (cx + quadratic_cost_control(p, du));
    }
    // also after the last knot, as ilqr.hh:168 does; result unused
    if (GENERIC) gm::discrete_step_any(p, x, u);
    else discrete_step(p, x, x + 3, x + 7, u);
  }

  if (a.mode == MODE_FORWARD) {
    if (a.cost_out) a.cost_out[b] = cost;
    return;
  }
  if (a.mode == MODE_WIDE) {
    a.cost_out[size_t(wide_j) * B + b] = cost;
    return;
  }
  rollout_finish(p, a, b, alpha, cost);
}

// ---------------------------------------------------------------------------
// The same rollout with THREE ROLE-SPECIALISED WARPS per 32 problems (reference model only).
// Within one knot the reference's chain  x (-) xbar -> u -> acceleration  does not feed the pose update
// (explicit Euler: pose+ = pose o Exp(dt v) needs only x), and the cost needs only (x, u):
//   warp A "control": delta = x (-) xbar, u = ubar + alpha k + K delta | acceleration, v+ = v + dt a
//   warp B "pose":    pose+ = pose o Exp(dt v)                         |
//   warp C "cost":    dx = x (-) x_d, dx^T Q dx, store the state       | du^T R du, running cost, store u
// with two CTA barriers per knot (x_i ready | u_i ready) and 6 kB of shared memory for the hand-over.
// The critical path per knot drops from the sum of the three chains to the longest one plus the
// acceleration; every value is computed by the same instruction sequence as in k_rollout, so the two
// kernels agree bit for bit (tests/test_gpu_parity.py::test_rollout_kernels_agree).
// ---------------------------------------------------------------------------
// The body shared by k_rollout_ws and the persistent tail kernel: every thread of a 96-thread CTA calls it (it
// contains CTA barriers); `lane` selects the problem slot, `role` the warp's part, `b` the problem (idle slots
// shadow a valid problem with valid = false: they must reach the barriers and write nothing).
template <bool NC = true>
QD void rollout_ws_run(const DeviceParams &p, const RolloutArgs &a, const int lane, const int role, const bool valid,
                       const int b, double (*s_pose)[7][32], double (*s_vel)[32], double (*s_u)[32]) {
  const int B = a.pr.B, N = a.pr.N, Bd = a.pr.Bd;
  const int bd = (Bd == 1) ? 0 : b;
  const double *cur;
  double *cand;
  double alpha;
  if (a.mode == MODE_FORWARD) {
    cur = a.cur; cand = a.out; alpha = a.alpha_in[b];
  } else {
    const int s = a.st.sel[b];
    cur = s ? a.pr.buf1 : a.pr.buf0;
    cand = s ? a.pr.buf0 : a.pr.buf1;
    alpha = a.st.alpha[b];
  }
  const bool want_cost = (a.mode != MODE_FORWARD) || (a.cost_out != nullptr);

  double x[13], ubar[4];
  load_point(cur, 0, B, b, x, ubar);  // state = current_traj.front().state (ilqr.hh:156)
  if (role == 1) {
#pragma unroll
    for (int c = 0; c < 7; ++c) s_pose[0][c][lane] = x[c];
  } else if (role == 0) {
#pragma unroll
    for (int c = 0; c < 6; ++c) s_vel[c][lane] = x[7 + c];
  }
  double cost = 0.0, cx = 0.0;
  for (int i = 0; i < N; ++i) {
    __syncthreads();  // x_i = (s_pose[i & 1], s_vel) is complete
    if (role == 0) {
      double xbar[13], gk[4], gK[48];
#pragma unroll
      for (int c = 0; c < 13; ++c) xbar[c] = ldg_early<NC>(&cur[row_index(i, c, 17, B, b)]);
#pragma unroll
      for (int c = 0; c < 4; ++c) ubar[c] = ldg_early<NC>(&cur[row_index(i, 13 + c, 17, B, b)]);
#pragma unroll
      for (int e = 0; e < 4; ++e) gk[e] = ldg_early<NC>(&a.pr.gk[row_index(i, e, 4, B, b)]);
#pragma unroll
      for (int e = 0; e < 48; ++e) gK[e] = ldg_early<NC>(&a.pr.gK[row_index(i, e, 48, B, b)]);
#pragma unroll
      for (int c = 0; c < 7; ++c) x[c] = s_pose[i & 1][c][lane];
      double d[12];
      state_minus(x, xbar, d, nullptr);  // (state - current_traj[i].state).coeffs()
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        double Kd = gK[12 * j] * d[0];
#pragma unroll
        for (int s = 1; s < 12; ++s) Kd = QFMA(gK[12 * j + s], d[s], Kd);
        ubar[j] = QFMA(alpha, gk[j], ubar[j]) + Kd;  // from here on: the new control u_i
        s_u[j][lane] = ubar[j];
      }
    } else if (role == 1) {
      // pose part of discrete_step: pose+ = pose o Exp(dt * body velocity)
      double vel[6];
#pragma unroll
      for (int c = 0; c < 6; ++c) vel[c] = s_vel[c][lane];
      double R[9];
      quat_to_rot(x + 3, R);
      const double dv[3] = {p.dt * vel[0], p.dt * vel[1], p.dt * vel[2]};
      const double dw[3] = {p.dt * vel[3], p.dt * vel[4], p.dt * vel[5]};
      double Jl[9], te[3], qe[4], Rt[3], qn[4];
      so3_ljac(dw, Jl);
      m3_vec(Jl, dv, te);
      so3_exp(dw, qe);
      m3_vec(R, te, Rt);
      quat_compose(x + 3, qe, qn);
#pragma unroll
      for (int c = 0; c < 3; ++c) x[c] = Rt[c] + x[c];
#pragma unroll
      for (int c = 0; c < 4; ++c) x[3 + c] = qn[c];
#pragma unroll
      for (int c = 0; c < 7; ++c) s_pose[(i + 1) & 1][c][lane] = x[c];
    } else {
#pragma unroll
      for (int c = 0; c < 7; ++c) x[c] = s_pose[i & 1][c][lane];
#pragma unroll
      for (int c = 0; c < 6; ++c) x[7 + c] = s_vel[c][lane];
      if (valid) {
#pragma unroll
        for (int c = 0; c < 13; ++c) cand[row_index(i, c, 17, B, b)] = x[c];
      }
      if (want_cost) {
        double xd[13], dx[12];
#pragma unroll
        for (int c = 0; c < 13; ++c) xd[c] = a.pr.desired[row_index(i, c, 17, Bd, bd)];
        state_minus(x, xd, dx, nullptr);
        cx = quadratic_cost_state(p, dx);
      }
    }
    __syncthreads();  // u_i is in s_u; every reader of v_i is done
    if (role == 0) {
      // velocity part of discrete_step: v+ = v + dt * body_acceleration(x_i, u_i)
      double R[9], acc[6];
      quat_to_rot(x + 3, R);
      body_acceleration(p, R, x + 7, ubar, acc);
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        x[7 + c] = QFMA(p.dt, acc[c], x[7 + c]);
        s_vel[c][lane] = x[7 + c];
      }
    } else if (role == 2) {
      double u[4], du[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) u[j] = s_u[j][lane];
      if (valid) {
#pragma unroll
        for (int c = 0; c < 4; ++c) cand[row_index(i, 13 + c, 17, B, b)] = u[c];
      }
      if (want_cost) {
#pragma unroll
        for (int j = 0; j < 4; ++j) du[j] = u[j] - a.pr.desired[row_index(i, 13 + j, 17, Bd, bd)];
        cost = cost + (cx + quadratic_cost_control(p, du));
      }
    }
  }
  if (role != 2 || !valid) return;
  if (a.mode == MODE_FORWARD) {
    if (a.cost_out) a.cost_out[b] = cost;
    return;
  }
  rollout_finish(p, a, b, alpha, cost);
}
#ifndef QILQR_WS_MINB
#define QILQR_WS_MINB 2
#endif
__global__ void __launch_bounds__(96, QILQR_WS_MINB) k_rollout_ws(const __grid_constant__ DeviceParams p, const __grid_constant__ RolloutArgs a) {
  __shared__ double s_pose[2][7][32], s_vel[6][32], s_u[4][32];
  const int lane = threadIdx.x & 31, role = threadIdx.x >> 5;
  const int t0 = blockIdx.x * 32 + lane;
  const int t = t0 < a.n ? t0 : a.n - 1;
  const int b = a.list ? a.list[t] : t;
  const bool valid = t0 < a.n && (!a.check_phase || a.st.phase[b] == PHASE_SEARCH);
  if (!__syncthreads_or(valid)) return;  // nothing to roll out among these 32 problems
  rollout_ws_run(p, a, lane, role, valid, b, s_pose, s_vel, s_u);
}

// ---------------------------------------------------------------------------
// Eigen::LDLT<Matrix4d> semantics -- ilqr.hh:126-128: symmetric pivoting on the largest
// remaining |diagonal|, lower triangle only, D pseudo-inverse.  Eigen's kernel is
// left-looking, so the pivot order depends only on the ORIGINAL diagonal: all (predicated)
// symmetric swaps are applied up front and the factorisation itself is straight-line code
// on registers (no local memory, no branches).  Divisions by the pivots are done as
// multiplications by their reciprocals (<= 1 ulp away from Eigen's divisions).
// m: row-major 4x4, only the lower triangle is read.
// ---------------------------------------------------------------------------
struct Ldlt4 {
  double m[16];
  double l10, l20, l21, l30, l31, l32;
  double dinv[4];  // 1/d_i, or 0 where |d_i| <= DBL_MIN (Eigen's pseudo-inverse of D)
  double d[4];     // the pivots themselves (read by the STRICT build only, which divides)
  bool s01, s02, s03, s12, s13, s23;
};
QD void cswap(bool c, double &a, double &b) {
  const double t = a;
  a = c ? b : a;
  b = c ? t : b;
}
QD void ldlt4_compute(Ldlt4 &f) {
  const double *m = f.m;
  double a00 = m[0], a10 = m[4], a11 = m[5], a20 = m[8], a21 = m[9], a22 = m[10], a30 = m[12], a31 = m[13],
         a32 = m[14], a33 = m[15];
  {  // step 0: first maximum of |a00|,|a11|,|a22|,|a33|
    double best = fabs(a00);
    int p = 0;
    if (fabs(a11) > best) { best = fabs(a11); p = 1; }
    if (fabs(a22) > best) { best = fabs(a22); p = 2; }
    if (fabs(a33) > best) { p = 3; }
    f.s01 = p == 1; f.s02 = p == 2; f.s03 = p == 3;
    cswap(f.s01, a00, a11); cswap(f.s01, a20, a21); cswap(f.s01, a30, a31);
    cswap(f.s02, a00, a22); cswap(f.s02, a10, a21); cswap(f.s02, a30, a32);
    cswap(f.s03, a00, a33); cswap(f.s03, a10, a31); cswap(f.s03, a20, a32);
  }
  {  // step 1: first maximum of |a11|,|a22|,|a33|
    double best = fabs(a11);
    int p = 1;
    if (fabs(a22) > best) { best = fabs(a22); p = 2; }
    if (fabs(a33) > best) { p = 3; }
    f.s12 = p == 2; f.s13 = p == 3;
    cswap(f.s12, a10, a20); cswap(f.s12, a11, a22); cswap(f.s12, a31, a32);
    cswap(f.s13, a10, a30); cswap(f.s13, a11, a33); cswap(f.s13, a21, a32);
  }
  {  // step 2
    f.s23 = fabs(a33) > fabs(a22);
    cswap(f.s23, a20, a30); cswap(f.s23, a21, a31); cswap(f.s23, a22, a33);
  }
  const double kMin = 2.2250738585072014e-308;
  const double d0 = a00;
  const double r0 = 1.0 / d0;
  const bool v0 = fabs(d0) > 0.0;
  const double l10 = v0 ? QDIV(a10, d0, r0) : a10, l20 = v0 ? QDIV(a20, d0, r0) : a20, l30 = v0 ? QDIV(a30, d0, r0) : a30;
  double t0 = d0 * l10;
  const double d1 = QFMA(-l10, t0, a11);
  const double m21 = QFMA(-l20, t0, a21), m31 = QFMA(-l30, t0, a31);
  const double r1 = 1.0 / d1;
  const bool v1 = fabs(d1) > 0.0;
  const double l21 = v1 ? QDIV(m21, d1, r1) : m21, l31 = v1 ? QDIV(m31, d1, r1) : m31;
  t0 = d0 * l20;
  double t1 = d1 * l21;
  const double d2 = a22 - QFMA(l21, t1, l20 * t0);
  t0 = d0 * l30;
  const double t1b = d1 * l31;
  const double m32 = a32 - QFMA(l31, t1, l30 * (d0 * l20));
  const double r2 = 1.0 / d2;
  const bool v2 = fabs(d2) > 0.0;
  const double l32 = v2 ? QDIV(m32, d2, r2) : m32;
  const double t2 = d2 * l32;
  const double d3 = a33 - QFMA(l32, t2, QFMA(l31, t1b, l30 * t0));
  f.l10 = l10; f.l20 = l20; f.l21 = l21; f.l30 = l30; f.l31 = l31; f.l32 = l32;
  f.dinv[0] = fabs(d0) > kMin ? r0 : 0.0;
  f.dinv[1] = fabs(d1) > kMin ? r1 : 0.0;
  f.dinv[2] = fabs(d2) > kMin ? r2 : 0.0;
  f.dinv[3] = fabs(d3) > kMin ? 1.0 / d3 : 0.0;
  f.d[0] = d0; f.d[1] = d1; f.d[2] = d2; f.d[3] = d3;
}
QD void ldlt4_solve(const Ldlt4 &f, double *x /*4, in place*/) {
  cswap(f.s01, x[0], x[1]); cswap(f.s02, x[0], x[2]); cswap(f.s03, x[0], x[3]);
  cswap(f.s12, x[1], x[2]); cswap(f.s13, x[1], x[3]);
  cswap(f.s23, x[2], x[3]);
  x[1] = QFMA(-f.l10, x[0], x[1]);
  x[2] = QFMA(-f.l21, x[1], QFMA(-f.l20, x[0], x[2]));
  x[3] = QFMA(-f.l32, x[2], QFMA(-f.l31, x[1], QFMA(-f.l30, x[0], x[3])));
#pragma unroll
#ifdef QILQR_STRICT
  for (int i = 0; i < 4; ++i) x[i] = (f.dinv[i] != 0.0) ? x[i] / f.d[i] : 0.0;
#else
  for (int i = 0; i < 4; ++i) x[i] = x[i] * f.dinv[i];
#endif
  x[2] = QFMA(-f.l32, x[3], x[2]);
  x[1] = QFMA(-f.l31, x[3], QFMA(-f.l21, x[2], x[1]));
  x[0] = QFMA(-f.l30, x[3], QFMA(-f.l20, x[2], QFMA(-f.l10, x[1], x[0])));
  cswap(f.s23, x[2], x[3]);
  cswap(f.s13, x[1], x[3]); cswap(f.s12, x[1], x[2]);
  cswap(f.s03, x[0], x[3]); cswap(f.s02, x[0], x[2]); cswap(f.s01, x[0], x[1]);
}

// ---------------------------------------------------------------------------
// backwards_pass, one thread per problem ("T1").  12x12 matrices are stored as
// 4x4 grids of 3x3 blocks: block (I,J) at X + (4*I+J)*9.
// ---------------------------------------------------------------------------
#define QBLK(X, I, J) ((X) + (4 * (I) + (J)) * 9)
QD double &bel(double *X, int r, int c) { return X[(4 * (r / 3) + (c / 3)) * 9 + 3 * (r % 3) + (c % 3)]; }

struct BackwardArgs {
  Problem pr;
  SolveState st;
  const int *list;
  int n;
  int iter;             // the solver's super-step number (informational)
  int wide;             // 1: parallel step sizes after the first rejection; 2: from the first candidate on
  int solve_mode;       // 1: solve() bookkeeping (exit A); 0: plain backwards_pass
  const double *traj;   // solve_mode == 0 only
  double *terms_out;    // solve_mode == 0 only: [B][2]
};

// What solve() does with the result of backwards_pass for problem b (ilqr.hh:59-68), or the plain output of
// backwards_pass when !solve_mode.  One copy for every backward kernel.
QD void backward_finish(const DeviceParams &p, const BackwardArgs &a, int b, double QuTk, double kTQuuk) {
  if (!a.solve_mode) {
    a.terms_out[2 * size_t(b)] = QuTk;
    a.terms_out[2 * size_t(b) + 1] = kTQuuk;
    return;
  }
  const SolveState &st = a.st;
  st.qutk[b] = QuTk;
  st.ktquuk[b] = kTQuuk;
  const int it = st.bwd[b];  // this problem's iteration index i of ilqr.hh:58
  st.bwd[b] = it + 1;
  const double cost = st.cost[b];
  const double expected_new_cost = cost + (QuTk + kTQuuk / 2.0);  // ilqr.hh:64-65 with step = 1
  if (it > 0 && is_converged(p, cost, expected_new_cost)) {
    st.status[b] = QILQR_STATUS_CONVERGED_EXPECTED;  // ilqr.hh:66-68
    st.phase[b] = PHASE_DONE;
  } else if (it > 0 && p.ls_max_iters <= 0) {
    // line_search's loop (ilqr.hh:178-193) runs zero candidates and throws
    st.status[b] = QILQR_STATUS_LINE_SEARCH_FAILED;
    st.phase[b] = PHASE_DONE;
  } else {
    st.alpha[b] = 1.0;
    st.ls_iter[b] = 0;
    // the full step first, on its own: it is the one the search accepts nearly always (and iteration 0 takes it
    // unconditionally); with parallel step sizes (a.wide) a rejection sends the problem to PHASE_WIDE, unless
    // a.wide == 2 (QILQR_WIDE_FIRST=1: every search starts with a parallel round)
    st.phase[b] = (a.wide == 2 && it > 0) ? PHASE_WIDE : PHASE_SEARCH;
  }
}

// Cost derivatives at one knot (cost.hh:47-57) in block form.
//   J = blkdiag([[Ji, Qi], [0, Ji]], I6);  C.x = ((2 dx^T) Q) J;  C.xx = ((2 J^T) Q) J
QD void cost_derivatives(const DeviceParams &p, const double *dx, const double *Ji, const double *Qi,
                         double *Cx /*12*/, double *Cxx /*block-major 144*/) {
  // y = (2 dx)^T Q
  double y[12];
#pragma unroll
  for (int j = 0; j < 12; ++j) {
    double s = (2.0 * dx[0]) * p.Q[j];
#pragma unroll
    for (int i = 1; i < 12; ++i) s = QFMA(2.0 * dx[i], p.Q[12 * i + j], s);
    y[j] = s;
  }
  m3T_vec(Ji, y, Cx);
  m3T_vec(Qi, y, Cx + 3);
  m3T_vec_add(Ji, y + 3, Cx + 3);
#pragma unroll
  for (int j = 6; j < 12; ++j) Cx[j] = y[j];
  // P = (2 J^T) Q, row-major dense 12x12
  double P[144];
#pragma unroll
  for (int j = 0; j < 12; ++j) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      // row i (<3): sum_k 2*Ji[k][i] Q[k][j]
      P[12 * i + j] = QFMA(2.0 * Ji[6 + i], p.Q[24 + j], QFMA(2.0 * Ji[3 + i], p.Q[12 + j], (2.0 * Ji[i]) * p.Q[j]));
      // row 3+i: sum_k 2*Qi[k][i] Q[k][j] + sum_k 2*Ji[k][i] Q[3+k][j]
      double s = QFMA(2.0 * Qi[6 + i], p.Q[24 + j], QFMA(2.0 * Qi[3 + i], p.Q[12 + j], (2.0 * Qi[i]) * p.Q[j]));
      s = QFMA(2.0 * Ji[i], p.Q[36 + j], s);
      s = QFMA(2.0 * Ji[3 + i], p.Q[48 + j], s);
      s = QFMA(2.0 * Ji[6 + i], p.Q[60 + j], s);
      P[12 * (3 + i) + j] = s;
    }
#pragma unroll
    for (int i = 6; i < 12; ++i) P[12 * i + j] = 2.0 * p.Q[12 * i + j];
  }
  // Cxx = P J
#pragma unroll
  for (int i = 0; i < 12; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      bel(Cxx, i, j) = QFMA(P[12 * i + 2], Ji[6 + j], QFMA(P[12 * i + 1], Ji[3 + j], P[12 * i] * Ji[j]));
      double s = QFMA(P[12 * i + 2], Qi[6 + j], QFMA(P[12 * i + 1], Qi[3 + j], P[12 * i] * Qi[j]));
      s = QFMA(P[12 * i + 3], Ji[j], s);
      s = QFMA(P[12 * i + 4], Ji[3 + j], s);
      s = QFMA(P[12 * i + 5], Ji[6 + j], s);
      bel(Cxx, i, 3 + j) = s;
    }
#pragma unroll
    for (int j = 6; j < 12; ++j) bel(Cxx, i, j) = P[12 * i + j];
  }
}

#ifndef QILQR_USER_MODEL_TU  // (not needed in a run-time compiled user-model unit)
__global__ void __launch_bounds__(64) k_backward_t1(const __grid_constant__ DeviceParams p, const __grid_constant__ BackwardArgs a) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.n) return;
  const int b = a.list ? a.list[t] : t;
  const int B = a.pr.B, N = a.pr.N, Bd = a.pr.Bd;
  const int bd = (Bd == 1) ? 0 : b;
  const double *traj = a.solve_mode ? (a.st.sel[b] ? a.pr.buf1 : a.pr.buf0) : a.traj;

  double V[144], vx[12];
#pragma unroll
  for (int i = 0; i < 144; ++i) V[i] = 0.0;
#pragma unroll
  for (int i = 0; i < 12; ++i) vx[i] = 0.0;
  double QuTk = 0.0, kTQuuk = 0.0;

  for (int i = N - 1; i >= 0; --i) {
    double x[13], u[4], xd[13], ud[4];
    load_point(traj, i, B, b, x, u);
    load_point(a.pr.desired, i, Bd, bd, xd, ud);

    ABlocks A;
    dynamics_blocks(p, x + 3, x + 7, A);

    double dx[12], Jli[9], Ji[9], Qi[9];
    Angle ang;
    state_minus(x, xd, dx, Jli, &ang);
    se3_rjacinv_blocks(dx, Jli, ang, Ji, Qi);
    double Cx[12], Qxx[144];
    cost_derivatives(p, dx, Ji, Qi, Cx, Qxx);  // Qxx starts as C.xx
    double Cu[4];
    {
      double du[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) du[j] = u[j] - ud[j];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        double s = (2.0 * du[0]) * p.R[j];
#pragma unroll
        for (int l = 1; l < 4; ++l) s = QFMA(2.0 * du[l], p.R[4 * l + j], s);
        Cu[j] = s;
      }
    }

    // M = A^T V   (J_x^T v_xx, ilqr.hh:121,123)
    double M[144];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      m3_mulT(A.Re, QBLK(V, 0, c), QBLK(M, 0, c));
      m3_mulT(A.Te, QBLK(V, 0, c), QBLK(M, 1, c));
      m3_maddT(A.Re, QBLK(V, 1, c), QBLK(M, 1, c));
      m3_maddT(A.dG, QBLK(V, 2, c), QBLK(M, 1, c));
      m3_mulT(A.dJr, QBLK(V, 0, c), QBLK(M, 2, c));
#pragma unroll
      for (int e = 0; e < 9; ++e) QBLK(M, 2, c)[e] += QBLK(V, 2, c)[e];
      m3_mulT(A.dQb, QBLK(V, 0, c), QBLK(M, 3, c));
      m3_maddT(A.dJr, QBLK(V, 1, c), QBLK(M, 3, c));
      m3_maddT(A.Wd, QBLK(V, 3, c), QBLK(M, 3, c));
    }
    // Q.xx = C.xx + M A
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      double T[9];
      m3_mul(QBLK(M, r, 0), A.Re, T);
#pragma unroll
      for (int e = 0; e < 9; ++e) QBLK(Qxx, r, 0)[e] += T[e];
      m3_mul(QBLK(M, r, 0), A.Te, T);
      m3_madd(QBLK(M, r, 1), A.Re, T);
      m3_madd(QBLK(M, r, 2), A.dG, T);
#pragma unroll
      for (int e = 0; e < 9; ++e) QBLK(Qxx, r, 1)[e] += T[e];
      m3_mul(QBLK(M, r, 0), A.dJr, T);
#pragma unroll
      for (int e = 0; e < 9; ++e) QBLK(Qxx, r, 2)[e] += T[e] + QBLK(M, r, 2)[e];
      m3_mul(QBLK(M, r, 0), A.dQb, T);
      m3_madd(QBLK(M, r, 1), A.dJr, T);
      m3_madd(QBLK(M, r, 3), A.Wd, T);
#pragma unroll
      for (int e = 0; e < 9; ++e) QBLK(Qxx, r, 3)[e] += T[e];
    }
    // Q.xu = M B (C.xu = 0);  B rows 8..11 = p.Bu
    double Qxu[48];
#pragma unroll
    for (int s = 0; s < 12; ++s)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        double acc = bel(M, s, 8) * p.Bu[j];
#pragma unroll
        for (int c = 1; c < 4; ++c) acc = QFMA(bel(M, s, 8 + c), p.Bu[4 * c + j], acc);
        Qxu[4 * s + j] = acc;
      }
    // Q.uu = C.uu + (B^T V) B
    Ldlt4 f;
    double Quu[16];
    {
      double BtV[16];  // only columns 8..11 of B^T V are needed
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          double acc = p.Bu[j] * bel(V, 8, 8 + c);
#pragma unroll
          for (int r = 1; r < 4; ++r) acc = QFMA(p.Bu[4 * r + j], bel(V, 8 + r, 8 + c), acc);
          BtV[4 * j + c] = acc;
        }
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int l = 0; l < 4; ++l) {
          double acc = BtV[4 * j] * p.Bu[l];
#pragma unroll
          for (int c = 1; c < 4; ++c) acc = QFMA(BtV[4 * j + c], p.Bu[4 * c + l], acc);
          Quu[4 * j + l] = QFMA(2.0, p.R[4 * j + l], acc);
        }
      if (p.quu_reg != 0.0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) Quu[5 * j] += p.quu_reg;
      }
    }
    // Q.x = C.x + A^T v_x ;  Q.u = C.u + B^T v_x
    double Qx[12], Qu[4];
    {
      double T[12];
      m3T_vec(A.Re, vx, T);
      m3T_vec(A.Te, vx, T + 3);
      m3T_vec_add(A.Re, vx + 3, T + 3);
      m3T_vec_add(A.dG, vx + 6, T + 3);
      m3T_vec(A.dJr, vx, T + 6);
#pragma unroll
      for (int e = 0; e < 3; ++e) T[6 + e] += vx[6 + e];
      m3T_vec(A.dQb, vx, T + 9);
      m3T_vec_add(A.dJr, vx + 3, T + 9);
      m3T_vec_add(A.Wd, vx + 9, T + 9);
#pragma unroll
      for (int e = 0; e < 12; ++e) Qx[e] = Cx[e] + T[e];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        double acc = p.Bu[j] * vx[8];
#pragma unroll
        for (int r = 1; r < 4; ++r) acc = QFMA(p.Bu[4 * r + j], vx[8 + r], acc);
        Qu[j] = Cu[j] + acc;
      }
    }
    // gains: K = -Quu^-1 Qxu^T, k = -Quu^-1 Qu   (ilqr.hh:126-130)
#pragma unroll
    for (int e = 0; e < 16; ++e) f.m[e] = Quu[e];
    ldlt4_compute(f);
    double K[48], k[4];
#pragma unroll
    for (int s = 0; s < 12; ++s) {
      double rhs[4] = {Qxu[4 * s], Qxu[4 * s + 1], Qxu[4 * s + 2], Qxu[4 * s + 3]};
      ldlt4_solve(f, rhs);
#pragma unroll
      for (int j = 0; j < 4; ++j) K[12 * j + s] = -rhs[j];
    }
    {
      double rhs[4] = {Qu[0], Qu[1], Qu[2], Qu[3]};
      ldlt4_solve(f, rhs);
#pragma unroll
      for (int j = 0; j < 4; ++j) k[j] = -rhs[j];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) a.pr.gk[row_index(i, j, 4, B, b)] = k[j];
#pragma unroll
    for (int e = 0; e < 48; ++e) a.pr.gK[row_index(i, e, 48, B, b)] = K[e];

    // v_x = Q.x - (K^T Q.uu) k ; v_xx = Q.xx - (K^T Q.uu) K   (ilqr.hh:132-133)
    double KtQ[48];  // [s][l]
#pragma unroll
    for (int s = 0; s < 12; ++s)
#pragma unroll
      for (int l = 0; l < 4; ++l) {
        double acc = K[s] * Quu[l];
#pragma unroll
        for (int j = 1; j < 4; ++j) acc = QFMA(K[12 * j + s], Quu[4 * j + l], acc);
        KtQ[4 * s + l] = acc;
      }
#pragma unroll
    for (int s = 0; s < 12; ++s) {
      double acc = KtQ[4 * s] * k[0];
#pragma unroll
      for (int l = 1; l < 4; ++l) acc = QFMA(KtQ[4 * s + l], k[l], acc);
      vx[s] = Qx[s] - acc;
    }
#pragma unroll
    for (int s = 0; s < 12; ++s)
#pragma unroll
      for (int c = 0; c < 12; ++c) {
        double acc = KtQ[4 * s] * K[c];
#pragma unroll
        for (int l = 1; l < 4; ++l) acc = QFMA(KtQ[4 * s + l], K[12 * l + c], acc);
        bel(V, s, c) = bel(Qxx, s, c) - acc;
      }
    if (p.symmetrize_vxx) {
#pragma unroll
      for (int s = 0; s < 12; ++s)
#pragma unroll
        for (int c = s + 1; c < 12; ++c) {
          const double m = 0.5 * (bel(V, s, c) + bel(V, c, s));
          bel(V, s, c) = m;
          bel(V, c, s) = m;
        }
#pragma unroll
      for (int s = 0; s < 12; ++s) bel(V, s, s) = 0.5 * (bel(V, s, s) + bel(V, s, s));
    }
    // expected cost reduction terms (ilqr.hh:136-140)
    {
      double acc = Qu[0] * k[0];
#pragma unroll
      for (int j = 1; j < 4; ++j) acc = QFMA(Qu[j], k[j], acc);
      QuTk = QuTk + acc;
      double z[4];
#pragma unroll
      for (int l = 0; l < 4; ++l) {
        double s = k[0] * Quu[l];
#pragma unroll
        for (int j = 1; j < 4; ++j) s = QFMA(k[j], Quu[4 * j + l], s);
        z[l] = s;
      }
      double acc2 = z[0] * k[0];
#pragma unroll
      for (int l = 1; l < 4; ++l) acc2 = QFMA(z[l], k[l], acc2);
      kTQuuk = kTQuuk + acc2;
    }
  }

  backward_finish(p, a, b, QuTk, kTQuuk);
}

#endif  // QILQR_USER_MODEL_TU

// ---------------------------------------------------------------------------
// Parallel line search: after a MODE_WIDE round, pick the FIRST (largest) step size that passes the
// Armijo test of ilqr.hh:182-188 -- exactly the one the sequential search would accept.  The accepted
// problem moves to PHASE_SEARCH with alpha[b] set (the ordinary rollout then writes its trajectory and
// accepts it); otherwise the next round starts palpha steps further down, or the search is exhausted.
// `rollouts` and `ls_iter` count what the sequential search would have evaluated.
// ---------------------------------------------------------------------------
__global__ void k_select_alpha(const __grid_constant__ DeviceParams p, SolveState st, const int *list, int n, int B,
                               const double *wide_cost, int palpha, int epoch) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int b = list ? list[t] : t;
  if (st.phase[b] != PHASE_WIDE) return;  // `list` is the list of alive problems
  const double cur_cost = st.cost[b], qutk = st.qutk[b], ktq = st.ktquuk[b];
  double alpha = st.alpha[b];
  const int ls0 = st.ls_iter[b];
  int ls = ls0;
  bool last_finite = true;
  for (int j = 0; j < palpha; ++j) {
    if (ls >= p.ls_max_iters) break;
    const double cost = wide_cost[size_t(j) * B + b];
    last_finite = isfinite(cost);
    const double desired = p.desired_reduction_frac * (alpha * qutk + alpha * alpha * ktq / 2.0);
    if (cost - cur_cost < desired) {
      st.alpha[b] = alpha;
      st.ls_iter[b] = ls;
      st.rollouts[b] += ls - ls0;
      if (j == 0) {
        // the round's first step size: its trajectory is already in the candidate buffer -- accept it here, as
        // rollout_finish would after one more (identical) rollout
        st.rollouts[b] += 1;
        st.new_cost[b] = cost;
        accept_candidate(p, st, B, b, epoch, MODE_SOLVE, st.bwd[b] - 1, cur_cost, cost);
      } else {
        st.phase[b] = PHASE_SEARCH;  // rolled out (and accepted) at this alpha by the next launch
      }
      return;
    }
    alpha *= p.step_update;
    ++ls;
  }
  st.rollouts[b] += ls - ls0;
  st.alpha[b] = alpha;
  st.ls_iter[b] = ls;
  if (ls >= p.ls_max_iters) {
    st.status[b] = (last_finite) ? QILQR_STATUS_LINE_SEARCH_FAILED : QILQR_STATUS_NONFINITE;
    st.phase[b] = PHASE_DONE;
  }
}

// ---------------------------------------------------------------------------
// Ordered compaction of a problem list by phase (single block; the lists are
// at most a few 10^4 entries and this runs twice per solver iteration).
//   out_search <- entries with phase == SEARCH,  out_active <- phase == ACTIVE
//   counts[0] = |out_search|, counts[1] = |out_active|   (mapped pinned host memory)
// ---------------------------------------------------------------------------
// mask_s / mask_a: bit ph set = problems in phase ph go to out_search / out_active (a problem may go to both)
QD bool phase_in(int mask, int ph) { return ph >= 0 && ((mask >> ph) & 1); }
constexpr int kMaskSearch = 1 << PHASE_SEARCH, kMaskActive = 1 << PHASE_ACTIVE, kMaskWide = 1 << PHASE_WIDE;
constexpr int kMaskAlive = kMaskSearch | kMaskActive | kMaskWide;  // everything but PHASE_DONE
__global__ void __launch_bounds__(1024) k_compact(const int *list_in, int n_in, const int *phase, int *out_search,
                                                 int *out_active, volatile int *counts, int phase_s = kMaskSearch,
                                                 int phase_a = kMaskActive, int seq = 0) {
  __shared__ int warp_s[32], warp_a[32];
  __shared__ int base_s, base_a;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;  // any multiple of 32 threads
  if (tid == 0) { base_s = 0; base_a = 0; }
  __syncthreads();
  for (int start = 0; start < n_in; start += blockDim.x) {
    const int idx = start + tid;
    int b = -1, ph = -1;
    if (idx < n_in) {
      b = list_in ? list_in[idx] : idx;
      ph = phase[b];
    }
    const unsigned ms = __ballot_sync(0xffffffffu, phase_in(phase_s, ph));
    const unsigned ma = __ballot_sync(0xffffffffu, phase_in(phase_a, ph));
    if (lane == 0) { warp_s[wid] = __popc(ms); warp_a[wid] = __popc(ma); }
    __syncthreads();
    if (wid == 0) {
      int vs = lane < nw ? warp_s[lane] : 0, va = lane < nw ? warp_a[lane] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int ts = __shfl_up_sync(0xffffffffu, vs, o), ta = __shfl_up_sync(0xffffffffu, va, o);
        if (lane >= o) { vs += ts; va += ta; }
      }
      warp_s[lane] = vs;  // inclusive
      warp_a[lane] = va;
    }
    __syncthreads();
    const int off_s = base_s + (wid ? warp_s[wid - 1] : 0) + __popc(ms & ((1u << lane) - 1));
    const int off_a = base_a + (wid ? warp_a[wid - 1] : 0) + __popc(ma & ((1u << lane) - 1));
    if (phase_in(phase_s, ph)) out_search[off_s] = b;
    if (phase_in(phase_a, ph)) out_active[off_a] = b;
    __syncthreads();
    if (tid == 0) { base_s += warp_s[31]; base_a += warp_a[31]; }
    __syncthreads();
  }
  if (tid == 0) {
    counts[0] = base_s;
    counts[1] = base_a;
    __threadfence_system();
    if (seq) { counts[2] = seq; __threadfence_system(); }  // the host polls this word (mapped pinned memory)
  }
}

// Multi-block version for long lists: pass 1 counts per 1024-entry chunk, pass 2 scans the chunk
// counts (at most 2048 chunks) and scatters in order.
__global__ void __launch_bounds__(1024) k_compact_count(const int *list_in, int n_in, const int *phase, int *chunk_counts,
                                                       int phase_s, int phase_a) {
  __shared__ int ws[32], wa[32];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int idx = blockIdx.x * 1024 + tid;
  int ph = -1;
  if (idx < n_in) ph = phase[list_in ? list_in[idx] : idx];
  const unsigned ms = __ballot_sync(0xffffffffu, phase_in(phase_s, ph));
  const unsigned ma = __ballot_sync(0xffffffffu, phase_in(phase_a, ph));
  if (lane == 0) { ws[wid] = __popc(ms); wa[wid] = __popc(ma); }
  __syncthreads();
  if (wid == 0) {
    int vs = ws[lane], va = wa[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { vs += __shfl_xor_sync(0xffffffffu, vs, o); va += __shfl_xor_sync(0xffffffffu, va, o); }
    if (lane == 0) { chunk_counts[2 * blockIdx.x] = vs; chunk_counts[2 * blockIdx.x + 1] = va; }
  }
}
__global__ void __launch_bounds__(1024) k_compact_scatter(const int *list_in, int n_in, const int *phase,
                                                         const int *chunk_counts, int *out_search, int *out_active,
                                                         volatile int *counts, int phase_s, int phase_a, int seq) {
  __shared__ int ws[32], wa[32];
  __shared__ int base_s, base_a;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  // exclusive prefix of the chunk counts before this block (and the grand total for the last block)
  int ps = 0, pa = 0, ts = 0, ta = 0;
  for (int j = tid; j < int(gridDim.x); j += 1024) {
    const int cs = chunk_counts[2 * j], ca = chunk_counts[2 * j + 1];
    ts += cs; ta += ca;
    if (j < int(blockIdx.x)) { ps += cs; pa += ca; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ps += __shfl_xor_sync(0xffffffffu, ps, o); pa += __shfl_xor_sync(0xffffffffu, pa, o);
    ts += __shfl_xor_sync(0xffffffffu, ts, o); ta += __shfl_xor_sync(0xffffffffu, ta, o);
  }
  if (lane == 0) { ws[wid] = ps; wa[wid] = pa; }
  __syncthreads();
  if (tid == 0) {
    int a = 0, b2 = 0;
    for (int w = 0; w < 32; ++w) { a += ws[w]; b2 += wa[w]; }
    base_s = a; base_a = b2;
  }
  __syncthreads();
  // totals: every warp holds partial (ts, ta); reduce through shared memory in block 0 only
  __shared__ int tsw[32], taw[32];
  if (lane == 0) { tsw[wid] = ts; taw[wid] = ta; }
  __syncthreads();
  if (blockIdx.x == 0 && tid == 0) {
    int a = 0, b2 = 0;
    for (int w = 0; w < 32; ++w) { a += tsw[w]; b2 += taw[w]; }
    counts[0] = a; counts[1] = b2;
    __threadfence_system();
    if (seq) { counts[2] = seq; __threadfence_system(); }
  }
  const int idx = blockIdx.x * 1024 + tid;
  int b = -1, ph = -1;
  if (idx < n_in) { b = list_in ? list_in[idx] : idx; ph = phase[b]; }
  const unsigned ms = __ballot_sync(0xffffffffu, phase_in(phase_s, ph));
  const unsigned ma = __ballot_sync(0xffffffffu, phase_in(phase_a, ph));
  __syncthreads();
  if (lane == 0) { ws[wid] = __popc(ms); wa[wid] = __popc(ma); }
  __syncthreads();
  if (wid == 0) {
    int vs = ws[lane], va = wa[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int xs = __shfl_up_sync(0xffffffffu, vs, o), xa = __shfl_up_sync(0xffffffffu, va, o);
      if (lane >= o) { vs += xs; va += xa; }
    }
    ws[lane] = vs; wa[lane] = va;
  }
  __syncthreads();
  const int off_s = base_s + (wid ? ws[wid - 1] : 0) + __popc(ms & ((1u << lane) - 1));
  const int off_a = base_a + (wid ? wa[wid - 1] : 0) + __popc(ma & ((1u << lane) - 1));
  if (phase_in(phase_s, ph)) out_search[off_s] = b;
  if (phase_in(phase_a, ph)) out_active[off_a] = b;
}

// ---------------------------------------------------------------------------
// Solve set-up / tear-down
// ---------------------------------------------------------------------------
// max_iters: the loop bound of ilqr.hh:58 for i = 0 (a non-positive or NaN bound runs no iteration at all)
__global__ void k_init_state(SolveState st, int B, double max_iters) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  st.new_cost[b] = 0.0; st.qutk[b] = 0.0; st.ktquuk[b] = 0.0; st.alpha[b] = 1.0;
  st.ls_iter[b] = 0; st.status[b] = QILQR_STATUS_NOT_RUN; st.bwd[b] = 0; st.rollouts[b] = 0;
  st.ndebug[b] = 0; st.sel[b] = 0; st.phase[b] = (0.0 < max_iters) ? PHASE_ACTIVE : PHASE_DONE;
  st.accepted_iter[b] = -1;
}
// Problems still running after the loop bound hit max_iters (ilqr.hh:58,86); writes results.
// totals[0] += sum of backward passes, totals[1] += sum of rollouts (the solver's statistics).
__global__ void k_finalize(SolveState st, int B, qilqr_result_t *res, unsigned long long *totals) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  int nb = 0, nr = 0;
  if (b < B) {
    int s = st.status[b];
    if (s == QILQR_STATUS_NOT_RUN) { s = QILQR_STATUS_MAX_ITERS; st.status[b] = s; }
    nb = st.bwd[b];
    nr = st.rollouts[b];
    if (res) {
      res[b].status = s;
      res[b].backward_passes = nb;
      res[b].rollouts = nr;
      res[b].num_debug = st.ndebug[b];
      res[b].final_cost = st.cost[b];
    }
  }
  if (!totals) return;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    nb += __shfl_xor_sync(0xffffffffu, nb, o);
    nr += __shfl_xor_sync(0xffffffffu, nr, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&totals[0], (unsigned long long)nb);
    atomicAdd(&totals[1], (unsigned long long)nr);
  }
}
// Copy the current trajectory of the problems whose result sits in buf1 back to buf0.
__global__ void k_collect(Problem pr, const int *sel) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= pr.B || !sel[b]) return;
  const int rows = pr.N * 17;
  for (int r = blockIdx.y; r < rows; r += gridDim.y) pr.buf0[size_t(r) * pr.B + b] = pr.buf1[size_t(r) * pr.B + b];
}
// Tail compaction.  Once only a few problems of a large batch are still iterating, their data sits 8*B bytes
// apart in every row of the SoA arrays (one 32-byte sector and, across rows, one page-table entry per value).
// k_tail_gather moves the m remaining problems (map[t], ordered) into a dense mini-batch of pitch m -- current
// trajectory, per-problem desired trajectory if any, solver state, cost history -- the remaining iterations run
// there with unchanged kernels, and k_tail_scatter puts trajectory, last gains, state and history back.
// Same arithmetic on the same values: results are bit-identical with and without the compaction.
__global__ void k_tail_gather(Problem big, SolveState sb, Problem mini, SolveState sm, double *mini_desired,
                              const int *map, int m) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m) return;
  const int b = map[t], B = big.B;
  const double *src = sb.sel[b] ? big.buf1 : big.buf0;
  const int rows = big.N * 17;
  for (int r = blockIdx.y; r < rows; r += gridDim.y) {
    mini.buf0[size_t(r) * m + t] = src[size_t(r) * B + b];
    if (mini_desired) mini_desired[size_t(r) * m + t] = big.desired[size_t(r) * B + b];
  }
  for (int r = blockIdx.y; r < sb.hist_cap; r += gridDim.y)
    if (sb.cost_hist) sm.cost_hist[size_t(r) * m + t] = sb.cost_hist[size_t(r) * B + b];
  if (sb.phase[b] != PHASE_ACTIVE) {
    // in the middle of its line search: the gains of its last backward pass are still needed
    for (int r = blockIdx.y; r < big.N * 4; r += gridDim.y) mini.gk[size_t(r) * m + t] = big.gk[size_t(r) * B + b];
    for (int r = blockIdx.y; r < big.N * 48; r += gridDim.y) mini.gK[size_t(r) * m + t] = big.gK[size_t(r) * B + b];
  }
  if (blockIdx.y != 0) return;
  sm.cost[t] = sb.cost[b]; sm.new_cost[t] = sb.new_cost[b]; sm.qutk[t] = sb.qutk[b]; sm.ktquuk[t] = sb.ktquuk[b];
  sm.alpha[t] = sb.alpha[b];
  sm.ls_iter[t] = sb.ls_iter[b]; sm.status[t] = sb.status[b]; sm.bwd[t] = sb.bwd[b]; sm.rollouts[t] = sb.rollouts[b];
  sm.ndebug[t] = sb.ndebug[b]; sm.phase[t] = sb.phase[b]; sm.accepted_iter[t] = sb.accepted_iter[b];
  sm.sel[t] = 0;  // the gathered trajectory is the mini-batch's buffer 0
}
__global__ void k_tail_scatter(Problem big, SolveState sb, Problem mini, SolveState sm, const int *map, int m) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m) return;
  const int b = map[t], B = big.B, N = big.N;
  double *dst = sb.sel[b] ? big.buf1 : big.buf0;  // the big batch's current buffer of b stays the current one
  const double *src = sm.sel[t] ? mini.buf1 : mini.buf0;
  for (int r = blockIdx.y; r < N * 17; r += gridDim.y) dst[size_t(r) * B + b] = src[size_t(r) * m + t];
  for (int r = blockIdx.y; r < N * 4; r += gridDim.y) big.gk[size_t(r) * B + b] = mini.gk[size_t(r) * m + t];
  for (int r = blockIdx.y; r < N * 48; r += gridDim.y) big.gK[size_t(r) * B + b] = mini.gK[size_t(r) * m + t];
  for (int r = blockIdx.y; r < sb.hist_cap; r += gridDim.y)
    if (sb.cost_hist) sb.cost_hist[size_t(r) * B + b] = sm.cost_hist[size_t(r) * m + t];
  if (blockIdx.y != 0) return;
  sb.cost[b] = sm.cost[t]; sb.new_cost[b] = sm.new_cost[t]; sb.qutk[b] = sm.qutk[t]; sb.ktquuk[b] = sm.ktquuk[t];
  sb.alpha[b] = sm.alpha[t];
  sb.ls_iter[b] = sm.ls_iter[t]; sb.status[b] = sm.status[t]; sb.bwd[b] = sm.bwd[t]; sb.rollouts[b] = sm.rollouts[t];
  sb.ndebug[b] = sm.ndebug[t]; sb.phase[b] = sm.phase[t]; sb.accepted_iter[b] = sm.accepted_iter[t];
}
// ILQRDebug capture (ilqr.hh:78-80): copy the trajectory accepted in iteration `iter`.
__global__ void k_debug_capture(Problem pr, SolveState st, const int *list, int n, int iter, double *debug /*[cap][N*17][B]*/, int cap) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int b = list ? list[t] : t;
  if (st.accepted_iter[b] != iter) return;
  const int slot = st.ndebug[b] - 1;
  if (slot < 0 || slot >= cap) return;
  const double *src = st.sel[b] ? pr.buf1 : pr.buf0;
  const int rows = pr.N * 17;
  for (int r = blockIdx.y; r < rows; r += gridDim.y)
    debug[(size_t(slot) * rows + r) * pr.B + b] = src[size_t(r) * pr.B + b];
}

// Sampled ILQRDebug stream (ilqr.hh:78-80 at batch scale): for the problems of `sample` (any order), keep the
// trajectories of the iterations i with i % every == 0 in a ring of `ring` slots per problem -- the last `ring`
// sampled iterations survive.  One block row per sampled problem; its ring is a contiguous array-of-structs block
//   traj [S][ring][N][18]   (column 0 = time_s, taken from `time_aos` [B][N][18] when given, else the knot index)
//   iters / costs [S][ring] the iteration index and new_cost of each slot (-1 / 0 where empty), count [S] the number
//                           of sampled iterations so far (count > ring: the ring has wrapped).
struct DebugRing {
  const int *sample;   // [S] problem indices
  int S, ring, every;
  double *traj;
  int *iters;
  double *costs;
  int *count;
  const double *time_aos;
};
__global__ void k_debug_ring_capture(Problem pr, SolveState st, DebugRing dr, int epoch) {
  const int s = blockIdx.x;
  if (s >= dr.S) return;
  const int b = dr.sample[s];
  if (b < 0 || b >= pr.B || st.accepted_iter[b] != epoch) return;  // no trajectory accepted in this super-step
  const int it = st.ndebug[b] - 1;  // index of the iteration that has just completed (ILQRDebug entry number)
  if (it < 0 || it % dr.every != 0) return;
  const int k = dr.count[s];  // (every thread of the block reads the same value; thread 0 bumps it at the end)
  const int slot = k % dr.ring;
  const double *src = st.sel[b] ? pr.buf1 : pr.buf0;
  double *dst = dr.traj + (size_t(s) * dr.ring + slot) * pr.N * 18;
  for (int e = threadIdx.x; e < pr.N * 18; e += blockDim.x) {
    const int i = e / 18, c = e % 18;
    dst[e] = (c == 0) ? (dr.time_aos ? dr.time_aos[(size_t(b) * pr.N + i) * 18] : double(i))
                      : src[(size_t(i) * 17 + (c - 1)) * pr.B + b];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    dr.iters[size_t(s) * dr.ring + slot] = it;
    dr.costs[size_t(s) * dr.ring + slot] = st.cost[b];
    dr.count[s] = k + 1;
  }
}

// ---------------------------------------------------------------------------
// AoS [B][N][18] <-> SoA [N][17][B] transposition through shared memory.
// Block: 32 problems x PACK_K consecutive knots (PACK_K * 18 contiguous doubles per problem on the AoS side, 32
// contiguous problems per row on the SoA side); grid.x = ceil(B/32), grid.y strides over groups of knots.
// ---------------------------------------------------------------------------
constexpr int PACK_K = 4, PACK_W = PACK_K * 18, PACK_THREADS = 256;
__global__ void __launch_bounds__(PACK_THREADS) k_pack(const double *aos, double *soa, int B, int N) {
  __shared__ double tile[32][PACK_W + 1];
  const int b0 = blockIdx.x * 32, tid = threadIdx.x;
  for (int i0 = blockIdx.y * PACK_K; i0 < N; i0 += gridDim.y * PACK_K) {
    const int w = min(PACK_K, N - i0) * 18;
    for (int e = tid; e < 32 * PACK_W; e += PACK_THREADS) {
      const int bl = e / PACK_W, col = e % PACK_W;
      if (col < w && b0 + bl < B) tile[bl][col] = aos[(size_t(b0 + bl) * N + i0) * 18 + col];
    }
    __syncthreads();
    for (int e = tid; e < 32 * PACK_W; e += PACK_THREADS) {
      const int bl = e % 32, col = e / 32, c = col % 18;
      if (c >= 1 && col < w && b0 + bl < B)
        soa[(size_t(i0 + col / 18) * 17 + (c - 1)) * B + b0 + bl] = tile[bl][col];
    }
    __syncthreads();
  }
}
// Leaves column 0 (time_s) of the AoS buffer untouched.
__global__ void __launch_bounds__(PACK_THREADS) k_unpack(const double *soa, double *aos, int B, int N) {
  __shared__ double tile[32][PACK_W + 1];
  const int b0 = blockIdx.x * 32, tid = threadIdx.x;
  for (int i0 = blockIdx.y * PACK_K; i0 < N; i0 += gridDim.y * PACK_K) {
    const int w = min(PACK_K, N - i0) * 18;
    for (int e = tid; e < 32 * PACK_W; e += PACK_THREADS) {
      const int bl = e % 32, col = e / 32, c = col % 18;
      if (c >= 1 && col < w && b0 + bl < B)
        tile[bl][col] = soa[(size_t(i0 + col / 18) * 17 + (c - 1)) * B + b0 + bl];
    }
    __syncthreads();
    for (int e = tid; e < 32 * PACK_W; e += PACK_THREADS) {
      const int bl = e / PACK_W, col = e % PACK_W;
      if (col % 18 >= 1 && col < w && b0 + bl < B) aos[(size_t(b0 + bl) * N + i0) * 18 + col] = tile[bl][col];
    }
    __syncthreads();
  }
}
// Generic [B][N][W] <-> [N][W][B] (gains), one thread per element, writes coalesced.
__global__ void k_transpose_bnw_to_nwb(const double *in, double *out, int B, int N, int W) {
  const size_t total = size_t(B) * N * W;
  for (size_t e = size_t(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += size_t(gridDim.x) * blockDim.x) {
    const int b = int(e % B);
    const size_t r = e / B;  // i*W + w
    out[e] = in[size_t(b) * N * W + r];
  }
}
__global__ void k_transpose_nwb_to_bnw(const double *in, double *out, int B, int N, int W) {
  const size_t total = size_t(B) * N * W;
  for (size_t e = size_t(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += size_t(gridDim.x) * blockDim.x) {
    const int b = int(e % B);
    const size_t r = e / B;
    out[size_t(b) * N * W + r] = in[e];
  }
}

// CostFunction::operator() with dense derivative outputs (cost.hh:36-61), AoS, for the API.
__global__ void k_api_cost(const __grid_constant__ DeviceParams p, int B, const double *x, const double *u,
                           const double *x_d, const double *u_d, double *cost, double *C_x, double *C_u,
                           double *C_xx, double *C_uu, double *C_xu) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double xs[13], xd[13], du[4], dx[12], Jli[9];
  for (int i = 0; i < 13; ++i) { xs[i] = x[size_t(b) * 13 + i]; xd[i] = x_d[size_t(b) * 13 + i]; }
  for (int i = 0; i < 4; ++i) du[i] = u[size_t(b) * 4 + i] - u_d[size_t(b) * 4 + i];
  Angle ang;
  state_minus(xs, xd, dx, Jli, &ang);
  cost[b] = quadratic_cost(p, dx, du);
  if (C_x || C_xx) {
    double Ji[9], Qi[9], Cx[12], Cxx[144];
    se3_rjacinv_blocks(dx, Jli, ang, Ji, Qi);
    cost_derivatives(p, dx, Ji, Qi, Cx, Cxx);
    if (C_x) for (int i = 0; i < 12; ++i) C_x[size_t(b) * 12 + i] = Cx[i];
    if (C_xx)
      for (int r = 0; r < 12; ++r)
        for (int c = 0; c < 12; ++c) C_xx[size_t(b) * 144 + 12 * r + c] = bel(Cxx, r, c);
  }
  if (C_u)
    for (int j = 0; j < 4; ++j) {
      double s = (2.0 * du[0]) * p.R[j];
      for (int l = 1; l < 4; ++l) s = QFMA(2.0 * du[l], p.R[4 * l + j], s);
      C_u[size_t(b) * 4 + j] = s;
    }
  if (C_uu) for (int e = 0; e < 16; ++e) C_uu[size_t(b) * 16 + e] = 2.0 * p.R[e];
  if (C_xu) for (int e = 0; e < 48; ++e) C_xu[size_t(b) * 48 + e] = 0.0;
}

// Receding-horizon step (BASELINE config 5): apply the first control of the solution to the plant
// (one discrete_dynamics step, optionally with an additive body-velocity disturbance), then shift the
// solution by one knot as the warm start of the next solve (last knot duplicated) and set its first
// state to the new plant state.  solve() uses initial_traj.front().state as x0 (ilqr.hh:156).
__global__ void __launch_bounds__(128)
k_mpc_advance(const __grid_constant__ DeviceParams p, double *traj, double *plant /*[13][B]*/,
              const double *disturbance /*[6][B] or nullptr*/, double *applied_u /*[4][B] or nullptr*/, int B, int N) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double x[13], u[4];
#pragma unroll
  for (int c = 0; c < 13; ++c) x[c] = plant[size_t(c) * B + b];
#pragma unroll
  for (int c = 0; c < 4; ++c) u[c] = traj[row_index(0, 13 + c, 17, B, b)];
  gm::discrete_step_any(p, x, u);
  if (disturbance) {
#pragma unroll
    for (int c = 0; c < 6; ++c) x[7 + c] += disturbance[size_t(c) * B + b];
  }
#pragma unroll
  for (int c = 0; c < 13; ++c) plant[size_t(c) * B + b] = x[c];
  if (applied_u) {
#pragma unroll
    for (int c = 0; c < 4; ++c) applied_u[size_t(c) * B + b] = u[c];
  }
  for (int i = 0; i + 1 < N; ++i)
    for (int c = 0; c < 17; ++c) traj[row_index(i, c, 17, B, b)] = traj[row_index(i + 1, c, 17, B, b)];
#pragma unroll
  for (int c = 0; c < 13; ++c) traj[row_index(0, c, 17, B, b)] = x[c];
}

// Open-loop rollout under a constant control from per-problem initial states.
__global__ void __launch_bounds__(128)
k_rollout_constant(const __grid_constant__ DeviceParams p, const double *x0 /*[13][B]*/, double u0, double u1,
                   double u2, double u3, double *traj, int B, int N) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double x[13];
  const double u[4] = {u0, u1, u2, u3};
#pragma unroll
  for (int c = 0; c < 13; ++c) x[c] = x0[size_t(c) * B + b];
  for (int i = 0; i < N; ++i) {
    store_point(traj, i, B, b, x, u);
    gm::discrete_step_any(p, x, u);
  }
}

// Open-loop rollout of array-of-structs initial states under given controls (the usual "nominal controls" initial
// guess of an iLQR solver): x0 [B][13], controls [Bc][N][4] with Bc = 1 (shared) or B -> trajectory SoA [N][17][B].
__global__ void __launch_bounds__(128)
k_rollout_controls(const __grid_constant__ DeviceParams p, const double *x0_aos, const double *controls, int Bc,
                   double *traj, int B, int N) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double x[13];
#pragma unroll
  for (int c = 0; c < 13; ++c) x[c] = x0_aos[size_t(b) * 13 + c];
  const double *uc = controls + (Bc == 1 ? size_t(0) : size_t(b) * N * 4);
  for (int i = 0; i < N; ++i) {
    const double u[4] = {uc[4 * i], uc[4 * i + 1], uc[4 * i + 2], uc[4 * i + 3]};
    store_point(traj, i, B, b, x, u);
    gm::discrete_step_any(p, x, u);
  }
}
// time_s column of an array-of-structs trajectory batch: t_i = i * dt
__global__ void k_fill_time(double *aos, int B, int N, double dt) {
  const size_t e = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= size_t(B) * N) return;
  aos[e * 18] = double(e % N) * dt;
}
// controls of a structure-of-arrays trajectory batch -> [B][N][4]
__global__ void k_extract_controls(const double *soa, double *out, int B, int N) {
  const size_t e = size_t(blockIdx.x) * blockDim.x + threadIdx.x;  // over N * 4 * B, problem fastest
  if (e >= size_t(B) * N * 4) return;
  const int b = int(e % B);
  const size_t r = e / B;  // i * 4 + j
  const int i = int(r / 4), j = int(r % 4);
  out[(size_t(b) * N + i) * 4 + j] = soa[(size_t(i) * 17 + 13 + j) * B + b];
}

}  // namespace qilqr
