/* =============================================================================
 * include/qilqr.h -- C ABI of libqilqr_b200.so: a batched, B200-native (sm_100a)
 * replacement for the hot path of nitishthatte/QuadrotorILQR
 * (ILQR<QuadrotorModel>::solve and the public methods under it).
 *
 * The reference has no FFI for this path; its only boundary is the C++ template
 * ILQR<ModelT> (src/ilqr.hh:25-206) and the pybind module over it
 * (src/quadrotor_ilqr_binding.cc:20-49).  Each entry point below names the
 * reference interface it replaces.  One call = a BATCH of independent problems
 * (the reference solves one per call); batch = 1 reproduces the reference call.
 *
 * Conventions
 *   - all numbers are IEEE double; matrices row-major;
 *   - state      = 13 doubles  t(3), unit quaternion (x,y,z,w), body velocity
 *                  lin(3) ang(3)          (QuadrotorModel::State, quadrotor_model.hh:11-14;
 *                  manif/Eigen coefficient order)
 *   - tangent    = 12 doubles  pose-lin(3) pose-ang(3) vel-lin(3) vel-ang(3)
 *                  (QuadrotorModel::StateBlocks, quadrotor_model.hh:29-37)
 *   - traj point = 18 doubles  time_s, state(13), control(4)   (trajectory.hh:9-14)
 *   - "_host" entry points take HOST pointers in array-of-structs order
 *     [batch][knot][18]; they copy to the GPU, run the CUDA kernels and copy
 *     back.  "_device" entry points take DEVICE pointers in the resident
 *     structure-of-arrays order described at qilqr_solve_device.
 *   - every function returns a qilqr_error_t (0 = ok).  Nothing here ever
 *     computes on the CPU: without a CUDA device qilqr_create fails.
 * ========================================================================== */
#ifndef QILQR_H_
#define QILQR_H_

#ifdef __CUDACC_RTC__ /* compiled by NVRTC as part of a user-model translation unit: no host headers */
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
#else
#include <stddef.h>
#include <stdint.h>
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define QILQR_STATE_DIM 12   /* QuadrotorModel::STATE_DIM   quadrotor_model.hh:16 */
#define QILQR_CONTROL_DIM 4  /* QuadrotorModel::CONTROL_DIM quadrotor_model.hh:40 */
#define QILQR_STATE_STORAGE 13
#define QILQR_POINT_STORAGE 18
#define QILQR_SOA_ROWS 17    /* state(13) + control(4); time_s stays on the host */

typedef enum {
  QILQR_OK = 0,
  QILQR_ERR_INVALID_ARGUMENT = 1,
  QILQR_ERR_INERTIA_NOT_PD = 2,   /* std::runtime_error, quadrotor_model.cc:21-24 */
  QILQR_ERR_OUT_OF_RANGE = 3,     /* std::out_of_range, cost.hh:39-40 (trajectory longer than desired) */
  QILQR_ERR_NO_DEVICE = 4,        /* no CUDA device / wrong architecture: there is no CPU fallback */
  QILQR_ERR_CUDA = 5,
  QILQR_ERR_OUT_OF_MEMORY = 6,
  QILQR_ERR_LINE_SEARCH = 7       /* std::runtime_error, ilqr.hh:191-193 (single-problem helpers only) */
} qilqr_error_t;

/* How one problem's solve() ended (ilqr.hh:53-87). */
typedef enum {
  QILQR_STATUS_NOT_RUN = 0,
  QILQR_STATUS_CONVERGED_EXPECTED = 1, /* exit at ilqr.hh:66-68 (predicted reduction small) */
  QILQR_STATUS_CONVERGED_ACTUAL = 2,   /* exit at ilqr.hh:82-84 (actual reduction small)    */
  QILQR_STATUS_MAX_ITERS = 3,          /* loop bound, ilqr.hh:58,86                          */
  QILQR_STATUS_LINE_SEARCH_FAILED = 4, /* the reference throws here, ilqr.hh:191-193          */
  QILQR_STATUS_NONFINITE = 5           /* same exit (the reference throws the same error), but the
                                          last candidate cost was NaN/Inf: a numerical blow-up,
                                          which makes every Armijo comparison false            */
} qilqr_status_t;

/* QuadrotorModel constructor arguments (quadrotor_model.hh:8-10). */
typedef struct {
  double mass_kg;
  double inertia[9];
  double arm_length_m;
  double torque_to_thrust_ratio_m;
  double g_mpss;
} qilqr_model_t;

/* ILQROptions, field for field (ilqr_options.hh:4-22), followed by batch-only
 * extensions whose zero value reproduces the reference. */
typedef struct {
  /* LineSearchParams */
  double step_update;
  double desired_reduction_frac;
  int32_t line_search_max_iters;
  int32_t populate_debug; /* bool ILQROptions::populate_debug */
  /* ConvergenceCriteria */
  double rtol;
  double atol;
  double max_iters; /* a double in the reference (ilqr_options.hh:14) */
  /* extensions */
  int32_t symmetrize_vxx;      /* V_xx <- (V_xx + V_xx^T)/2 after ilqr.hh:133 (needed for long horizons) */
  int32_t num_parallel_alphas; /* 0/1: sequential backtracking; P>1: P step sizes per round as
                                  parallel rollouts, first passing one wins (same result) */
  double quu_regularization;   /* Q_uu += mu*I before ilqr.hh:126; 0 = reference */
} qilqr_options_t;

/* Per-problem outcome of a solve. */
typedef struct {
  int32_t status;          /* qilqr_status_t */
  int32_t backward_passes; /* calls of backwards_pass (ilqr.hh:59) */
  int32_t rollouts;        /* calls of forward_sim (ilqr.hh:71,180) */
  int32_t num_debug;       /* completed iterations = ILQRDebug entries (ilqr.hh:78-80) */
  double final_cost;       /* cost of the returned trajectory */
} qilqr_result_t;

typedef struct qilqr_solver qilqr_solver_t;

/* ILQR<QuadrotorModel>{model, CostFunction{Q, R, .}, dt_s, options} (ilqr.hh:27-32,
 * quadrotor_ilqr_binding.cc:20-32).  The desired trajectory, which the reference's
 * CostFunction owns (cost.hh:30-34), is an argument of each call instead so that
 * every problem of a batch may have its own.  `device` = CUDA ordinal. */
int qilqr_create(const qilqr_model_t *model, const double *Q /*[144]*/, const double *R /*[16]*/,
                 double dt_s, const qilqr_options_t *options, int device, qilqr_solver_t **out);
void qilqr_destroy(qilqr_solver_t *solver);
int qilqr_set_options(qilqr_solver_t *solver, const qilqr_options_t *options);

/* The ModelT concept (ilqr.hh:25-44: ILQR<ModelT> sees only discrete_dynamics(x, u, dt, diffs*) with
 * dense J_x, J_u, and minus()).  The reference ships one model; these flags select a second dynamics
 * function on the same state manifold, solved by model-agnostic kernels that consume dense Jacobians
 * (not in the reference; defined by QuadrotorModelVariant in oracle/qilqr_oracle.hpp):
 *   QILQR_MODEL_RK4       RK4 over continuous_dynamics, the scheme commented out at
 *                         quadrotor_model.cc:51-63, with the chain rule through its stages
 *   QILQR_MODEL_CORIOLIS  adds the transport term -omega x v to the body linear acceleration
 *   QILQR_MODEL_GENERIC   runs the model-agnostic kernels even for the reference model (cross-check:
 *                         same results as the quadrotor-specific kernels to rounding)
 * 0 (the default) is the reference's QuadrotorModel on the quadrotor-specific kernels. */
enum { QILQR_MODEL_REFERENCE = 0, QILQR_MODEL_RK4 = 1, QILQR_MODEL_CORIOLIS = 2, QILQR_MODEL_GENERIC = 4 };
int qilqr_set_model_variant(qilqr_solver_t *solver, int model_flags);
/* A USER-SUPPLIED MODEL behind the ModelT concept (ilqr.hh:25-44, 110-112: the solver sees a model only through
 * discrete_dynamics(x, u, dt, diffs*)).  `cuda_source` is CUDA C++ that defines
 *
 *   extern "C" __device__ void qilqr_user_discrete_dynamics(const double *params, const double *x, const double *u,
 *                                                           double dt, double *x_next, double *J_x, double *J_u);
 *
 * (x, x_next: 13 doubles; u: 4; J_x: 12x12 and J_u: 12x4 row-major, nullptr when only the step is wanted)
 * on the state manifold of QuadrotorModel::State (SE(3) x R^6; Jacobians with respect to right-plus perturbations,
 * as QuadrotorModel::DynamicsDifferentials, quadrotor_model.hh:42-45).  It is compiled at run time (NVRTC, sm_100a,
 * -fmad=false) together with the model-agnostic kernels -- dense linearisation, rollout -- and may use this
 * library's device functions (so3_exp, se3_plus_blocks, quat_compose, ...: qilqr_device.cuh is included ahead of
 * it).  `params` (at most 64 doubles) is handed to every call.  Every later solve / forward_sim / backwards_pass of
 * this handle uses the model; compile errors come back through qilqr_last_error_message.  Needs libnvrtc.so.12 and
 * the csrc/ directory next to the shared library (or QILQR_CSRC_DIR). */
int qilqr_set_user_model(qilqr_solver_t *solver, const char *cuda_source, const double *params, int n_params);
const char *qilqr_error_string(int err);
/* How this build rounds: "production: explicit fused multiply-adds, reciprocal multiplies" or
 * "strict: no fused multiply-adds, true divisions (-DQILQR_STRICT)"; both are compiled with -fmad=false. */
const char *qilqr_build_info(void);
const char *qilqr_last_error_message(const qilqr_solver_t *solver);
/* Number of this library's kernels launched by the solver so far (for bench.py's gpu_launches). */
int64_t qilqr_kernel_launch_count(const qilqr_solver_t *solver);
/* The CUDA stream (cudaStream_t) the solver launches on, for event timing by the caller. */
void *qilqr_stream(const qilqr_solver_t *solver);

/* ---------------------------------------------------------------------------
 * Host-buffer entry points (array-of-structs).  desired_count is 1 (one desired
 * trajectory shared by the whole batch) or batch.
 * ------------------------------------------------------------------------- */

/* ILQR::solve (ilqr.hh:53-87) for `batch` problems of `n_knots` knots.
 *   out_k [batch][n][4], out_K [batch][n][4][12]: gains of the last backward pass (NULL to skip)
 *   cost_hist [batch][hist_cap]: new_cost after each completed iteration (NULL to skip)
 *   debug_traj [batch][debug_cap][n][18]: ILQRDebug trajectories (ilqr_debug.hh:9-22); needs
 *     populate_debug and debug_cap >= ceil(max_iters) (NULL to skip)
 *   results [batch]. */
int qilqr_solve_host(qilqr_solver_t *solver, int batch, int n_knots, const double *desired,
                     int desired_count, const double *initial, double *out_traj, double *out_k,
                     double *out_K, double *cost_hist, int hist_cap, double *debug_traj,
                     int debug_cap, qilqr_result_t *results);

/* The same solve in two halves, for callers that keep several batches in flight.  qilqr_solve_host_begin
 * returns as soon as the host has nothing left to sequence: the upload, the throughput-bound part of the solve
 * and the hand-over of the last few problems (the ones that creep to max_iters) to a kernel that finishes them on
 * its own.  qilqr_solve_host_finish waits for that kernel and copies trajectories and results back.  Between the
 * two calls the handle accepts no other call; run the next batch on another handle meanwhile.  out_traj and
 * results must stay valid until _finish returns.  Same results as qilqr_solve_host, bit for bit. */
int qilqr_solve_host_begin(qilqr_solver_t *solver, int batch, int n_knots, const double *desired, int desired_count,
                           const double *initial, double *out_traj, qilqr_result_t *results);
int qilqr_solve_host_finish(qilqr_solver_t *solver);

/* The usual calling convention of an iLQR solver: the initial guess is a control sequence.  x0 [batch][13],
 * controls [control_count][n_knots][4] with control_count = 1 (one nominal sequence for the whole batch) or batch;
 * the initial trajectory ILQR::solve starts from (ilqr.hh:53-56) is their open-loop rollout under
 * QuadrotorModel::discrete_dynamics, made on the device, with time_s = knot * dt_s.  Identical, bit for bit, to
 * rolling the trajectory out first (qilqr_forward_sim_host with zero gains) and calling qilqr_solve_host -- but
 * 13 + 4 n_knots doubles go up per problem instead of 18 n_knots.  out_traj [batch][n_knots][18] and / or
 * out_controls [batch][n_knots][4] (either may be NULL, not both): a receding-horizon caller needs only the
 * controls.  _begin / qilqr_solve_host_finish: as for qilqr_solve_host_begin. */
int qilqr_solve_from_controls_host(qilqr_solver_t *solver, int batch, int n_knots, const double *desired,
                                   int desired_count, const double *x0, const double *controls, int control_count,
                                   double *out_traj, double *out_controls, qilqr_result_t *results);
int qilqr_solve_from_controls_host_begin(qilqr_solver_t *solver, int batch, int n_knots, const double *desired,
                                         int desired_count, const double *x0, const double *controls,
                                         int control_count, double *out_traj, double *out_controls,
                                         qilqr_result_t *results);

/* ---------------------------------------------------------------------------
 * ILQRDebug at batch scale (ilqr.hh:78-80, ilqr_debug.hh:9-22, ilqr_debug.proto:7-14).  A full capture is
 * iterations x n_knots x 144 bytes per problem (35 GB for 65536 problems): qilqr_solve_host's debug_traj argument
 * is meant for single problems and small batches.  For large batches:
 *
 *   - the per-iteration COST of every problem is always kept on the device (8 bytes per completed iteration, up to
 *     128 iterations; QILQR_ALWAYS_HIST_CAP) and read back on demand with qilqr_last_cost_history_host:
 *     out [count][out_cap] for problems first .. first+count-1 of the last solve, zero-padded; *stored_cap = entries
 *     kept per problem (a problem's valid entries: results[b].num_debug);
 *   - TRAJECTORIES are captured for a sample only: qilqr_set_debug_sampling selects the problems (indices into the
 *     batch, any order) and the iterations (i % every_kth_iteration == 0), kept in a ring of `ring_slots` slots
 *     per sampled problem (the last ring_slots sampled iterations survive).  With options.populate_debug set, every
 *     following solve fills the rings; qilqr_read_debug_samples_host copies them back in one asynchronous transfer
 *     per array (pass pinned memory): traj [S][ring_slots][n_knots][18], iters [S][ring_slots] (iteration index of a
 *     slot, -1 = empty), costs [S][ring_slots] (new_cost of that iteration), counts [S] (sampled iterations seen;
 *     slot of the j-th one = j % ring_slots).  num_problems = 0 switches sampling off.
 * ------------------------------------------------------------------------- */
int qilqr_set_debug_sampling(qilqr_solver_t *solver, int every_kth_iteration, const int32_t *problems,
                             int num_problems, int ring_slots);
int qilqr_read_debug_samples_host(qilqr_solver_t *solver, double *traj, int32_t *iters, double *costs,
                                  int32_t *counts);
int qilqr_last_cost_history_host(qilqr_solver_t *solver, int first, int count, double *out, int out_cap,
                                 int *stored_cap);

/* ILQR::forward_sim (ilqr.hh:149-172): alpha [batch]. */
int qilqr_forward_sim_host(qilqr_solver_t *solver, int batch, int n_knots, const double *current,
                           const double *k, const double *K, const double *alpha, double *out_traj);
/* ILQR::cost_trajectory (ilqr.hh:89-95): cost [batch].  n_desired_knots < n_knots ->
 * QILQR_ERR_OUT_OF_RANGE (cost.hh:39-40). */
int qilqr_cost_trajectory_host(qilqr_solver_t *solver, int batch, int n_knots, const double *desired,
                               int desired_count, int n_desired_knots, const double *traj,
                               double *cost);
/* ILQR::backwards_pass (ilqr.hh:97-147): k [batch][n][4], K [batch][n][4][12],
 * terms [batch][2] = {QuTk, kTQuuk} (detail::CostReductionTerms, ilqr.hh:13-16). */
int qilqr_backwards_pass_host(qilqr_solver_t *solver, int batch, int n_knots, const double *desired,
                              int desired_count, const double *traj, double *k, double *K,
                              double *terms);
/* ILQR::line_search (ilqr.hh:174-194): per problem new trajectory, cost, accepted step and
 * status[b] = 0 or QILQR_ERR_LINE_SEARCH. */
int qilqr_line_search_host(qilqr_solver_t *solver, int batch, int n_knots, const double *desired,
                           int desired_count, const double *current, const double *current_cost,
                           const double *k, const double *K, const double *terms, double *out_traj,
                           double *new_cost, double *step, int32_t *status);

/* QuadrotorModel::discrete_dynamics (quadrotor_model.cc:33-49): x [batch][13], u [batch][4] ->
 * x_next [batch][13], J_x [batch][144], J_u [batch][48] (either may be NULL). */
int qilqr_discrete_dynamics_host(qilqr_solver_t *solver, int batch, const double *x, const double *u,
                                 double *x_next, double *J_x, double *J_u);
/* QuadrotorModel::continuous_dynamics (quadrotor_model.cc:65-122): xdot [batch][12]. */
int qilqr_continuous_dynamics_host(qilqr_solver_t *solver, int batch, const double *x,
                                   const double *u, double *xdot, double *J_x, double *J_u);
/* minus(State, State, diffs) (quadrotor_model.cc:215-250): out [batch][12],
 * J_lhs/J_rhs [batch][144] (may be NULL). */
int qilqr_state_minus_host(qilqr_solver_t *solver, int batch, const double *lhs, const double *rhs,
                           double *out, double *J_lhs, double *J_rhs);
/* add(State, StateTangent, diffs) (quadrotor_model.cc:174-206): out [batch][13]. */
int qilqr_state_add_host(qilqr_solver_t *solver, int batch, const double *x, const double *tangent,
                         double *out, double *J_lhs, double *J_rhs);
/* CostFunction::operator() (cost.hh:36-61): cost [batch], C_x [batch][12], C_u [batch][4],
 * C_xx [batch][144], C_uu [batch][16], C_xu [batch][48] (derivatives may be NULL). */
int qilqr_cost_host(qilqr_solver_t *solver, int batch, const double *x, const double *u,
                    const double *x_d, const double *u_d, double *cost, double *C_x, double *C_u,
                    double *C_xx, double *C_uu, double *C_xu);

/* ---------------------------------------------------------------------------
 * Device-resident entry points (structure-of-arrays, problem index fastest):
 *   trajectory  double[n_knots][17][batch]   rows 0-12 state, 13-16 control
 *   desired     double[n_knots][17][desired_count]  (desired_count = 1 or batch)
 *   k           double[n_knots][4][batch]
 *   K           double[n_knots][48][batch]   row = 12*control + state
 *   cost_hist   double[hist_cap][batch]
 * All pointers are device pointers on the solver's device; calls are asynchronous
 * on qilqr_stream() except that qilqr_solve_device returns after the last
 * iteration has been issued and its results are complete on the stream.
 * ------------------------------------------------------------------------- */
int qilqr_solve_device(qilqr_solver_t *solver, int batch, int n_knots, const double *d_desired,
                       int desired_count, double *d_traj_inout, double *d_k, double *d_K,
                       double *d_cost_hist, int hist_cap, qilqr_result_t *d_results);
/* qilqr_solve_device in two halves (see qilqr_solve_host_begin): _begin returns once the device finishes the solve
 * on its own, _finish waits and completes d_traj_inout / d_results.  The buffers passed to _begin must not be
 * touched in between. */
int qilqr_solve_device_begin(qilqr_solver_t *solver, int batch, int n_knots, const double *d_desired,
                             int desired_count, double *d_traj_inout, double *d_k, double *d_K,
                             double *d_cost_hist, int hist_cap, qilqr_result_t *d_results);
int qilqr_solve_device_finish(qilqr_solver_t *solver);
/* AoS <-> SoA transposition kernels (device to device). */
int qilqr_pack_trajectory_device(qilqr_solver_t *solver, int batch, int n_knots,
                                 const double *d_aos /*[batch][n][18]*/, double *d_soa);
int qilqr_unpack_trajectory_device(qilqr_solver_t *solver, int batch, int n_knots,
                                   const double *d_soa, const double *d_time_aos_or_null,
                                   double *d_aos);
/* Open-loop rollout helper: x0 [13][batch] (SoA), u_const[4] -> trajectory SoA. */
int qilqr_rollout_constant_control_device(qilqr_solver_t *solver, int batch, int n_knots,
                                          const double *d_x0_soa, const double *u_const /*host[4]*/,
                                          double *d_traj_soa);

/* Receding-horizon MPC on top of solve(initial_traj) (ilqr.hh:53-55; BASELINE config 5), all on the device.
 * One closed-loop step = solve from the warm start, apply the first control to the plant
 * (QuadrotorModel::discrete_dynamics, optionally plus an additive body-velocity disturbance
 * d_disturbance[6][batch]), shift the solution one knot (last knot duplicated) and put the new plant
 * state in knot 0.  d_plant_state [13][batch] and d_traj_inout are updated in place.
 * qilqr_mpc_run_device repeats that `steps` times; d_state_log [steps][13][batch] and
 * d_control_log [steps][4][batch] (either may be NULL) receive the plant state after and the control
 * applied at every step; totals[0..3] (host, may be NULL) = sum of backward passes, sum of rollouts,
 * number of re-solves that did not converge (status 3 or 4), number of re-solves. */
int qilqr_mpc_advance_device(qilqr_solver_t *solver, int batch, int n_knots, double *d_traj_inout,
                             double *d_plant_state, const double *d_disturbance, double *d_applied_u);
int qilqr_mpc_run_device(qilqr_solver_t *solver, int steps, int batch, int n_knots, const double *d_desired,
                         int desired_count, double *d_traj_inout, double *d_plant_state,
                         const double *d_disturbance, double *d_state_log, double *d_control_log,
                         int64_t *totals);

/* Aggregate counters of the last qilqr_solve_* call. */
typedef struct {
  int64_t solver_iterations;   /* super-steps sequenced by the host (backward pass + rollout + compaction over the
                                  problem lists); the iterations the persistent tail kernel runs on its own once few
                                  problems are alive are not counted */
  int64_t problem_iterations;  /* sum over problems of backward passes */
  int64_t problem_rollouts;    /* sum over problems of rollouts */
  int64_t kernel_launches;     /* kernels launched by this call */
  double backward_ms;          /* CUDA-event time spent in backward-pass kernels */
  double rollout_ms;           /* CUDA-event time spent in rollout/line-search kernels */
  int64_t backward_problem_knots; /* problem-knots processed by backward kernels */
  int64_t rollout_problem_knots;  /* problem-knots processed by rollout kernels */
  double bulk_wall_ms;            /* host wall time of the solve loop while more than hi_threshold problems were active */
  double tail_wall_ms;            /* ... and after the switch to the high-priority tail stream */
  /* the part of backward_ms / rollout_ms / backward_problem_knots spent in launches of the throughput-bound bulk
   * (more than `hi_threshold` problems alive); the rest is the latency-bound tail of the few problems that creep to
   * max_iters */
  double backward_ms_bulk;
  double rollout_ms_bulk;
  int64_t backward_problem_knots_bulk;
  int64_t backward_launches_bulk; /* backward passes (linearisation + Riccati launch pairs) issued in the bulk */
} qilqr_solve_stats_t;
int qilqr_last_solve_stats(const qilqr_solver_t *solver, qilqr_solve_stats_t *out);
/* Enable/disable per-kernel CUDA-event timing (adds a few microseconds per launch). */
int qilqr_set_profiling(qilqr_solver_t *solver, int enabled);


/* Measurement aid for the roofline denominator: sustained FP64 FMA throughput of `device`
 * in TFLOP/s (2 flops per DFMA), from a register-resident DFMA kernel timed with CUDA events. */
int qilqr_measure_fp64_peak(int device, double *tflops);

/* The QuadrotorModel constructor's validity test (quadrotor_model.cc:20-24): QILQR_OK, or
 * QILQR_ERR_INERTIA_NOT_PD when the inertia is not symmetric positive definite.  Host only. */
int qilqr_check_model(const qilqr_model_t *model);

#ifdef __cplusplus
}
#endif
#endif /* QILQR_H_ */
