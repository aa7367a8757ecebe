// =============================================================================
// qilqr_capi.cu -- host side of libqilqr_b200.so: the C ABI of include/qilqr.h.
// Owns the device workspace, drives ILQR::solve's control flow (ilqr.hh:53-87)
// for a whole batch with per-problem termination, and offers the reference's
// public methods as batched calls.  No CPU fallback: every entry point launches
// CUDA kernels or fails.
// =============================================================================
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <atomic>
#include <chrono>
#include <mutex>
#include <vector>

#include "qilqr_api_kernels.cuh"
#include "qilqr_kernels.cuh"
#include "qilqr_backward_g4.cuh"
#include "qilqr_backward_split.cuh"
#include "qilqr_backward_dense.cuh"
#include "qilqr_tail_persistent.cuh"
#include "qilqr_riccati_g16.cuh"

using namespace qilqr;

namespace {

struct DeviceBuffer {
  void *ptr = nullptr;
  size_t bytes = 0;
  cudaError_t ensure(size_t need) {
    if (need <= bytes) return cudaSuccess;
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    bytes = 0;
    cudaError_t e = cudaMalloc(&ptr, need);
    if (e == cudaSuccess) bytes = need;
    return e;
  }
  void release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    bytes = 0;
  }
  template <class T> T *as() const { return static_cast<T *>(ptr); }
};

struct TimedSpan {
  cudaEvent_t start, stop;
  int kind;   // 0 backward, 1 rollout
  bool bulk;  // launched while the solve was in its throughput-bound bulk (more than hi_threshold problems alive)
};

}  // namespace

namespace {
struct SolveCtx {  // what a solve leaves pending when its tail runs on the device (qilqr_solve_device_begin / _finish)
  bool pending = false;
  bool compacted = false, on_hi = false;
  int B = 0, N = 0, m_tail = 0;
  Problem pr_big{}, pr{};
  SolveState st_big{}, st{};
  qilqr_result_t *d_results = nullptr;
  int64_t launches0 = 0;
  std::chrono::steady_clock::time_point t_begin, t_switch;
};

struct HostCall {  // where the results of a host-buffer solve go (qilqr_solve_host / _begin / _finish)
  bool pending = false;
  int B = 0, N = 0, hist_cap = 0, debug_cap = 0;
  double *out_traj = nullptr, *out_k = nullptr, *out_K = nullptr, *cost_hist = nullptr, *debug_traj = nullptr;
  double *out_controls = nullptr;  // qilqr_solve_from_controls_host: [B][N][4]
  qilqr_result_t *results = nullptr;
};
}  // namespace

struct qilqr_solver {
  DeviceParams p{};
  qilqr_options_t opt{};
  int device = 0;
  cudaStream_t stream = nullptr;     // bulk work, lowest priority
  cudaStream_t stream_hi = nullptr;  // the latency-bound tail of a solve (few problems left), highest priority
  cudaStream_t cur = nullptr;        // the one solve_core is currently launching on
  cudaEvent_t ev_switch = nullptr;
  int seq = 0;                       // sequence number of the last compaction (see wait_counts)
  int lists_B = 0;                   // batch the list buffer is laid out for
  int hi_threshold = 2048;           // switch to stream_hi when at most this many problems are active
  int bulk_poll_us = 20;             // QILQR_BULK_POLL_US: sleep this long between polls of the list lengths while the
                                     // solve is in its bulk (the kernels ahead take milliseconds); 0: spin.  Keeps the host
                                     // threads of many pipelined handles / ranks from needing a core each.
  bool tail_stream = true;           // QILQR_TAIL_STREAM=0: the tail stays on the handle's main stream (one stream per handle)
  bool in_tail = false;              // the current solve has entered its latency-bound tail
  std::string last_error;
  int64_t launches = 0;
  qilqr_solve_stats_t stats{};
  bool profiling = false;
  bool q_block_diagonal = true;  // Q = blkdiag(Q_pp, Q_vv): the 4-lanes-per-problem backward kernel applies
  int g4_kpp = 4;                // knots linearised per phase by the quad kernel (1, 2 or 4)
  bool force_t1 = false;         // QILQR_BACKWARD=t1: the one-thread-per-problem kernel (cross-checks)
  bool split_backward = true;    // linearise kernel + TMA-fed Riccati kernel (default) vs the fused quad kernel
  bool rollout_ws = true;        // role-specialised rollout kernel (3 warps per 32 problems); QILQR_ROLLOUT=thread: one thread per problem
  int ws_threshold = 4096;       // ... used for launches of at most this many problems (latency-bound); QILQR_WS_THRESHOLD
  bool wide_first = false;       // QILQR_WIDE_FIRST=1: parallel step sizes from the first candidate of every line search
  int g16_threshold = 1184;      // Riccati sweep with 16 lanes per problem for launches of at most this many problems: one
                                 // tile of 8 per SM (beyond that k_riccati_g4's lower instruction count wins)
                                 // (k_riccati_g16; QILQR_G16_THRESHOLD, 0 = never)
  bool generic_path = false;     // model-agnostic kernels (dense J_x, J_u): any model variant, or forced for cross-checks
  std::vector<TimedSpan> spans;
  std::vector<cudaEvent_t> event_pool;

  // workspace
  DeviceBuffer buf1, gk, gK, state_d, state_i, lists, desired_soa, traj_soa, stage_a, stage_b, stage_c, results_d,
      hist_d, debug_d, misc, wide_d, rec_d, totals_d, stage_d;
  // dense mini-batch for the tail of a solve (k_tail_gather / k_tail_scatter)
  DeviceBuffer tail_traj, tail_gains, tail_des, tail_sd, tail_si, tail_hist, tail_map, tail_lists, rec_tail_d, tail_scratch;
  bool tail_compaction = true;  // QILQR_TAIL_COMPACTION=0 keeps the stragglers in the big batch's layout
  int compact_threshold = 0;    // QILQR_COMPACT_THRESHOLD: gather at this many alive problems; 0 = max(hi_threshold, B / 4)
  bool persistent_tail = false;  // QILQR_PERSISTENT_TAIL=1: one kernel finishes the solve on the device once at most
  int persist_threshold = 64;    // `persist_threshold` problems are alive (frees the host thread: begin / finish API)
  int always_hist_cap = 128;    // per-problem cost history kept on the device even when the caller passes no buffer
  int last_hist_cap = 0, last_hist_B = 0;  // ... of the last solve (qilqr_last_cost_history_host)
  bool last_hist_internal = false;
  int *h_counts = nullptr;  // mapped pinned: [0] = alive, [1] = active, [2] = sequence number of the compaction
  int *d_counts = nullptr;
  long long *h_totals = nullptr;  // pinned: sums over the batch of backward passes and rollouts of the last solve
  // sampled ILQRDebug stream (qilqr_set_debug_sampling): which problems / iterations, and the device ring buffers
  DeviceBuffer dbg_sample, dbg_traj, dbg_iters, dbg_costs, dbg_count;
  int dbg_S = 0, dbg_ring = 0, dbg_every = 0, dbg_N = 0;  // dbg_S == 0: sampling off; dbg_N: knots of the captured solve
  const double *dbg_time_aos = nullptr;                    // host path: the input AoS on the device (time_s column)
  // user-supplied model compiled at run time (qilqr_set_user_model): the generic kernels around the caller's
  // discrete_dynamics device function
  cudaLibrary_t user_lib = nullptr;
  cudaKernel_t user_linearise = nullptr, user_rollout = nullptr;
  bool user_model = false;
  SolveCtx ctx;                   // a solve whose tail is still running on the device (begin / finish API)
  HostCall host_call;
  std::mutex mu;                  // one call at a time per handle (the workspace is shared by every entry point)
};

namespace {

#define QCUDA(S, expr)                                                                            \
  do {                                                                                            \
    cudaError_t e__ = (expr);                                                                     \
    if (e__ != cudaSuccess) {                                                                     \
      (S)->last_error = std::string(#expr) + ": " + cudaGetErrorString(e__);                      \
      return (e__ == cudaErrorMemoryAllocation) ? QILQR_ERR_OUT_OF_MEMORY : QILQR_ERR_CUDA;       \
    }                                                                                             \
  } while (0)

int fail(qilqr_solver *S, int code, const char *msg) {
  if (S) S->last_error = msg;
  return code;
}
// Every entry point that touches the handle's workspace: one call at a time per handle (the reference's solver
// methods are const and re-entrant; here the workspace is shared, so concurrent callers are serialised), and
// nothing may run between ..._begin and ..._finish.
#define QENTER(S)                                                                                          \
  std::lock_guard<std::mutex> lock__((S)->mu);                                                             \
  if ((S)->ctx.pending || (S)->host_call.pending)                                                          \
    return fail((S), QILQR_ERR_INVALID_ARGUMENT, "a solve begun with qilqr_solve_*_begin is still pending on this handle")

cudaEvent_t get_event(qilqr_solver *S) {
  if (!S->event_pool.empty()) {
    cudaEvent_t e = S->event_pool.back();
    S->event_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
struct SpanGuard {
  qilqr_solver *S;
  TimedSpan sp{};
  bool on;
  SpanGuard(qilqr_solver *s, int kind) : S(s), on(s->profiling) {
    if (on) {
      sp.start = get_event(S);
      sp.stop = get_event(S);
      sp.kind = kind;
      sp.bulk = !S->in_tail;
      cudaEventRecord(sp.start, S->cur);
    }
  }
  ~SpanGuard() {
    if (on) {
      cudaEventRecord(sp.stop, S->cur);
      S->spans.push_back(sp);
    }
  }
};
void drain_spans(qilqr_solver *S) {  // call after a stream synchronise
  for (auto &sp : S->spans) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, sp.start, sp.stop);
    (sp.kind == 0 ? S->stats.backward_ms : S->stats.rollout_ms) += ms;
    if (sp.bulk) (sp.kind == 0 ? S->stats.backward_ms_bulk : S->stats.rollout_ms_bulk) += ms;
    S->event_pool.push_back(sp.start);
    S->event_pool.push_back(sp.stop);
  }
  S->spans.clear();
}

void apply_options(qilqr_solver *S) {
  S->p.step_update = S->opt.step_update;
  S->p.desired_reduction_frac = S->opt.desired_reduction_frac;
  S->p.rtol = S->opt.rtol;
  S->p.atol = S->opt.atol;
  S->p.max_iters = S->opt.max_iters;
  S->p.quu_reg = S->opt.quu_regularization;
  S->p.ls_max_iters = S->opt.line_search_max_iters;
  S->p.symmetrize_vxx = S->opt.symmetrize_vxx;
}

// Eigen::LLT<Matrix3d> of the inertia + the isApprox(transpose) test (quadrotor_model.cc:20-24)
bool factor_inertia(const double *I, double *L) {
  for (int i = 0; i < 9; ++i) L[i] = 0.0;
  for (int k = 0; k < 3; ++k) {
    double x = I[4 * k];
    for (int j = 0; j < k; ++j) x -= L[3 * k + j] * L[3 * k + j];
    if (!(x > 0.0)) return false;
    x = std::sqrt(x);
    L[4 * k] = x;
    for (int i = k + 1; i < 3; ++i) {
      double s = I[3 * i + k];
      for (int j = 0; j < k; ++j) s -= L[3 * i + j] * L[3 * k + j];
      L[3 * i + k] = s / x;
    }
  }
  double d2 = 0, n2 = 0;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      const double d = I[3 * i + j] - I[3 * j + i];
      d2 += d * d;
      n2 += I[3 * i + j] * I[3 * i + j];
    }
  return d2 <= 1e-24 * n2;
}
void llt_solve(const double *L, const double *b, double *x) {
  double y[3];
  y[0] = b[0] / L[0];
  y[1] = (b[1] - L[3] * y[0]) / L[4];
  y[2] = (b[2] - L[6] * y[0] - L[7] * y[1]) / L[8];
  x[2] = y[2] / L[8];
  x[1] = (y[1] - L[7] * x[2]) / L[4];
  x[0] = (y[0] - L[3] * x[1] - L[6] * x[2]) / L[0];
}

struct StateLayout {  // carve SolveState out of two flat buffers
  static constexpr int kDoubles = 5, kInts = 8;
};
SolveState make_state_from(double *d, int *i, int B, double *hist, int hist_cap);
SolveState make_state(qilqr_solver *S, int B, double *hist, int hist_cap) {
  return make_state_from(S->state_d.as<double>(), S->state_i.as<int>(), B, hist, hist_cap);
}
SolveState make_state_from(double *d, int *i, int B, double *hist, int hist_cap) {
  SolveState st;
  st.cost = d; st.new_cost = d + B; st.qutk = d + 2 * size_t(B); st.ktquuk = d + 3 * size_t(B); st.alpha = d + 4 * size_t(B);
  st.ls_iter = i; st.status = i + B; st.bwd = i + 2 * size_t(B); st.rollouts = i + 3 * size_t(B);
  st.ndebug = i + 4 * size_t(B); st.sel = i + 5 * size_t(B); st.phase = i + 6 * size_t(B);
  st.accepted_iter = i + 7 * size_t(B);
  st.cost_hist = hist;
  st.hist_cap = hist_cap;
  return st;
}
int ensure_state(qilqr_solver *S, int B) {
  QCUDA(S, S->state_d.ensure(sizeof(double) * StateLayout::kDoubles * size_t(B)));
  QCUDA(S, S->state_i.ensure(sizeof(int) * StateLayout::kInts * size_t(B)));
  QCUDA(S, S->lists.ensure(sizeof(int) * (5 * size_t(B) + 2 * (size_t(B) / 1024 + 2))));
  S->lists_B = B;
  return QILQR_OK;
}

inline unsigned blocks_for(int n, int per) { return unsigned((n + per - 1) / per); }

// Ordered compaction of `list` (n entries) by phase into two lists; the two lengths land in S->h_counts
// after the next synchronisation of S->cur.
void launch_compact(qilqr_solver *S, const int *list, int n, const int *phase, int *out_s, int *out_a,
                    int phase_s = kMaskSearch, int phase_a = kMaskActive) {
  const int seq = ++S->seq;
  if (n <= 4096) {
    // a small CTA is enough for the short lists of a solve's tail (and fits next to the bulk kernels' CTAs)
    const int threads = n <= 512 ? 128 : 1024;
    k_compact<<<1, threads, 0, S->cur>>>(list, n, phase, out_s, out_a, S->d_counts, phase_s, phase_a, seq);
    ++S->launches;
    return;
  }
  const int chunks = (n + 1023) / 1024;
  int *chunk_counts = S->lists.as<int>() + 5 * size_t(S->lists_B);
  k_compact_count<<<chunks, 1024, 0, S->cur>>>(list, n, phase, chunk_counts, phase_s, phase_a);
  k_compact_scatter<<<chunks, 1024, 0, S->cur>>>(list, n, phase, chunk_counts, out_s, out_a, S->d_counts, phase_s, phase_a, seq);
  S->launches += 2;
}

// Wait for the list lengths of the last launch_compact.  The compaction kernel publishes them, followed
// by a sequence number, in mapped pinned memory; polling that word (yielding the core between polls)
// costs a few microseconds less per iteration than a stream synchronisation and keeps the host threads
// of several pipelined handles from burning cores in the driver's spin loop.
int wait_counts(qilqr_solver *S, cudaStream_t st) {
  if (S->profiling) {
    QCUDA(S, cudaStreamSynchronize(st));
    return QILQR_OK;
  }
  volatile int *seq = S->h_counts + 2;
  // Tail (high-priority stream): the wait is tens of microseconds and on the critical path -> spin.
  // Bulk: the kernels ahead take milliseconds -> sleep between polls, so that many pipelined handles do not
  // need a core each.
  const bool spin = S->in_tail || S->bulk_poll_us <= 0;
  for (unsigned polls = 0; *seq != S->seq; ++polls) {
    if ((polls & (spin ? 0x3ffu : 0x3fu)) == (spin ? 0x3ffu : 0x3fu)) {
      const cudaError_t e = cudaStreamQuery(st);
      if (e != cudaSuccess && e != cudaErrorNotReady) QCUDA(S, e);
    }
    if (spin) std::this_thread::yield();
    else std::this_thread::sleep_for(std::chrono::microseconds(S->bulk_poll_us));
  }
  std::atomic_thread_fence(std::memory_order_acquire);  // the list lengths are read after the sequence word
  return QILQR_OK;
}

// Opt-in to more than 48 kB of dynamic shared memory and the maximum shared-memory carve-out.  Function
// attributes are per device (per context): qilqr_create calls this after cudaSetDevice for every handle.
size_t riccati_extra_smem() {  // QILQR_RICCATI_EXTRA_SMEM: occupancy experiments only
  static const size_t extra = std::getenv("QILQR_RICCATI_EXTRA_SMEM") ? std::atoi(std::getenv("QILQR_RICCATI_EXTRA_SMEM")) : 0;
  return extra;
}
template <class K>
cudaError_t opt_in_smem(K kernel, size_t bytes) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes));
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}
cudaError_t configure_kernels() {
  cudaError_t e;
  if ((e = opt_in_smem(k_backward_g4<1>, sizeof(double) * g4::smem_doubles(1))) != cudaSuccess) return e;
  if ((e = opt_in_smem(k_backward_g4<2>, sizeof(double) * g4::smem_doubles(2))) != cudaSuccess) return e;
  if ((e = opt_in_smem(k_backward_g4<4>, sizeof(double) * g4::smem_doubles(4))) != cudaSuccess) return e;
  if ((e = opt_in_smem(k_riccati_g4<false>, sizeof(double) * g4::split_smem_doubles(false) + riccati_extra_smem())) != cudaSuccess) return e;
  if ((e = opt_in_smem(k_riccati_g4<true>, sizeof(double) * g4::split_smem_doubles(true) + riccati_extra_smem())) != cudaSuccess) return e;
  if ((e = opt_in_smem(k_riccati_dense, sizeof(double) * dn::smem_doubles())) != cudaSuccess) return e;
  if ((e = opt_in_smem(k_riccati_g16, sizeof(double) * g16::smem_doubles())) != cudaSuccess) return e;
  if ((e = opt_in_smem(k_tail_persistent<false>, sizeof(double) * tp::smem_doubles(false))) != cudaSuccess) return e;
  if ((e = opt_in_smem(k_tail_persistent<true>, sizeof(double) * tp::smem_doubles(true))) != cudaSuccess) return e;
  return cudaSuccess;
}

template <int KPP>
void launch_g4(qilqr_solver *S, const BackwardArgs &ba) {
  const size_t smem = sizeof(double) * g4::smem_doubles(KPP);
  k_backward_g4<KPP><<<blocks_for(ba.n, 8), 32, smem, S->cur>>>(S->p, ba);
}
// ILQR::backwards_pass for the problems in ba.list: linearisation kernel + TMA-fed Riccati kernel
template <bool DENSEQ>
int launch_split(qilqr_solver *S, const BackwardArgs &ba) {
  const int n8 = (ba.n + 7) & ~7;
  const size_t need = sizeof(double) * size_t(n8) * ba.pr.N * g4::rect(DENSEQ);
  if (S->rec_d.ensure(need) != cudaSuccess) return QILQR_ERR_OUT_OF_MEMORY;
  const size_t smem = sizeof(double) * g4::split_smem_doubles(DENSEQ) + riccati_extra_smem();
  const size_t threads = size_t(n8) * ba.pr.N;
  k_linearise<DENSEQ><<<unsigned((threads + 127) / 128), 128, 0, S->cur>>>(S->p, ba, S->rec_d.as<double>());
  if (!DENSEQ && ba.n <= S->g16_threshold)  // latency-bound launch: 16 lanes per problem (bit-identical results)
    k_riccati_g16<<<n8 / 8, 128, sizeof(double) * g16::smem_doubles(), S->cur>>>(S->p, ba, S->rec_d.as<double>());
  else
    k_riccati_g4<DENSEQ><<<n8 / 8, 32, smem, S->cur>>>(S->p, ba, S->rec_d.as<double>());
  ++S->launches;
  return QILQR_OK;
}
// The model-agnostic backward pass: dense records from the configured model, dense Riccati sweep.
int launch_dense(qilqr_solver *S, const BackwardArgs &ba) {
  const int n8 = (ba.n + 7) & ~7;
  const size_t need = sizeof(double) * size_t(n8) * ba.pr.N * dn::DREC;
  if (S->rec_d.ensure(need) != cudaSuccess) return QILQR_ERR_OUT_OF_MEMORY;
  const size_t smem = sizeof(double) * dn::smem_doubles();
  const size_t threads = size_t(n8) * ba.pr.N;
  if (S->user_model) {
    double *rec = S->rec_d.as<double>();
    void *args[] = {const_cast<DeviceParams *>(&S->p), const_cast<BackwardArgs *>(&ba), &rec};
    if (cudaLaunchKernel(reinterpret_cast<const void *>(S->user_linearise), dim3(unsigned((threads + 127) / 128)), dim3(128),
                         args, 0, S->cur) != cudaSuccess)
      return QILQR_ERR_CUDA;
  } else {
    k_linearise_dense<<<unsigned((threads + 127) / 128), 128, 0, S->cur>>>(S->p, ba, S->rec_d.as<double>());
  }
  k_riccati_dense<<<n8 / 8, 32, smem, S->cur>>>(S->p, ba, S->rec_d.as<double>());
  ++S->launches;
  return QILQR_OK;
}
// forward_sim (+ cost, + line-search bookkeeping) with the dynamics of the configured model
void launch_rollout(qilqr_solver *S, const RolloutArgs &ra, int threads, cudaStream_t st) {
  if (S->user_model) {
    void *args[] = {const_cast<DeviceParams *>(&S->p), const_cast<RolloutArgs *>(&ra)};
    cudaLaunchKernel(reinterpret_cast<const void *>(S->user_rollout), dim3(blocks_for(threads, 128)), dim3(128), args, 0, st);
  } else if (S->generic_path) k_rollout<true><<<blocks_for(threads, 128), 128, 0, st>>>(S->p, ra);
  // (the parallel step-size rounds stay on the one-thread-per-problem kernel; the two kernels are bit-identical)
  else if (S->rollout_ws && ra.mode != MODE_WIDE && threads <= S->ws_threshold) k_rollout_ws<<<blocks_for(threads, 32), 96, 0, st>>>(S->p, ra);
  else k_rollout<false><<<blocks_for(threads, 128), 128, 0, st>>>(S->p, ra);
}
int launch_backward(qilqr_solver *S, const BackwardArgs &ba) {
  if (S->generic_path) {
    const int rc = launch_dense(S, ba);
    if (rc) return fail(S, rc, "out of device memory for the linearisation records");
    return QILQR_OK;
  }
  if (S->split_backward && !S->force_t1) {
    const int rc = S->q_block_diagonal ? launch_split<false>(S, ba) : launch_split<true>(S, ba);
    if (rc) return fail(S, rc, "out of device memory for the linearisation records");
  } else if (!S->q_block_diagonal || S->force_t1) {
    k_backward_t1<<<blocks_for(ba.n, 64), 64, 0, S->cur>>>(S->p, ba);
  } else if (S->g4_kpp == 1) {
    launch_g4<1>(S, ba);
  } else if (S->g4_kpp == 2) {
    launch_g4<2>(S, ba);
  } else {
    launch_g4<4>(S, ba);
  }
  return QILQR_OK;
}

// ---------------------------------------------------------------------------
// The batched solve loop (device-resident data).
//
// Every problem runs ILQR::solve's own control flow (ilqr.hh:53-87) and counts its own iterations; the host only
// sequences "super-steps" over lists of problem indices:
//
//   backward pass          over the ACTIVE list (problems that need one)        k_linearise + k_riccati_g4
//   [parallel step sizes]  over the ALIVE list, for the problems in PHASE_WIDE  k_rollout<MODE_WIDE> + k_select_alpha
//   rollout                over the ALIVE list, for the problems in PHASE_SEARCH k_rollout / k_rollout_ws
//   compaction             ALIVE list -> next ALIVE and ACTIVE lists (ordered), lengths to the host
//
// A problem whose candidate is rejected by the Armijo test stays in PHASE_SEARCH with a smaller step and is simply
// rolled out again in the next super-step, while the others go on to their next backward pass: a backtracking
// problem costs only itself a round, there is one host round trip per super-step, and the lists stay sorted (SoA
// accesses stay coalesced).  Problems never interact, so the results do not depend on this scheduling.
// ---------------------------------------------------------------------------
int solve_finish(qilqr_solver *S, SolveCtx &cx);

int solve_core(qilqr_solver *S, int B, int N, const double *d_desired, int Bd, double *d_traj, double *d_k,
               double *d_K, double *d_hist, int hist_cap, qilqr_result_t *d_results, double *d_debug,
               int debug_cap, bool async_tail = false) {
  if (B <= 0 || N <= 0 || (Bd != 1 && Bd != B)) return fail(S, QILQR_ERR_INVALID_ARGUMENT, "bad batch/knots/desired_count");
  SolveCtx &cx = S->ctx;
  if (cx.pending) return fail(S, QILQR_ERR_INVALID_ARGUMENT, "a solve begun with qilqr_solve_device_begin is still pending");
  QCUDA(S, cudaSetDevice(S->device));
  int rc = ensure_state(S, B);
  if (rc) return rc;
  QCUDA(S, S->buf1.ensure(sizeof(double) * size_t(N) * 17 * B));
  if (!d_k) {
    QCUDA(S, S->gk.ensure(sizeof(double) * size_t(N) * 4 * B));
    d_k = S->gk.as<double>();
  }
  if (!d_K) {
    QCUDA(S, S->gK.ensure(sizeof(double) * size_t(N) * 48 * B));
    d_K = S->gK.as<double>();
  }
  S->last_hist_internal = false;
  if (!d_hist && S->always_hist_cap > 0) {  // always-on cost history (8 bytes per completed iteration and problem)
    hist_cap = S->always_hist_cap;
    if (S->opt.max_iters > 0 && S->opt.max_iters < hist_cap) hist_cap = int(std::ceil(S->opt.max_iters));
    QCUDA(S, S->hist_d.ensure(sizeof(double) * size_t(hist_cap) * B));
    d_hist = S->hist_d.as<double>();
    S->last_hist_internal = true;
    // entries a problem never reaches read back as zero
    QCUDA(S, cudaMemsetAsync(d_hist, 0, sizeof(double) * size_t(hist_cap) * B, S->stream));
  }
  if (d_hist == S->hist_d.as<double>()) S->last_hist_internal = true;  // (the host path's own staging buffer)
  S->last_hist_cap = d_hist ? hist_cap : 0;
  S->last_hist_B = B;
  cudaStream_t st_ = S->stream;
  S->cur = st_;  // (an earlier solve that failed in its tail may have left the high-priority stream selected)
  S->in_tail = false;
  Problem pr{B, N, Bd, d_traj, S->buf1.as<double>(), d_desired, d_k, d_K};
  SolveState st = make_state(S, B, d_hist, d_hist ? hist_cap : 0);
  int *listL[2] = {S->lists.as<int>(), S->lists.as<int>() + B};                      // alive
  int *listA[2] = {S->lists.as<int>() + 2 * size_t(B), S->lists.as<int>() + 3 * size_t(B)};  // active

  S->stats = qilqr_solve_stats_t{};
  cx = SolveCtx{};
  cx.B = B; cx.N = N; cx.d_results = d_results;
  cx.launches0 = S->launches;
  cx.t_begin = cx.t_switch = std::chrono::steady_clock::now();

  k_init_state<<<blocks_for(B, 256), 256, 0, st_>>>(st, B, S->p.max_iters);
  k_cost_trajectory<<<blocks_for(B, 128), 128, 0, st_>>>(S->p, d_traj, d_desired, B, N, Bd, st.cost);
  S->launches += 2;

  const bool ring_debug = S->opt.populate_debug && S->dbg_S > 0;
  const bool capture_debug = (S->opt.populate_debug && d_debug && debug_cap > 0) || ring_debug;
  DebugRing dr{};
  if (ring_debug) {
    const size_t slots = size_t(S->dbg_S) * S->dbg_ring;
    QCUDA(S, S->dbg_traj.ensure(sizeof(double) * slots * N * 18));
    QCUDA(S, S->dbg_iters.ensure(sizeof(int) * slots));
    QCUDA(S, S->dbg_costs.ensure(sizeof(double) * slots));
    QCUDA(S, S->dbg_count.ensure(sizeof(int) * size_t(S->dbg_S)));
    QCUDA(S, cudaMemsetAsync(S->dbg_iters.ptr, 0xff, sizeof(int) * slots, st_));
    QCUDA(S, cudaMemsetAsync(S->dbg_costs.ptr, 0, sizeof(double) * slots, st_));
    QCUDA(S, cudaMemsetAsync(S->dbg_count.ptr, 0, sizeof(int) * size_t(S->dbg_S), st_));
    dr = DebugRing{S->dbg_sample.as<int>(), S->dbg_S, S->dbg_ring, S->dbg_every, S->dbg_traj.as<double>(),
                   S->dbg_iters.as<int>(), S->dbg_costs.as<double>(), S->dbg_count.as<int>(), S->dbg_time_aos};
    S->dbg_N = N;
  }
  const int P_alpha = S->opt.num_parallel_alphas > 1 ? S->opt.num_parallel_alphas : 1;
  if (P_alpha > 1) QCUDA(S, S->wide_d.ensure(sizeof(double) * size_t(P_alpha) * B));
  const bool can_persist = S->persistent_tail && P_alpha <= 1 && !S->generic_path && !S->force_t1 &&
                           S->split_backward && !capture_debug && !S->profiling;
  int n_alive = (0.0 < S->opt.max_iters) ? B : 0, n_active = n_alive;
  const int *alive = nullptr, *active = nullptr;  // nullptr = identity list
  int cur = 0;
  cx.pr_big = pr; cx.st_big = st;
  int B_eff = B;
  bool persistent_launched = false;
  for (int epoch = 0; n_alive > 0; ++epoch) {
    if (!cx.on_hi && n_alive <= S->hi_threshold && B > S->hi_threshold) {
      // Few problems left: every further super-step is a chain of tiny, latency-bound launches.  Move them
      // to the high-priority stream so that they are not queued behind another handle's bulk kernels.
      if (S->tail_stream) {
        cudaEventRecord(S->ev_switch, st_);
        cudaStreamWaitEvent(S->stream_hi, S->ev_switch, 0);
        st_ = S->stream_hi;
        S->cur = st_;
      }
      cx.on_hi = true;
      S->in_tail = true;
      cx.t_switch = std::chrono::steady_clock::now();
    }
    {
      // A quarter of the batch left (never fewer than hi_threshold): gather the survivors into a dense mini-batch.
      // With n of B problems alive a warp's 32 list entries span 32 B / n problem slots, so that below a quarter
      // every 8-byte element sits alone in its 32-byte sector (4x the DRAM traffic) -- measured on the headline
      // batch: gathering at 16384 instead of 2048 alive problems is worth 4 % of the pipelined rate.
      const int compact_at = S->compact_threshold > 0 ? S->compact_threshold : std::max(S->hi_threshold, B / 4);
      if (!cx.compacted && n_alive <= compact_at && B > compact_at && B > S->hi_threshold && S->tail_compaction &&
          !capture_debug && alive) {
        const int m = n_alive;
        const size_t md = size_t(m);
        QCUDA(S, S->tail_traj.ensure(sizeof(double) * 2 * N * 17 * md));
        QCUDA(S, S->tail_gains.ensure(sizeof(double) * N * 52 * md));
        QCUDA(S, S->tail_sd.ensure(sizeof(double) * StateLayout::kDoubles * md));
        QCUDA(S, S->tail_si.ensure(sizeof(int) * StateLayout::kInts * md));
        QCUDA(S, S->tail_map.ensure(sizeof(int) * md));
        QCUDA(S, S->tail_lists.ensure(sizeof(int) * 4 * md));
        if (Bd != 1) QCUDA(S, S->tail_des.ensure(sizeof(double) * N * 17 * md));
        if (st.cost_hist) QCUDA(S, S->tail_hist.ensure(sizeof(double) * size_t(st.hist_cap) * md));
        QCUDA(S, cudaMemcpyAsync(S->tail_map.ptr, alive, sizeof(int) * md, cudaMemcpyDeviceToDevice, st_));
        double *tt = S->tail_traj.as<double>(), *tg = S->tail_gains.as<double>();
        double *mini_des = (Bd != 1) ? S->tail_des.as<double>() : nullptr;
        Problem pm{m, N, Bd != 1 ? m : 1, tt, tt + size_t(N) * 17 * md, mini_des ? mini_des : d_desired, tg,
                   tg + size_t(N) * 4 * md};
        SolveState sm = make_state_from(S->tail_sd.as<double>(), S->tail_si.as<int>(), m,
                                        st.cost_hist ? S->tail_hist.as<double>() : nullptr, st.hist_cap);
        dim3 grid(blocks_for(m, 128), 32);
        k_tail_gather<<<grid, 128, 0, st_>>>(cx.pr_big, cx.st_big, pm, sm, mini_des, S->tail_map.as<int>(), m);
        ++S->launches;
        pr = pm;
        st = sm;
        B_eff = m;
        cx.m_tail = m;
        cx.compacted = true;
        // the lists of the mini-batch: everything is alive, the active ones need a compaction
        int *tl = S->tail_lists.as<int>();
        listL[0] = tl; listL[1] = tl + md; listA[0] = tl + 2 * md; listA[1] = tl + 3 * md;
        cur = 0;
        launch_compact(S, nullptr, m, st.phase, listL[1], listA[0], kMaskAlive, kMaskActive);
        if ((rc = wait_counts(S, st_))) return rc;
        alive = nullptr;
        active = listA[0];
        n_active = S->h_counts[1];
      }
    }
    if (can_persist && n_alive <= S->persist_threshold) {
      // the rest of the solve for these problems in one launch (qilqr_tail_persistent.cuh); the host is done
      const int tiles = (n_alive + 7) / 8;
      const bool dq = !S->q_block_diagonal;
      if (S->rec_tail_d.ensure(sizeof(double) * size_t(tiles) * 8 * N * g4::rect(dq)) != cudaSuccess)
        return fail(S, QILQR_ERR_OUT_OF_MEMORY, "out of device memory for the linearisation records");
      const size_t smem = sizeof(double) * tp::smem_doubles(dq);
      if (S->tail_scratch.ensure(sizeof(double) * size_t(N) * 52 * size_t(pr.B)) != cudaSuccess)
        return fail(S, QILQR_ERR_OUT_OF_MEMORY, "out of device memory for the tail kernel's scratch gains");
      double *scr = S->tail_scratch.as<double>();
      if (dq) k_tail_persistent<true><<<tiles, 96, smem, st_>>>(S->p, pr, st, S->rec_tail_d.as<double>(), scr, alive, n_alive, epoch);
      else k_tail_persistent<false><<<tiles, 96, smem, st_>>>(S->p, pr, st, S->rec_tail_d.as<double>(), scr, alive, n_alive, epoch);
      ++S->launches;
      persistent_launched = true;
      break;
    }
    if (n_active > 0) {
      BackwardArgs ba{pr, st, active, n_active, epoch, P_alpha > 1 ? (S->wide_first ? 2 : 1) : 0, 1, nullptr, nullptr};
      {
        SpanGuard g(S, 0);
        if ((rc = launch_backward(S, ba))) return rc;
      }
      S->stats.backward_problem_knots += int64_t(n_active) * N;
      if (!S->in_tail) {
        S->stats.backward_problem_knots_bulk += int64_t(n_active) * N;
        ++S->stats.backward_launches_bulk;
      }
      ++S->launches;
    }
    if (P_alpha > 1 && epoch > 0 && (S->wide_first || n_alive > n_active)) {  // someone is backtracking
      // parallel line search: P_alpha step sizes per problem as independent cost-only rollouts, then the first
      // (largest) one that passes the Armijo test -- what the sequential search would accept
      RolloutArgs rw{pr, st, alive, n_alive, epoch, MODE_WIDE, nullptr, nullptr, nullptr, S->wide_d.as<double>(), P_alpha, 1, 0};
      {
        SpanGuard g(S, 1);
        launch_rollout(S, rw, n_alive * P_alpha, st_);
      }
      k_select_alpha<<<blocks_for(n_alive, 128), 128, 0, st_>>>(S->p, st, alive, n_alive, B_eff, S->wide_d.as<double>(), P_alpha, epoch);
      S->launches += 2;
    }
    {
      // forward_sim + cost + Armijo / convergence bookkeeping for every problem that is searching at its alpha[b]
      RolloutArgs ra{pr, st, alive, n_alive, epoch, MODE_SOLVE, nullptr, nullptr, nullptr, nullptr, 1, 1, P_alpha > 1 ? 1 : 0};
      {
        SpanGuard g(S, 1);
        launch_rollout(S, ra, n_alive, st_);
      }
      ++S->launches;
      if (d_debug && debug_cap > 0 && S->opt.populate_debug) {
        dim3 grid(blocks_for(n_alive, 128), 32);
        k_debug_capture<<<grid, 128, 0, st_>>>(pr, st, alive, n_alive, epoch, d_debug, debug_cap);
        ++S->launches;
      }
      if (ring_debug) {
        k_debug_ring_capture<<<S->dbg_S, 128, 0, st_>>>(pr, st, dr, epoch);
        ++S->launches;
      }
    }
    launch_compact(S, alive, n_alive, st.phase, listL[1 - cur], listA[1 - cur], kMaskAlive, kMaskActive);
    if ((rc = wait_counts(S, st_))) return rc;
    drain_spans(S);
    ++S->stats.solver_iterations;
    cur = 1 - cur;
    alive = listL[cur];
    active = listA[cur];
    n_alive = S->h_counts[0];
    n_active = S->h_counts[1];
  }
  cx.pr = pr; cx.st = st;
  cx.pending = true;
  if (async_tail && persistent_launched) return QILQR_OK;  // qilqr_solve_device_finish completes it
  return solve_finish(S, cx);
}

// Scatter the tail's results back, write the result records, rejoin the main stream and wait.
int solve_finish(qilqr_solver *S, SolveCtx &cx) {
  if (!cx.pending) return QILQR_OK;
  cx.pending = false;
  QCUDA(S, cudaSetDevice(S->device));
  cudaStream_t st_ = S->cur;
  const int B = cx.B;
  if (cx.compacted) {
    dim3 grid(blocks_for(cx.m_tail, 128), 32);
    k_tail_scatter<<<grid, 128, 0, st_>>>(cx.pr_big, cx.st_big, cx.pr, cx.st, S->tail_map.as<int>(), cx.m_tail);
    ++S->launches;
  }
  cudaMemsetAsync(S->totals_d.ptr, 0, 2 * sizeof(long long), st_);
  k_finalize<<<blocks_for(B, 256), 256, 0, st_>>>(cx.st_big, B, cx.d_results, S->totals_d.as<unsigned long long>());
  cudaMemcpyAsync(S->h_totals, S->totals_d.ptr, 2 * sizeof(long long), cudaMemcpyDeviceToHost, st_);
  {
    dim3 grid(blocks_for(B, 128), 64);
    k_collect<<<grid, 128, 0, st_>>>(cx.pr_big, cx.st_big.sel);
  }
  S->launches += 2;
  if (cx.on_hi && S->tail_stream) {  // rejoin the solver's main stream
    cudaEventRecord(S->ev_switch, st_);
    cudaStreamWaitEvent(S->stream, S->ev_switch, 0);
    S->cur = S->stream;
  }
  S->in_tail = false;
  QCUDA(S, cudaStreamSynchronize(S->stream));
  QCUDA(S, cudaGetLastError());
  S->stats.problem_iterations = S->h_totals[0];
  S->stats.problem_rollouts = S->h_totals[1];
  S->stats.rollout_problem_knots = S->h_totals[1] * int64_t(cx.N);
  S->stats.kernel_launches = S->launches - cx.launches0;
  {
    const auto t_end = std::chrono::steady_clock::now();
    if (!cx.on_hi) cx.t_switch = t_end;
    S->stats.bulk_wall_ms = std::chrono::duration<double, std::milli>(cx.t_switch - cx.t_begin).count();
    S->stats.tail_wall_ms = std::chrono::duration<double, std::milli>(t_end - cx.t_switch).count();
  }
  return QILQR_OK;
}

int pack_traj(qilqr_solver *S, int B, int N, const double *d_aos, double *d_soa) {
  dim3 grid(blocks_for(B, 32), unsigned(std::min((N + PACK_K - 1) / PACK_K, 16)));
  k_pack<<<grid, PACK_THREADS, 0, S->stream>>>(d_aos, d_soa, B, N);
  ++S->launches;
  return QILQR_OK;
}
int unpack_traj(qilqr_solver *S, int B, int N, const double *d_soa, double *d_aos) {
  dim3 grid(blocks_for(B, 32), unsigned(std::min((N + PACK_K - 1) / PACK_K, 16)));
  k_unpack<<<grid, PACK_THREADS, 0, S->stream>>>(d_soa, d_aos, B, N);
  ++S->launches;
  return QILQR_OK;
}
int transpose_to_soa(qilqr_solver *S, const double *in, double *out, int B, int N, int W) {
  const size_t total = size_t(B) * N * W;
  const unsigned blocks = unsigned(std::min<size_t>((total + 255) / 256, 148 * 16));
  k_transpose_bnw_to_nwb<<<blocks, 256, 0, S->stream>>>(in, out, B, N, W);
  ++S->launches;
  return QILQR_OK;
}
int transpose_to_aos(qilqr_solver *S, const double *in, double *out, int B, int N, int W) {
  const size_t total = size_t(B) * N * W;
  const unsigned blocks = unsigned(std::min<size_t>((total + 255) / 256, 148 * 16));
  k_transpose_nwb_to_bnw<<<blocks, 256, 0, S->stream>>>(in, out, B, N, W);
  ++S->launches;
  return QILQR_OK;
}

// Upload an AoS host trajectory batch and pack it to SoA.
int upload_traj(qilqr_solver *S, DeviceBuffer &stage, int B, int N, const double *h_aos, double *d_soa) {
  const size_t bytes = sizeof(double) * size_t(B) * N * 18;
  QCUDA(S, stage.ensure(bytes));
  QCUDA(S, cudaMemcpyAsync(stage.ptr, h_aos, bytes, cudaMemcpyHostToDevice, S->stream));
  return pack_traj(S, B, N, stage.as<double>(), d_soa);
}

}  // namespace

// =============================================================================
// C ABI
// =============================================================================
extern "C" {

const char *qilqr_error_string(int err) {
  switch (err) {
    case QILQR_OK: return "ok";
    case QILQR_ERR_INVALID_ARGUMENT: return "invalid argument";
    case QILQR_ERR_INERTIA_NOT_PD: return "Inertia matrix is not positive definite!";
    case QILQR_ERR_OUT_OF_RANGE: return "trajectory longer than the desired trajectory";
    case QILQR_ERR_NO_DEVICE: return "no sm_100 CUDA device (this library has no CPU fallback)";
    case QILQR_ERR_CUDA: return "CUDA error";
    case QILQR_ERR_OUT_OF_MEMORY: return "out of device memory";
    case QILQR_ERR_LINE_SEARCH: return "Reached maximum number of line search iterations";
    default: return "unknown error";
  }
}

const char *qilqr_build_info(void) {
#if defined(QILQR_STRICT) && defined(QILQR_PORTABLE_LIBM)
  return "strict+portable-libm: no fused multiply-adds, true divisions, sin/cos/atan2 from qilqr_portable_libm.h "
         "(-DQILQR_STRICT -DQILQR_PORTABLE_LIBM -fmad=false)";
#elif defined(QILQR_STRICT)
  return "strict: no fused multiply-adds, true divisions (-DQILQR_STRICT -fmad=false)";
#else
  return "production: explicit fused multiply-adds, reciprocal multiplies (-fmad=false)";
#endif
}

int qilqr_create(const qilqr_model_t *model, const double *Q, const double *R, double dt_s,
                 const qilqr_options_t *options, int device, qilqr_solver_t **out) {
  if (!model || !Q || !R || !options || !out) return QILQR_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  DeviceParams p{};
  if (!factor_inertia(model->inertia, p.L)) return QILQR_ERR_INERTIA_NOT_PD;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return QILQR_ERR_NO_DEVICE;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10) return QILQR_ERR_NO_DEVICE;
  if (cudaSetDevice(device) != cudaSuccess) return QILQR_ERR_NO_DEVICE;

  qilqr_solver *S = new qilqr_solver();
  S->device = device;
  p.mass = model->mass_kg;
  p.g = model->g_mpss;
  p.dt = dt_s;
  for (int i = 0; i < 9; ++i) p.inertia[i] = model->inertia[i];
  for (int i = 0; i < 3; ++i) p.Linv[i] = 1.0 / p.L[4 * i];
  const double a = model->arm_length_m, r = model->torque_to_thrust_ratio_m;
  const double ma[12] = {0, -a, 0, a, a, 0.0, -a, 0.0, -r, r, -r, r};  // quadrotor_model.cc:15-18
  for (int i = 0; i < 12; ++i) p.moment_arms[i] = ma[i];
  for (int j = 0; j < 4; ++j) {
    p.JuC[j] = 1.0 / model->mass_kg;  // quadrotor_model.cc:115-116
    const double col[3] = {ma[j], ma[4 + j], ma[8 + j]};
    double sol[3];
    llt_solve(p.L, col, sol);  // quadrotor_model.cc:118-119
    for (int i = 0; i < 3; ++i) p.JuC[4 * (1 + i) + j] = sol[i];
  }
  for (int i = 0; i < 16; ++i) p.Bu[i] = dt_s * p.JuC[i];
  for (int i = 0; i < 144; ++i) p.Q[i] = Q[i];
  for (int i = 0; i < 16; ++i) p.R[i] = R[i];
  p.q_diagonal = 1;
  for (int i = 0; i < 12; ++i)
    for (int j = 0; j < 12; ++j)
      if (i != j && Q[12 * i + j] != 0.0) p.q_diagonal = 0;
  if (const char *e = std::getenv("QILQR_Q_DIAGONAL")) p.q_diagonal = p.q_diagonal && std::atoi(e) != 0;
  S->p = p;
  for (int i = 0; i < 6; ++i)
    for (int j = 6; j < 12; ++j)
      if (Q[12 * i + j] != 0.0 || Q[12 * j + i] != 0.0) S->q_block_diagonal = false;
  if (const char *e = std::getenv("QILQR_BACKWARD")) {  // debugging aid: "t1" forces the per-thread kernel
    if (std::string(e) == "t1") S->force_t1 = true;
    if (std::string(e) == "split") S->split_backward = true;
    if (std::string(e) == "fused") S->split_backward = false;
  }
  if (const char *e = std::getenv("QILQR_ROLLOUT")) S->rollout_ws = std::string(e) != "thread";
  if (const char *e = std::getenv("QILQR_TAIL_COMPACTION")) S->tail_compaction = std::atoi(e) != 0;
  if (const char *e = std::getenv("QILQR_COMPACT_THRESHOLD")) S->compact_threshold = std::max(0, std::atoi(e));
  if (const char *e = std::getenv("QILQR_PERSISTENT_TAIL")) S->persistent_tail = std::atoi(e) != 0;
  if (const char *e = std::getenv("QILQR_PERSIST_THRESHOLD")) S->persist_threshold = std::atoi(e);
  if (const char *e = std::getenv("QILQR_ALWAYS_HIST_CAP")) S->always_hist_cap = std::max(0, std::atoi(e));
  if (const char *e = std::getenv("QILQR_WS_THRESHOLD")) S->ws_threshold = std::atoi(e);
  if (const char *e = std::getenv("QILQR_G16_THRESHOLD")) S->g16_threshold = std::atoi(e);
  if (const char *e = std::getenv("QILQR_WIDE_FIRST")) S->wide_first = std::atoi(e) != 0;
  if (const char *e = std::getenv("QILQR_KPP")) {
    const int v = std::atoi(e);
    if (v == 1 || v == 2 || v == 4) S->g4_kpp = v;
  }
  S->opt = *options;
  apply_options(S);
  int prio_least = 0, prio_greatest = 0;
  cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
  if (const char *e = std::getenv("QILQR_HI_THRESHOLD")) S->hi_threshold = std::atoi(e);
  if (const char *e = std::getenv("QILQR_BULK_POLL_US")) S->bulk_poll_us = std::atoi(e);
  if (const char *e = std::getenv("QILQR_TAIL_STREAM")) S->tail_stream = std::atoi(e) != 0;
  if (cudaStreamCreateWithPriority(&S->stream, cudaStreamNonBlocking, prio_least) != cudaSuccess ||
      cudaStreamCreateWithPriority(&S->stream_hi, cudaStreamNonBlocking, prio_greatest) != cudaSuccess ||
      cudaEventCreateWithFlags(&S->ev_switch, cudaEventDisableTiming) != cudaSuccess ||
      cudaHostAlloc(reinterpret_cast<void **>(&S->h_counts), 4 * sizeof(int), cudaHostAllocMapped) != cudaSuccess ||
      cudaHostGetDevicePointer(reinterpret_cast<void **>(&S->d_counts), S->h_counts, 0) != cudaSuccess ||
      cudaHostAlloc(reinterpret_cast<void **>(&S->h_totals), 2 * sizeof(long long), cudaHostAllocDefault) != cudaSuccess ||
      S->totals_d.ensure(2 * sizeof(long long)) != cudaSuccess || configure_kernels() != cudaSuccess) {
    qilqr_destroy(S);
    return QILQR_ERR_CUDA;
  }
  S->h_totals[0] = S->h_totals[1] = 0;
  S->cur = S->stream;
  std::memset(S->h_counts, 0, 4 * sizeof(int));
  *out = S;
  return QILQR_OK;
}

void qilqr_destroy(qilqr_solver_t *S) {
  if (!S) return;
  cudaSetDevice(S->device);
  if (S->stream) cudaStreamSynchronize(S->stream);
  if (S->stream_hi) cudaStreamSynchronize(S->stream_hi);
  for (DeviceBuffer *b : {&S->buf1, &S->gk, &S->gK, &S->state_d, &S->state_i, &S->lists, &S->desired_soa,
                          &S->traj_soa, &S->stage_a, &S->stage_b, &S->stage_c, &S->results_d, &S->hist_d,
                          &S->debug_d, &S->misc, &S->wide_d, &S->rec_d, &S->totals_d, &S->tail_traj, &S->tail_gains,
                          &S->tail_des, &S->tail_sd, &S->tail_si, &S->tail_hist, &S->tail_map, &S->tail_lists,
                          &S->rec_tail_d, &S->dbg_sample, &S->dbg_traj, &S->dbg_iters, &S->dbg_costs, &S->dbg_count, &S->stage_d, &S->tail_scratch})
    b->release();
  for (auto e : S->event_pool) cudaEventDestroy(e);
  if (S->user_lib) cudaLibraryUnload(S->user_lib);
  if (S->h_counts) cudaFreeHost(S->h_counts);
  if (S->h_totals) cudaFreeHost(S->h_totals);
  if (S->ev_switch) cudaEventDestroy(S->ev_switch);
  if (S->stream_hi) cudaStreamDestroy(S->stream_hi);
  if (S->stream) cudaStreamDestroy(S->stream);
  delete S;
}

int qilqr_set_options(qilqr_solver_t *S, const qilqr_options_t *options) {
  if (!S || !options) return QILQR_ERR_INVALID_ARGUMENT;
  S->opt = *options;
  apply_options(S);
  return QILQR_OK;
}
// ---- user-supplied model: NVRTC ---------------------------------------------------------------
namespace {
struct Nvrtc {  // libnvrtc, loaded on first use (the library itself does not link against it)
  void *lib = nullptr;
  int (*CreateProgram)(void **, const char *, const char *, int, const char *const *, const char *const *) = nullptr;
  int (*DestroyProgram)(void **) = nullptr;
  int (*CompileProgram)(void *, int, const char *const *) = nullptr;
  int (*AddNameExpression)(void *, const char *) = nullptr;
  int (*GetLoweredName)(void *, const char *, const char **) = nullptr;
  int (*GetCUBINSize)(void *, size_t *) = nullptr;
  int (*GetCUBIN)(void *, char *) = nullptr;
  int (*GetProgramLogSize)(void *, size_t *) = nullptr;
  int (*GetProgramLog)(void *, char *) = nullptr;
  bool ok = false;
  Nvrtc() {
    for (const char *name : {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"}) {
      lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
      if (lib) break;
    }
    if (!lib) return;
#define QSYM(field, sym) field = reinterpret_cast<decltype(field)>(dlsym(lib, sym)); if (!field) return
    QSYM(CreateProgram, "nvrtcCreateProgram");
    QSYM(DestroyProgram, "nvrtcDestroyProgram");
    QSYM(CompileProgram, "nvrtcCompileProgram");
    QSYM(AddNameExpression, "nvrtcAddNameExpression");
    QSYM(GetLoweredName, "nvrtcGetLoweredName");
    QSYM(GetCUBINSize, "nvrtcGetCUBINSize");
    QSYM(GetCUBIN, "nvrtcGetCUBIN");
    QSYM(GetProgramLogSize, "nvrtcGetProgramLogSize");
    QSYM(GetProgramLog, "nvrtcGetProgramLog");
#undef QSYM
    ok = true;
  }
};
std::string source_dir() {  // <dir of this shared library>/csrc, or QILQR_CSRC_DIR
  if (const char *e = std::getenv("QILQR_CSRC_DIR")) return e;
  Dl_info info;
  if (dladdr(reinterpret_cast<const void *>(&qilqr_build_info), &info) && info.dli_fname) {
    std::string path = info.dli_fname;
    const size_t slash = path.rfind('/');
    return (slash == std::string::npos ? std::string(".") : path.substr(0, slash)) + "/csrc";
  }
  return "csrc";
}
}  // namespace

int qilqr_set_user_model(qilqr_solver_t *S, const char *cuda_source, const double *params, int n_params) {
  if (!S || !cuda_source || n_params < 0 || n_params > 64 || (n_params && !params)) return QILQR_ERR_INVALID_ARGUMENT;
  QENTER(S);
  static Nvrtc nv;
  if (!nv.ok) return fail(S, QILQR_ERR_CUDA, "libnvrtc.so.12 could not be loaded: a user-supplied model needs the CUDA run-time compiler");
  QCUDA(S, cudaSetDevice(S->device));
  const std::string dir = source_dir();
  // the caller's function sits between the device library (which it may use) and the generic kernels (which call it)
  const std::string unit = std::string("#define QILQR_USER_MODEL_TU 1\n#include \"qilqr_device.cuh\"\n#line 1 \"user_model.cu\"\n") +
                           cuda_source + "\n#include \"qilqr_backward_dense.cuh\"\n";
  void *prog = nullptr;
  if (nv.CreateProgram(&prog, unit.c_str(), "qilqr_user_model_unit.cu", 0, nullptr, nullptr) != 0)
    return fail(S, QILQR_ERR_CUDA, "nvrtcCreateProgram failed");
  const std::string inc1 = "-I" + dir, inc2 = "-I" + dir + "/../../include";
  const char *opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-fmad=false", "-default-device", inc1.c_str(), inc2.c_str()};
  const char *n_lin = "qilqr::k_linearise_dense", *n_roll = "qilqr::k_rollout<true>";
  nv.AddNameExpression(prog, n_lin);
  nv.AddNameExpression(prog, n_roll);
  const int rc = nv.CompileProgram(prog, int(sizeof(opts) / sizeof(opts[0])), opts);
  if (rc != 0) {
    size_t n = 0;
    nv.GetProgramLogSize(prog, &n);
    std::string log(n, '\0');
    if (n) nv.GetProgramLog(prog, &log[0]);
    nv.DestroyProgram(&prog);
    S->last_error = "the user model did not compile:\n" + log;
    return QILQR_ERR_INVALID_ARGUMENT;
  }
  const char *l_lin = nullptr, *l_roll = nullptr;
  size_t nb = 0;
  if (nv.GetLoweredName(prog, n_lin, &l_lin) != 0 || nv.GetLoweredName(prog, n_roll, &l_roll) != 0 ||
      nv.GetCUBINSize(prog, &nb) != 0 || nb == 0) {
    nv.DestroyProgram(&prog);
    return fail(S, QILQR_ERR_CUDA, "NVRTC produced no code for the user-model kernels");
  }
  std::vector<char> cubin(nb);
  nv.GetCUBIN(prog, cubin.data());
  const std::string s_lin = l_lin, s_roll = l_roll;
  nv.DestroyProgram(&prog);
  cudaLibrary_t lib = nullptr;
  QCUDA(S, cudaLibraryLoadData(&lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
  cudaKernel_t k_lin = nullptr, k_roll = nullptr;
  void *d_params = nullptr;
  size_t param_bytes = 0;
  if (cudaLibraryGetKernel(&k_lin, lib, s_lin.c_str()) != cudaSuccess || cudaLibraryGetKernel(&k_roll, lib, s_roll.c_str()) != cudaSuccess ||
      cudaLibraryGetGlobal(&d_params, &param_bytes, lib, "_ZN5qilqr11user_paramsE") != cudaSuccess) {
    cudaLibraryUnload(lib);
    return fail(S, QILQR_ERR_CUDA, "the user-model kernels were not found in the compiled unit");
  }
  if (n_params) QCUDA(S, cudaMemcpy(d_params, params, sizeof(double) * size_t(n_params), cudaMemcpyHostToDevice));
  if (S->user_lib) cudaLibraryUnload(S->user_lib);
  S->user_lib = lib;
  S->user_linearise = k_lin;
  S->user_rollout = k_roll;
  S->user_model = true;
  S->generic_path = true;
  return QILQR_OK;
}

int qilqr_set_model_variant(qilqr_solver_t *S, int model_flags) {
  if (!S || model_flags < 0 || model_flags > 7) return QILQR_ERR_INVALID_ARGUMENT;
  S->user_model = false;
  S->p.integrator = (model_flags & QILQR_MODEL_RK4) ? 1 : 0;
  S->p.coriolis = (model_flags & QILQR_MODEL_CORIOLIS) ? 1 : 0;
  S->generic_path = model_flags != 0;
  return QILQR_OK;
}
const char *qilqr_last_error_message(const qilqr_solver_t *S) { return S ? S->last_error.c_str() : ""; }
int64_t qilqr_kernel_launch_count(const qilqr_solver_t *S) { return S ? S->launches : 0; }
void *qilqr_stream(const qilqr_solver_t *S) { return S ? static_cast<void *>(S->stream) : nullptr; }
int qilqr_last_solve_stats(const qilqr_solver_t *S, qilqr_solve_stats_t *out) {
  if (!S || !out) return QILQR_ERR_INVALID_ARGUMENT;
  *out = S->stats;
  return QILQR_OK;
}
int qilqr_set_profiling(qilqr_solver_t *S, int enabled) {
  if (!S) return QILQR_ERR_INVALID_ARGUMENT;
  S->profiling = enabled != 0;
  return QILQR_OK;
}

// ---- device-resident API ------------------------------------------------------
int qilqr_solve_device(qilqr_solver_t *S, int batch, int n_knots, const double *d_desired, int desired_count,
                       double *d_traj_inout, double *d_k, double *d_K, double *d_cost_hist, int hist_cap,
                       qilqr_result_t *d_results) {
  if (!S || !d_desired || !d_traj_inout) return QILQR_ERR_INVALID_ARGUMENT;
  std::lock_guard<std::mutex> lock(S->mu);
  return solve_core(S, batch, n_knots, d_desired, desired_count, d_traj_inout, d_k, d_K, d_cost_hist, hist_cap,
                    d_results, nullptr, 0);
}
int qilqr_solve_device_begin(qilqr_solver_t *S, int batch, int n_knots, const double *d_desired, int desired_count,
                             double *d_traj_inout, double *d_k, double *d_K, double *d_cost_hist, int hist_cap,
                             qilqr_result_t *d_results) {
  if (!S || !d_desired || !d_traj_inout) return QILQR_ERR_INVALID_ARGUMENT;
  std::lock_guard<std::mutex> lock(S->mu);
  return solve_core(S, batch, n_knots, d_desired, desired_count, d_traj_inout, d_k, d_K, d_cost_hist, hist_cap,
                    d_results, nullptr, 0, /*async_tail=*/true);
}
int qilqr_solve_device_finish(qilqr_solver_t *S) {
  if (!S) return QILQR_ERR_INVALID_ARGUMENT;
  std::lock_guard<std::mutex> lock(S->mu);
  return solve_finish(S, S->ctx);
}
int qilqr_pack_trajectory_device(qilqr_solver_t *S, int batch, int n_knots, const double *d_aos, double *d_soa) {
  if (!S || !d_aos || !d_soa) return QILQR_ERR_INVALID_ARGUMENT;
  QENTER(S);
  QCUDA(S, cudaSetDevice(S->device));
  return pack_traj(S, batch, n_knots, d_aos, d_soa);
}
int qilqr_unpack_trajectory_device(qilqr_solver_t *S, int batch, int n_knots, const double *d_soa,
                                   const double *d_time_aos, double *d_aos) {
  if (!S || !d_aos || !d_soa) return QILQR_ERR_INVALID_ARGUMENT;
  QENTER(S);
  QCUDA(S, cudaSetDevice(S->device));
  const size_t bytes = sizeof(double) * size_t(batch) * n_knots * 18;
  if (d_time_aos && d_time_aos != d_aos)
    QCUDA(S, cudaMemcpyAsync(d_aos, d_time_aos, bytes, cudaMemcpyDeviceToDevice, S->stream));
  else if (!d_time_aos)
    QCUDA(S, cudaMemsetAsync(d_aos, 0, bytes, S->stream));
  return unpack_traj(S, batch, n_knots, d_soa, d_aos);
}
int qilqr_rollout_constant_control_device(qilqr_solver_t *S, int batch, int n_knots, const double *d_x0_soa,
                                          const double *u, double *d_traj_soa) {
  if (!S || !d_x0_soa || !u || !d_traj_soa) return QILQR_ERR_INVALID_ARGUMENT;
  QENTER(S);
  QCUDA(S, cudaSetDevice(S->device));
  k_rollout_constant<<<blocks_for(batch, 128), 128, 0, S->stream>>>(S->p, d_x0_soa, u[0], u[1], u[2], u[3],
                                                                   d_traj_soa, batch, n_knots);
  ++S->launches;
  QCUDA(S, cudaGetLastError());
  return QILQR_OK;
}

int qilqr_mpc_advance_device(qilqr_solver_t *S, int B, int N, double *d_traj, double *d_plant, const double *d_dist,
                             double *d_applied_u) {
  if (!S || !d_traj || !d_plant || B <= 0 || N <= 0) return QILQR_ERR_INVALID_ARGUMENT;
  QENTER(S);
  QCUDA(S, cudaSetDevice(S->device));
  k_mpc_advance<<<blocks_for(B, 128), 128, 0, S->stream>>>(S->p, d_traj, d_plant, d_dist, d_applied_u, B, N);
  ++S->launches;
  QCUDA(S, cudaGetLastError());
  return QILQR_OK;
}

int qilqr_mpc_run_device(qilqr_solver_t *S, int steps, int B, int N, const double *d_desired, int Bd, double *d_traj,
                         double *d_plant, const double *d_dist, double *d_state_log, double *d_control_log,
                         int64_t *totals) {
  if (!S || !d_desired || !d_traj || !d_plant || steps < 0) return QILQR_ERR_INVALID_ARGUMENT;
  QENTER(S);
  QCUDA(S, cudaSetDevice(S->device));
  QCUDA(S, S->results_d.ensure(sizeof(qilqr_result_t) * size_t(B)));
  std::vector<qilqr_result_t> res(totals ? B : 0);
  int64_t acc[4] = {0, 0, 0, 0};
  qilqr_solve_stats_t total_stats{};
  for (int t = 0; t < steps; ++t) {
    int rc = solve_core(S, B, N, d_desired, Bd, d_traj, nullptr, nullptr, nullptr, 0,
                        S->results_d.as<qilqr_result_t>(), nullptr, 0);
    if (rc) return rc;
    total_stats.solver_iterations += S->stats.solver_iterations;
    total_stats.problem_iterations += S->stats.problem_iterations;
    total_stats.problem_rollouts += S->stats.problem_rollouts;
    total_stats.kernel_launches += S->stats.kernel_launches + 1;
    total_stats.backward_ms += S->stats.backward_ms;
    total_stats.rollout_ms += S->stats.rollout_ms;
    total_stats.backward_problem_knots += S->stats.backward_problem_knots;
    total_stats.rollout_problem_knots += S->stats.rollout_problem_knots;
    total_stats.backward_ms_bulk += S->stats.backward_ms_bulk;
    total_stats.rollout_ms_bulk += S->stats.rollout_ms_bulk;
    total_stats.backward_problem_knots_bulk += S->stats.backward_problem_knots_bulk;
    total_stats.backward_launches_bulk += S->stats.backward_launches_bulk;
    total_stats.bulk_wall_ms += S->stats.bulk_wall_ms;
    total_stats.tail_wall_ms += S->stats.tail_wall_ms;
    double *u_log = d_control_log ? d_control_log + size_t(t) * 4 * B : nullptr;
    k_mpc_advance<<<blocks_for(B, 128), 128, 0, S->stream>>>(S->p, d_traj, d_plant, d_dist, u_log, B, N);
    ++S->launches;
    if (d_state_log)
      QCUDA(S, cudaMemcpyAsync(d_state_log + size_t(t) * 13 * B, d_plant, sizeof(double) * 13 * size_t(B),
                               cudaMemcpyDeviceToDevice, S->stream));
    if (totals) {
      QCUDA(S, cudaMemcpyAsync(res.data(), S->results_d.ptr, sizeof(qilqr_result_t) * size_t(B),
                               cudaMemcpyDeviceToHost, S->stream));
      QCUDA(S, cudaStreamSynchronize(S->stream));
      for (int b = 0; b < B; ++b) {
        acc[0] += res[b].backward_passes;
        acc[1] += res[b].rollouts;
        acc[2] += (res[b].status == QILQR_STATUS_MAX_ITERS || res[b].status >= QILQR_STATUS_LINE_SEARCH_FAILED);
      }
      acc[3] += B;
    }
  }
  QCUDA(S, cudaStreamSynchronize(S->stream));
  QCUDA(S, cudaGetLastError());
  S->stats = total_stats;
  if (totals) for (int i = 0; i < 4; ++i) totals[i] = acc[i];
  return QILQR_OK;
}

// ---- host-buffer API -----------------------------------------------------------
namespace {
// Second half of qilqr_solve_host: wait for the solve, bring the results back in the caller's layout.
int solve_host_finish(qilqr_solver *S) {
  HostCall &hc = S->host_call;
  if (!hc.pending) return fail(S, QILQR_ERR_INVALID_ARGUMENT, "no host solve is pending");
  hc.pending = false;
  int rc = solve_finish(S, S->ctx);
  if (rc) return rc;
  cudaStream_t st_ = S->stream;
  const int B = hc.B, N = hc.N;
  const size_t traj_bytes = sizeof(double) * size_t(B) * N * 18;
  if (hc.out_traj) {
    // results back: stage_a still holds the input AoS (time_s column), unpack over it
    unpack_traj(S, B, N, S->traj_soa.as<double>(), S->stage_a.as<double>());
    QCUDA(S, cudaMemcpyAsync(hc.out_traj, S->stage_a.ptr, traj_bytes, cudaMemcpyDeviceToHost, st_));
  }
  if (hc.out_controls) {
    const size_t n4 = size_t(B) * N * 4;
    QCUDA(S, S->stage_d.ensure(sizeof(double) * n4));
    k_extract_controls<<<unsigned((n4 + 255) / 256), 256, 0, st_>>>(S->traj_soa.as<double>(), S->stage_d.as<double>(), B, N);
    ++S->launches;
    QCUDA(S, cudaMemcpyAsync(hc.out_controls, S->stage_d.ptr, sizeof(double) * n4, cudaMemcpyDeviceToHost, st_));
  }
  if (hc.results) QCUDA(S, cudaMemcpyAsync(hc.results, S->results_d.ptr, sizeof(qilqr_result_t) * size_t(B), cudaMemcpyDeviceToHost, st_));
  if (hc.out_k) {
    QCUDA(S, S->stage_c.ensure(sizeof(double) * size_t(N) * 48 * B));
    transpose_to_aos(S, S->gk.as<double>(), S->stage_c.as<double>(), B, N, 4);
    QCUDA(S, cudaMemcpyAsync(hc.out_k, S->stage_c.ptr, sizeof(double) * size_t(N) * 4 * B, cudaMemcpyDeviceToHost, st_));
  }
  if (hc.out_K) {
    if (hc.out_k) QCUDA(S, cudaStreamSynchronize(st_));  // stage_c is reused
    QCUDA(S, S->stage_c.ensure(sizeof(double) * size_t(N) * 48 * B));
    transpose_to_aos(S, S->gK.as<double>(), S->stage_c.as<double>(), B, N, 48);
    QCUDA(S, cudaMemcpyAsync(hc.out_K, S->stage_c.ptr, sizeof(double) * size_t(N) * 48 * B, cudaMemcpyDeviceToHost, st_));
  }
  if (hc.cost_hist) {  // [hist_cap][B] -> [B][hist_cap]
    QCUDA(S, S->stage_b.ensure(sizeof(double) * size_t(hc.hist_cap) * B));
    transpose_to_aos(S, S->hist_d.as<double>(), S->stage_b.as<double>(), B, 1, hc.hist_cap);
    QCUDA(S, cudaMemcpyAsync(hc.cost_hist, S->stage_b.ptr, sizeof(double) * size_t(hc.hist_cap) * B, cudaMemcpyDeviceToHost, st_));
  }
  QCUDA(S, cudaStreamSynchronize(st_));
  if (hc.debug_traj) {
    // [cap][N][17][B] on the device -> [B][cap][N][18] on the host, one slot at a time
    double *d_debug = S->debug_d.as<double>();
    std::vector<qilqr_result_t> res(B);
    QCUDA(S, cudaMemcpy(res.data(), S->results_d.ptr, sizeof(qilqr_result_t) * size_t(B), cudaMemcpyDeviceToHost));
    int max_nd = 0;
    for (int b = 0; b < B; ++b) max_nd = std::max(max_nd, res[b].num_debug);
    max_nd = std::min(max_nd, hc.debug_cap);
    std::vector<double> slot(size_t(B) * N * 18);
    for (int sidx = 0; sidx < max_nd; ++sidx) {
      // stage_a still has time_s in column 0
      unpack_traj(S, B, N, d_debug + size_t(sidx) * N * 17 * B, S->stage_a.as<double>());
      QCUDA(S, cudaMemcpyAsync(slot.data(), S->stage_a.ptr, traj_bytes, cudaMemcpyDeviceToHost, st_));
      QCUDA(S, cudaStreamSynchronize(st_));
      for (int b = 0; b < B; ++b)
        if (sidx < res[b].num_debug)
          std::memcpy(hc.debug_traj + (size_t(b) * hc.debug_cap + sidx) * N * 18, slot.data() + size_t(b) * N * 18,
                      sizeof(double) * N * 18);
    }
  }
  QCUDA(S, cudaGetLastError());
  return QILQR_OK;
}

// First half: upload, transpose to structure-of-arrays, sequence the solve (all of it, or -- async -- up to the
// point where the device finishes it on its own).
struct ControlsInput {  // qilqr_solve_from_controls_host: the initial trajectory is rolled out on the device
  const double *x0 = nullptr;        // [B][13]
  const double *controls = nullptr;  // [Bc][N][4]
  int Bc = 0;
  double *out_controls = nullptr;
};
int solve_host_begin(qilqr_solver *S, int B, int N, const double *desired, int Bd, const double *initial,
                     double *out_traj, double *out_k, double *out_K, double *cost_hist, int hist_cap,
                     double *debug_traj, int debug_cap, qilqr_result_t *results, bool async,
                     const ControlsInput *ci = nullptr) {
  if (!S || !desired || (!initial && !ci) || (!out_traj && !(ci && ci->out_controls))) return QILQR_ERR_INVALID_ARGUMENT;
  if (B <= 0 || N <= 0 || (Bd != 1 && Bd != B)) return fail(S, QILQR_ERR_INVALID_ARGUMENT, "bad batch/knots/desired_count");
  if (S->ctx.pending || S->host_call.pending) return fail(S, QILQR_ERR_INVALID_ARGUMENT, "a solve is still pending on this handle");
  QCUDA(S, cudaSetDevice(S->device));
  cudaStream_t st_ = S->stream;
  QCUDA(S, S->traj_soa.ensure(sizeof(double) * size_t(N) * 17 * B));
  QCUDA(S, S->desired_soa.ensure(sizeof(double) * size_t(N) * 17 * Bd));
  int rc = upload_traj(S, S->stage_b, Bd, N, desired, S->desired_soa.as<double>());
  if (rc) return rc;
  if (ci) {
    // x0 and the nominal controls go up (13 + 4 N doubles per problem, or 13 when the controls are shared); the
    // initial trajectory is their open-loop rollout, made where it is needed
    QCUDA(S, S->stage_a.ensure(sizeof(double) * size_t(B) * N * 18));
    QCUDA(S, S->stage_c.ensure(sizeof(double) * (size_t(B) * 13 + size_t(ci->Bc) * N * 4)));
    double *d_x0 = S->stage_c.as<double>(), *d_u = d_x0 + size_t(B) * 13;
    QCUDA(S, cudaMemcpyAsync(d_x0, ci->x0, sizeof(double) * size_t(B) * 13, cudaMemcpyHostToDevice, st_));
    QCUDA(S, cudaMemcpyAsync(d_u, ci->controls, sizeof(double) * size_t(ci->Bc) * N * 4, cudaMemcpyHostToDevice, st_));
    k_rollout_controls<<<blocks_for(B, 128), 128, 0, st_>>>(S->p, d_x0, d_u, ci->Bc, S->traj_soa.as<double>(), B, N);
    if (out_traj) k_fill_time<<<blocks_for(B * N, 256), 256, 0, st_>>>(S->stage_a.as<double>(), B, N, S->p.dt);
    S->launches += 2;
  } else {
    rc = upload_traj(S, S->stage_a, B, N, initial, S->traj_soa.as<double>());
    if (rc) return rc;
  }
  double *d_k = nullptr, *d_K = nullptr, *d_hist = nullptr, *d_debug = nullptr;
  if (out_k) { QCUDA(S, S->gk.ensure(sizeof(double) * size_t(N) * 4 * B)); d_k = S->gk.as<double>(); }
  if (out_K) { QCUDA(S, S->gK.ensure(sizeof(double) * size_t(N) * 48 * B)); d_K = S->gK.as<double>(); }
  if (cost_hist && hist_cap > 0) {
    QCUDA(S, S->hist_d.ensure(sizeof(double) * size_t(hist_cap) * B));
    QCUDA(S, cudaMemsetAsync(S->hist_d.ptr, 0, sizeof(double) * size_t(hist_cap) * B, st_));
    d_hist = S->hist_d.as<double>();
  }
  const bool want_debug = debug_traj && debug_cap > 0 && S->opt.populate_debug;
  if (want_debug) {
    const size_t need = sizeof(double) * size_t(debug_cap) * N * 17 * B;
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    if (need > S->debug_d.bytes && need > free_b)
      return fail(S, QILQR_ERR_OUT_OF_MEMORY, "the ILQRDebug capture of every problem does not fit in device memory: sample "
                                               "problems with qilqr_set_debug_sampling or solve in smaller batches");
    QCUDA(S, S->debug_d.ensure(need));
    d_debug = S->debug_d.as<double>();
    QCUDA(S, cudaMemsetAsync(d_debug, 0, need, st_));  // slots a problem never reaches
  }
  QCUDA(S, S->results_d.ensure(sizeof(qilqr_result_t) * size_t(B)));
  S->dbg_time_aos = S->stage_a.as<double>();  // the uploaded input AoS: time_s of the sampled ILQRDebug trajectories
  HostCall &hc = S->host_call;
  hc = HostCall{};
  hc.B = B; hc.N = N;
  hc.out_traj = out_traj; hc.out_k = out_k; hc.out_K = out_K;
  hc.cost_hist = d_hist ? cost_hist : nullptr; hc.hist_cap = hist_cap;
  hc.debug_traj = want_debug ? debug_traj : nullptr; hc.debug_cap = debug_cap;
  hc.results = results;
  hc.out_controls = ci ? ci->out_controls : nullptr;
  rc = solve_core(S, B, N, S->desired_soa.as<double>(), Bd, S->traj_soa.as<double>(), d_k, d_K, d_hist, hist_cap,
                  S->results_d.as<qilqr_result_t>(), d_debug, debug_cap, /*async_tail=*/true);
  S->dbg_time_aos = nullptr;
  if (rc) return rc;
  hc.pending = true;
  return async ? QILQR_OK : solve_host_finish(S);
}
}  // namespace

int qilqr_solve_host(qilqr_solver_t *S, int B, int N, const double *desired, int Bd, const double *initial,
                     double *out_traj, double *out_k, double *out_K, double *cost_hist, int hist_cap,
                     double *debug_traj, int debug_cap, qilqr_result_t *results) {
  if (!S) return QILQR_ERR_INVALID_ARGUMENT;
  std::lock_guard<std::mutex> lock(S->mu);
  return solve_host_begin(S, B, N, desired, Bd, initial, out_traj, out_k, out_K, cost_hist, hist_cap, debug_traj,
                          debug_cap, results, false);
}
int qilqr_solve_host_begin(qilqr_solver_t *S, int B, int N, const double *desired, int Bd, const double *initial,
                           double *out_traj, qilqr_result_t *results) {
  if (!S) return QILQR_ERR_INVALID_ARGUMENT;
  std::lock_guard<std::mutex> lock(S->mu);
  return solve_host_begin(S, B, N, desired, Bd, initial, out_traj, nullptr, nullptr, nullptr, 0, nullptr, 0, results, true);
}
int qilqr_solve_from_controls_host(qilqr_solver_t *S, int B, int N, const double *desired, int Bd, const double *x0,
                                   const double *controls, int control_count, double *out_traj, double *out_controls,
                                   qilqr_result_t *results) {
  if (!S || !x0 || !controls || (control_count != 1 && control_count != B)) return QILQR_ERR_INVALID_ARGUMENT;
  std::lock_guard<std::mutex> lock(S->mu);
  ControlsInput ci{x0, controls, control_count, out_controls};
  return solve_host_begin(S, B, N, desired, Bd, nullptr, out_traj, nullptr, nullptr, nullptr, 0, nullptr, 0, results, false, &ci);
}
int qilqr_solve_from_controls_host_begin(qilqr_solver_t *S, int B, int N, const double *desired, int Bd, const double *x0,
                                         const double *controls, int control_count, double *out_traj,
                                         double *out_controls, qilqr_result_t *results) {
  if (!S || !x0 || !controls || (control_count != 1 && control_count != B)) return QILQR_ERR_INVALID_ARGUMENT;
  std::lock_guard<std::mutex> lock(S->mu);
  ControlsInput ci{x0, controls, control_count, out_controls};
  return solve_host_begin(S, B, N, desired, Bd, nullptr, out_traj, nullptr, nullptr, nullptr, 0, nullptr, 0, results, true, &ci);
}
int qilqr_solve_host_finish(qilqr_solver_t *S) {
  if (!S) return QILQR_ERR_INVALID_ARGUMENT;
  std::lock_guard<std::mutex> lock(S->mu);
  return solve_host_finish(S);
}

// ---- ILQRDebug at batch scale: sampling policy + ring buffer, always-on cost history ----------------------
int qilqr_set_debug_sampling(qilqr_solver_t *S, int every_kth_iteration, const int32_t *problems, int num_problems,
                             int ring_slots) {
  if (!S) return QILQR_ERR_INVALID_ARGUMENT;
  QENTER(S);
  if (num_problems <= 0 || !problems) {  // off
    S->dbg_S = S->dbg_ring = S->dbg_every = 0;
    return QILQR_OK;
  }
  if (every_kth_iteration < 1 || ring_slots < 1) return fail(S, QILQR_ERR_INVALID_ARGUMENT, "every_kth_iteration and ring_slots must be >= 1");
  QCUDA(S, cudaSetDevice(S->device));
  QCUDA(S, S->dbg_sample.ensure(sizeof(int) * size_t(num_problems)));
  QCUDA(S, cudaMemcpyAsync(S->dbg_sample.ptr, problems, sizeof(int) * size_t(num_problems), cudaMemcpyHostToDevice, S->stream));
  QCUDA(S, cudaStreamSynchronize(S->stream));
  S->dbg_S = num_problems;
  S->dbg_ring = ring_slots;
  S->dbg_every = every_kth_iteration;
  S->dbg_N = 0;
  return QILQR_OK;
}
int qilqr_read_debug_samples_host(qilqr_solver_t *S, double *traj, int32_t *iters, double *costs, int32_t *counts) {
  if (!S) return QILQR_ERR_INVALID_ARGUMENT;
  QENTER(S);
  if (S->dbg_S <= 0 || S->dbg_N <= 0) return fail(S, QILQR_ERR_INVALID_ARGUMENT, "no sampled ILQRDebug capture: call qilqr_set_debug_sampling, then solve with populate_debug");
  QCUDA(S, cudaSetDevice(S->device));
  const size_t slots = size_t(S->dbg_S) * S->dbg_ring;
  // one asynchronous copy per array on the solver's stream (pinned destinations make them true DMA transfers)
  if (traj) QCUDA(S, cudaMemcpyAsync(traj, S->dbg_traj.ptr, sizeof(double) * slots * S->dbg_N * 18, cudaMemcpyDeviceToHost, S->stream));
  if (iters) QCUDA(S, cudaMemcpyAsync(iters, S->dbg_iters.ptr, sizeof(int) * slots, cudaMemcpyDeviceToHost, S->stream));
  if (costs) QCUDA(S, cudaMemcpyAsync(costs, S->dbg_costs.ptr, sizeof(double) * slots, cudaMemcpyDeviceToHost, S->stream));
  if (counts) QCUDA(S, cudaMemcpyAsync(counts, S->dbg_count.ptr, sizeof(int) * size_t(S->dbg_S), cudaMemcpyDeviceToHost, S->stream));
  QCUDA(S, cudaStreamSynchronize(S->stream));
  return QILQR_OK;
}
int qilqr_last_cost_history_host(qilqr_solver_t *S, int first, int count, double *out, int out_cap, int *stored_cap) {
  if (!S || first < 0 || count < 0) return QILQR_ERR_INVALID_ARGUMENT;
  QENTER(S);
  if (stored_cap) *stored_cap = S->last_hist_cap;
  if (!out || count == 0) return QILQR_OK;
  if (S->last_hist_cap <= 0 || !S->last_hist_internal) return fail(S, QILQR_ERR_INVALID_ARGUMENT, "the last solve kept no internal cost history (the caller passed its own buffer, or QILQR_ALWAYS_HIST_CAP=0)");
  if (first + count > S->last_hist_B || out_cap < 1) return fail(S, QILQR_ERR_INVALID_ARGUMENT, "cost history range out of the last batch");
  QCUDA(S, cudaSetDevice(S->device));
  const int cap = std::min(out_cap, S->last_hist_cap);
  // [cap][B] on the device -> [count][out_cap] on the host
  std::vector<double> rows(size_t(cap) * count);
  QCUDA(S, cudaMemcpy2DAsync(rows.data(), sizeof(double) * count, S->hist_d.as<double>() + first,
                             sizeof(double) * S->last_hist_B, sizeof(double) * count, cap, cudaMemcpyDeviceToHost, S->stream));
  QCUDA(S, cudaStreamSynchronize(S->stream));
  for (int b = 0; b < count; ++b)
    for (int i = 0; i < out_cap; ++i) out[size_t(b) * out_cap + i] = i < cap ? rows[size_t(i) * count + b] : 0.0;
  return QILQR_OK;
}

int qilqr_forward_sim_host(qilqr_solver_t *S, int B, int N, const double *current, const double *k, const double *K,
                           const double *alpha, double *out_traj) {
  if (!S || !current || !k || !K || !alpha || !out_traj || B <= 0 || N <= 0) return QILQR_ERR_INVALID_ARGUMENT;
  QENTER(S);
  QCUDA(S, cudaSetDevice(S->device));
  cudaStream_t st_ = S->stream;
  QCUDA(S, S->traj_soa.ensure(sizeof(double) * size_t(N) * 17 * B));
  QCUDA(S, S->buf1.ensure(sizeof(double) * size_t(N) * 17 * B));
  QCUDA(S, S->gk.ensure(sizeof(double) * size_t(N) * 4 * B));
  QCUDA(S, S->gK.ensure(sizeof(double) * size_t(N) * 48 * B));
  QCUDA(S, S->stage_c.ensure(sizeof(double) * size_t(N) * 48 * B));
  QCUDA(S, S->misc.ensure(sizeof(double) * size_t(B)));
  int rc = upload_traj(S, S->stage_a, B, N, current, S->traj_soa.as<double>());
  if (rc) return rc;
  QCUDA(S, cudaMemcpyAsync(S->stage_c.ptr, k, sizeof(double) * size_t(N) * 4 * B, cudaMemcpyHostToDevice, st_));
  transpose_to_soa(S, S->stage_c.as<double>(), S->gk.as<double>(), B, N, 4);
  QCUDA(S, cudaStreamSynchronize(st_));
  QCUDA(S, cudaMemcpyAsync(S->stage_c.ptr, K, sizeof(double) * size_t(N) * 48 * B, cudaMemcpyHostToDevice, st_));
  transpose_to_soa(S, S->stage_c.as<double>(), S->gK.as<double>(), B, N, 48);
  QCUDA(S, cudaMemcpyAsync(S->misc.ptr, alpha, sizeof(double) * size_t(B), cudaMemcpyHostToDevice, st_));
  Problem pr{B, N, 1, nullptr, nullptr, nullptr, S->gk.as<double>(), S->gK.as<double>()};
  RolloutArgs ra{pr, SolveState{}, nullptr, B, 0, MODE_FORWARD, S->traj_soa.as<double>(), S->buf1.as<double>(),
                 S->misc.as<double>(), nullptr, 1};
  launch_rollout(S, ra, B, st_);
  ++S->launches;
  unpack_traj(S, B, N, S->buf1.as<double>(), S->stage_a.as<double>());
  QCUDA(S, cudaMemcpyAsync(out_traj, S->stage_a.ptr, sizeof(double) * size_t(B) * N * 18, cudaMemcpyDeviceToHost, st_));
  QCUDA(S, cudaStreamSynchronize(st_));
  QCUDA(S, cudaGetLastError());
  return QILQR_OK;
}

int qilqr_cost_trajectory_host(qilqr_solver_t *S, int B, int N, const double *desired, int Bd, int n_desired,
                               const double *traj, double *cost) {
  if (!S || !desired || !traj || !cost || B <= 0 || N <= 0 || (Bd != 1 && Bd != B)) return QILQR_ERR_INVALID_ARGUMENT;
  if (n_desired < N) return fail(S, QILQR_ERR_OUT_OF_RANGE, "vector::_M_range_check (cost.hh:39-40)");
  QENTER(S);
  QCUDA(S, cudaSetDevice(S->device));
  cudaStream_t st_ = S->stream;
  QCUDA(S, S->traj_soa.ensure(sizeof(double) * size_t(N) * 17 * B));
  QCUDA(S, S->desired_soa.ensure(sizeof(double) * size_t(n_desired) * 17 * Bd));
  QCUDA(S, S->misc.ensure(sizeof(double) * size_t(B)));
  int rc = upload_traj(S, S->stage_b, Bd, n_desired, desired, S->desired_soa.as<double>());
  if (rc) return rc;
  rc = upload_traj(S, S->stage_a, B, N, traj, S->traj_soa.as<double>());
  if (rc) return rc;
  k_cost_trajectory<<<blocks_for(B, 128), 128, 0, st_>>>(S->p, S->traj_soa.as<double>(), S->desired_soa.as<double>(),
                                                         B, N, Bd, S->misc.as<double>());
  ++S->launches;
  QCUDA(S, cudaMemcpyAsync(cost, S->misc.ptr, sizeof(double) * size_t(B), cudaMemcpyDeviceToHost, st_));
  QCUDA(S, cudaStreamSynchronize(st_));
  QCUDA(S, cudaGetLastError());
  return QILQR_OK;
}

int qilqr_backwards_pass_host(qilqr_solver_t *S, int B, int N, const double *desired, int Bd, const double *traj,
                              double *k, double *K, double *terms) {
  if (!S || !desired || !traj || !k || !K || !terms || B <= 0 || N <= 0 || (Bd != 1 && Bd != B))
    return QILQR_ERR_INVALID_ARGUMENT;
  QENTER(S);
  QCUDA(S, cudaSetDevice(S->device));
  cudaStream_t st_ = S->stream;
  QCUDA(S, S->traj_soa.ensure(sizeof(double) * size_t(N) * 17 * B));
  QCUDA(S, S->desired_soa.ensure(sizeof(double) * size_t(N) * 17 * Bd));
  QCUDA(S, S->gk.ensure(sizeof(double) * size_t(N) * 4 * B));
  QCUDA(S, S->gK.ensure(sizeof(double) * size_t(N) * 48 * B));
  QCUDA(S, S->stage_c.ensure(sizeof(double) * size_t(N) * 48 * B));
  QCUDA(S, S->misc.ensure(sizeof(double) * 2 * size_t(B)));
  int rc = upload_traj(S, S->stage_b, Bd, N, desired, S->desired_soa.as<double>());
  if (rc) return rc;
  rc = upload_traj(S, S->stage_a, B, N, traj, S->traj_soa.as<double>());
  if (rc) return rc;
  Problem pr{B, N, Bd, nullptr, nullptr, S->desired_soa.as<double>(), S->gk.as<double>(), S->gK.as<double>()};
  BackwardArgs ba{pr, SolveState{}, nullptr, B, 0, PHASE_SEARCH, 0, S->traj_soa.as<double>(), S->misc.as<double>()};
  if ((rc = launch_backward(S, ba))) return rc;
  ++S->launches;
  transpose_to_aos(S, S->gk.as<double>(), S->stage_c.as<double>(), B, N, 4);
  QCUDA(S, cudaMemcpyAsync(k, S->stage_c.ptr, sizeof(double) * size_t(N) * 4 * B, cudaMemcpyDeviceToHost, st_));
  QCUDA(S, cudaStreamSynchronize(st_));
  transpose_to_aos(S, S->gK.as<double>(), S->stage_c.as<double>(), B, N, 48);
  QCUDA(S, cudaMemcpyAsync(K, S->stage_c.ptr, sizeof(double) * size_t(N) * 48 * B, cudaMemcpyDeviceToHost, st_));
  QCUDA(S, cudaMemcpyAsync(terms, S->misc.ptr, sizeof(double) * 2 * size_t(B), cudaMemcpyDeviceToHost, st_));
  QCUDA(S, cudaStreamSynchronize(st_));
  QCUDA(S, cudaGetLastError());
  return QILQR_OK;
}

int qilqr_line_search_host(qilqr_solver_t *S, int B, int N, const double *desired, int Bd, const double *current,
                           const double *current_cost, const double *k, const double *K, const double *terms,
                           double *out_traj, double *new_cost, double *step, int32_t *status) {
  if (!S || !desired || !current || !current_cost || !k || !K || !terms || !out_traj || !new_cost || !step ||
      !status || B <= 0 || N <= 0 || (Bd != 1 && Bd != B))
    return QILQR_ERR_INVALID_ARGUMENT;
  QENTER(S);
  if (S->p.ls_max_iters <= 0) {
    // line_search's loop (ilqr.hh:178-193) evaluates no candidate at all and throws
    std::memcpy(out_traj, current, sizeof(double) * size_t(B) * N * 18);
    for (int b = 0; b < B; ++b) { status[b] = QILQR_ERR_LINE_SEARCH; step[b] = 1.0; new_cost[b] = current_cost[b]; }
    return QILQR_OK;
  }
  QCUDA(S, cudaSetDevice(S->device));
  cudaStream_t st_ = S->stream;
  int rc = ensure_state(S, B);
  if (rc) return rc;
  QCUDA(S, S->traj_soa.ensure(sizeof(double) * size_t(N) * 17 * B));
  QCUDA(S, S->buf1.ensure(sizeof(double) * size_t(N) * 17 * B));
  QCUDA(S, S->desired_soa.ensure(sizeof(double) * size_t(N) * 17 * Bd));
  QCUDA(S, S->gk.ensure(sizeof(double) * size_t(N) * 4 * B));
  QCUDA(S, S->gK.ensure(sizeof(double) * size_t(N) * 48 * B));
  QCUDA(S, S->stage_c.ensure(sizeof(double) * size_t(N) * 48 * B));
  rc = upload_traj(S, S->stage_b, Bd, N, desired, S->desired_soa.as<double>());
  if (rc) return rc;
  rc = upload_traj(S, S->stage_a, B, N, current, S->traj_soa.as<double>());
  if (rc) return rc;
  QCUDA(S, cudaMemcpyAsync(S->stage_c.ptr, k, sizeof(double) * size_t(N) * 4 * B, cudaMemcpyHostToDevice, st_));
  transpose_to_soa(S, S->stage_c.as<double>(), S->gk.as<double>(), B, N, 4);
  QCUDA(S, cudaStreamSynchronize(st_));
  QCUDA(S, cudaMemcpyAsync(S->stage_c.ptr, K, sizeof(double) * size_t(N) * 48 * B, cudaMemcpyHostToDevice, st_));
  transpose_to_soa(S, S->stage_c.as<double>(), S->gK.as<double>(), B, N, 48);
  Problem pr{B, N, Bd, S->traj_soa.as<double>(), S->buf1.as<double>(), S->desired_soa.as<double>(),
             S->gk.as<double>(), S->gK.as<double>()};
  SolveState st = make_state(S, B, nullptr, 0);
  k_init_state<<<blocks_for(B, 256), 256, 0, st_>>>(st, B, 1.0);
  ++S->launches;
  std::vector<double> q(B), kq(B);
  for (int b = 0; b < B; ++b) { q[b] = terms[2 * b]; kq[b] = terms[2 * b + 1]; }
  QCUDA(S, cudaMemcpyAsync(st.cost, current_cost, sizeof(double) * B, cudaMemcpyHostToDevice, st_));
  QCUDA(S, cudaMemcpyAsync(st.qutk, q.data(), sizeof(double) * B, cudaMemcpyHostToDevice, st_));
  QCUDA(S, cudaMemcpyAsync(st.ktquuk, kq.data(), sizeof(double) * B, cudaMemcpyHostToDevice, st_));
  QCUDA(S, cudaStreamSynchronize(st_));
  int *listS[2] = {S->lists.as<int>(), S->lists.as<int>() + B};
  int *dummy = S->lists.as<int>() + 2 * size_t(B);
  const int *search = nullptr;
  int n_search = B, s = 0;
  while (n_search > 0) {
    RolloutArgs ra{pr, st, search, n_search, 1, MODE_LINE_SEARCH, nullptr, nullptr, nullptr, nullptr, 1};
    launch_rollout(S, ra, n_search, st_);
    // phase: rejected problems keep PHASE_ACTIVE(0)... mark searching ones explicitly below
    k_compact<<<1, 1024, 0, st_>>>(search, n_search, st.phase, dummy, listS[s], S->d_counts);
    S->launches += 2;
    QCUDA(S, cudaStreamSynchronize(st_));
    n_search = S->h_counts[1];  // still PHASE_ACTIVE = not yet accepted and not failed
    search = listS[s];
    s = 1 - s;
  }
  dim3 grid(blocks_for(B, 128), 64);
  k_collect<<<grid, 128, 0, st_>>>(pr, st.sel);
  ++S->launches;
  unpack_traj(S, B, N, S->traj_soa.as<double>(), S->stage_a.as<double>());
  QCUDA(S, cudaMemcpyAsync(out_traj, S->stage_a.ptr, sizeof(double) * size_t(B) * N * 18, cudaMemcpyDeviceToHost, st_));
  std::vector<int> hstatus(B);
  std::vector<double> halpha(B);
  QCUDA(S, cudaMemcpyAsync(new_cost, st.cost, sizeof(double) * B, cudaMemcpyDeviceToHost, st_));
  QCUDA(S, cudaMemcpyAsync(halpha.data(), st.alpha, sizeof(double) * B, cudaMemcpyDeviceToHost, st_));
  QCUDA(S, cudaMemcpyAsync(hstatus.data(), st.status, sizeof(int) * B, cudaMemcpyDeviceToHost, st_));
  QCUDA(S, cudaStreamSynchronize(st_));
  for (int b = 0; b < B; ++b) {
    status[b] = (hstatus[b] >= QILQR_STATUS_LINE_SEARCH_FAILED) ? QILQR_ERR_LINE_SEARCH : QILQR_OK;
    step[b] = halpha[b];
  }
  QCUDA(S, cudaGetLastError());
  return QILQR_OK;
}

// ---- model / cost entry points (array-of-structs, one thread per problem) -----------
namespace {
struct Staged {
  std::vector<void *> ptrs;
  ~Staged() { for (void *p : ptrs) cudaFree(p); }
  double *in(qilqr_solver *S, const double *h, size_t n, int &rc) {
    if (!h) return nullptr;
    void *d = nullptr;
    if (cudaMalloc(&d, n * sizeof(double)) != cudaSuccess) { rc = QILQR_ERR_OUT_OF_MEMORY; return nullptr; }
    ptrs.push_back(d);
    if (cudaMemcpyAsync(d, h, n * sizeof(double), cudaMemcpyHostToDevice, S->stream) != cudaSuccess) rc = QILQR_ERR_CUDA;
    return static_cast<double *>(d);
  }
  double *out(const double *h, size_t n, int &rc) {
    if (!h) return nullptr;
    void *d = nullptr;
    if (cudaMalloc(&d, n * sizeof(double)) != cudaSuccess) { rc = QILQR_ERR_OUT_OF_MEMORY; return nullptr; }
    ptrs.push_back(d);
    return static_cast<double *>(d);
  }
};
int fetch(qilqr_solver *S, double *h, const double *d, size_t n) {
  if (!h) return QILQR_OK;
  QCUDA(S, cudaMemcpyAsync(h, d, n * sizeof(double), cudaMemcpyDeviceToHost, S->stream));
  return QILQR_OK;
}
}  // namespace

int qilqr_discrete_dynamics_host(qilqr_solver_t *S, int B, const double *x, const double *u, double *x_next,
                                 double *J_x, double *J_u) {
  if (!S || !x || !u || !x_next || B <= 0) return QILQR_ERR_INVALID_ARGUMENT;
  QENTER(S);
  QCUDA(S, cudaSetDevice(S->device));
  int rc = 0;
  Staged sg;
  double *dx = sg.in(S, x, size_t(B) * 13, rc), *du = sg.in(S, u, size_t(B) * 4, rc);
  double *dn = sg.out(x_next, size_t(B) * 13, rc), *dJx = sg.out(J_x, size_t(B) * 144, rc), *dJu = sg.out(J_u, size_t(B) * 48, rc);
  if (rc) return rc;
  k_api_discrete_dynamics<<<blocks_for(B, 64), 64, 0, S->stream>>>(S->p, B, dx, du, dn, dJx, dJu);
  ++S->launches;
  if ((rc = fetch(S, x_next, dn, size_t(B) * 13)) || (rc = fetch(S, J_x, dJx, size_t(B) * 144)) ||
      (rc = fetch(S, J_u, dJu, size_t(B) * 48)))
    return rc;
  QCUDA(S, cudaStreamSynchronize(S->stream));
  QCUDA(S, cudaGetLastError());
  return QILQR_OK;
}

int qilqr_continuous_dynamics_host(qilqr_solver_t *S, int B, const double *x, const double *u, double *xdot,
                                   double *J_x, double *J_u) {
  if (!S || !x || !u || !xdot || B <= 0) return QILQR_ERR_INVALID_ARGUMENT;
  QENTER(S);
  QCUDA(S, cudaSetDevice(S->device));
  int rc = 0;
  Staged sg;
  double *dx = sg.in(S, x, size_t(B) * 13, rc), *du = sg.in(S, u, size_t(B) * 4, rc);
  double *dd = sg.out(xdot, size_t(B) * 12, rc), *dJx = sg.out(J_x, size_t(B) * 144, rc), *dJu = sg.out(J_u, size_t(B) * 48, rc);
  if (rc) return rc;
  k_api_continuous_dynamics<<<blocks_for(B, 64), 64, 0, S->stream>>>(S->p, B, dx, du, dd, dJx, dJu);
  ++S->launches;
  if ((rc = fetch(S, xdot, dd, size_t(B) * 12)) || (rc = fetch(S, J_x, dJx, size_t(B) * 144)) ||
      (rc = fetch(S, J_u, dJu, size_t(B) * 48)))
    return rc;
  QCUDA(S, cudaStreamSynchronize(S->stream));
  QCUDA(S, cudaGetLastError());
  return QILQR_OK;
}

int qilqr_state_minus_host(qilqr_solver_t *S, int B, const double *lhs, const double *rhs, double *out, double *J_lhs,
                           double *J_rhs) {
  if (!S || !lhs || !rhs || !out || B <= 0) return QILQR_ERR_INVALID_ARGUMENT;
  QENTER(S);
  QCUDA(S, cudaSetDevice(S->device));
  int rc = 0;
  Staged sg;
  double *dl = sg.in(S, lhs, size_t(B) * 13, rc), *dr = sg.in(S, rhs, size_t(B) * 13, rc);
  double *dout = sg.out(out, size_t(B) * 12, rc), *dJl = sg.out(J_lhs, size_t(B) * 144, rc), *dJr = sg.out(J_rhs, size_t(B) * 144, rc);
  if (rc) return rc;
  k_api_state_minus<<<blocks_for(B, 64), 64, 0, S->stream>>>(B, dl, dr, dout, dJl, dJr);
  ++S->launches;
  if ((rc = fetch(S, out, dout, size_t(B) * 12)) || (rc = fetch(S, J_lhs, dJl, size_t(B) * 144)) ||
      (rc = fetch(S, J_rhs, dJr, size_t(B) * 144)))
    return rc;
  QCUDA(S, cudaStreamSynchronize(S->stream));
  QCUDA(S, cudaGetLastError());
  return QILQR_OK;
}

int qilqr_state_add_host(qilqr_solver_t *S, int B, const double *x, const double *tangent, double *out, double *J_lhs,
                         double *J_rhs) {
  if (!S || !x || !tangent || !out || B <= 0) return QILQR_ERR_INVALID_ARGUMENT;
  QENTER(S);
  QCUDA(S, cudaSetDevice(S->device));
  int rc = 0;
  Staged sg;
  double *dx = sg.in(S, x, size_t(B) * 13, rc), *dt = sg.in(S, tangent, size_t(B) * 12, rc);
  double *dout = sg.out(out, size_t(B) * 13, rc), *dJl = sg.out(J_lhs, size_t(B) * 144, rc), *dJr = sg.out(J_rhs, size_t(B) * 144, rc);
  if (rc) return rc;
  k_api_state_add<<<blocks_for(B, 64), 64, 0, S->stream>>>(B, dx, dt, dout, dJl, dJr);
  ++S->launches;
  if ((rc = fetch(S, out, dout, size_t(B) * 13)) || (rc = fetch(S, J_lhs, dJl, size_t(B) * 144)) ||
      (rc = fetch(S, J_rhs, dJr, size_t(B) * 144)))
    return rc;
  QCUDA(S, cudaStreamSynchronize(S->stream));
  QCUDA(S, cudaGetLastError());
  return QILQR_OK;
}

int qilqr_cost_host(qilqr_solver_t *S, int B, const double *x, const double *u, const double *x_d, const double *u_d,
                    double *cost, double *C_x, double *C_u, double *C_xx, double *C_uu, double *C_xu) {
  if (!S || !x || !u || !x_d || !u_d || !cost || B <= 0) return QILQR_ERR_INVALID_ARGUMENT;
  QENTER(S);
  QCUDA(S, cudaSetDevice(S->device));
  int rc = 0;
  Staged sg;
  double *dx = sg.in(S, x, size_t(B) * 13, rc), *du = sg.in(S, u, size_t(B) * 4, rc);
  double *dxd = sg.in(S, x_d, size_t(B) * 13, rc), *dud = sg.in(S, u_d, size_t(B) * 4, rc);
  double *dc = sg.out(cost, size_t(B), rc), *dCx = sg.out(C_x, size_t(B) * 12, rc), *dCu = sg.out(C_u, size_t(B) * 4, rc);
  double *dCxx = sg.out(C_xx, size_t(B) * 144, rc), *dCuu = sg.out(C_uu, size_t(B) * 16, rc), *dCxu = sg.out(C_xu, size_t(B) * 48, rc);
  if (rc) return rc;
  k_api_cost<<<blocks_for(B, 64), 64, 0, S->stream>>>(S->p, B, dx, du, dxd, dud, dc, dCx, dCu, dCxx, dCuu, dCxu);
  ++S->launches;
  if ((rc = fetch(S, cost, dc, size_t(B))) || (rc = fetch(S, C_x, dCx, size_t(B) * 12)) ||
      (rc = fetch(S, C_u, dCu, size_t(B) * 4)) || (rc = fetch(S, C_xx, dCxx, size_t(B) * 144)) ||
      (rc = fetch(S, C_uu, dCuu, size_t(B) * 16)) || (rc = fetch(S, C_xu, dCxu, size_t(B) * 48)))
    return rc;
  QCUDA(S, cudaStreamSynchronize(S->stream));
  QCUDA(S, cudaGetLastError());
  return QILQR_OK;
}

namespace {
__global__ void __launch_bounds__(256) k_dfma_peak(double *out, int iters, double a, double b) {
  double x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = double(threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = fma(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  if (s == 123.456) out[0] = s;  // never true; keeps the chain alive
}
}  // namespace

int qilqr_check_model(const qilqr_model_t *model) {
  if (!model) return QILQR_ERR_INVALID_ARGUMENT;
  double L[9];
  return factor_inertia(model->inertia, L) ? QILQR_OK : QILQR_ERR_INERTIA_NOT_PD;
}

int qilqr_measure_fp64_peak(int device, double *tflops) {
  if (!tflops) return QILQR_ERR_INVALID_ARGUMENT;
  if (cudaSetDevice(device) != cudaSuccess) return QILQR_ERR_NO_DEVICE;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return QILQR_ERR_NO_DEVICE;
  double *d = nullptr;
  if (cudaMalloc(&d, 8) != cudaSuccess) return QILQR_ERR_OUT_OF_MEMORY;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    k_dfma_peak<<<blocks, threads>>>(d, iters, 0.999999, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 64.0 * double(iters) * double(blocks) * double(threads);
    if (rep > 0 && ms > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  *tflops = best;
  return cudaGetLastError() == cudaSuccess ? QILQR_OK : QILQR_ERR_CUDA;
}

}  // extern "C"
