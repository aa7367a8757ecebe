// =============================================================================
// qilqr_tail_persistent.cuh -- the latency-bound TAIL of a solve in ONE kernel launch.
//
// Once only a few problems of a batch are still iterating (the handful that creep to max_iters), every further
// iteration of the host-driven loop is a chain of tiny launches and a host round trip.  Problems never
// interact, so instead one CTA takes 8 of the remaining problems (one Riccati tile) and runs the rest of
// ILQR::solve (ilqr.hh:58-85) for them on its own, with block-level barriers only:
//
//   while a problem of the tile is alive:
//     rollout(s) for the problems in line search            (the three role warps of k_rollout_ws)
//     linearise the problems that need a backward pass      (96 threads, records -> global tile)
//     Riccati sweep                                         (warp 0: the quad kernel's step;
//                                                            warps 1-2 stage the next record tile)
//
// It is the same device code as the per-iteration kernels (linearise_to_record, riccati_step, rollout_ws_run,
// rollout_finish, backward_finish) and the library is compiled with -fmad=false, so its results are bit-identical
// to the host-driven loop (tests/test_gpu_parity.py::test_persistent_tail_matches_host_loop).  The host does
// not take part: it sleeps on the stream (or, with the begin / finish API, goes on to the next batch).
//
// Used when at most `persist_threshold` problems are alive (default 64: a few CTAs; launching it for thousands
// of problems would pin 255-register CTAs on every SM while most of their problems have long finished).
// Reference model, sequential line search; the host loop remains for model variants, parallel step sizes and
// ILQRDebug capture.
// =============================================================================
#pragma once
#include "qilqr_backward_split.cuh"

namespace qilqr {

namespace tp {
__host__ __device__ constexpr int smem_doubles(bool denseq) {
  return 2 * g4::tile_doubles(denseq) + 8 * g4::XS + g4::QVV_TILE + (2 * 7 + 6 + 4) * 32;
}
}  // namespace tp

template <bool DENSEQ>
__global__ void __launch_bounds__(96, 1) k_tail_persistent(const __grid_constant__ DeviceParams p, const Problem pr,
                                                           const SolveState st, double *rec_g, double *scratch_gains,
                                                           const int *list, const int n, const int epoch) {
  using namespace g4;
  constexpr int TILE = tile_doubles(DENSEQ);
  extern __shared__ __align__(128) double smem[];
  double *bufs = smem;
  double *xch_all = smem + 2 * TILE;
  double *s2Qvv = xch_all + 8 * XS;
  double(*s_pose)[7][32] = reinterpret_cast<double(*)[7][32]>(s2Qvv + QVV_TILE);
  double(*s_vel)[32] = reinterpret_cast<double(*)[32]>(s2Qvv + QVV_TILE + 2 * 7 * 32);
  double(*s_u)[32] = reinterpret_cast<double(*)[32]>(s2Qvv + QVV_TILE + (2 * 7 + 6) * 32);
  __shared__ int s_prob[8], s_act[8];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x, N = pr.N, B = pr.B, Bd = pr.Bd;
  double *rec_tile = rec_g + size_t(tile) * N * TILE;
  init_qvv_tile(p, s2Qvv, tid, 96);
  // the 8 problems of this tile; slots beyond n shadow the last problem and never write
  if (tid < 8) {
    const int idx = min(tile * 8 + tid, n - 1);
    s_prob[tid] = list ? list[idx] : idx;
  }
  __syncthreads();
  const int n_real = min(8, n - tile * 8);

  for (;;) {
    // ---- line search: one more rollout for every problem that is searching (ilqr.hh:70-77, 174-194) ----
    for (;;) {
      const bool searching = lane < n_real && st.phase[s_prob[lane & 7]] == PHASE_SEARCH;
      if (!__syncthreads_or(searching)) break;
      RolloutArgs ra{pr, st, nullptr, n, epoch, MODE_SOLVE, nullptr, nullptr, nullptr, nullptr, 1, 0};
      // slots that do not search shadow the tile's first problem without writing anything
      rollout_ws_run<false>(p, ra, lane, warp, searching, s_prob[searching ? lane : 0], s_pose, s_vel, s_u);
      __syncthreads();  // the phases set by the line-search bookkeeping are visible to every thread
    }
    // ---- which problems need a backward pass ----
    if (tid < 8) s_act[tid] = tid < n_real && st.phase[s_prob[tid]] == PHASE_ACTIVE;
    __syncthreads();
    int any = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) any |= s_act[q];
    if (!any) break;

    // ---- linearisation of their knots (every knot independent) ----
    for (int item = tid; item < 8 * N; item += 96) {
      const int q = item & 7, knot = item >> 3;
      if (!s_act[q]) continue;  // (the sweep below runs on stale records for those slots and writes nothing)
      const int t = s_prob[q];
      const double *traj = st.sel[t] ? pr.buf1 : pr.buf0;
      double x[13], u[4], xd[13], ud[4];
      load_point(traj, knot, B, t, x, u);
      load_point(pr.desired, knot, Bd, (Bd == 1) ? 0 : t, xd, ud);
      linearise_to_record<8, DENSEQ>(p, x, u, xd, ud, rec_tile + size_t(knot) * TILE + q);
    }
    __syncthreads();

    // ---- Riccati sweep: warp 0 computes, warps 1-2 stage the next knot's record tile ----
    {
      BackwardArgs ba{pr, st, nullptr, n, epoch, 0, 1, nullptr, nullptr};
      const int c = lane & 3, q = lane >> 2;
      const int t = s_prob[q];
      const bool act = s_act[q] != 0;
      double V0[9], V1[9], V2[9], V3[9], vx[12], V88[16];
      double QuTk = 0.0, kTQuuk = 0.0;
      if (warp == 0) {
#pragma unroll
        for (int e = 0; e < 16; ++e) V88[e] = 0.0;
#pragma unroll
        for (int e = 0; e < 9; ++e) V0[e] = V1[e] = V2[e] = V3[e] = 0.0;
#pragma unroll
        for (int e = 0; e < 12; ++e) vx[e] = 0.0;
      }
      for (int e = tid; e < TILE; e += 96) bufs[e] = rec_tile[size_t(N - 1) * TILE + e];
      __syncthreads();
#pragma unroll 1
      for (int i = N - 1; i >= 0; --i) {
        const int s = (N - 1 - i) & 1;
        if (warp == 0) {
          // (quads whose problem is not in its backward pass sweep stale records: their gains go to a scratch array
          //  laid out like the real one)
          double *gk0 = act ? pr.gk : scratch_gains, *gK0 = act ? pr.gK : scratch_gains + size_t(N) * 4 * B;
          riccati_step<8, DENSEQ>(p, ba, bufs + s * TILE + q, s2Qvv + ((q + 4) & 7), xch_all + q * XS, c, gk0 + size_t(c) * B + t,
                                  gK0 + size_t(3 * c) * B + t, i, B, V0, V1, V2, V3, vx, V88, QuTk, kTQuuk);
        } else if (i > 0) {
          for (int e = tid - 32; e < TILE; e += 64) bufs[(s ^ 1) * TILE + e] = rec_tile[size_t(i - 1) * TILE + e];
        }
        __syncthreads();
      }
      if (warp == 0 && c == 0 && act) backward_finish(p, ba, t, QuTk, kTQuuk);
    }
    __syncthreads();  // the new phases are visible to every thread
  }
}

}  // namespace qilqr
