// =============================================================================
// qilqr_tail_persistent.cuh -- the latency-bound TAIL of a solve in ONE kernel launch.
//
// After the tail compaction (k_tail_gather) the m <= hi_threshold problems that are still iterating live
// in a dense mini-batch, and problems never interact.  Instead of ~85 more host iterations of five tiny
// launches each -- every one of which, on a GPU kept busy by other solver handles, waits for its turn
// behind their kernels -- one CTA takes 8 problems (one Riccati tile) and runs the rest of
// ILQR::solve (ilqr.hh:58-85) for them on its own, with block-level barriers only:
//
//   per outer iteration:   linearise the active problems' knots          (96 threads, records -> global tile)
//                          Riccati sweep                                  (warp 0: the quad kernel's step;
//                                                                          warps 1-2 stage the next record tile)
//                          rollout(s) while a problem is in line search   (the three role warps of k_rollout_ws)
//
// It is the same device code as the per-iteration kernels (linearise_to_record, riccati_step,
// rollout_ws_run, rollout_finish): identical decisions, values equal to the host-driven loop up to the
// compiler's FMA contraction, which differs between kernels (<= 1e-14 relative; bit-identical when built
// with -fmad=false) -- tests/test_gpu_parity.py::test_persistent_tail_matches_host_loop.
//
// EXPERIMENTAL, off by default (QILQR_PERSISTENT_TAIL=1).  Measured on B200 (profiles/r1_experiments_log.md):
// one batch at a time it is 3 % faster than the host loop; with 8-24 pipelined handles the resident tail CTAs
// (255 registers x 96 threads, 34 kB) keep the tail's wall time bounded (65 ms at P = 24 against 129 ms) but
// take occupancy from the other handles' bulk kernels, and the batch rate drops by 7 %.  Reference model,
// sequential line search; the host loop remains for model variants, parallel step sizes and ILQRDebug capture.
// =============================================================================
#pragma once
#include "qilqr_backward_split.cuh"

namespace qilqr {

namespace tp {
__host__ __device__ constexpr int smem_doubles(bool denseq) {
  return 2 * g4::tile_doubles(denseq) + 8 * g4::XS + 36 + (2 * 7 + 6 + 4) * 32;
}
}  // namespace tp

template <bool DENSEQ>
__global__ void __launch_bounds__(96, 1) k_tail_persistent(const __grid_constant__ DeviceParams p, const Problem pr,
                                                           const SolveState st, double *rec_g, const int m,
                                                           const int first_iter) {
  using namespace g4;
  constexpr int TILE = tile_doubles(DENSEQ);
  extern __shared__ __align__(128) double smem[];
  double *bufs = smem;
  double *xch_all = smem + 2 * TILE;
  double *s2Qvv = xch_all + 8 * XS;
  double(*s_pose)[7][32] = reinterpret_cast<double(*)[7][32]>(s2Qvv + 36);
  double(*s_vel)[32] = reinterpret_cast<double(*)[32]>(s2Qvv + 36 + 2 * 7 * 32);
  double(*s_u)[32] = reinterpret_cast<double(*)[32]>(s2Qvv + 36 + (2 * 7 + 6) * 32);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x, N = pr.N, B = pr.B, Bd = pr.Bd;
  double *rec_tile = rec_g + size_t(tile) * N * TILE;
  for (int e = tid; e < 36; e += 96) s2Qvv[e] = 2.0 * p.Q[12 * (6 + e / 6) + 6 + e % 6];

  // the problem a thread looks after when it tests phases (slot tid & 7); slots beyond m shadow the last problem
  const int my_slot = tid & 7;
  const int my_t = min(tile * 8 + my_slot, m - 1);
  const bool my_real = tile * 8 + my_slot < m;

  for (int iter = first_iter; iter < p.max_iters; ++iter) {  // ilqr.hh:58 (max_iters is a double)
    if (!__syncthreads_or(my_real && st.phase[my_t] == PHASE_ACTIVE)) break;

    // ---- linearisation of the active problems (every knot independent) ----
    for (int item = tid; item < 8 * N; item += 96) {
      const int q = item & 7, knot = item >> 3;
      const int t = min(tile * 8 + q, m - 1);  // slots beyond m replicate the last problem (as k_linearise pads)
      if (st.phase[t] != PHASE_ACTIVE) continue;  // finished problems keep the records of their last pass
      const double *traj = st.sel[t] ? pr.buf1 : pr.buf0;
      double x[13], u[4], xd[13], ud[4];
      load_point(traj, knot, B, t, x, u);
      load_point(pr.desired, knot, Bd, (Bd == 1) ? 0 : t, xd, ud);
      linearise_to_record<8, DENSEQ>(p, x, u, xd, ud, rec_tile + size_t(knot) * TILE + q);
    }
    __syncthreads();

    // ---- Riccati sweep: warp 0 computes, warps 1-2 stage the next knot's record tile ----
    {
      BackwardArgs ba{pr, st, nullptr, m, iter, PHASE_SEARCH, 1, nullptr, nullptr};
      const int c = lane & 3, q = lane >> 2;
      const int t = min(tile * 8 + q, m - 1);  // quads beyond m redo the last problem (identical values)
      const bool real = tile * 8 + q < m;
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
          riccati_step<8, DENSEQ>(p, ba, bufs + s * TILE + q, s2Qvv, xch_all + q * XS, c, real, i, B, t, V0, V1, V2, V3,
                                  vx, V88, QuTk, kTQuuk);
        } else if (i > 0) {
          for (int e = tid - 32; e < TILE; e += 64) bufs[(s ^ 1) * TILE + e] = rec_tile[size_t(i - 1) * TILE + e];
        }
        __syncthreads();
      }
      if (warp == 0 && c == 0 && real && st.phase[t] == PHASE_ACTIVE) {
        st.qutk[t] = QuTk;
        st.ktquuk[t] = kTQuuk;
        st.bwd[t] += 1;
        const double cost = st.cost[t];
        const double expected_new_cost = cost + (QuTk + kTQuuk / 2.0);  // ilqr.hh:64-65 with step = 1
        if (iter > 0 && is_converged(p, cost, expected_new_cost)) {
          st.status[t] = QILQR_STATUS_CONVERGED_EXPECTED;  // ilqr.hh:66-68
          st.phase[t] = PHASE_DONE;
        } else {
          st.alpha[t] = 1.0;
          st.ls_iter[t] = 0;
          st.phase[t] = PHASE_SEARCH;
        }
      }
    }
    __syncthreads();  // the new phases are visible to every thread

    // ---- line search: one more rollout for every problem still searching (ilqr.hh:70-77, 174-194) ----
    for (;;) {
      const bool searching = lane < 8 && tile * 8 + lane < m && st.phase[min(tile * 8 + lane, m - 1)] == PHASE_SEARCH;
      if (!__syncthreads_or(searching)) break;
      RolloutArgs ra{pr, st, nullptr, m, iter, MODE_SOLVE, nullptr, nullptr, nullptr, nullptr, 1};
      // slots that do not search shadow the tile's first problem without writing anything
      const int b = searching ? tile * 8 + lane : tile * 8;
      rollout_ws_run(p, ra, lane, warp, searching, b, s_pose, s_vel, s_u);
      __syncthreads();  // ... and so are the phases set by the line-search bookkeeping
    }
  }
}

}  // namespace qilqr
