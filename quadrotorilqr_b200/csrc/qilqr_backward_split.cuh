// =============================================================================
// qilqr_backward_split.cuh -- ILQR::backwards_pass (ilqr.hh:97-147) as two kernels:
//
//   k_linearise   one thread per (problem, knot): all knots of all problems are linearised in
//                 parallel (dynamics blocks, cost gradient, pose block of the Gauss-Newton Hessian)
//                 into 101-double records, stored as tiles  rec[tile of 8 problems][knot][elem][8]
//   k_riccati_g4  one warp per tile of 8 problems (4 lanes each): walks the horizon backwards; the
//                 next knot's 6.4 kB record tile is fetched by a TMA bulk copy (cp.async.bulk +
//                 mbarrier) into the one record buffer as soon as the current knot's Riccati step
//                 (qilqr_riccati_step.cuh) has read the record for the last time, and lands while
//                 the rest of the step runs on registers / shared memory.
//
// Compared with the fused kernel (qilqr_backward_g4.cuh) this removes the Lie-group code and the
// per-problem record storage from the sequential kernel (smaller code, less shared memory per warp
// -> more resident warps), makes the scalar linearisation fully parallel, and costs one round trip
// of the records through HBM (808 B per problem-knot).
// =============================================================================
#pragma once
#ifndef __CUDACC_RTC__
#include <cstdint>
#endif

#include "qilqr_backward_g4.cuh"

namespace qilqr {
namespace g4 {

// record elements per knot: 101 for Q = blkdiag(Q_pp, Q_vv) (100 values + 1 padding element, see R_CPP_GROUP), 173 for
// a Q with pose/velocity coupling
__host__ __device__ constexpr int rect(bool denseq) { return denseq ? 173 : 101; }
// doubles per (tile of 8 problems, knot): 6464 B / 11072 B, multiples of 16
__host__ __device__ constexpr int tile_doubles(bool denseq) { return rect(denseq) * 8; }
constexpr int XS = 228;               // exchange stride per problem, = 4 (mod 16)
__host__ __device__ constexpr int split_smem_doubles(bool denseq) { return tile_doubles(denseq) + 8 * XS + QVV_TILE + 2; }

QD uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
QD void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
QD void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
QD void bulk_copy_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
QD void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
}  // namespace g4

// One thread per (slot, knot); slots are padded to whole tiles of 8 (padding replicates the last
// problem so that every record is finite).  Consecutive threads = consecutive slots.
template <bool DENSEQ>
__global__ void __launch_bounds__(128) k_linearise(const __grid_constant__ DeviceParams p,
                                                   const __grid_constant__ BackwardArgs a, double *rec_g) {
  using namespace g4;
  constexpr int TILE = tile_doubles(DENSEQ);
  const int n8 = (a.n + 7) & ~7;
  const size_t id = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int N = a.pr.N;
  if (id >= size_t(n8) * N) return;
  const int t = int(id % n8), i = int(id / n8);
  const int tt = t < a.n ? t : a.n - 1;
  const int b = a.list ? a.list[tt] : tt;
  const int B = a.pr.B, Bd = a.pr.Bd;
  const int bd = (Bd == 1) ? 0 : b;
  const double *traj = a.solve_mode ? (a.st.sel[b] ? a.pr.buf1 : a.pr.buf0) : a.traj;
  double x[13], u[4], xd[13], ud[4];
  load_point(traj, i, B, b, x, u);
  load_point(a.pr.desired, i, Bd, bd, xd, ud);
  double *dst = rec_g + (size_t(t >> 3) * N + i) * TILE + (t & 7);
  linearise_to_record<8, DENSEQ>(p, x, u, xd, ud, dst);
}

namespace g4 {
// riccati_step's record_done hook of the kernel below: lane 0 starts the bulk copy of the next knot's record tile into
// the (single) record buffer.  fence.proxy.async orders the warp's generic-proxy reads of the buffer, which the
// __syncwarp before the hook has collected, before the async-proxy write.
struct NextRecord {
  static constexpr bool kActive = true;
  double *bufs;
  const double *next;
  uint64_t *mbar;
  int lane;
  uint32_t bytes;
  QD void operator()() const {
    if (next && lane == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(mbar, bytes);
      bulk_copy_g2s(bufs, next, bytes, mbar);
    }
  }
};
}  // namespace g4
// One warp per tile of 8 problems, 4 lanes per problem.  ONE record buffer: the TMA bulk copy of knot i-1 is issued
// as soon as the warp has made its last read of record i (after step 3 of the Riccati step) and lands during steps
// 4-6, which do not touch the record.  (A second buffer with the copy issued at the top of the knot measured 2 %
// slower, and 9 or 10 resident warps per SM at 224 / 200 registers -- which the smaller footprint would allow -- no
// faster: profiles/README.md.)
template <bool DENSEQ>
__global__ void __maxnreg__(255) k_riccati_g4(const __grid_constant__ DeviceParams p,
                                                                       const __grid_constant__ BackwardArgs a, const double *rec_g) {
  using namespace g4;
  constexpr int TILE = tile_doubles(DENSEQ);
  extern __shared__ __align__(128) double smem[];
  const int lane = threadIdx.x, c = lane & 3, q = lane >> 2;
  const int tile = blockIdx.x;
  const int t = tile * 8 + q;
  const bool valid = t < a.n;
  const int tt = valid ? t : a.n - 1;
  const int b = a.list ? a.list[tt] : tt;
  const int B = a.pr.B, N = a.pr.N;
  double *bufs = smem;
  double *xch = smem + TILE + q * XS;
  double *s2Qvv = smem + TILE + 8 * XS;
  uint64_t *mbar = reinterpret_cast<uint64_t *>(s2Qvv + QVV_TILE);
  init_qvv_tile(p, s2Qvv, lane, 32);
  if (lane == 0) {
    mbar_init(&mbar[0], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const double *src = rec_g + size_t(tile) * N * TILE;
  constexpr uint32_t kBytes = TILE * sizeof(double);
  if (lane == 0) {
    mbar_expect_tx(&mbar[0], kBytes);
    bulk_copy_g2s(bufs, src + size_t(N - 1) * TILE, kBytes, &mbar[0]);
  }
  uint32_t phase0 = 0;
  double V0[9], V1[9], V2[9], V3[9], vx[12], V88[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) V88[e] = 0.0;
#pragma unroll
  for (int e = 0; e < 9; ++e) V0[e] = V1[e] = V2[e] = V3[e] = 0.0;
#pragma unroll
  for (int e = 0; e < 12; ++e) vx[e] = 0.0;
  double QuTk = 0.0, kTQuuk = 0.0;
  double *gk_lane = a.pr.gk + size_t(c) * B + b, *gK_lane = a.pr.gK + size_t(3 * c) * B + b;
#pragma unroll 1
  for (int i = N - 1; i >= 0; --i) {
    mbar_wait(&mbar[0], phase0);
    phase0 ^= 1;
    g4::NextRecord hook{bufs, i > 0 ? src + size_t(i - 1) * TILE : nullptr, &mbar[0], lane, kBytes};
    riccati_step<8, DENSEQ, g4::NextRecord>(p, a, bufs + q, s2Qvv + ((q + 4) & 7), xch, c, gk_lane, gK_lane, i, B, V0, V1, V2, V3, vx, V88, QuTk, kTQuuk, hook);
  }
  if (!valid || c != 0) return;
  backward_finish(p, a, b, QuTk, kTQuuk);
}

}  // namespace qilqr
