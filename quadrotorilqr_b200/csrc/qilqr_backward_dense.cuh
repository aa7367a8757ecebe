// =============================================================================
// qilqr_backward_dense.cuh -- ILQR::backwards_pass (ilqr.hh:97-147) for ANY model behind the
// ModelT concept (ilqr.hh:25-44): the Riccati sweep consumes dense J_x (12x12), J_u (12x4)
// and dense cost differentials, exactly the quantities the reference's template sees, and
// nothing of the quadrotor's block structure.
//
//   k_linearise_dense  one thread per (problem, knot): discrete_dynamics with differentials of
//                      the configured model variant (qilqr_model_generic.cuh) + CostFunction
//                      differentials (cost.hh:47-57), written as 316-double records in tiles
//                      rec[tile of 8 problems][knot][element][8]
//   k_riccati_dense    one warp per tile, 4 lanes per problem; lane c owns the 12x3 column
//                      block c of V_xx, Q_xx and of the gains in registers.  Record tiles
//                      (20 kB) arrive by TMA bulk copies, double-buffered; the products
//                      whose operands live on other lanes go through a 193-double exchange
//                      area per problem in shared memory (52.8 kB per warp: 4 warps per SM):
//        step 1  lane c:  M[:,c] = J_x^T V[:,c],  (J_u^T V)[:,c],  Q_x[c],  Q_u   -> smem
//        step 2  lane c:  Q_xx[:,c] = C_xx[:,c] + M J_x[:,c];  Q_xu[c,:] = M[c,:] J_u;
//                         Q_uu = C_uu + (J_u^T V) J_u   (replicated)
//        step 3  LDLT (replicated), k, K[:,c], (K^T Q_uu)[c,:], V_x[c]                -> smem
//        step 4  lane c:  V[:,c] = Q_xx[:,c] - (K^T Q_uu) K[:,c]   (+ optional symmetrisation)
//      Association and summation order follow the reference ((J_x^T V_xx) J_x, k ascending).
//
// The quadrotor-specific kernels (qilqr_backward_split.cuh) execute ~2.4x fewer FLOPs by
// skipping structural zeros; this path is what a user-supplied model costs.
// =============================================================================
#pragma once
#include "qilqr_backward_split.cuh"
#include "qilqr_model_generic.cuh"

namespace qilqr {
namespace dn {
// record: J_x (144), J_u (48), C.x (12), C.u (4), and the pose/pose, pose/velocity, velocity/pose 6x6 blocks
// of C.xx (the velocity/velocity block is the constant 2 Q_vv and stays out of the record)
constexpr int D_A = 0, D_B = 144, D_CX = 192, D_CU = 204, D_CPP = 208, D_CPV = 244, D_CVP = 280, DREC = 316;
constexpr int TILE = DREC * 8;  // doubles per (tile of 8 problems, knot): 20224 B
// exchange area per problem: M (144) and J_u^T V (48); (K^T Q_uu) and v_x reuse the start of M. Stride odd.
constexpr int E_M = 0, E_BTV = 144, E_KTQ = 0, E_VX = 48, DXS = 193;
__host__ __device__ constexpr int smem_doubles() { return 2 * TILE + 8 * DXS + 36 + 2; }
struct DenseLayout {  // cost differentials inside a dense record (see g4::G4Layout)
  __host__ __device__ static constexpr int cx(int j) { return D_CX + j; }
  __host__ __device__ static constexpr int cu(int j) { return D_CU + j; }
  __host__ __device__ static constexpr int cpp(int i, int j) { return D_CPP + 6 * i + j; }
  __host__ __device__ static constexpr int cpv(int i, int j) { return D_CPV + 6 * i + j; }
  __host__ __device__ static constexpr int cvp(int i, int j) { return D_CVP + 6 * i + j; }
};
}  // namespace dn

__global__ void __launch_bounds__(128) k_linearise_dense(const __grid_constant__ DeviceParams p,
                                                         const __grid_constant__ BackwardArgs a, double *rec_g) {
  using namespace dn;
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
  {
    double xn[13];
    gm::discrete_with_jacobians(p, x, u, xn, dst + D_A * 8, dst + D_B * 8, 8);
  }
  g4::cost_to_record<8, true, DenseLayout>(p, x, u, xd, ud, dst);
}

__global__ void __launch_bounds__(32) k_riccati_dense(const __grid_constant__ DeviceParams p,
                                                      const __grid_constant__ BackwardArgs a, const double *rec_g) {
  using namespace dn;
  using namespace g4;
  extern __shared__ __align__(128) double smem[];
  const int lane = threadIdx.x, c = lane & 3, q = lane >> 2;
  const int tile = blockIdx.x;
  const int t = tile * 8 + q;
  const bool valid = t < a.n;
  const int tt = valid ? t : a.n - 1;
  const int b = a.list ? a.list[tt] : tt;
  const int B = a.pr.B, N = a.pr.N;
  double *bufs = smem;
  double *xch = smem + 2 * TILE + q * DXS;
  double *s2Qvv = smem + 2 * TILE + 8 * DXS;  // C.xx velocity block = 2 Q_vv (cost.hh:52 with J = blkdiag(.., I))
  uint64_t *mbar = reinterpret_cast<uint64_t *>(s2Qvv + 36);
  for (int e = lane; e < 36; e += 32) s2Qvv[e] = 2.0 * p.Q[12 * (6 + e / 6) + 6 + e % 6];
  if (lane == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const double *src = rec_g + size_t(tile) * N * TILE;
  constexpr uint32_t kBytes = TILE * sizeof(double);
  if (lane == 0) {
    mbar_expect_tx(&mbar[0], kBytes);
    bulk_copy_g2s(bufs, src + size_t(N - 1) * TILE, kBytes, &mbar[0]);
  }
  uint32_t phase0 = 0, phase1 = 0;

  double Vc[36], vx[12];  // Vc[3 r + j] = V_xx[r][3 c + j]
#pragma unroll
  for (int e = 0; e < 36; ++e) Vc[e] = 0.0;
#pragma unroll
  for (int e = 0; e < 12; ++e) vx[e] = 0.0;
  double QuTk = 0.0, kTQuuk = 0.0;

#pragma unroll 1
  for (int i = N - 1; i >= 0; --i) {
    const int s = (N - 1 - i) & 1;
    if (i > 0 && lane == 0) {
      // every lane left the previous knot's last read of the other buffer through a __syncwarp
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(&mbar[s ^ 1], kBytes);
      bulk_copy_g2s(bufs + (s ^ 1) * TILE, src + size_t(i - 1) * TILE, kBytes, &mbar[s ^ 1]);
    }
    if (s == 0) { mbar_wait(&mbar[0], phase0); phase0 ^= 1; }
    else        { mbar_wait(&mbar[1], phase1); phase1 ^= 1; }
    const double *rec = bufs + s * TILE + q;  // element e of this problem's record at rec[8 e]
#define RA(k, r) rec[(D_A + 12 * (k) + (r)) * 8]
#define RB(k, j) rec[(D_B + 4 * (k) + (j)) * 8]

    // ---- step 1: M[:,c] = J_x^T V[:,c]; (J_u^T V)[:,c]; Q_x rows of c; Q_u ----
    double Mc[36];
#pragma unroll
    for (int r = 0; r < 12; ++r) {
      double m0, m1, m2;
      {
        const double av = RA(0, r);
        m0 = av * Vc[0]; m1 = av * Vc[1]; m2 = av * Vc[2];
      }
#pragma unroll
      for (int k = 1; k < 12; ++k) {
        const double av = RA(k, r);
        m0 = QFMA(av, Vc[3 * k], m0); m1 = QFMA(av, Vc[3 * k + 1], m1); m2 = QFMA(av, Vc[3 * k + 2], m2);
      }
      Mc[3 * r] = m0; Mc[3 * r + 1] = m1; Mc[3 * r + 2] = m2;
    }
    double BtVc[12], Qu[4];  // BtVc[3 j + jj] = (J_u^T V)[j][3 c + jj]
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const double bv = RB(0, j);
      BtVc[3 * j] = bv * Vc[0]; BtVc[3 * j + 1] = bv * Vc[1]; BtVc[3 * j + 2] = bv * Vc[2];
      Qu[j] = bv * vx[0];
    }
#pragma unroll
    for (int k = 1; k < 12; ++k)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const double bv = RB(k, j);
        BtVc[3 * j] = QFMA(bv, Vc[3 * k], BtVc[3 * j]);
        BtVc[3 * j + 1] = QFMA(bv, Vc[3 * k + 1], BtVc[3 * j + 1]);
        BtVc[3 * j + 2] = QFMA(bv, Vc[3 * k + 2], BtVc[3 * j + 2]);
        Qu[j] = QFMA(bv, vx[k], Qu[j]);
      }
#pragma unroll
    for (int j = 0; j < 4; ++j) Qu[j] = rec[(D_CU + j) * 8] + Qu[j];  // Q.u = C.u + J_u^T v_x
    const double *recAc = rec + (D_A + 3 * c) * 8;  // column block c of J_x: J_x[k][3c+j] at recAc[(12 k + j) 8]
    double Ac[36], Qxc[3];
#pragma unroll
    for (int k = 0; k < 12; ++k)
#pragma unroll
      for (int j = 0; j < 3; ++j) Ac[3 * k + j] = recAc[(12 * k + j) * 8];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      double sx = Ac[j] * vx[0];
#pragma unroll
      for (int k = 1; k < 12; ++k) sx = QFMA(Ac[3 * k + j], vx[k], sx);
      Qxc[j] = rec[(D_CX + 3 * c + j) * 8] + sx;  // Q.x = C.x + J_x^T v_x
    }
    __syncwarp();  // the exchange area is free: every lane has finished step 4 of the previous knot
#pragma unroll
    for (int r = 0; r < 12; ++r)
#pragma unroll
      for (int j = 0; j < 3; ++j) xch[E_M + 12 * r + 3 * c + j] = Mc[3 * r + j];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int jj = 0; jj < 3; ++jj) xch[E_BTV + 12 * j + 3 * c + jj] = BtVc[3 * j + jj];
    __syncwarp();

    // ---- step 2: Q_xx[:,c], Q_xu rows of c, Q_uu ----
    const double *ctop = rec + ((c < 2) ? D_CPP + 3 * c : D_CPV + 3 * (c - 2)) * 8;
    const double *cbot = (c < 2) ? rec + (D_CVP + 3 * c) * 8 : s2Qvv + 3 * (c - 2);
    const int cbot_rs = (c < 2) ? 48 : 6, cbot_cs = (c < 2) ? 8 : 1;
    double Qxxc[36];
#pragma unroll
    for (int r = 0; r < 12; ++r) {
      double s0, s1, s2;
      {
        const double mv = xch[E_M + 12 * r];
        s0 = mv * Ac[0]; s1 = mv * Ac[1]; s2 = mv * Ac[2];
      }
#pragma unroll
      for (int k = 1; k < 12; ++k) {
        const double mv = xch[E_M + 12 * r + k];
        s0 = QFMA(mv, Ac[3 * k], s0); s1 = QFMA(mv, Ac[3 * k + 1], s1); s2 = QFMA(mv, Ac[3 * k + 2], s2);
      }
      // C.xx[r][3c + j]: rows 0..5 from the pp | pv blocks of the record, rows 6..11 from the vp block | 2 Q_vv table
      const double *cxx = (r < 6) ? ctop + 48 * r : cbot + cbot_rs * (r - 6);
      const int cs = (r < 6) ? 8 : cbot_cs;
      Qxxc[3 * r] = cxx[0] + s0; Qxxc[3 * r + 1] = cxx[cs] + s1; Qxxc[3 * r + 2] = cxx[2 * cs] + s2;
    }
    double Qxuc[12], Quu[16];  // Qxuc[4 i + j] = Q_xu[3 c + i][j]
    {
      const double *Mrow = xch + E_M + 36 * c;  // rows 3c..3c+2 of M
#pragma unroll
      for (int k = 0; k < 12; ++k) {
        double bv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) bv[j] = RB(k, j);
#pragma unroll
        for (int i2 = 0; i2 < 3; ++i2) {
          const double mv = Mrow[12 * i2 + k];
#pragma unroll
          for (int j = 0; j < 4; ++j) Qxuc[4 * i2 + j] = (k == 0) ? mv * bv[j] : QFMA(mv, bv[j], Qxuc[4 * i2 + j]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const double tv = xch[E_BTV + 12 * j + k];
#pragma unroll
          for (int l = 0; l < 4; ++l) Quu[4 * j + l] = (k == 0) ? tv * bv[l] : QFMA(tv, bv[l], Quu[4 * j + l]);
        }
      }
#pragma unroll
      for (int e = 0; e < 16; ++e) Quu[e] = QFMA(2.0, p.R[e], Quu[e]);  // C.uu = 2 R (cost.hh:54)
      if (p.quu_reg != 0.0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) Quu[5 * j] += p.quu_reg;
      }
    }

    // ---- step 3: gains ----
    Ldlt4 f;
#pragma unroll
    for (int e = 0; e < 16; ++e) f.m[e] = Quu[e];
    ldlt4_compute(f);
    double kk[4] = {Qu[0], Qu[1], Qu[2], Qu[3]};
    ldlt4_solve(f, kk);
#pragma unroll
    for (int j = 0; j < 4; ++j) kk[j] = -kk[j];
    double Kc[12];  // Kc[3 j + i] = K[j][3 c + i]
#pragma unroll
    for (int i2 = 0; i2 < 3; ++i2) {
      double rhs[4] = {Qxuc[4 * i2], Qxuc[4 * i2 + 1], Qxuc[4 * i2 + 2], Qxuc[4 * i2 + 3]};
      ldlt4_solve(f, rhs);
#pragma unroll
      for (int j = 0; j < 4; ++j) Kc[3 * j + i2] = -rhs[j];
    }
    if (valid) {
      a.pr.gk[row_index(i, c, 4, B, b)] = kk[c == 0 ? 0 : c == 1 ? 1 : c == 2 ? 2 : 3];
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i2 = 0; i2 < 3; ++i2) a.pr.gK[row_index(i, 12 * j + 3 * c + i2, 48, B, b)] = Kc[3 * j + i2];
    }
    __syncwarp();  // every lane has finished reading M (step 2): its first 60 doubles are reused for K^T Q_uu and v_x
    double KtQc[12];  // (K^T Q_uu)[3 c + i][l]
#pragma unroll
    for (int i2 = 0; i2 < 3; ++i2)
#pragma unroll
      for (int l = 0; l < 4; ++l) {
        double acc = Kc[i2] * Quu[l];
#pragma unroll
        for (int j = 1; j < 4; ++j) acc = QFMA(Kc[3 * j + i2], Quu[4 * j + l], acc);
        KtQc[4 * i2 + l] = acc;
      }
#pragma unroll
    for (int i2 = 0; i2 < 3; ++i2) {
      double acc = KtQc[4 * i2] * kk[0];
#pragma unroll
      for (int l = 1; l < 4; ++l) acc = QFMA(KtQc[4 * i2 + l], kk[l], acc);
      xch[E_VX + 3 * c + i2] = Qxc[i2] - acc;  // v_x = Q.x - (K^T Q.uu) k
#pragma unroll
      for (int l = 0; l < 4; ++l) xch[E_KTQ + 4 * (3 * c + i2) + l] = KtQc[4 * i2 + l];
    }
    {  // expected cost reduction terms (ilqr.hh:136-140)
      double acc = Qu[0] * kk[0];
#pragma unroll
      for (int j = 1; j < 4; ++j) acc = QFMA(Qu[j], kk[j], acc);
      QuTk = QuTk + acc;
      double z[4];
#pragma unroll
      for (int l = 0; l < 4; ++l) {
        double sz = kk[0] * Quu[l];
#pragma unroll
        for (int j = 1; j < 4; ++j) sz = QFMA(kk[j], Quu[4 * j + l], sz);
        z[l] = sz;
      }
      double acc2 = z[0] * kk[0];
#pragma unroll
      for (int l = 1; l < 4; ++l) acc2 = QFMA(z[l], kk[l], acc2);
      kTQuuk = kTQuuk + acc2;
    }
    __syncwarp();

    // ---- step 4: V[:,c] = Q_xx[:,c] - (K^T Q_uu) K[:,c];  v_x ----
#pragma unroll
    for (int r = 0; r < 12; ++r) {
      const double t0 = xch[E_KTQ + 4 * r], t1 = xch[E_KTQ + 4 * r + 1], t2 = xch[E_KTQ + 4 * r + 2],
                   t3 = xch[E_KTQ + 4 * r + 3];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        double acc = t0 * Kc[j];
        acc = QFMA(t1, Kc[3 + j], acc);
        acc = QFMA(t2, Kc[6 + j], acc);
        acc = QFMA(t3, Kc[9 + j], acc);
        Vc[3 * r + j] = Qxxc[3 * r + j] - acc;
      }
    }
#pragma unroll
    for (int e = 0; e < 12; ++e) vx[e] = xch[E_VX + e];
    if (p.symmetrize_vxx) {  // V <- (V + V^T) / 2 through the M area
      __syncwarp();  // ... once every lane has read K^T Q_uu and v_x, which live at its start
#pragma unroll
      for (int r = 0; r < 12; ++r)
#pragma unroll
        for (int j = 0; j < 3; ++j) xch[E_M + 12 * r + 3 * c + j] = Vc[3 * r + j];
      __syncwarp();
#pragma unroll
      for (int r = 0; r < 12; ++r)
#pragma unroll
        for (int j = 0; j < 3; ++j) Vc[3 * r + j] = 0.5 * (Vc[3 * r + j] + xch[E_M + 12 * (3 * c + j) + r]);
    }
#undef RA
#undef RB
  }

  if (!valid || c != 0) return;
  backward_finish(p, a, b, QuTk, kTQuuk);
}

}  // namespace qilqr
