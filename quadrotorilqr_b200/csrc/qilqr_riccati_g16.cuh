// =============================================================================
// qilqr_riccati_g16.cuh -- the Riccati sweep of ILQR::backwards_pass (ilqr.hh:118-140) with SIXTEEN LANES PER
// PROBLEM, for launches that are latency-bound: the handful of problems that creep to max_iters at the end of a
// batch, long horizons at moderate batch sizes (BASELINE config 4: N = 1000, batch 4096), receding-horizon
// re-solves (config 5).
//
// k_riccati_g4 (qilqr_riccati_step.cuh) gives every problem 4 lanes; one knot costs a lone warp ~6500 cycles
// (~1500 FP64 instructions per lane at 2 issue cycles each, plus the dependent chain of the 4x4 factorisation),
// i.e. 3.3 us per knot whatever the batch size once the GPU is not full.  Here lane (r, c) of a 16-lane group owns
// the single 3x3 block V_xx[r][c]; a CTA of 4 warps takes one record tile of 8 problems (2 problems per warp, one
// warp per SM sub-partition), so that the same knot costs ~1/3 of the instructions per lane and 4 schedulers work
// on a tile instead of one:
//
//   step 1  lane (I, c):  M[I][c] = sum_k A[k][I]^T V[k][c]        (<= 3 block products, operands chosen by I)
//           lane l:       Q_uu entry l = 2 R + (B^T V_xx B)        (one entry each instead of all 16)
//   step 2  all lanes:    Q_u, factorisation of Q_uu, k, Delta-J terms (replicated: a dependent chain)
//           lane (r, .):  Q_x block r
//   step 3  lane (r, J):  Q_xx[r][J] = C_xx[r][J] + sum_I M[r][I] A[I][J]
//   step 4  lane (r, c<3): column 3r+c of Q_xu, the gain column K[:, 3r+c], (K^T Q_uu) row, v_x entry
//   step 5  lane (r, J):  V'[r][J] = Q_xx[r][J] - (K^T Q_uu)[r] K[:, J]   -- already where the next knot needs it
//
// EVERY OUTPUT ELEMENT IS PRODUCED BY THE SAME SEQUENCE OF ROUNDED OPERATIONS AS IN riccati_step (the same helper
// functions in the same order; lanes that have no term at some position skip it instead of adding a zero), and the
// library is compiled with -fmad=false: the two kernels are bit-identical (tests/test_gpu_riccati_g16.py), which
// is what allows choosing between them by launch size without the results depending on the batch a problem is in.
// Q must have no pose/velocity coupling (as for the 100-double records); otherwise the solver keeps k_riccati_g4.
// =============================================================================
#pragma once
#include "qilqr_backward_split.cuh"

namespace qilqr {
namespace g16 {
using namespace g4;

// per-problem exchange area (doubles): 3x3 blocks at 9 * (4 r + c) (conflict-free for one lane per block)
// (v_x is double-buffered by knot parity: a lane writes the new entry while others may still read the old vector)
constexpr int E_V = 0, E_M = 144, E_K = 288, E_KQ = 336, E_VX = 384, E_QUU = 408, E_SIZE = 424;
constexpr int ES = E_SIZE;  // stride per problem
// + 2 Q_vv (36), B rows 8..11 (16) and R (16) for lane-indexed access, 2 mbarriers
__host__ __device__ constexpr int smem_doubles() { return 2 * tile_doubles(false) + 8 * ES + 36 + 32 + 2; }

QD int boff(int r, int c) { return 9 * (4 * r + c); }

template <int RS>
QD void ld9s_at(const double *rec, int elem, double *r) {
#pragma unroll
  for (int e = 0; e < 9; ++e) r[e] = rec[(elem + e) * RS];
}
}  // namespace g16

// One CTA (4 warps) per tile of 8 problems; warp w takes problems 2w and 2w+1, 16 lanes each.
__global__ void __launch_bounds__(128) k_riccati_g16(const __grid_constant__ DeviceParams p,
                                                     const __grid_constant__ BackwardArgs a, const double *rec_g) {
  using namespace g16;
  constexpr int TILE = tile_doubles(false);
  constexpr int RS = 8;
  extern __shared__ __align__(128) double smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int q = 2 * warp + (lane >> 4);  // problem slot within the tile
  const int l = lane & 15, r = l >> 2, c = l & 3;
  const int tile = blockIdx.x;
  const int t = tile * 8 + q;
  const bool valid = t < a.n;
  const int tt = valid ? t : a.n - 1;
  const int b = a.list ? a.list[tt] : tt;
  const int B = a.pr.B, N = a.pr.N;
  double *bufs = smem;
  double *ex = smem + 2 * TILE + q * ES;
  double *s2Qvv = smem + 2 * TILE + 8 * ES;
  double *sBu = s2Qvv + 36, *sR = sBu + 16;
  uint64_t *mbar = reinterpret_cast<uint64_t *>(sR + 16);
  for (int e = tid; e < 36; e += 128) s2Qvv[e] = 2.0 * p.Q[12 * (6 + e / 6) + 6 + e % 6];
  if (tid < 16) { sBu[tid] = p.Bu[tid]; sR[tid] = p.R[tid]; }
  // V_xx = 0, v_x = 0 entering the last knot (ilqr.hh:103-106)
  for (int e = l; e < 144; e += 16) ex[E_V + e] = 0.0;
  if (l < 12) ex[E_VX + l] = 0.0;
  if (tid == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const double *src = rec_g + size_t(tile) * N * TILE;
  constexpr uint32_t kBytes = TILE * sizeof(double);
  if (tid == 0) {
    mbar_expect_tx(&mbar[0], kBytes);
    bulk_copy_g2s(bufs, src + size_t(N - 1) * TILE, kBytes, &mbar[0]);
  }
  uint32_t phase0 = 0, phase1 = 0;

  double Vb[9];  // the lane's block V_xx[r][c]
#pragma unroll
  for (int e = 0; e < 9; ++e) Vb[e] = 0.0;
  double QuTk = 0.0, kTQuuk = 0.0;

  // record elements of the A-blocks this lane multiplies with, by role (see the step comments)
  const int a1_col = (r == 0) ? R_RE : (r == 1) ? R_TE : (r == 2) ? R_DJR : R_DQB;  // A[0][r] (step 1) = A[0][J] (step 3)
  const int a2_col = (r == 1) ? R_RE : R_DJR;                                       // A[1][r] for r odd
  const int a1_row = (c == 0) ? R_RE : (c == 1) ? R_TE : (c == 2) ? R_DJR : R_DQB;
  const int a2_row = (c == 1) ? R_RE : R_DJR;

#pragma unroll 1
  for (int i = N - 1; i >= 0; --i) {
    const int s = (N - 1 - i) & 1;
    if (i > 0 && tid == 0) {
      // the other buffer was last read during the previous knot, which every thread has left through the
      // __syncthreads at the end of the step: order those reads before the asynchronous write
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(&mbar[s ^ 1], kBytes);
      bulk_copy_g2s(bufs + (s ^ 1) * TILE, src + size_t(i - 1) * TILE, kBytes, &mbar[s ^ 1]);
    }
    if (s == 0) { mbar_wait(&mbar[0], phase0); phase0 ^= 1; }
    else        { mbar_wait(&mbar[1], phase1); phase1 ^= 1; }
    const double *rec = bufs + s * TILE + q;  // element e of this problem's record at rec[e * 8]
    const double dgz[3] = {rec[R_GZ * RS], rec[(R_GZ + 1) * RS], rec[(R_GZ + 2) * RS]};
    const double ndgz[3] = {-dgz[0], -dgz[1], -dgz[2]};

    // ---------------- step 1: M[r][c] = (A^T V)[r][c]; one entry of Q_uu ----------------
    {
      double A1[9], V0c[9], Mb[9];
      ld9s_at<RS>(rec, a1_col, A1);
      ld9(ex + E_V + boff(0, c), V0c);
      m3_mulT(A1, V0c, Mb);
      if (r & 1) {  // rows 1 and 3: + A[1][r]^T V[1][c]
        double A2[9], V1c[9];
        ld9s_at<RS>(rec, a2_col, A2);
        ld9(ex + E_V + boff(1, c), V1c);
        m3_maddT(A2, V1c, Mb);
      }
      if (r == 1) {  // + dG^T V[2][c]
        double V2c[9];
        ld9(ex + E_V + boff(2, c), V2c);
        m3_hat_madd(ndgz, V2c, Mb);
      } else if (r == 2) {  // + I V[2][c] (the lane's own block)
#pragma unroll
        for (int e = 0; e < 9; ++e) Mb[e] += Vb[e];
      } else if (r == 3) {  // + Wd^T V[3][c] (own block)
        double Wd[9];
        ld9s_at<RS>(rec, R_WD, Wd);
        m3_maddT(Wd, Vb, Mb);
      }
      st9(ex + E_M + boff(r, c), Mb);
      // Q_uu[r][c] = 2 R[r][c] + ((B^T V_xx) B)[r][c] with B rows 8..11 = Bu: entry (jj, l) = (r, c) of riccati_step
      double acc = 0.0;
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        double btv = 0.0;  // BtV[r][cc] = sum_rr Bu[rr][r] V_xx[8 + rr][8 + cc]
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
          // V_xx[8 + rr][8 + cc]: block (rr ? 3 : 2, cc ? 3 : 2), element ((rr + 2) % 3, (cc + 2) % 3)
          const double v = ex[E_V + boff(rr == 0 ? 2 : 3, cc == 0 ? 2 : 3) + 3 * ((rr + 2) % 3) + (cc + 2) % 3];
          btv = (rr == 0) ? sBu[r] * v : QFMA(sBu[4 * rr + r], v, btv);
        }
        acc = (cc == 0) ? btv * sBu[c] : QFMA(btv, sBu[4 * cc + c], acc);
      }
      double quu = QFMA(2.0, sR[4 * r + c], acc);
      if (p.quu_reg != 0.0 && r == c) quu += p.quu_reg;
      ex[E_QUU + 4 * r + c] = quu;
    }
    __syncwarp();

    // ---------------- step 2: Q_u, Q_x block r, factorisation, k (replicated) ----------------
    double Quu[16], Qu[4], k[4], Qxr[3];
    Ldlt4 f;
    {
      double vx[12];
#pragma unroll
      for (int e = 0; e < 12; ++e) vx[e] = ex[E_VX + 12 * s + e];
#pragma unroll
      for (int e = 0; e < 16; ++e) Quu[e] = ex[E_QUU + e];
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        double acc = p.Bu[jj] * vx[8];
#pragma unroll
        for (int rr = 1; rr < 4; ++rr) acc = QFMA(p.Bu[4 * rr + jj], vx[8 + rr], acc);
        Qu[jj] = rec[(R_CU + jj) * RS] + acc;
      }
      // Q.x block r = C.x[3r..] + (A^T v_x)[3r..]
      {
        double A1[9], T[3];
        ld9s_at<RS>(rec, a1_col, A1);
        m3T_vec(A1, vx, T);
        if (r & 1) {
          double A2[9];
          ld9s_at<RS>(rec, a2_col, A2);
          m3T_vec_add(A2, vx + 3, T);
        }
        if (r == 1) {  // += dG^T vx[6:9] = hat(-dgz) vx[6:9]
          const double *v = vx + 6;
          T[0] = QFMA(ndgz[1], v[2], QFMA(-ndgz[2], v[1], T[0]));
          T[1] = QFMA(-ndgz[0], v[2], QFMA(ndgz[2], v[0], T[1]));
          T[2] = QFMA(ndgz[0], v[1], QFMA(-ndgz[1], v[0], T[2]));
        } else if (r == 2) {
#pragma unroll
          for (int e = 0; e < 3; ++e) T[e] += vx[6 + e];
        } else if (r == 3) {
          double Wd[9];
          ld9s_at<RS>(rec, R_WD, Wd);
          m3T_vec_add(Wd, vx + 9, T);
        }
#pragma unroll
        for (int e = 0; e < 3; ++e) Qxr[e] = rec[(R_CX + 3 * r + e) * RS] + T[e];
      }
#pragma unroll
      for (int e = 0; e < 16; ++e) f.m[e] = Quu[e];
      ldlt4_compute(f);
      double rhs[4] = {Qu[0], Qu[1], Qu[2], Qu[3]};
      ldlt4_solve(f, rhs);
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) k[jj] = -rhs[jj];
    }

    // ---------------- step 3: Q_xx[r][J], J = c ----------------
    double Qb[9];
    {
      // C_xx[r][c] (cost.hh:52): pose block from the record, velocity block = 2 Q_vv, no coupling
      if (r < 2 && c < 2) {
#pragma unroll
        for (int ri = 0; ri < 3; ++ri)
#pragma unroll
          for (int cj = 0; cj < 3; ++cj) Qb[3 * ri + cj] = rec[(R_CPP + R_CPP_GROUP * r + 6 * ri + 3 * c + cj) * RS];
      } else if (r >= 2 && c >= 2) {
#pragma unroll
        for (int ri = 0; ri < 3; ++ri)
#pragma unroll
          for (int cj = 0; cj < 3; ++cj) Qb[3 * ri + cj] = s2Qvv[6 * (3 * (r - 2) + ri) + 3 * (c - 2) + cj];
      } else {
#pragma unroll
        for (int e = 0; e < 9; ++e) Qb[e] = 0.0;
      }
      double X0[9], A1[9], T[9];
      ld9(ex + E_M + boff(r, 0), X0);
      ld9s_at<RS>(rec, a1_row, A1);
      m3_mul(X0, A1, T);
      if (c & 1) {  // columns 1 and 3: + M[r][1] A[1][c]
        double X1[9], A2[9];
        ld9(ex + E_M + boff(r, 1), X1);
        ld9s_at<RS>(rec, a2_row, A2);
        m3_madd(X1, A2, T);
      }
      if (c == 1) {  // + M[r][2] dG
        double X2[9];
        ld9(ex + E_M + boff(r, 2), X2);
        m3_madd_hat(X2, dgz, T);
#pragma unroll
        for (int e = 0; e < 9; ++e) Qb[e] += T[e];
      } else if (c == 2) {  // + M[r][2] I
        double X2[9];
        ld9(ex + E_M + boff(r, 2), X2);
#pragma unroll
        for (int e = 0; e < 9; ++e) Qb[e] += T[e] + X2[e];
      } else if (c == 3) {  // + M[r][3] Wd
        double X3[9], Wd[9];
        ld9(ex + E_M + boff(r, 3), X3);
        ld9s_at<RS>(rec, R_WD, Wd);
        m3_madd(X3, Wd, T);
#pragma unroll
        for (int e = 0; e < 9; ++e) Qb[e] += T[e];
      } else {
#pragma unroll
        for (int e = 0; e < 9; ++e) Qb[e] += T[e];
      }
    }

    // ---------------- step 4: lanes (r, c < 3): gain column 3r + c ----------------
    if (c < 3) {
      // Q.xu[3r + c][:] = M[3r + c, 8] B[8, :] + M[3r + c, 9:12] B[9:12, :]
      const double m8 = ex[E_M + boff(r, 2) + 3 * c + 2];
      const double m9 = ex[E_M + boff(r, 3) + 3 * c], m10 = ex[E_M + boff(r, 3) + 3 * c + 1],
                   m11 = ex[E_M + boff(r, 3) + 3 * c + 2];
      double rhs[4];
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        double acc = m8 * p.Bu[jj];
        acc = QFMA(m9, p.Bu[4 + jj], acc);
        acc = QFMA(m10, p.Bu[8 + jj], acc);
        acc = QFMA(m11, p.Bu[12 + jj], acc);
        rhs[jj] = acc;
      }
      ldlt4_solve(f, rhs);
      double Ks[4], KQ[4];
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) Ks[jj] = -rhs[jj];
#pragma unroll
      for (int ll = 0; ll < 4; ++ll) {
        double acc = Ks[0] * Quu[ll];
#pragma unroll
        for (int jj = 1; jj < 4; ++jj) acc = QFMA(Ks[jj], Quu[4 * jj + ll], acc);
        KQ[ll] = acc;
      }
      double acc = KQ[0] * k[0];
#pragma unroll
      for (int ll = 1; ll < 4; ++ll) acc = QFMA(KQ[ll], k[ll], acc);
      const double qx = (c == 0) ? Qxr[0] : (c == 1) ? Qxr[1] : Qxr[2];
      const int sidx = 3 * r + c;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        ex[E_K + 12 * jj + sidx] = Ks[jj];
        ex[E_KQ + 4 * sidx + jj] = KQ[jj];
        if (valid) a.pr.gK[row_index(i, 12 * jj + sidx, 48, B, b)] = Ks[jj];
      }
      ex[E_VX + 12 * (s ^ 1) + sidx] = qx - acc;  // the next knot reads the other copy
    } else if (valid) {
      const double kc = (r == 0) ? k[0] : (r == 1) ? k[1] : (r == 2) ? k[2] : k[3];
      a.pr.gk[row_index(i, r, 4, B, b)] = kc;
    }
    // expected cost reduction terms (ilqr.hh:136-140), replicated
    {
      double acc = Qu[0] * k[0];
#pragma unroll
      for (int jj = 1; jj < 4; ++jj) acc = QFMA(Qu[jj], k[jj], acc);
      QuTk = QuTk + acc;
      double acc2 = 0.0;
#pragma unroll
      for (int ll = 0; ll < 4; ++ll) {
        double z = k[0] * Quu[ll];
#pragma unroll
        for (int jj = 1; jj < 4; ++jj) z = QFMA(k[jj], Quu[4 * jj + ll], z);
        acc2 = (ll == 0) ? z * k[0] : QFMA(z, k[ll], acc2);
      }
      kTQuuk = kTQuuk + acc2;
    }
    __syncwarp();

    // ---------------- step 5: V'[r][c] = Q_xx[r][c] - (K^T Q_uu)[r] K[:, c] ----------------
    {
      double KQr[12], Kc[12];  // KQr[4 s + l] = (K^T Q_uu)[3r + s][l];  Kc[3 l + cc] = K[l][3c + cc]
#pragma unroll
      for (int e = 0; e < 12; ++e) KQr[e] = ex[E_KQ + 12 * r + e];
#pragma unroll
      for (int ll = 0; ll < 4; ++ll)
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) Kc[3 * ll + cc] = ex[E_K + 12 * ll + 3 * c + cc];
#pragma unroll
      for (int ss = 0; ss < 3; ++ss)
#pragma unroll
        for (int cc = 0; cc < 3; ++cc) {
          double a0 = KQr[4 * ss] * Kc[cc];
#pragma unroll
          for (int ll = 1; ll < 4; ++ll) a0 = QFMA(KQr[4 * ss + ll], Kc[3 * ll + cc], a0);
          Vb[3 * ss + cc] = Qb[3 * ss + cc] - a0;
        }
      st9(ex + E_V + boff(r, c), Vb);
    }
    if (p.symmetrize_vxx) {  // V <- (V + V^T)/2: block (r, c) needs (block (c, r))^T
      __syncwarp();
      double Vt[9];
      ld9(ex + E_V + boff(c, r), Vt);
      __syncwarp();
#pragma unroll
      for (int ri = 0; ri < 3; ++ri)
#pragma unroll
        for (int cj = 0; cj < 3; ++cj) Vb[3 * ri + cj] = 0.5 * (Vb[3 * ri + cj] + Vt[3 * cj + ri]);
      st9(ex + E_V + boff(r, c), Vb);
    }
    __syncthreads();  // V', v_x in place for the next knot; every warp is done with this knot's record buffer
  }

  if (!valid || l != 0) return;
  backward_finish(p, a, b, QuTk, kTQuuk);
}

}  // namespace qilqr
