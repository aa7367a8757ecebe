// =============================================================================
// qilqr_backward_g4.cuh -- ILQR::backwards_pass (ilqr.hh:97-147) with FOUR LANES PER
// PROBLEM (a "quad"; 8 problems per warp, one warp per CTA).
//
// Why: one thread per problem needs V_xx, Q_xx and A^T V_xx live at once (>400 doubles)
// and spills ~5 KB per thread.  Here lane c of a quad owns the 12x3 column block c of
// V_xx (36 doubles) and the work of one knot is split so that every 12x12 product is a
// set of 3x3 block products on registers:
//
//   phase L (per KPP knots): lane j linearises knot i-j on its own -- dynamics blocks, cost
//            gradient and the 6x6 pose block of the Gauss-Newton Hessian -- and leaves a
//            101-double record in shared memory (scalar Lie-group code at full lane use)
//   step 1   lane c:  M[:,c] = A^T V[:,c]                      (8 block products)  -> smem
//   step 2   all:     Q_uu, Q_u, Q_x, LDLT, k                  (small, replicated)
//   step 3   lane r:  Q_xx[r,:] = C_xx[r,:] + M[r,:] A ; Q_xu[r] = M[r,:] B   (8 block products)
//   step 4   lane r:  K[:,r] = -Q_uu^-1 Q_xu[r]^T ; (K^T Q_uu)[r]           -> smem
//   step 5   lane r:  V'[r,:] = Q_xx[r,:] - (K^T Q_uu)[r] K                 -> smem
//   step 6   lane c:  V[:,c] <- V'[:,c]  (transpose through smem; optional symmetrisation)
//
// The association and summation order follow the reference ((J_x^T V) J_x, k ascending),
// structural zeros of A and B are skipped (they contribute exact zeros).  Requires the
// pose/velocity coupling blocks of Q to be zero (Q = blkdiag(Q_pp, Q_vv)); the solver uses
// the one-thread-per-problem kernel otherwise.
//
// Shared memory per problem (doubles): KPP records of 101 + an exchange area of 224, padded
// to a stride = 4 (mod 16) so that the quad-strided accesses below are bank-conflict free.
// =============================================================================
#pragma once
#include "qilqr_kernels.cuh"

namespace qilqr {

namespace g4 {
constexpr int R_RE = 0, R_TE = 9, R_DJR = 18, R_DQB = 27, R_GZ = 36, R_WD = 39, R_CX = 48, R_CU = 60, R_CPP = 64;
constexpr int REC = 101;  // = 5 (mod 16)
constexpr int X_M = 0, X_K = 148, X_VX = 196, X_Q = 208, XCH = 224;
__host__ __device__ constexpr int stride(int kpp) {
  const int base = kpp * REC + XCH;
  return base + ((4 - base % 16) + 16) % 16;
}
__host__ __device__ constexpr int smem_doubles(int kpp) { return 8 * stride(kpp) + 36; }
QD int moff(int I, int J) { return X_M + I * 37 + J * 9; }

QD void ld9(const double *s, double *r) {
#pragma unroll
  for (int e = 0; e < 9; ++e) r[e] = s[e];
}
QD void st9(double *s, const double *r) {
#pragma unroll
  for (int e = 0; e < 9; ++e) s[e] = r[e];
}
// C += hat(t) * M
QD void m3_hat_madd(const double *t, const double *M, double *C) {
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    C[0 + j] += fma(t[1], M[6 + j], -(t[2] * M[3 + j]));
    C[3 + j] += fma(t[2], M[0 + j], -(t[0] * M[6 + j]));
    C[6 + j] += fma(t[0], M[3 + j], -(t[1] * M[0 + j]));
  }
}
// C += M * hat(w)
QD void m3_madd_hat(const double *M, const double *w, double *C) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    C[3 * i + 0] += fma(M[3 * i + 1], w[2], -(M[3 * i + 2] * w[1]));
    C[3 * i + 1] += fma(M[3 * i + 2], w[0], -(M[3 * i + 0] * w[2]));
    C[3 * i + 2] += fma(M[3 * i + 0], w[1], -(M[3 * i + 1] * w[0]));
  }
}

// Linearise one knot into a shared-memory record (requires Q_pv = Q_vp = 0).
QD void linearise_to_record(const DeviceParams &p, const double *x, const double *u, const double *xd,
                            const double *ud, double *rec) {
  {
    ABlocks A;
    dynamics_blocks(p, x + 3, x + 7, A);
    st9(rec + R_RE, A.Re);
    st9(rec + R_TE, A.Te);
    st9(rec + R_DJR, A.dJr);
    st9(rec + R_DQB, A.dQb);
    rec[R_GZ + 0] = A.dG[7];
    rec[R_GZ + 1] = A.dG[2];
    rec[R_GZ + 2] = A.dG[3];
    st9(rec + R_WD, A.Wd);
  }
  double dx[12], Jli[9], Ji[9], Qi[9];
  Angle ang;
  state_minus(x, xd, dx, Jli, &ang);
  se3_rjacinv_blocks(dx, Jli, ang, Ji, Qi);
  // y = (2 dx)^T Q with Q = blkdiag(Qpp, Qvv)
  double y[12];
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    double s = (2.0 * dx[0]) * p.Q[j];
#pragma unroll
    for (int i = 1; i < 6; ++i) s = fma(2.0 * dx[i], p.Q[12 * i + j], s);
    y[j] = s;
    double s2 = (2.0 * dx[6]) * p.Q[72 + 6 + j];
#pragma unroll
    for (int i = 7; i < 12; ++i) s2 = fma(2.0 * dx[i], p.Q[12 * i + 6 + j], s2);
    y[6 + j] = s2;
  }
  double Cx[6];
  m3T_vec(Ji, y, Cx);
  m3T_vec(Qi, y, Cx + 3);
  m3T_vec_add(Ji, y + 3, Cx + 3);
#pragma unroll
  for (int j = 0; j < 6; ++j) { rec[R_CX + j] = Cx[j]; rec[R_CX + 6 + j] = y[6 + j]; }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    double s = (2.0 * (u[0] - ud[0])) * p.R[j];
#pragma unroll
    for (int l = 1; l < 4; ++l) s = fma(2.0 * (u[l] - ud[l]), p.R[4 * l + j], s);
    rec[R_CU + j] = s;
  }
  // Cpp = ((2 J6^T) Qpp) J6, J6 = [[Ji, Qi], [0, Ji]]
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    double Pa[6], Pb[6];  // rows i and 3+i of (2 J6^T) Qpp
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      Pa[j] = fma(2.0 * Ji[6 + i], p.Q[24 + j], fma(2.0 * Ji[3 + i], p.Q[12 + j], (2.0 * Ji[i]) * p.Q[j]));
      double s = fma(2.0 * Qi[6 + i], p.Q[24 + j], fma(2.0 * Qi[3 + i], p.Q[12 + j], (2.0 * Qi[i]) * p.Q[j]));
      s = fma(2.0 * Ji[i], p.Q[36 + j], s);
      s = fma(2.0 * Ji[3 + i], p.Q[48 + j], s);
      s = fma(2.0 * Ji[6 + i], p.Q[60 + j], s);
      Pb[j] = s;
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      rec[R_CPP + 6 * i + j] = fma(Pa[2], Ji[6 + j], fma(Pa[1], Ji[3 + j], Pa[0] * Ji[j]));
      rec[R_CPP + 6 * (3 + i) + j] = fma(Pb[2], Ji[6 + j], fma(Pb[1], Ji[3 + j], Pb[0] * Ji[j]));
      double s = fma(Pa[2], Qi[6 + j], fma(Pa[1], Qi[3 + j], Pa[0] * Qi[j]));
      s = fma(Pa[3], Ji[j], s);
      s = fma(Pa[4], Ji[3 + j], s);
      s = fma(Pa[5], Ji[6 + j], s);
      rec[R_CPP + 6 * i + 3 + j] = s;
      double s2 = fma(Pb[2], Qi[6 + j], fma(Pb[1], Qi[3 + j], Pb[0] * Qi[j]));
      s2 = fma(Pb[3], Ji[j], s2);
      s2 = fma(Pb[4], Ji[3 + j], s2);
      s2 = fma(Pb[5], Ji[6 + j], s2);
      rec[R_CPP + 6 * (3 + i) + 3 + j] = s2;
    }
  }
}
}  // namespace g4

template <int KPP>
__global__ void __launch_bounds__(32) k_backward_g4(const __grid_constant__ DeviceParams p,
                                                    const __grid_constant__ BackwardArgs a) {
  using namespace g4;
  extern __shared__ double smem[];
  constexpr int S = stride(KPP);
  const int lane = threadIdx.x, c = lane & 3, q = lane >> 2;
  const int t = blockIdx.x * 8 + q;
  const bool valid = t < a.n;
  const int tt = valid ? t : a.n - 1;  // tail quads redo the last problem and write nothing
  const int b = a.list ? a.list[tt] : tt;
  const int B = a.pr.B, N = a.pr.N, Bd = a.pr.Bd;
  const int bd = (Bd == 1) ? 0 : b;
  const double *traj = a.solve_mode ? (a.st.sel[b] ? a.pr.buf1 : a.pr.buf0) : a.traj;
  double *recs = smem + q * S;
  double *xch = recs + KPP * REC;
  double *s2Qvv = smem + 8 * S;
  for (int e = lane; e < 36; e += 32) s2Qvv[e] = 2.0 * p.Q[12 * (6 + e / 6) + 6 + e % 6];

  double V0[9], V1[9], V2[9], V3[9], vx[12];
#pragma unroll
  for (int e = 0; e < 9; ++e) V0[e] = V1[e] = V2[e] = V3[e] = 0.0;
#pragma unroll
  for (int e = 0; e < 12; ++e) vx[e] = 0.0;
  double QuTk = 0.0, kTQuuk = 0.0;
  __syncwarp();

  for (int i = N - 1; i >= 0; i -= KPP) {
    // ---------------- phase L: lane j linearises knot i - j ----------------
    if (c < KPP && i - c >= 0) {
      double x[13], u[4], xd[13], ud[4];
      load_point(traj, i - c, B, b, x, u);
      load_point(a.pr.desired, i - c, Bd, bd, xd, ud);
      linearise_to_record(p, x, u, xd, ud, recs + c * REC);
    }
    __syncwarp();

#pragma unroll 1
    for (int j = 0; j < KPP; ++j) {
      const int ii = i - j;
      if (ii < 0) break;
      const double *rec = recs + j * REC;
      const double dgz[3] = {rec[R_GZ], rec[R_GZ + 1], rec[R_GZ + 2]};
      const double ndgz[3] = {-dgz[0], -dgz[1], -dgz[2]};

      // ---------------- step 1: M[:,c] = A^T V[:,c] ----------------
      {
        double Ab[9], Mb[9];
        ld9(rec + R_RE, Ab);
        m3_mulT(Ab, V0, Mb);
        st9(xch + moff(0, c), Mb);
        double Tb[9];
        ld9(rec + R_TE, Tb);
        m3_mulT(Tb, V0, Mb);
        m3_maddT(Ab, V1, Mb);
        m3_hat_madd(ndgz, V2, Mb);  // dG^T = hat(dgz)^T = hat(-dgz)
        st9(xch + moff(1, c), Mb);
        ld9(rec + R_DJR, Ab);
        m3_mulT(Ab, V0, Mb);
#pragma unroll
        for (int e = 0; e < 9; ++e) Mb[e] += V2[e];
        st9(xch + moff(2, c), Mb);
        ld9(rec + R_DQB, Tb);
        m3_mulT(Tb, V0, Mb);
        m3_maddT(Ab, V1, Mb);
        ld9(rec + R_WD, Tb);
        m3_maddT(Tb, V3, Mb);
        st9(xch + moff(3, c), Mb);
        // V[8:12, 8:12] for Q_uu: lane 2 owns column 8, lane 3 columns 9..11
        if (c == 2) {
          xch[X_Q + 0] = V2[8];
          xch[X_Q + 4] = V3[2];
          xch[X_Q + 8] = V3[5];
          xch[X_Q + 12] = V3[8];
        } else if (c == 3) {
#pragma unroll
          for (int jj = 0; jj < 3; ++jj) {
            xch[X_Q + 1 + jj] = V2[6 + jj];
            xch[X_Q + 5 + jj] = V3[jj];
            xch[X_Q + 9 + jj] = V3[3 + jj];
            xch[X_Q + 13 + jj] = V3[6 + jj];
          }
        }
      }
      __syncwarp();

      // ---------------- step 2 (replicated): Q_uu, Q_u, Q_x, factorisation, k ----------------
      double Quu[16], Qu[4], Qx[12], k[4];
      Ldlt4 f;
      {
        double V88[16], BtV[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) V88[e] = xch[X_Q + e];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            double acc = p.Bu[jj] * V88[cc];
#pragma unroll
            for (int r = 1; r < 4; ++r) acc = fma(p.Bu[4 * r + jj], V88[4 * r + cc], acc);
            BtV[4 * jj + cc] = acc;
          }
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
#pragma unroll
          for (int l = 0; l < 4; ++l) {
            double acc = BtV[4 * jj] * p.Bu[l];
#pragma unroll
            for (int cc = 1; cc < 4; ++cc) acc = fma(BtV[4 * jj + cc], p.Bu[4 * cc + l], acc);
            Quu[4 * jj + l] = 2.0 * p.R[4 * jj + l] + acc;
          }
        if (p.quu_reg != 0.0) {
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) Quu[5 * jj] += p.quu_reg;
        }
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          double acc = p.Bu[jj] * vx[8];
#pragma unroll
          for (int r = 1; r < 4; ++r) acc = fma(p.Bu[4 * r + jj], vx[8 + r], acc);
          Qu[jj] = rec[R_CU + jj] + acc;
        }
        // Q.x = C.x + A^T v_x
        double Ab[9], T[12];
        ld9(rec + R_RE, Ab);
        m3T_vec(Ab, vx, T);
        double Tb[9];
        ld9(rec + R_TE, Tb);
        m3T_vec(Tb, vx, T + 3);
        m3T_vec_add(Ab, vx + 3, T + 3);
        {  // += dG^T vx[6:9] = hat(-dgz) vx[6:9]
          const double *v = vx + 6;
          T[3] += fma(ndgz[1], v[2], -(ndgz[2] * v[1]));
          T[4] += fma(ndgz[2], v[0], -(ndgz[0] * v[2]));
          T[5] += fma(ndgz[0], v[1], -(ndgz[1] * v[0]));
        }
        ld9(rec + R_DJR, Ab);
        m3T_vec(Ab, vx, T + 6);
#pragma unroll
        for (int e = 0; e < 3; ++e) T[6 + e] += vx[6 + e];
        ld9(rec + R_DQB, Tb);
        m3T_vec(Tb, vx, T + 9);
        m3T_vec_add(Ab, vx + 3, T + 9);
        ld9(rec + R_WD, Tb);
        m3T_vec_add(Tb, vx + 9, T + 9);
#pragma unroll
        for (int e = 0; e < 12; ++e) Qx[e] = rec[R_CX + e] + T[e];
#pragma unroll
        for (int e = 0; e < 16; ++e) f.m[e] = Quu[e];
        ldlt4_compute(f);
        double rhs[4] = {Qu[0], Qu[1], Qu[2], Qu[3]};
        ldlt4_solve(f, rhs);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) k[jj] = -rhs[jj];
      }

      // ---------------- step 3: row block r = c of Q_xx and Q_xu ----------------
      double Q0[9], Q1[9], Q2[9], Q3[9], Qxu[12];
      {
        // C_xx[r,:]: rows 0..5 come from the record's pose block, rows 6..11 from 2 Q_vv
        const bool lo = c < 2;
        const double *src = lo ? (rec + R_CPP + 18 * c) : (s2Qvv + 18 * (c - 2));
#pragma unroll
        for (int ri = 0; ri < 3; ++ri)
#pragma unroll
          for (int cj = 0; cj < 3; ++cj) {
            const double a0 = src[6 * ri + cj], a1 = src[6 * ri + 3 + cj];
            Q0[3 * ri + cj] = lo ? a0 : 0.0;
            Q1[3 * ri + cj] = lo ? a1 : 0.0;
            Q2[3 * ri + cj] = lo ? 0.0 : a0;
            Q3[3 * ri + cj] = lo ? 0.0 : a1;
          }
        double X0[9], X1[9], X2[9], X3[9], Ab[9], T[9];
        ld9(xch + moff(c, 0), X0);
        ld9(rec + R_RE, Ab);
        m3_mul(X0, Ab, T);
#pragma unroll
        for (int e = 0; e < 9; ++e) Q0[e] += T[e];
        double Tb[9];
        ld9(rec + R_TE, Tb);
        m3_mul(X0, Tb, T);
        ld9(xch + moff(c, 1), X1);
        m3_madd(X1, Ab, T);
        ld9(xch + moff(c, 2), X2);
        m3_madd_hat(X2, dgz, T);
#pragma unroll
        for (int e = 0; e < 9; ++e) Q1[e] += T[e];
        ld9(rec + R_DJR, Ab);
        m3_mul(X0, Ab, T);
#pragma unroll
        for (int e = 0; e < 9; ++e) Q2[e] += T[e] + X2[e];
        ld9(rec + R_DQB, Tb);
        m3_mul(X0, Tb, T);
        m3_madd(X1, Ab, T);
        ld9(xch + moff(c, 3), X3);
        ld9(rec + R_WD, Tb);
        m3_madd(X3, Tb, T);
#pragma unroll
        for (int e = 0; e < 9; ++e) Q3[e] += T[e];
        // Q.xu[r] = M[r, 8] B[8,:] + M[r, 9:12] B[9:12,:]
#pragma unroll
        for (int s = 0; s < 3; ++s)
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            double acc = X2[3 * s + 2] * p.Bu[jj];
#pragma unroll
            for (int cc = 0; cc < 3; ++cc) acc = fma(X3[3 * s + cc], p.Bu[4 * (1 + cc) + jj], acc);
            Qxu[4 * s + jj] = acc;
          }
      }

      // ---------------- step 4: K[:, 3r..3r+2], (K^T Q_uu)[r], v_x'[r] ----------------
      double KQ[12];  // [s][l]
      {
        double Ks[12];  // [j][s]
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          double rhs[4] = {Qxu[4 * s], Qxu[4 * s + 1], Qxu[4 * s + 2], Qxu[4 * s + 3]};
          ldlt4_solve(f, rhs);
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) Ks[3 * jj + s] = -rhs[jj];
        }
#pragma unroll
        for (int s = 0; s < 3; ++s)
#pragma unroll
          for (int l = 0; l < 4; ++l) {
            double acc = Ks[s] * Quu[l];
#pragma unroll
            for (int jj = 1; jj < 4; ++jj) acc = fma(Ks[3 * jj + s], Quu[4 * jj + l], acc);
            KQ[4 * s + l] = acc;
          }
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          double acc = KQ[4 * s] * k[0];
#pragma unroll
          for (int l = 1; l < 4; ++l) acc = fma(KQ[4 * s + l], k[l], acc);
          const double qx = (c == 0) ? Qx[s] : (c == 1) ? Qx[3 + s] : (c == 2) ? Qx[6 + s] : Qx[9 + s];
          xch[X_VX + 3 * c + s] = qx - acc;
        }
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
#pragma unroll
          for (int s = 0; s < 3; ++s) {
            xch[X_K + 12 * jj + 3 * c + s] = Ks[3 * jj + s];
            if (valid) a.pr.gK[row_index(ii, 12 * jj + 3 * c + s, 48, B, b)] = Ks[3 * jj + s];
          }
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
          if (valid && c == jj) a.pr.gk[row_index(ii, jj, 4, B, b)] = k[jj];
      }
      __syncwarp();

      // ---------------- step 5: V'[r,:] = Q_xx[r,:] - (K^T Q_uu)[r] K ----------------
      {
        double Kall[48];
#pragma unroll
        for (int e = 0; e < 48; ++e) Kall[e] = xch[X_K + e];
#pragma unroll
        for (int s = 0; s < 3; ++s)
#pragma unroll
          for (int cc = 0; cc < 3; ++cc) {
            double a0 = KQ[4 * s] * Kall[cc], a1 = KQ[4 * s] * Kall[3 + cc], a2 = KQ[4 * s] * Kall[6 + cc],
                   a3 = KQ[4 * s] * Kall[9 + cc];
#pragma unroll
            for (int l = 1; l < 4; ++l) {
              a0 = fma(KQ[4 * s + l], Kall[12 * l + cc], a0);
              a1 = fma(KQ[4 * s + l], Kall[12 * l + 3 + cc], a1);
              a2 = fma(KQ[4 * s + l], Kall[12 * l + 6 + cc], a2);
              a3 = fma(KQ[4 * s + l], Kall[12 * l + 9 + cc], a3);
            }
            Q0[3 * s + cc] -= a0;
            Q1[3 * s + cc] -= a1;
            Q2[3 * s + cc] -= a2;
            Q3[3 * s + cc] -= a3;
          }
#pragma unroll
        for (int e = 0; e < 12; ++e) vx[e] = xch[X_VX + e];
        st9(xch + moff(c, 0), Q0);
        st9(xch + moff(c, 1), Q1);
        st9(xch + moff(c, 2), Q2);
        st9(xch + moff(c, 3), Q3);
        // expected cost reduction terms (ilqr.hh:136-140), replicated
        double acc = Qu[0] * k[0];
#pragma unroll
        for (int jj = 1; jj < 4; ++jj) acc = fma(Qu[jj], k[jj], acc);
        QuTk = QuTk + acc;
        double acc2 = 0.0;
#pragma unroll
        for (int l = 0; l < 4; ++l) {
          double z = k[0] * Quu[l];
#pragma unroll
          for (int jj = 1; jj < 4; ++jj) z = fma(k[jj], Quu[4 * jj + l], z);
          acc2 = (l == 0) ? z * k[0] : fma(z, k[l], acc2);
        }
        kTQuuk = kTQuuk + acc2;
      }
      __syncwarp();

      // ---------------- step 6: V[:,c] <- V'[:,c] ----------------
      ld9(xch + moff(0, c), V0);
      ld9(xch + moff(1, c), V1);
      ld9(xch + moff(2, c), V2);
      ld9(xch + moff(3, c), V3);
      if (p.symmetrize_vxx) {  // V <- (V + V^T)/2: block (K,c) needs (block (c,K))^T, which this lane owns
#pragma unroll
        for (int ri = 0; ri < 3; ++ri)
#pragma unroll
          for (int cj = 0; cj < 3; ++cj) {
            V0[3 * ri + cj] = 0.5 * (V0[3 * ri + cj] + Q0[3 * cj + ri]);
            V1[3 * ri + cj] = 0.5 * (V1[3 * ri + cj] + Q1[3 * cj + ri]);
            V2[3 * ri + cj] = 0.5 * (V2[3 * ri + cj] + Q2[3 * cj + ri]);
            V3[3 * ri + cj] = 0.5 * (V3[3 * ri + cj] + Q3[3 * cj + ri]);
          }
      }
      __syncwarp();
    }
  }

  if (!valid || c != 0) return;
  if (!a.solve_mode) {
    a.terms_out[2 * size_t(b)] = QuTk;
    a.terms_out[2 * size_t(b) + 1] = kTQuuk;
    return;
  }
  const SolveState &st = a.st;
  st.qutk[b] = QuTk;
  st.ktquuk[b] = kTQuuk;
  st.bwd[b] += 1;
  const double cost = st.cost[b];
  const double expected_new_cost = cost + (QuTk + kTQuuk / 2.0);  // ilqr.hh:64-65 with step = 1
  if (a.iter > 0 && is_converged(p, cost, expected_new_cost)) {
    st.status[b] = QILQR_STATUS_CONVERGED_EXPECTED;  // ilqr.hh:66-68
    st.phase[b] = PHASE_DONE;
  } else {
    st.alpha[b] = 1.0;
    st.ls_iter[b] = 0;
    st.phase[b] = a.search_phase;
  }
}

}  // namespace qilqr
