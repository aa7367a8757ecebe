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
// C_pp is stored as two row groups (rows 0-2, rows 3-5) of 18 elements, ONE PADDING ELEMENT APART: in the Riccati step
// lanes c = 0 and c = 1 of every quad read the same position of their own row group at once, and with tiles of 8
// problems per element an even element distance would put both groups on the same 16 shared-memory banks
// (ncu: 2-3 excess wavefronts on each of those 18 loads); an odd distance puts them on the two halves.
constexpr int R_CPP_GROUP = 19;
constexpr int R_CPV = 101, R_CVP = 137;  // only in records of a Q with pose/velocity coupling (173 doubles)
constexpr int REC = 101;  // elements 0..100 of a record in the fused kernel's shared memory (= 5 mod 16)
constexpr int X_M = 0, X_K = 148, X_VX = 196, X_Q = 208, XCH = 224;
__host__ __device__ constexpr int stride(int kpp) {
  const int base = kpp * REC + XCH;
  return base + ((4 - base % 16) + 16) % 16;
}
__host__ __device__ constexpr int smem_doubles(int kpp) { return 8 * stride(kpp) + 36; }
// The 2*Q_vv table that lanes c = 2, 3 of a quad read while lanes c = 0, 1 read their row group of C_pp from the
// record.  RS = 1 (records per problem): 36 doubles, row-major.  RS = 8 (record tiles [element][8 problems]): laid
// out LIKE the C_pp part of a tile -- element (row group g, j) at ((R_CPP_GROUP g + j) * 8 + slot), the same value in
// all 8 slots -- and quad q reads slot (q + 4) & 7: within each half-warp the table lanes then sit on the 16 banks
// that the half-warp's record lanes leave free (ncu: the remaining excess wavefronts of the step came from here).
constexpr int QVV_TILE = (R_CPP_GROUP + 18) * 8;
template <int RS>
QD int qvv_group(int g) { return (RS == 1 ? 18 : R_CPP_GROUP) * g * RS; }
QD void init_qvv_tile(const DeviceParams &p, double *tile, int tid, int nthreads) {
  for (int e = tid; e < 36 * 8; e += nthreads) {
    const int slot = e & 7, k = e >> 3, g = k / 18, j = k % 18;  // k: row-major index into the 6x6 block
    tile[(R_CPP_GROUP * g + j) * 8 + slot] = 2.0 * p.Q[12 * (6 + k / 6) + 6 + k % 6];
  }
}
QD int moff(int I, int J) { return X_M + I * 37 + J * 9; }

QD void ld9(const double *s, double *r) {
#pragma unroll
  for (int e = 0; e < 9; ++e) r[e] = s[e];
}
QD void st9(double *s, const double *r) {
#pragma unroll
  for (int e = 0; e < 9; ++e) s[e] = r[e];
}
template <int RS>
QD void st9s(double *s, const double *r) {
#pragma unroll
  for (int e = 0; e < 9; ++e) s[e * RS] = r[e];
}
// C += hat(t) * M, term by term in the order of the dense product (row i of hat(t) = [0 -t2 t1; t2 0 -t0; -t1 t0 0];
// the structural zero of each row contributes an exact zero and is skipped)
QD void m3_hat_madd(const double *t, const double *M, double *C) {
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    C[0 + j] = QFMA(t[1], M[6 + j], QFMA(-t[2], M[3 + j], C[0 + j]));
    C[3 + j] = QFMA(-t[0], M[6 + j], QFMA(t[2], M[0 + j], C[3 + j]));
    C[6 + j] = QFMA(t[0], M[3 + j], QFMA(-t[1], M[0 + j], C[6 + j]));
  }
}
// C += M * hat(w), likewise (column j of hat(w) = [0 w2 -w1; -w2 0 w0; w1 -w0 0])
QD void m3_madd_hat(const double *M, const double *w, double *C) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    C[3 * i + 0] = QFMA(M[3 * i + 2], -w[1], QFMA(M[3 * i + 1], w[2], C[3 * i + 0]));
    C[3 * i + 1] = QFMA(M[3 * i + 2], w[0], QFMA(M[3 * i + 0], -w[2], C[3 * i + 1]));
    C[3 * i + 2] = QFMA(M[3 * i + 1], -w[0], QFMA(M[3 * i + 0], w[1], C[3 * i + 2]));
  }
}

// Where the cost differentials of one knot go inside a record (element index; the element lives at
// rec[index * RS]): C.x, C.u, and the pose/pose, pose/velocity, velocity/pose 6x6 blocks of C.xx.
struct G4Layout {
  __host__ __device__ static constexpr int cx(int j) { return R_CX + j; }
  __host__ __device__ static constexpr int cu(int j) { return R_CU + j; }
  __host__ __device__ static constexpr int cpp(int i, int j) { return R_CPP + 6 * i + j + (i >= 3 ? R_CPP_GROUP - 18 : 0); }
  __host__ __device__ static constexpr int cpv(int i, int j) { return R_CPV + 6 * i + j; }
  __host__ __device__ static constexpr int cvp(int i, int j) { return R_CVP + 6 * i + j; }
};

// CostFunction differentials (cost.hh:47-57) of one knot into a record.  DENSEQ = false requires
// Q_pv = Q_vp = 0 and writes only C.x, C.u and the pose block of C.xx; DENSEQ = true handles any Q.
template <int RS, bool DENSEQ, class L>
QD void cost_to_record(const DeviceParams &p, const double *x, const double *u, const double *xd,
                       const double *ud, double *rec) {
  double dx[12], Jli[9], Ji[9], Qi[9];
  Angle ang;
  state_minus(x, xd, dx, Jli, &ang);
  se3_rjacinv_blocks(dx, Jli, ang, Ji, Qi);
  // y = (2 dx)^T Q   (with Q = blkdiag(Qpp, Qvv) when !DENSEQ: the skipped terms are exact zeros)
  double y[12];
  if (DENSEQ) {
#pragma unroll
    for (int j = 0; j < 12; ++j) {
      double s = (2.0 * dx[0]) * p.Q[j];
#pragma unroll
      for (int i = 1; i < 12; ++i) s = QFMA(2.0 * dx[i], p.Q[12 * i + j], s);
      y[j] = s;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      double s = (2.0 * dx[0]) * p.Q[j];
#pragma unroll
      for (int i = 1; i < 6; ++i) s = QFMA(2.0 * dx[i], p.Q[12 * i + j], s);
      y[j] = s;
      double s2 = (2.0 * dx[6]) * p.Q[72 + 6 + j];
#pragma unroll
      for (int i = 7; i < 12; ++i) s2 = QFMA(2.0 * dx[i], p.Q[12 * i + 6 + j], s2);
      y[6 + j] = s2;
    }
  }
  double Cx[6];
  m3T_vec(Ji, y, Cx);
  m3T_vec(Qi, y, Cx + 3);
  m3T_vec_add(Ji, y + 3, Cx + 3);
#pragma unroll
  for (int j = 0; j < 6; ++j) { rec[L::cx(j) * RS] = Cx[j]; rec[L::cx(6 + j) * RS] = y[6 + j]; }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    double s = (2.0 * (u[0] - ud[0])) * p.R[j];
#pragma unroll
    for (int l = 1; l < 4; ++l) s = QFMA(2.0 * (u[l] - ud[l]), p.R[4 * l + j], s);
    rec[L::cu(j) * RS] = s;
  }
  // C_xx = ((2 J^T) Q) J with J = blkdiag(J6, I), J6 = [[Ji, Qi], [0, Ji]]:
  //   rows 0..5 of P = (2 J^T) Q are (2 J6^T) Q[0:6, :];  C_pp = P[:, 0:6] J6,  C_pv = P[:, 6:12]
  constexpr int PC = DENSEQ ? 12 : 6;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    double Pa[PC], Pb[PC];  // rows i and 3+i of P
#pragma unroll
    for (int j = 0; j < PC; ++j) {
      Pa[j] = QFMA(2.0 * Ji[6 + i], p.Q[24 + j], QFMA(2.0 * Ji[3 + i], p.Q[12 + j], (2.0 * Ji[i]) * p.Q[j]));
      double s = QFMA(2.0 * Qi[6 + i], p.Q[24 + j], QFMA(2.0 * Qi[3 + i], p.Q[12 + j], (2.0 * Qi[i]) * p.Q[j]));
      s = QFMA(2.0 * Ji[i], p.Q[36 + j], s);
      s = QFMA(2.0 * Ji[3 + i], p.Q[48 + j], s);
      s = QFMA(2.0 * Ji[6 + i], p.Q[60 + j], s);
      Pb[j] = s;
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      rec[L::cpp(i, j) * RS] = QFMA(Pa[2], Ji[6 + j], QFMA(Pa[1], Ji[3 + j], Pa[0] * Ji[j]));
      rec[L::cpp(3 + i, j) * RS] = QFMA(Pb[2], Ji[6 + j], QFMA(Pb[1], Ji[3 + j], Pb[0] * Ji[j]));
      double s = QFMA(Pa[2], Qi[6 + j], QFMA(Pa[1], Qi[3 + j], Pa[0] * Qi[j]));
      s = QFMA(Pa[3], Ji[j], s);
      s = QFMA(Pa[4], Ji[3 + j], s);
      s = QFMA(Pa[5], Ji[6 + j], s);
      rec[L::cpp(i, 3 + j) * RS] = s;
      double s2 = QFMA(Pb[2], Qi[6 + j], QFMA(Pb[1], Qi[3 + j], Pb[0] * Qi[j]));
      s2 = QFMA(Pb[3], Ji[j], s2);
      s2 = QFMA(Pb[4], Ji[3 + j], s2);
      s2 = QFMA(Pb[5], Ji[6 + j], s2);
      rec[L::cpp(3 + i, 3 + j) * RS] = s2;
    }
    if (DENSEQ) {
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        rec[L::cpv(i, j) * RS] = Pa[6 + (DENSEQ ? j : 0)];
        rec[L::cpv(3 + i, j) * RS] = Pb[6 + (DENSEQ ? j : 0)];
      }
    }
  }
  if (DENSEQ) {
    // C_vp = P[6:12, 0:6] J6 with P[6+i, :] = 2 Q[6+i, :]
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      double Pr[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) Pr[k] = 2.0 * p.Q[12 * (6 + i) + k];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        rec[L::cvp(i, j) * RS] = QFMA(Pr[2], Ji[6 + j], QFMA(Pr[1], Ji[3 + j], Pr[0] * Ji[j]));
        double s = QFMA(Pr[2], Qi[6 + j], QFMA(Pr[1], Qi[3 + j], Pr[0] * Qi[j]));
        s = QFMA(Pr[3], Ji[j], s);
        s = QFMA(Pr[4], Ji[3 + j], s);
        s = QFMA(Pr[5], Ji[6 + j], s);
        rec[L::cvp(i, 3 + j) * RS] = s;
      }
    }
  }
}

// Linearise one knot into a record whose element e lives at rec[e * RS].  DENSEQ = false requires
// Q_pv = Q_vp = 0 (100-double record); DENSEQ = true handles any Q (172-double record).
template <int RS = 1, bool DENSEQ = false>
QD void linearise_to_record(const DeviceParams &p, const double *x, const double *u, const double *xd,
                            const double *ud, double *rec) {
  {
    ABlocks A;
    dynamics_blocks(p, x + 3, x + 7, A);
    st9s<RS>(rec + R_RE * RS, A.Re);
    st9s<RS>(rec + R_TE * RS, A.Te);
    st9s<RS>(rec + R_DJR * RS, A.dJr);
    st9s<RS>(rec + R_DQB * RS, A.dQb);
    rec[(R_GZ + 0) * RS] = A.dG[7];
    rec[(R_GZ + 1) * RS] = A.dG[2];
    rec[(R_GZ + 2) * RS] = A.dG[3];
    st9s<RS>(rec + R_WD * RS, A.Wd);
  }
  cost_to_record<RS, DENSEQ, G4Layout>(p, x, u, xd, ud, rec);
}
}  // namespace g4

}  // namespace qilqr
#include "qilqr_riccati_step.cuh"
namespace qilqr {

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

  double V0[9], V1[9], V2[9], V3[9], vx[12], V88[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) V88[e] = 0.0;
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
      riccati_step<1>(p, a, recs + j * REC, s2Qvv, xch, c, a.pr.gk + size_t(c) * B + b, a.pr.gK + size_t(3 * c) * B + b, ii, B,
                      V0, V1, V2, V3, vx, V88, QuTk, kTQuuk);
    }
  }

  if (!valid || c != 0) return;
  backward_finish(p, a, b, QuTk, kTQuuk);
}

}  // namespace qilqr
