// =============================================================================
// qilqr_device.cuh -- FP64 device library for the quadrotor-on-SE(3) iLQR hot path.
//
// Lie-group primitives equivalent to the manif calls the reference makes
// (call sites: quadrotor_model.cc:68-78,89,183-186,204-205,211-212,217-218,232-235;
// cost.hh:43), the quadrotor dynamics with the block structure of its Jacobians
// (quadrotor_model.cc:33-122,174-200,266-276) and the quadratic tracking cost
// (cost.hh:36-61).  Everything is fully unrolled scalar code on registers; no
// dense 12x12 Jacobian is ever materialised on the hot path: the kernels work on
// the six non-trivial 3x3 blocks of A = df/dx and the 16 non-zeros of B = df/du.
//
// Storage: quaternion (x,y,z,w); 3x3 matrices row-major double[9].
// =============================================================================
#pragma once
#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#include <math.h>
#endif

#define QD __device__ __forceinline__

// ---------------------------------------------------------------------------
// Rounding contract.  The library is compiled with -fmad=false: the compiler never contracts a
// multiply and an add on its own, so the same inline function rounds the same way in every kernel it
// is inlined into (bulk, tail and API kernels are bit-identical by construction).  Every fused
// multiply-add is therefore written out:
//   QFMA(a, b, c)   a*b + c in ONE rounding (production) -- the summation order is the reference's
//                   (Eigen's coefficient-wise dot products, k ascending), only the rounding is fused;
//   QDIV(x, d, r)   x / d, computed as x * r with r = 1/d precomputed (production).
// The opt-in STRICT build (-DQILQR_STRICT, libqilqr_b200_strict.so; tools/full_batch_parity.py) turns
// both back into what the reference's x86-64 build executes -- a rounded product followed by a rounded
// sum, and a true division -- which leaves the CUDA math library (sin, cos, atan2 against glibc's) as
// the only rounding difference to the CPU oracle.
// ---------------------------------------------------------------------------
#ifdef QILQR_STRICT
#define QFMA(a, b, c) __dadd_rn(__dmul_rn((a), (b)), (c))
#define QDIV(x, d, r) ((x) / (d))
#else
#define QFMA(a, b, c) fma((a), (b), (c))
#define QDIV(x, d, r) ((x) * (r))
#endif

// sin / cos / atan2: the CUDA math library, or -- for the bit-for-bit comparison with the oracle only
// (-DQILQR_PORTABLE_LIBM, see qilqr_portable_libm.h) -- a portable implementation shared with the host
#ifdef QILQR_PORTABLE_LIBM
#include "qilqr_portable_libm.h"
#define QSINCOS(x, s, c) qilqr_plibm::sincos((x), (s), (c))
#define QATAN2(y, x) qilqr_plibm::atan2((y), (x))
#else
#define QSINCOS(x, s, c) sincos((x), (s), (c))
#define QATAN2(y, x) atan2((y), (x))
#endif

namespace qilqr {

// Branch-free selection (the compiler turns chained `c == k ? ... : ...` on doubles into branches otherwise).
QD double selp(double a, double b, bool take_a) {
  double r;
  asm("{ .reg .pred q; setp.ne.s32 q, %3, 0; selp.f64 %0, %1, %2, q; }" : "=d"(r) : "d"(a), "d"(b), "r"(int(take_a)));
  return r;
}
// v[c] for c in 0..3
QD double sel4(int c, double v0, double v1, double v2, double v3) {
  return selp(selp(v3, v2, (c & 1) != 0), selp(v1, v0, (c & 1) != 0), (c & 2) != 0);
}

constexpr double kEps = 1e-14;  // manif Constants<double>::eps

// Solver constants, passed to every kernel by value (constant bank, so matrix
// entries are FMA immediates-from-constant rather than loads).
struct DeviceParams {
  double mass, g, dt;
  double inertia[9];
  double L[9];         // Cholesky factor of inertia (Eigen LLT, quadrotor_model.cc:20)
  double Linv[3];      // 1 / L[i][i]
  double moment_arms[12];  // 3x4, quadrotor_model.cc:15-18
  double JuC[16];      // rows 8..11 of the continuous J_u (quadrotor_model.cc:113-119)
  double Bu[16];       // rows 8..11 of B = dt * JuC (quadrotor_model.cc:45,272); rows 0..7 are zero
  double Q[144];
  double R[16];
  double step_update, desired_reduction_frac, rtol, atol, max_iters, quu_reg;
  int ls_max_iters;
  int symmetrize_vxx;
  // model variant behind the ModelT concept (qilqr_model_generic.cuh); 0/0 = the reference's QuadrotorModel
  int integrator;  // 0: explicit Euler (quadrotor_model.cc:33-49), 1: RK4 (the scheme commented out at :51-63)
  int coriolis;    // 1: adds -omega x v to the body linear acceleration
  int q_diagonal;  // Q has no off-diagonal entry: the cost skips the products with structural zeros
};

// ---------------------------------------------------------------------------
// 3x3 helpers
// ---------------------------------------------------------------------------
QD void m3_mul(const double *A, const double *B, double *C) {  // C = A B
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      C[3 * i + j] = QFMA(A[3 * i + 2], B[6 + j], QFMA(A[3 * i + 1], B[3 + j], A[3 * i] * B[j]));
}
QD void m3_madd(const double *A, const double *B, double *C) {  // C += A B
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      C[3 * i + j] = QFMA(A[3 * i + 2], B[6 + j], QFMA(A[3 * i + 1], B[3 + j], QFMA(A[3 * i], B[j], C[3 * i + j])));
}
QD void m3_mulT(const double *A, const double *B, double *C) {  // C = A^T B
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      C[3 * i + j] = QFMA(A[6 + i], B[6 + j], QFMA(A[3 + i], B[3 + j], A[i] * B[j]));
}
QD void m3_maddT(const double *A, const double *B, double *C) {  // C += A^T B
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      C[3 * i + j] = QFMA(A[6 + i], B[6 + j], QFMA(A[3 + i], B[3 + j], QFMA(A[i], B[j], C[3 * i + j])));
}
QD void m3_vec(const double *A, const double *v, double *r) {  // r = A v
#pragma unroll
  for (int i = 0; i < 3; ++i) r[i] = QFMA(A[3 * i + 2], v[2], QFMA(A[3 * i + 1], v[1], A[3 * i] * v[0]));
}
QD void m3T_vec(const double *A, const double *v, double *r) {  // r = A^T v
#pragma unroll
  for (int i = 0; i < 3; ++i) r[i] = QFMA(A[6 + i], v[2], QFMA(A[3 + i], v[1], A[i] * v[0]));
}
QD void m3T_vec_add(const double *A, const double *v, double *r) {  // r += A^T v
#pragma unroll
  for (int i = 0; i < 3; ++i) r[i] = QFMA(A[6 + i], v[2], QFMA(A[3 + i], v[1], QFMA(A[i], v[0], r[i])));
}
QD void m3_transpose(const double *A, double *T) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) T[3 * i + j] = A[3 * j + i];
}
// C = M * hat(w)  (column j of hat(w) has two non-zeros)
QD void m3_mul_hat(const double *M, const double *w, double *C) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    C[3 * i + 0] = QFMA(M[3 * i + 1], w[2], -(M[3 * i + 2] * w[1]));
    C[3 * i + 1] = QFMA(M[3 * i + 2], w[0], -(M[3 * i + 0] * w[2]));
    C[3 * i + 2] = QFMA(M[3 * i + 0], w[1], -(M[3 * i + 1] * w[0]));
  }
}
// C = hat(t) * M  (row i of hat(t) has two non-zeros)
QD void m3_hat_mul(const double *t, const double *M, double *C) {
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    C[0 + j] = QFMA(t[1], M[6 + j], -(t[2] * M[3 + j]));
    C[3 + j] = QFMA(t[2], M[0 + j], -(t[0] * M[6 + j]));
    C[6 + j] = QFMA(t[0], M[3 + j], -(t[1] * M[0 + j]));
  }
}

// Eigen toRotationMatrix (what manif's rotation() returns).  Every product and sum is rounded on its own
// (no FMA contraction), as in the reference build: besides matching it more closely, this keeps the
// entries independent of which of them a caller uses (a kernel that needs only the third row would
// otherwise get differently-contracted code after dead-code elimination).
QD void quat_to_rot(const double *q, double *R) {
  const double tx = 2.0 * q[0], ty = 2.0 * q[1], tz = 2.0 * q[2];
  const double twx = __dmul_rn(tx, q[3]), twy = __dmul_rn(ty, q[3]), twz = __dmul_rn(tz, q[3]);
  const double txx = __dmul_rn(tx, q[0]), txy = __dmul_rn(ty, q[0]), txz = __dmul_rn(tz, q[0]);
  const double tyy = __dmul_rn(ty, q[1]), tyz = __dmul_rn(tz, q[1]), tzz = __dmul_rn(tz, q[2]);
  R[0] = __dsub_rn(1.0, __dadd_rn(tyy, tzz)); R[1] = __dsub_rn(txy, twz);               R[2] = __dadd_rn(txz, twy);
  R[3] = __dadd_rn(txy, twz);               R[4] = __dsub_rn(1.0, __dadd_rn(txx, tzz)); R[5] = __dsub_rn(tyz, twx);
  R[6] = __dsub_rn(txz, twy);               R[7] = __dadd_rn(tyz, twx);               R[8] = __dsub_rn(1.0, __dadd_rn(txx, tyy));
}
// manif SO3::compose: Hamilton product + first-order renormalisation
// The products are rounded individually (no FMA contraction) so that conj(q) (x) q has an
// exactly zero vector part, as in the reference: x (-) x == 0 exactly (ilqr.hh:161 at knot 0).
QD void quat_compose(const double *a, const double *b, double *r) {
#define QM(i, j) __dmul_rn(a[i], b[j])
  double w = __dadd_rn(__dadd_rn(__dadd_rn(QM(3, 3), -QM(0, 0)), -QM(1, 1)), -QM(2, 2));
  double x = __dadd_rn(__dadd_rn(__dadd_rn(QM(3, 0), QM(0, 3)), QM(1, 2)), -QM(2, 1));
  double y = __dadd_rn(__dadd_rn(__dadd_rn(QM(3, 1), QM(1, 3)), QM(2, 0)), -QM(0, 2));
  double z = __dadd_rn(__dadd_rn(__dadd_rn(QM(3, 2), QM(2, 3)), QM(0, 1)), -QM(1, 0));
#undef QM
  const double sq = QFMA(w, w, QFMA(z, z, QFMA(y, y, x * x)));
  if (fabs(sq - 1.0) > kEps) {
    const double s = 2.0 / (1.0 + sq);
    x *= s; y *= s; z *= s; w *= s;
  }
  r[0] = x; r[1] = y; r[2] = z; r[3] = w;
}

// ---------------------------------------------------------------------------
// SO(3) tangent-space coefficient sets.  For w with theta = |w|:
//   Jl(w)    = I + a W + b W^2         Jr = Jl^T
//   Jr^-1(w) = I + W/2 + c W^2         Jl^-1 = (Jr^-1)^T
// ---------------------------------------------------------------------------
QD void hat_sq(const double *w, double *WW) {  // hat(w)^2, exactly as the matrix product
  const double xx = w[0] * w[0], yy = w[1] * w[1], zz = w[2] * w[2];
  const double xy = w[0] * w[1], xz = w[0] * w[2], yz = w[1] * w[2];
  WW[0] = -(zz + yy); WW[1] = xy;         WW[2] = xz;
  WW[3] = xy;         WW[4] = -(zz + xx); WW[5] = yz;
  WW[6] = xz;         WW[7] = yz;         WW[8] = -(yy + xx);
}
// J = I + a*hat(w) + b*hat(w)^2
QD void so3_jac_from_coeffs(const double *w, double a, double b, double *J) {
  double WW[9];
  hat_sq(w, WW);
  J[0] = QFMA(b, WW[0], 1.0);          J[1] = QFMA(b, WW[1], -a * w[2]); J[2] = QFMA(b, WW[2], a * w[1]);
  J[3] = QFMA(b, WW[3], a * w[2]);     J[4] = QFMA(b, WW[4], 1.0);       J[5] = QFMA(b, WW[5], -a * w[0]);
  J[6] = QFMA(b, WW[6], -a * w[1]);    J[7] = QFMA(b, WW[7], a * w[0]);  J[8] = QFMA(b, WW[8], 1.0);
}
// theta^2, theta, sin(theta), cos(theta) of a rotation vector, evaluated once and shared by the
// Jacobians that manif evaluates separately (ljac/ljacinv/fillQ all use the same theta).
struct Angle {
  double th2, th, s, c;  // th, s, c are meaningful only when th2 > kEps
};
QD Angle angle_of(const double *w) {
  Angle a;
  a.th2 = QFMA(w[2], w[2], QFMA(w[1], w[1], w[0] * w[0]));
  a.th = 0.0; a.s = 0.0; a.c = 1.0;
  if (a.th2 > kEps) {
    a.th = sqrt(a.th2);
    QSINCOS(a.th, &a.s, &a.c);
  }
  return a;
}
// manif SO3Tangent::ljac
QD void so3_ljac(const double *w, const Angle &a, double *J) {
  if (a.th2 <= kEps) { so3_jac_from_coeffs(w, 0.5, 0.0, J); return; }
  so3_jac_from_coeffs(w, (1.0 - a.c) / a.th2, (a.th - a.s) / (a.th2 * a.th), J);
}
QD void so3_ljac(const double *w, double *J) { so3_ljac(w, angle_of(w), J); }
// manif SO3Tangent::ljacinv;  rjacinv is its transpose
QD void so3_ljacinv(const double *w, const Angle &a, double *J) {
  if (a.th2 <= kEps) { so3_jac_from_coeffs(w, -0.5, 0.0, J); return; }
  so3_jac_from_coeffs(w, -0.5, 1.0 / a.th2 - (1.0 + a.c) / (2.0 * a.th * a.s), J);
}
QD void so3_ljacinv(const double *w, double *J) { so3_ljacinv(w, angle_of(w), J); }
// manif SO3Tangent::exp
QD void so3_exp(const double *w, double *q) {
  const double th2 = QFMA(w[2], w[2], QFMA(w[1], w[1], w[0] * w[0]));
  if (th2 > kEps) {
    const double th = sqrt(th2);
    double s, c;
    QSINCOS(0.5 * th, &s, &c);
    q[0] = s * (w[0] / th); q[1] = s * (w[1] / th); q[2] = s * (w[2] / th); q[3] = c;
  } else {
    q[0] = w[0] / 2.0; q[1] = w[1] / 2.0; q[2] = w[2] / 2.0; q[3] = 1.0;
  }
}
// manif SO3::log
QD void so3_log(const double *q, double *w) {
  const double s2 = QFMA(q[2], q[2], QFMA(q[1], q[1], q[0] * q[0]));
  double coeff;
  if (s2 > kEps) {
    const double s = sqrt(s2);
    const double two_angle = 2.0 * ((q[3] < 0.0) ? QATAN2(-s, -q[3]) : QATAN2(s, q[3]));
    coeff = two_angle / s;
  } else {
    coeff = 2.0;
  }
  w[0] = q[0] * coeff; w[1] = q[1] * coeff; w[2] = q[2] * coeff;
}

// manif SE3Tangent::fillQ (Barfoot's Q block, manif's arrangement), for tangent (v, w)
// (`a` = angle_of(w); the angle of -w is the same)
QD void se3_fillQ(const double *v, const double *w, const Angle &a, double *Q) {
  const double th2 = a.th2;
  double B, C, D;
  if (th2 <= kEps) {
    B = 1.0 / 6.0 + (1.0 / 120.0) * th2;
    C = -(1.0 / 24.0) + (1.0 / 720.0) * th2;
    D = -(1.0 / 60.0);
  } else {
    const double th = a.th, s = a.s, c = a.c;
    B = (th - s) / (th2 * th);
    C = (1.0 - th2 / 2.0 - c) / (th2 * th2);
    D = C - 3.0 * (th - s - th2 * th / 6.0) / (th2 * th2 * th);
  }
  // VW = hat(v) hat(w) = w v^T - (v.w) I  (entry by entry, as the matrix product gives it)
  double VW[9], WV[9], WVW[9], VWW[9], WVWW[9];
  VW[0] = QFMA(-v[1], w[1], -(v[2] * w[2])); VW[1] = v[1] * w[0];                      VW[2] = v[2] * w[0];
  VW[3] = v[0] * w[1];                      VW[4] = QFMA(-v[0], w[0], -(v[2] * w[2])); VW[5] = v[2] * w[1];
  VW[6] = v[0] * w[2];                      VW[7] = v[1] * w[2];                      VW[8] = QFMA(-v[0], w[0], -(v[1] * w[1]));
  m3_transpose(VW, WV);
  m3_mul_hat(WV, w, WVW);
  m3_mul_hat(VW, w, VWW);
  m3_mul_hat(WVW, w, WVWW);
  const double V[9] = {0.0, -v[2], v[1], v[2], 0.0, -v[0], -v[1], v[0], 0.0};
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int ij = 3 * i + j, ji = 3 * j + i;
      // ((A V + B (..)) - C (..)) - D (..), each product fused into the running sum
      const double t1 = 0.5 * V[ij];
      const double s2 = (WV[ij] + VW[ij]) + WVW[ij];
      const double s3 = QFMA(-3.0, WVW[ij], VWW[ij] - VWW[ji]);
      Q[ij] = QFMA(-D, WVWW[ij], QFMA(-C, s3, QFMA(B, s2, t1)));
    }
}
QD void se3_fillQ(const double *v, const double *w, double *Q) { se3_fillQ(v, w, angle_of(w), Q); }

// ---------------------------------------------------------------------------
// SE(3) pose = t[3], q[4].
// ---------------------------------------------------------------------------
// tau = Log(B^-1 o A)   (manif rminus; State - State, quadrotor_model.cc:215-219)
// Optionally returns Jl^-1(w) (3x3) so callers can build Jr^-1 blocks without
// recomputing the trigonometry.
QD void se3_rminus(const double *tA, const double *qA, const double *tB, const double *qB,
                   double *tau, double *Jlinv_out /* may be nullptr */, Angle *angle_out = nullptr) {
  // inverse(B) = (-R_B^T t_B, conj(q_B))
  double RB[9];
  quat_to_rot(qB, RB);
  double tinv[3];
  m3T_vec(RB, tB, tinv);
  const double qBc[4] = {-qB[0], -qB[1], -qB[2], qB[3]};
  // compose(inverse(B), A) = (R(conj q_B) t_A + tinv', q_Bc (x) q_A)
  double trel[3];
  m3T_vec(RB, tA, trel);
#pragma unroll
  for (int i = 0; i < 3; ++i) trel[i] = trel[i] - tinv[i];
  double qrel[4];
  quat_compose(qBc, qA, qrel);
  // log
  double w[3];
  so3_log(qrel, w);
  double Jli[9];
  const Angle ang = angle_of(w);
  if (angle_out) *angle_out = ang;
  so3_ljacinv(w, ang, Jli);
  m3_vec(Jli, trel, tau);
  tau[3] = w[0]; tau[4] = w[1]; tau[5] = w[2];
  if (Jlinv_out) {
#pragma unroll
    for (int i = 0; i < 9; ++i) Jlinv_out[i] = Jli[i];
  }
}

// Blocks of the SE(3) right-Jacobian inverse at tau (manif SE3Tangent::rjacinv):
//   Jr^-1(tau) = [[Ji, Qi], [0, Ji]],  Ji = Jr^-1(w) = Jl^-1(w)^T,  Qi = -Ji Q(-tau) Ji
QD void se3_rjacinv_blocks(const double *tau, const double *Jlinv, const Angle &ang, double *Ji, double *Qi) {
  m3_transpose(Jlinv, Ji);
  const double nv[3] = {-tau[0], -tau[1], -tau[2]}, nw[3] = {-tau[3], -tau[4], -tau[5]};
  double Qm[9], T[9];
  se3_fillQ(nv, nw, ang, Qm);
  m3_mul(Ji, Qm, T);
  m3_mul(T, Ji, Qi);
#pragma unroll
  for (int i = 0; i < 9; ++i) Qi[i] = -Qi[i];
}

// ---------------------------------------------------------------------------
// Quadrotor dynamics
// ---------------------------------------------------------------------------
// Eigen LLT solve with the precomputed factor: L L^T x = b
QD void inertia_solve(const DeviceParams &p, const double *b, double *x) {
  const double y0 = QDIV(b[0], p.L[0], p.Linv[0]);
  const double y1 = QDIV(QFMA(-p.L[3], y0, b[1]), p.L[4], p.Linv[1]);
  const double y2 = QDIV(QFMA(-p.L[7], y1, QFMA(-p.L[6], y0, b[2])), p.L[8], p.Linv[2]);
  x[2] = QDIV(y2, p.L[8], p.Linv[2]);
  x[1] = QDIV(QFMA(-p.L[7], x[2], y1), p.L[4], p.Linv[1]);
  x[0] = QDIV(QFMA(-p.L[6], x[2], QFMA(-p.L[3], x[1], y0)), p.L[0], p.Linv[0]);
}

// Body acceleration of continuous_dynamics (quadrotor_model.cc:65-78): acc[6]
QD void body_acceleration(const DeviceParams &p, const double *R /*rotation of q*/,
                          const double *vel, const double *u, double *acc) {
  const double usum = ((u[0] + u[1]) + u[2]) + u[3];
  acc[0] = -p.g * R[6];
  acc[1] = -p.g * R[7];
  acc[2] = QFMA(-p.g, R[8], usum / p.mass);
  double M[3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
    M[i] = QFMA(p.moment_arms[4 * i + 3], u[3],
               QFMA(p.moment_arms[4 * i + 2], u[2], QFMA(p.moment_arms[4 * i + 1], u[1], p.moment_arms[4 * i] * u[0])));
  const double *om = vel + 3;
  // M - (hat(omega) * inertia) * omega, in the reference's association (quadrotor_model.cc:76-77)
  double HI[9], HIw[3];
  m3_hat_mul(om, p.inertia, HI);
  m3_vec(HI, om, HIw);
  const double rhs[3] = {M[0] - HIw[0], M[1] - HIw[1], M[2] - HIw[2]};
  inertia_solve(p, rhs, acc + 3);
}

// One explicit-Euler step on the manifold without derivatives
// (discrete_dynamics, quadrotor_model.cc:33-49 with diffs == nullptr):
//   pose+ = pose o Exp(dt * vel),  vel+ = vel + dt * acc
QD void discrete_step(const DeviceParams &p, double *t, double *q, double *vel, const double *u) {
  double R[9];
  quat_to_rot(q, R);
  double acc[6];
  body_acceleration(p, R, vel, u, acc);
  const double dv[3] = {p.dt * vel[0], p.dt * vel[1], p.dt * vel[2]};
  const double dw[3] = {p.dt * vel[3], p.dt * vel[4], p.dt * vel[5]};
  double Jl[9], te[3], qe[4];
  so3_ljac(dw, Jl);
  m3_vec(Jl, dv, te);
  so3_exp(dw, qe);
  double Rt[3];
  m3_vec(R, te, Rt);
#pragma unroll
  for (int i = 0; i < 3; ++i) t[i] = Rt[i] + t[i];
  double qn[4];
  quat_compose(q, qe, qn);
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] = qn[i];
#pragma unroll
  for (int i = 0; i < 6; ++i) vel[i] = QFMA(p.dt, acc[i], vel[i]);
}

// The non-trivial 3x3 blocks of A = d(discrete_dynamics)/dx at (x, u)
// (quadrotor_model.cc:42-44 with the structure of :84-111, :188-195, :271-272):
//
//        [ Re   Te   dJr  dQb ]      Re|Te : Ad(Exp(dt v)^-1) = [[Re, Te], [0, Re]]
//    A = [ 0    Re   0    dJr ]      dJr|dQb: dt * Jr(dt v)   = dt [[Jr, Qb], [0, Jr]]
//        [ 0    dG   I    0   ]      dG   : dt * (-g hat(R^T e_z))
//        [ 0    0    0    Wd  ]      Wd   : I + dt * (-I^-1 (hat(w) I - hat(I w)))
//
// B = d(discrete_dynamics)/du has rows 0..7 zero and rows 8..11 = p.Bu (constant).
struct ABlocks {
  double Re[9], Te[9], dJr[9], dQb[9], dG[9], Wd[9];
};
// Jacobian blocks of X (+) tau = X o Exp(tau) (manif rplus; add(), quadrotor_model.cc:174-200):
//   d/dX = Ad(Exp(tau)^-1) = [[Re, Te], [0, Re]],   d/dtau = Jr(tau) = [[Jr, Qb], [0, Jr]]
// Also returns Exp(tau) = (te, qe) and R(qe).
QD void se3_plus_blocks(const double *tau, double *Re, double *Te, double *Jr, double *Qb, double *te,
                        double *qe) {
  const double *v = tau, *w = tau + 3;
  double Jl[9], RE[9];
  const Angle ang = angle_of(w);
  so3_ljac(w, ang, Jl);
  m3_vec(Jl, v, te);
  so3_exp(w, qe);
  quat_to_rot(qe, RE);
  m3_transpose(RE, Re);  // rotation of conj(qe)
  double tinv[3];
  m3T_vec(RE, te, tinv);
#pragma unroll
  for (int i = 0; i < 3; ++i) tinv[i] = -tinv[i];
  m3_hat_mul(tinv, Re, Te);
  m3_transpose(Jl, Jr);
  const double nv[3] = {-v[0], -v[1], -v[2]}, nw[3] = {-w[0], -w[1], -w[2]};
  se3_fillQ(nv, nw, ang, Qb);
}
// Continuous-time Jacobian pieces (quadrotor_model.cc:88-111):
//   gz = -g R^T e_z  (G = hat(gz)),   Wc = -I^-1 (hat(w) I - hat(I w))
QD void continuous_blocks(const DeviceParams &p, const double *q, const double *vel, double *gz, double *Wc) {
  double R[9];
  quat_to_rot(q, R);
  gz[0] = -p.g * R[6]; gz[1] = -p.g * R[7]; gz[2] = -p.g * R[8];
  const double *om = vel + 3;
  double Iw[3];
  m3_vec(p.inertia, om, Iw);
  double HI[9];
  m3_hat_mul(om, p.inertia, HI);
  const double HIw[9] = {0.0, -Iw[2], Iw[1], Iw[2], 0.0, -Iw[0], -Iw[1], Iw[0], 0.0};
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const double col[3] = {HI[j] - HIw[j], HI[3 + j] - HIw[3 + j], HI[6 + j] - HIw[6 + j]};
    double sol[3];
    inertia_solve(p, col, sol);
#pragma unroll
    for (int i = 0; i < 3; ++i) Wc[3 * i + j] = -sol[i];
  }
}
QD void dynamics_blocks(const DeviceParams &p, const double *q, const double *vel, ABlocks &A) {
  double tau[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) tau[i] = p.dt * vel[i];
  double Jr[9], Qb[9], te[3], qe[4];
  se3_plus_blocks(tau, A.Re, A.Te, Jr, Qb, te, qe);
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    A.dJr[i] = p.dt * Jr[i];
    A.dQb[i] = p.dt * Qb[i];
  }
  double g3[3], Wc[9];
  continuous_blocks(p, q, vel, g3, Wc);
  const double gz[3] = {p.dt * g3[0], p.dt * g3[1], p.dt * g3[2]};
  A.dG[0] = 0.0;    A.dG[1] = -gz[2]; A.dG[2] = gz[1];
  A.dG[3] = gz[2];  A.dG[4] = 0.0;    A.dG[5] = -gz[0];
  A.dG[6] = -gz[1]; A.dG[7] = gz[0];  A.dG[8] = 0.0;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) A.Wd[3 * i + j] = QFMA(p.dt, Wc[3 * i + j], (i == j) ? 1.0 : 0.0);
}

// ---------------------------------------------------------------------------
// Tracking cost (cost.hh:36-61)
// ---------------------------------------------------------------------------
// delta_x = x (-) x_d  as 12 coefficients [Log(x_d^-1 x); v - v_d]
QD void state_minus(const double *x /*13*/, const double *xd /*13*/, double *dx /*12*/, double *Jlinv,
                    Angle *angle_out = nullptr) {
  se3_rminus(x, x + 3, xd, xd + 3, dx, Jlinv, angle_out);
#pragma unroll
  for (int i = 0; i < 6; ++i) dx[6 + i] = x[7 + i] - xd[7 + i];
}
// cost = dx^T Q dx + du^T R du, evaluated as (dx^T Q) dx   (cost.hh:47-48)
QD double quadratic_cost_state(const DeviceParams &p, const double *dx) {
  double cx = 0.0;
  if (p.q_diagonal) {
    // dx^T Q with a diagonal Q: every other term of the dense chain below is (+-0) + (+-0), so column j of the
    // product is the rounded dx_j * Q_jj either way; the sum over j keeps its order
#pragma unroll
    for (int j = 0; j < 12; ++j) {
      const double y = dx[j] * p.Q[13 * j];
      cx = (j == 0) ? y * dx[0] : QFMA(y, dx[j], cx);
    }
    return cx;
  }
#pragma unroll
  for (int j = 0; j < 12; ++j) {
    double y = dx[0] * p.Q[j];
#pragma unroll
    for (int i = 1; i < 12; ++i) y = QFMA(dx[i], p.Q[12 * i + j], y);
    cx = (j == 0) ? y * dx[0] : QFMA(y, dx[j], cx);
  }
  return cx;
}
QD double quadratic_cost_control(const DeviceParams &p, const double *du) {
  double cu = 0.0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    double y = du[0] * p.R[j];
#pragma unroll
    for (int i = 1; i < 4; ++i) y = QFMA(du[i], p.R[4 * i + j], y);
    cu = (j == 0) ? y * du[0] : QFMA(y, du[j], cu);
  }
  return cu;
}
QD double quadratic_cost(const DeviceParams &p, const double *dx, const double *du) {
  return quadratic_cost_state(p, dx) + quadratic_cost_control(p, du);
}

}  // namespace qilqr
