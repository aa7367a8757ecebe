// =============================================================================
// qilqr_model_generic.cuh -- the model side of the GENERIC solver path: a second dynamics
// function behind the ModelT concept of ilqr.hh:25-44 (discrete_dynamics(x, u, dt, diffs*)
// with dense J_x (12x12) and J_u (12x4)).
//
// The reference has one model; its solver, however, is a template over the model, and
// SURVEY.md section 8(f)-4 asks that the CUDA design not be hard-wired to one dynamics
// function.  The variants here keep the state manifold SE(3) x R^6 and change the dynamics
// (oracle: QuadrotorModelVariant in oracle/qilqr_oracle.hpp, which is their definition):
//   p.integrator = 1  RK4 over continuous_dynamics -- the scheme the reference left commented
//                     out at quadrotor_model.cc:51-63 -- with the chain rule through its stages
//   p.coriolis   = 1  body-frame transport term -omega x v in the linear acceleration
// With both off the functions reproduce QuadrotorModel (quadrotor_model.cc:33-122).
//
// Jacobians are produced as dense row-major matrices (what a generic model hands the solver);
// internally the model uses the sparsity of ITS OWN factors (the 3x3 blocks of the Euler-step
// and continuous Jacobians), accumulating in ascending column order so that the results equal
// the dense products of the oracle up to the sign of zero.
// =============================================================================
#pragma once
#include "qilqr_device.cuh"

namespace qilqr {
namespace gm {

// xdot = continuous_dynamics(x, u) as 12 coefficients [body velocity; body acceleration]
QD void continuous(const DeviceParams &p, const double *x /*13*/, const double *u, double *xdot /*12*/) {
  double R[9];
  quat_to_rot(x + 3, R);
  body_acceleration(p, R, x + 7, u, xdot + 6);
  if (p.coriolis) {
    const double *v = x + 7, *w = x + 10;
    xdot[6] = xdot[6] + QFMA(v[1], w[2], -(v[2] * w[1]));
    xdot[7] = xdot[7] + QFMA(v[2], w[0], -(v[0] * w[2]));
    xdot[8] = xdot[8] + QFMA(v[0], w[1], -(v[1] * w[0]));
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) xdot[i] = x[7 + i];
}

// detail::euler_step (quadrotor_model.cc:266-276) without derivatives: xn = x (+) dt * k
QD void euler_state(const double *x /*13*/, const double *k /*12*/, double dt, double *xn /*13, may alias x*/) {
  const double dv[3] = {dt * k[0], dt * k[1], dt * k[2]};
  const double dw[3] = {dt * k[3], dt * k[4], dt * k[5]};
  double R[9], Jl[9], te[3], qe[4], Rt[3], qn[4];
  quat_to_rot(x + 3, R);
  so3_ljac(dw, Jl);
  m3_vec(Jl, dv, te);
  so3_exp(dw, qe);
  m3_vec(R, te, Rt);
  quat_compose(x + 3, qe, qn);
#pragma unroll
  for (int i = 0; i < 3; ++i) xn[i] = Rt[i] + x[i];
#pragma unroll
  for (int i = 0; i < 4; ++i) xn[3 + i] = qn[i];
#pragma unroll
  for (int i = 0; i < 6; ++i) xn[7 + i] = QFMA(dt, k[6 + i], x[7 + i]);
}

#ifdef QILQR_USER_MODEL_TU
}  // namespace gm
}  // namespace qilqr
// ---------------------------------------------------------------------------
// USER-SUPPLIED MODEL (qilqr_set_user_model): this translation unit is compiled at run time by NVRTC around a
// device function the caller provides -- the ModelT concept of ilqr.hh:25-44 on the state manifold SE(3) x R^6:
//   x, x_next: 13 doubles (t, q(x,y,z,w), body velocity); u: 4; J_x: 12x12 row-major, J_u: 12x4 row-major,
//   derivatives with respect to right-plus perturbations of the state (as QuadrotorModel::DynamicsDifferentials);
//   J_x / J_u are nullptr when only the step is wanted; params: the doubles passed to qilqr_set_user_model.
// ---------------------------------------------------------------------------
extern "C" __device__ void qilqr_user_discrete_dynamics(const double *params, const double *x, const double *u,
                                                        double dt, double *x_next, double *J_x, double *J_u);
namespace qilqr {
__device__ double user_params[64];
namespace gm {
QD void discrete_step_any(const DeviceParams &p, double *x /*13*/, const double *u) {
  double xn[13];
  qilqr_user_discrete_dynamics(user_params, x, u, p.dt, xn, nullptr, nullptr);
#pragma unroll
  for (int i = 0; i < 13; ++i) x[i] = xn[i];
}
QD void discrete_with_jacobians(const DeviceParams &p, const double *x, const double *u, double *xn, double *A,
                                double *B, int stride) {
  double Jx[144], Ju[48];
  qilqr_user_discrete_dynamics(user_params, x, u, p.dt, xn, A ? Jx : nullptr, B ? Ju : nullptr);
  if (A)
    for (int e = 0; e < 144; ++e) A[e * stride] = Jx[e];
  if (B)
    for (int e = 0; e < 48; ++e) B[e * stride] = Ju[e];
}
#else
// discrete_dynamics without derivatives for any variant; x is advanced in place
QD void discrete_step_any(const DeviceParams &p, double *x /*13*/, const double *u) {
  if (!p.integrator) {
    if (!p.coriolis) {
      discrete_step(p, x, x + 3, x + 7, u);
      return;
    }
    double k[12];
    continuous(p, x, u, k);
    euler_state(x, k, p.dt, x);
    return;
  }
  const double half = p.dt / 2.0;
  double k[12], xdot[12];
#pragma unroll
  for (int e = 0; e < 12; ++e) { k[e] = 0.0; xdot[e] = 0.0; }
#pragma unroll 1
  for (int i = 0; i < 4; ++i) {
    const double dti = (i == 0) ? 0.0 : (i == 3 ? p.dt : half);
    const double ci = (i == 0 || i == 3) ? 1.0 / 6.0 : 2.0 / 6.0;
    double xi[13];
    euler_state(x, k, dti, xi);
    continuous(p, xi, u, k);
#pragma unroll
    for (int e = 0; e < 12; ++e) xdot[e] = QFMA(ci, k[e], xdot[e]);
  }
  euler_state(x, xdot, p.dt, x);
}

// ---------------------------------------------------------------------------
// Differentials.  Dense matrices are row-major [12][n] in local memory (n = 12 or 4).
// ---------------------------------------------------------------------------
struct EulerBlocks {  // euler_step Jacobians: J_lhs = blkdiag([[Re,Te],[0,Re]], I6), J_rhs = blkdiag([[dJr,dQb],[0,dJr]], dt I6)
  double Re[9], Te[9], dJr[9], dQb[9], dt;
};
struct ContBlocks {  // continuous J_x: [0 I6] on rows 0..5; row block 2: G at cols 3..5 (+ transport blocks); row block 3: Wc at cols 9..11
  double G[9], Wc[9], nW[9], V[9];  // nW = -hat(omega), V = hat(v): only with p.coriolis
};

QD void euler_with_blocks(const double *x, const double *k, double dt, double *xn, EulerBlocks &E) {
  double tau[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) tau[i] = dt * k[i];
  double Jr[9], Qb[9], te[3], qe[4], R[9], Rt[3], qn[4];
  se3_plus_blocks(tau, E.Re, E.Te, Jr, Qb, te, qe);
#pragma unroll
  for (int i = 0; i < 9; ++i) { E.dJr[i] = Jr[i] * dt; E.dQb[i] = Qb[i] * dt; }
  E.dt = dt;
  quat_to_rot(x + 3, R);
  m3_vec(R, te, Rt);
  quat_compose(x + 3, qe, qn);
#pragma unroll
  for (int i = 0; i < 3; ++i) xn[i] = Rt[i] + x[i];
#pragma unroll
  for (int i = 0; i < 4; ++i) xn[3 + i] = qn[i];
#pragma unroll
  for (int i = 0; i < 6; ++i) xn[7 + i] = QFMA(dt, k[6 + i], x[7 + i]);
}
QD void continuous_with_blocks(const DeviceParams &p, const double *x, const double *u, double *xdot, ContBlocks &F) {
  continuous(p, x, u, xdot);
  double gz[3];
  continuous_blocks(p, x + 3, x + 7, gz, F.Wc);
  F.G[0] = 0.0;    F.G[1] = -gz[2]; F.G[2] = gz[1];
  F.G[3] = gz[2];  F.G[4] = 0.0;    F.G[5] = -gz[0];
  F.G[6] = -gz[1]; F.G[7] = gz[0];  F.G[8] = 0.0;
  const double *v = x + 7, *w = x + 10;
  F.nW[0] = 0.0;   F.nW[1] = w[2];  F.nW[2] = -w[1];
  F.nW[3] = -w[2]; F.nW[4] = 0.0;   F.nW[5] = w[0];
  F.nW[6] = w[1];  F.nW[7] = -w[0]; F.nW[8] = 0.0;
  F.V[0] = 0.0;   F.V[1] = -v[2]; F.V[2] = v[1];
  F.V[3] = v[2];  F.V[4] = 0.0;   F.V[5] = -v[0];
  F.V[6] = -v[1]; F.V[7] = v[0];  F.V[8] = 0.0;
}

// out[3][n] (+)= blk(3x3) * X[3][n], both strips inside row-major [12][n] matrices
template <int n, bool ACC>
QD void strip_mul(const double *blk, const double *X, double *out) {
#pragma unroll
  for (int c = 0; c < n; ++c) {
    const double x0 = X[c], x1 = X[n + c], x2 = X[2 * n + c];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      double s = ACC ? QFMA(blk[3 * i], x0, out[i * n + c]) : blk[3 * i] * x0;
      s = QFMA(blk[3 * i + 1], x1, s);
      out[i * n + c] = QFMA(blk[3 * i + 2], x2, s);
    }
  }
}
// T = J_rhs * D  (euler_diffs.J_x_rhs * X) for a [12][n] column strip
template <int n>
QD void euler_rhs_mul(const EulerBlocks &E, const double *D, double *T) {
  strip_mul<n, false>(E.dJr, D, T);
  strip_mul<n, true>(E.dQb, D + 3 * n, T);
  strip_mul<n, false>(E.dJr, D + 3 * n, T + 3 * n);
#pragma unroll
  for (int e = 6 * n; e < 12 * n; ++e) T[e] = E.dt * D[e];
}
// T = J_lhs[:, 3s..3s+2] + T   for the column strip s (0..3) of a 12x12 matrix
QD void euler_lhs_add_strip(const EulerBlocks &E, int s, double *T /*[12][3]*/) {
  if (s == 0) {
#pragma unroll
    for (int e = 0; e < 9; ++e) T[e] = E.Re[e] + T[e];
  } else if (s == 1) {
#pragma unroll
    for (int e = 0; e < 9; ++e) { T[e] = E.Te[e] + T[e]; T[9 + e] = E.Re[e] + T[9 + e]; }
  } else {  // identity block of rows 6..11
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      if (s == 2) T[3 * (6 + j) + j] = 1.0 + T[3 * (6 + j) + j];
      else T[3 * (9 + j) + j] = 1.0 + T[3 * (9 + j) + j];
    }
  }
}
// D = J_x^c * T   (continuous Jacobian times a [12][n] strip); D must not alias T
template <int n>
QD void cont_mul(const DeviceParams &p, const ContBlocks &F, const double *T, double *D) {
#pragma unroll
  for (int e = 0; e < 6 * n; ++e) D[e] = T[6 * n + e];
  strip_mul<n, false>(F.G, T + 3 * n, D + 6 * n);
  if (p.coriolis) {
    strip_mul<n, true>(F.nW, T + 6 * n, D + 6 * n);
    strip_mul<n, true>(F.V, T + 9 * n, D + 6 * n);
  }
  strip_mul<n, false>(F.Wc, T + 9 * n, D + 9 * n);
}
// column strip s of the continuous J_x (n = 3), or J_u (n = 4)
template <int n>
QD void cont_strip(const DeviceParams &p, const ContBlocks &F, int s, double *D) {
#pragma unroll
  for (int e = 0; e < 12 * n; ++e) D[e] = 0.0;
  if (n == 4) {
#pragma unroll
    for (int e = 0; e < 16; ++e) D[32 + e] = p.JuC[e];
    return;
  }
  if (s == 1) {
#pragma unroll
    for (int e = 0; e < 9; ++e) D[18 + e] = F.G[e];
  } else if (s == 2) {
#pragma unroll
    for (int j = 0; j < 3; ++j) D[3 * j + j] = 1.0;
    if (p.coriolis) {
#pragma unroll
      for (int e = 0; e < 9; ++e) D[18 + e] = F.nW[e];
    }
  } else if (s == 3) {
#pragma unroll
    for (int j = 0; j < 3; ++j) D[3 * (3 + j) + j] = 1.0;
#pragma unroll
    for (int e = 0; e < 9; ++e) D[27 + e] = F.Wc[e];
    if (p.coriolis) {
#pragma unroll
      for (int e = 0; e < 9; ++e) D[18 + e] = F.V[e];
    }
  }
}

// State recursion of discrete_dynamics for any variant, keeping the Jacobian blocks of every stage:
//   Fs[i]: continuous Jacobian at stage i (Euler: only Fs[0]);  Es[0..2]: Euler-step Jacobians of RK4
//   stages 1..3;  Es[3]: Jacobians of the final step x (+) dt xdot.
struct ContStore {  // ContBlocks without the redundancy of the three skew matrices (kept per stage in local memory)
  double gz[3], Wc[9], w[3], v[3];
};
QD void compact(const ContBlocks &F, ContStore &C) {
  C.gz[0] = F.G[7]; C.gz[1] = F.G[2]; C.gz[2] = F.G[3];
  C.w[0] = F.nW[5]; C.w[1] = F.nW[6]; C.w[2] = F.nW[1];
  C.v[0] = F.V[7];  C.v[1] = F.V[2];  C.v[2] = F.V[3];
#pragma unroll
  for (int e = 0; e < 9; ++e) C.Wc[e] = F.Wc[e];
}
QD void expand(const ContStore &C, ContBlocks &F) {
  const double *gz = C.gz, *w = C.w, *v = C.v;
  F.G[0] = 0.0;    F.G[1] = -gz[2]; F.G[2] = gz[1];
  F.G[3] = gz[2];  F.G[4] = 0.0;    F.G[5] = -gz[0];
  F.G[6] = -gz[1]; F.G[7] = gz[0];  F.G[8] = 0.0;
  F.nW[0] = 0.0;   F.nW[1] = w[2];  F.nW[2] = -w[1];
  F.nW[3] = -w[2]; F.nW[4] = 0.0;   F.nW[5] = w[0];
  F.nW[6] = w[1];  F.nW[7] = -w[0]; F.nW[8] = 0.0;
  F.V[0] = 0.0;   F.V[1] = -v[2]; F.V[2] = v[1];
  F.V[3] = v[2];  F.V[4] = 0.0;   F.V[5] = -v[0];
  F.V[6] = -v[1]; F.V[7] = v[0];  F.V[8] = 0.0;
#pragma unroll
  for (int e = 0; e < 9; ++e) F.Wc[e] = C.Wc[e];
}
struct StageBlocks {
  ContStore Fs[4];
  EulerBlocks Es[4];
};
QD void discrete_stages(const DeviceParams &p, const double *x, const double *u, double *xn, StageBlocks &W) {
  double k[12];
  if (!p.integrator) {  // quadrotor_model.cc:33-49
    ContBlocks F;
    continuous_with_blocks(p, x, u, k, F);
    compact(F, W.Fs[0]);
    euler_with_blocks(x, k, p.dt, xn, W.Es[3]);
    return;
  }
  const double half = p.dt / 2.0;
  double xdot[12];
#pragma unroll
  for (int e = 0; e < 12; ++e) { k[e] = 0.0; xdot[e] = 0.0; }
#pragma unroll 1
  for (int i = 0; i < 4; ++i) {
    const double dti = (i == 0) ? 0.0 : (i == 3 ? p.dt : half);
    const double ci = (i == 0 || i == 3) ? 1.0 / 6.0 : 2.0 / 6.0;
    double xi[13];
    // stage 0: x (+) 0 equals x up to the quaternion renormalisation of compose; its J_lhs = I, J_rhs = 0
    if (i == 0) euler_state(x, k, 0.0, xi);
    else euler_with_blocks(x, k, dti, xi, W.Es[i - 1]);
    ContBlocks F;
    continuous_with_blocks(p, xi, u, k, F);
    compact(F, W.Fs[i]);
#pragma unroll
    for (int e = 0; e < 12; ++e) xdot[e] = QFMA(ci, k[e], xdot[e]);
  }
  euler_with_blocks(x, xdot, p.dt, xn, W.Es[3]);
}
// One column strip of J_x (n = 3, s = 0..3) or all of J_u (n = 4) by the chain rule through the stages,
// on registers; element (r, j) of the strip is stored at out[(ld * r + j) * stride].
//   Euler: J = J_lhs + J_rhs J^c                                  (quadrotor_model.cc:42-45)
//   RK4:   dk_i = J^c_i (J_lhs_i + J_rhs_i dk_{i-1}) [+ J_u^c],  J = J_lhs + J_rhs sum_i c_i dk_i
template <int n>
QD void jacobian_strip(const DeviceParams &p, const StageBlocks &W, int s, double *out, int ld, int stride) {
  double D[12 * n], T[12 * n];
  ContBlocks F;
  expand(W.Fs[0], F);
  cont_strip<n>(p, F, s, D);
  if (!p.integrator) {
    euler_rhs_mul<n>(W.Es[3], D, T);
  } else {
    double S[12 * n];
    const double c0 = 1.0 / 6.0;
#pragma unroll
    for (int e = 0; e < 12 * n; ++e) S[e] = 0.0 + c0 * D[e];
#pragma unroll 1
    for (int i = 1; i < 4; ++i) {
      const double ci = (i == 3) ? 1.0 / 6.0 : 2.0 / 6.0;
      euler_rhs_mul<n>(W.Es[i - 1], D, T);
      if (n == 3) euler_lhs_add_strip(W.Es[i - 1], s, T);
      expand(W.Fs[i], F);
      cont_mul<n>(p, F, T, D);
      if (n == 4) {
#pragma unroll
        for (int e = 0; e < 16; ++e) D[32 + e] = D[32 + e] + p.JuC[e];
      }
#pragma unroll
      for (int e = 0; e < 12 * n; ++e) S[e] = QFMA(ci, D[e], S[e]);
    }
    euler_rhs_mul<n>(W.Es[3], S, T);
  }
  if (n == 3) euler_lhs_add_strip(W.Es[3], s, T);
#pragma unroll
  for (int r = 0; r < 12; ++r)
#pragma unroll
    for (int j = 0; j < n; ++j) out[(ld * r + j) * stride] = T[n * r + j];
}
// discrete_dynamics with differentials: xn [13], A = J_x at A[(12 r + c) stride], B = J_u at B[(4 r + c) stride]
// (either may be nullptr)
QD void discrete_with_jacobians(const DeviceParams &p, const double *x, const double *u, double *xn, double *A,
                                double *B, int stride) {
  StageBlocks W;
  discrete_stages(p, x, u, xn, W);
  if (A) {
#pragma unroll 1
    for (int s = 0; s < 4; ++s) jacobian_strip<3>(p, W, s, A + 3 * s * stride, 12, stride);
  }
  if (B) jacobian_strip<4>(p, W, 0, B, 4, stride);
}
// dense continuous J_x (API kernel)
QD void cont_jx_dense(const DeviceParams &p, const ContBlocks &F, double *J) {
  for (int e = 0; e < 144; ++e) J[e] = 0.0;
#pragma unroll
  for (int i = 0; i < 6; ++i) J[12 * i + 6 + i] = 1.0;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      J[12 * (6 + i) + 3 + j] = F.G[3 * i + j];
      J[12 * (9 + i) + 9 + j] = F.Wc[3 * i + j];
      if (p.coriolis) {
        J[12 * (6 + i) + 6 + j] = F.nW[3 * i + j];
        J[12 * (6 + i) + 9 + j] = F.V[3 * i + j];
      }
    }
}
#endif  // QILQR_USER_MODEL_TU

}  // namespace gm
}  // namespace qilqr
