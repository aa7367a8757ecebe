// =============================================================================
// qilqr_api_kernels.cuh -- batched single-call versions of the reference's model
// and cost interfaces with their DENSE Jacobian outputs (array-of-structs in and
// out, one thread per problem).  These serve the drop-in API and the parity
// tests of the device Lie library; the solver kernels never materialise dense
// Jacobians.
// =============================================================================
#pragma once
#include "qilqr_device.cuh"
#include "qilqr_model_generic.cuh"

namespace qilqr {

QD void zero_fill(double *p, int n) {
  for (int i = 0; i < n; ++i) p[i] = 0.0;
}
QD void put_block(double *M, int ld, int r0, int c0, const double *blk, double scale = 1.0) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) M[(r0 + i) * ld + c0 + j] = scale * blk[3 * i + j];
}

// QuadrotorModel::continuous_dynamics (quadrotor_model.cc:65-122)
__global__ void k_api_continuous_dynamics(const __grid_constant__ DeviceParams p, int B, const double *x,
                                          const double *u, double *xdot, double *J_x, double *J_u) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const double *xs = x + size_t(b) * 13, *us = u + size_t(b) * 4;
  double *o = xdot + size_t(b) * 12;
  if (p.coriolis) {  // model variant (qilqr_model_generic.cuh)
    double xl[13], k[12];
    for (int i = 0; i < 13; ++i) xl[i] = xs[i];
    gm::ContBlocks F;
    gm::continuous_with_blocks(p, xl, us, k, F);
    for (int i = 0; i < 12; ++i) o[i] = k[i];
    if (J_x) gm::cont_jx_dense(p, F, J_x + size_t(b) * 144);
  } else {
    double R[9], acc[6];
    quat_to_rot(xs + 3, R);
    body_acceleration(p, R, xs + 7, us, acc);
    for (int i = 0; i < 6; ++i) { o[i] = xs[7 + i]; o[6 + i] = acc[i]; }
    if (J_x) {
      double *J = J_x + size_t(b) * 144;
      zero_fill(J, 144);
      for (int i = 0; i < 6; ++i) J[i * 12 + 6 + i] = 1.0;
      double gz[3], Wc[9];
      continuous_blocks(p, xs + 3, xs + 7, gz, Wc);
      const double G[9] = {0.0, -gz[2], gz[1], gz[2], 0.0, -gz[0], -gz[1], gz[0], 0.0};
      put_block(J, 12, 6, 3, G);
      put_block(J, 12, 9, 9, Wc);
    }
  }
  if (J_u) {
    double *J = J_u + size_t(b) * 48;
    zero_fill(J, 48);
    for (int e = 0; e < 16; ++e) J[32 + e] = p.JuC[e];
  }
}

// QuadrotorModel::discrete_dynamics (quadrotor_model.cc:33-49)
__global__ void k_api_discrete_dynamics(const __grid_constant__ DeviceParams p, int B, const double *x,
                                        const double *u, double *x_next, double *J_x, double *J_u) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double xs[13], us[4];
  for (int i = 0; i < 13; ++i) xs[i] = x[size_t(b) * 13 + i];
  for (int i = 0; i < 4; ++i) us[i] = u[size_t(b) * 4 + i];
  if (p.integrator || p.coriolis) {  // model variant (qilqr_model_generic.cuh)
    if (J_x || J_u) {
      double xn[13];
      gm::discrete_with_jacobians(p, xs, us, xn, J_x ? J_x + size_t(b) * 144 : nullptr,
                                  J_u ? J_u + size_t(b) * 48 : nullptr, 1);
    }
    gm::discrete_step_any(p, xs, us);
    for (int i = 0; i < 13; ++i) x_next[size_t(b) * 13 + i] = xs[i];
    return;
  }
  if (J_x) {
    ABlocks A;
    dynamics_blocks(p, xs + 3, xs + 7, A);
    double *J = J_x + size_t(b) * 144;
    zero_fill(J, 144);
    put_block(J, 12, 0, 0, A.Re);
    put_block(J, 12, 0, 3, A.Te);
    put_block(J, 12, 0, 6, A.dJr);
    put_block(J, 12, 0, 9, A.dQb);
    put_block(J, 12, 3, 3, A.Re);
    put_block(J, 12, 3, 9, A.dJr);
    put_block(J, 12, 6, 3, A.dG);
    for (int i = 0; i < 3; ++i) J[(6 + i) * 12 + 6 + i] = 1.0;
    put_block(J, 12, 9, 9, A.Wd);
  }
  if (J_u) {
    double *J = J_u + size_t(b) * 48;
    zero_fill(J, 48);
    for (int e = 0; e < 16; ++e) J[32 + e] = p.Bu[e];
  }
  discrete_step(p, xs, xs + 3, xs + 7, us);
  for (int i = 0; i < 13; ++i) x_next[size_t(b) * 13 + i] = xs[i];
}

// minus(State, State, diffs) (quadrotor_model.cc:215-250)
__global__ void k_api_state_minus(int B, const double *lhs, const double *rhs, double *out, double *J_lhs,
                                  double *J_rhs) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double l[13], r[13], d[12], Jli[9];
  for (int i = 0; i < 13; ++i) { l[i] = lhs[size_t(b) * 13 + i]; r[i] = rhs[size_t(b) * 13 + i]; }
  Angle ang;
  state_minus(l, r, d, Jli, &ang);
  for (int i = 0; i < 12; ++i) out[size_t(b) * 12 + i] = d[i];
  if (J_lhs) {
    double Ji[9], Qi[9];
    se3_rjacinv_blocks(d, Jli, ang, Ji, Qi);
    double *J = J_lhs + size_t(b) * 144;
    zero_fill(J, 144);
    put_block(J, 12, 0, 0, Ji);
    put_block(J, 12, 0, 3, Qi);
    put_block(J, 12, 3, 3, Ji);
    for (int i = 6; i < 12; ++i) J[i * 12 + i] = 1.0;
  }
  if (J_rhs) {
    // -Jl^-1(t) = -[[Jli, -Jli Q(t) Jli], [0, Jli]]
    double Qm[9], T[9], Ql[9];
    se3_fillQ(d, d + 3, Qm);
    m3_mul(Jli, Qm, T);
    m3_mul(T, Jli, Ql);  // Jli Q Jli ; the block is -(that), and J_rhs negates again
    double *J = J_rhs + size_t(b) * 144;
    zero_fill(J, 144);
    put_block(J, 12, 0, 0, Jli, -1.0);
    put_block(J, 12, 0, 3, Ql, 1.0);
    put_block(J, 12, 3, 3, Jli, -1.0);
    for (int i = 6; i < 12; ++i) J[i * 12 + i] = -1.0;
  }
}

// add(State, StateTangent, diffs) (quadrotor_model.cc:174-206)
__global__ void k_api_state_add(int B, const double *x, const double *tangent, double *out, double *J_lhs,
                                double *J_rhs) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double xs[13], tg[12];
  for (int i = 0; i < 13; ++i) xs[i] = x[size_t(b) * 13 + i];
  for (int i = 0; i < 12; ++i) tg[i] = tangent[size_t(b) * 12 + i];
  double Re[9], Te[9], Jr[9], Qb[9], te[3], qe[4];
  se3_plus_blocks(tg, Re, Te, Jr, Qb, te, qe);
  double R[9], Rt[3], qn[4];
  quat_to_rot(xs + 3, R);
  m3_vec(R, te, Rt);
  quat_compose(xs + 3, qe, qn);
  double *o = out + size_t(b) * 13;
  for (int i = 0; i < 3; ++i) o[i] = Rt[i] + xs[i];
  for (int i = 0; i < 4; ++i) o[3 + i] = qn[i];
  for (int i = 0; i < 6; ++i) o[7 + i] = xs[7 + i] + tg[6 + i];
  if (J_lhs) {
    double *J = J_lhs + size_t(b) * 144;
    zero_fill(J, 144);
    put_block(J, 12, 0, 0, Re);
    put_block(J, 12, 0, 3, Te);
    put_block(J, 12, 3, 3, Re);
    for (int i = 6; i < 12; ++i) J[i * 12 + i] = 1.0;
  }
  if (J_rhs) {
    double *J = J_rhs + size_t(b) * 144;
    zero_fill(J, 144);
    put_block(J, 12, 0, 0, Jr);
    put_block(J, 12, 0, 3, Qb);
    put_block(J, 12, 3, 3, Jr);
    for (int i = 6; i < 12; ++i) J[i * 12 + i] = 1.0;
  }
}

}  // namespace qilqr
