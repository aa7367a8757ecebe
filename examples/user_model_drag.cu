// Example of a user-supplied model for qilqr_set_user_model (include/qilqr.h): the reference's quadrotor
// (quadrotor_model.cc:33-122, explicit Euler on SE(3) x R^6, diagonal inertia) plus a linear drag -c_d * v on the
// body-frame linear velocity -- a dynamics function the library does not ship.  The solver sees it only through
// discrete_dynamics(x, u, dt, diffs*) with dense Jacobians, as ILQR<ModelT> does (ilqr.hh:110-112).
//
// params: [0] mass, [1..3] Ixx Iyy Izz, [4] arm length, [5] torque-to-thrust ratio, [6] g, [7] c_d
// The Lie-group helpers (qilqr_device.cuh) are in scope: se3_plus_blocks, quat_to_rot, quat_compose, m3_vec, ...
using namespace qilqr;

extern "C" __device__ void qilqr_user_discrete_dynamics(const double *params, const double *x, const double *u,
                                                        double dt, double *x_next, double *J_x, double *J_u) {
  const double m = params[0], I[3] = {params[1], params[2], params[3]}, a = params[4], r = params[5], g = params[6],
               cd = params[7];
  const double *q = x + 3, *v = x + 7, *w = x + 10;
  double R[9];
  quat_to_rot(q, R);
  // body acceleration: -g R^T e_z + (sum u / m) e_z - c_d v ;  I^-1 (M - w x I w)
  const double usum = ((u[0] + u[1]) + u[2]) + u[3];
  const double M[3] = {a * (u[3] - u[1]), a * (u[0] - u[2]), r * ((u[1] - u[0]) + (u[3] - u[2]))};
  const double Iw[3] = {I[0] * w[0], I[1] * w[1], I[2] * w[2]};
  double acc[6];
  acc[0] = -g * R[6] - cd * v[0];
  acc[1] = -g * R[7] - cd * v[1];
  acc[2] = -g * R[8] + usum / m - cd * v[2];
  acc[3] = (M[0] - (w[1] * Iw[2] - w[2] * Iw[1])) / I[0];
  acc[4] = (M[1] - (w[2] * Iw[0] - w[0] * Iw[2])) / I[1];
  acc[5] = (M[2] - (w[0] * Iw[1] - w[1] * Iw[0])) / I[2];
  // pose+ = pose o Exp(dt * body velocity), velocity+ = velocity + dt * acceleration
  double tau[6], Re[9], Te[9], Jr[9], Qb[9], te[3], qe[4], Rt[3];
  for (int i = 0; i < 6; ++i) tau[i] = dt * x[7 + i];
  se3_plus_blocks(tau, Re, Te, Jr, Qb, te, qe);
  m3_vec(R, te, Rt);
  for (int i = 0; i < 3; ++i) x_next[i] = Rt[i] + x[i];
  quat_compose(q, qe, x_next + 3);
  for (int i = 0; i < 6; ++i) x_next[7 + i] = x[7 + i] + dt * acc[i];
  if (J_x) {
    for (int e = 0; e < 144; ++e) J_x[e] = 0.0;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        J_x[12 * i + j] = Re[3 * i + j];                 // d pose+ / d pose = Ad(Exp(tau)^-1)
        J_x[12 * i + 3 + j] = Te[3 * i + j];
        J_x[12 * (3 + i) + 3 + j] = Re[3 * i + j];
        J_x[12 * i + 6 + j] = dt * Jr[3 * i + j];         // d pose+ / d velocity = dt Jr(tau)
        J_x[12 * i + 9 + j] = dt * Qb[3 * i + j];
        J_x[12 * (3 + i) + 9 + j] = dt * Jr[3 * i + j];
      }
    // d v_lin+ / d attitude = dt * (-g hat(R^T e_z)); d v_lin+ / d v_lin = (1 - dt c_d) I
    const double gz[3] = {-g * R[6], -g * R[7], -g * R[8]};
    const double G[9] = {0.0, -gz[2], gz[1], gz[2], 0.0, -gz[0], -gz[1], gz[0], 0.0};
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) J_x[12 * (6 + i) + 3 + j] = dt * G[3 * i + j];
    for (int i = 0; i < 3; ++i) J_x[12 * (6 + i) + 6 + i] = 1.0 - dt * cd;
    // d w+ / d w = I + dt * (-I^-1 (hat(w) I - hat(I w)))
    const double HI[9] = {0.0, -w[2] * I[1], w[1] * I[2], w[2] * I[0], 0.0, -w[0] * I[2], -w[1] * I[0], w[0] * I[1], 0.0};
    const double HIw[9] = {0.0, -Iw[2], Iw[1], Iw[2], 0.0, -Iw[0], -Iw[1], Iw[0], 0.0};
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
        J_x[12 * (9 + i) + 9 + j] = ((i == j) ? 1.0 : 0.0) - dt * (HI[3 * i + j] - HIw[3 * i + j]) / I[i];
  }
  if (J_u) {
    for (int e = 0; e < 48; ++e) J_u[e] = 0.0;
    const double arms[12] = {0.0, -a, 0.0, a, a, 0.0, -a, 0.0, -r, r, -r, r};  // quadrotor_model.cc:15-18
    for (int j = 0; j < 4; ++j) {
      J_u[4 * 8 + j] = dt / m;
      for (int i = 0; i < 3; ++i) J_u[4 * (9 + i) + j] = dt * arms[4 * i + j] / I[i];
    }
  }
}
