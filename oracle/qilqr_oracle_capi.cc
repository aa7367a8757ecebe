// =============================================================================
// oracle/qilqr_oracle_capi.cc -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// Flat C entry points (for ctypes) over qilqr_oracle.hpp.  See that header for
// provenance and parity status.  Layouts:
//   state      13 doubles: t(3), quaternion (x,y,z,w), body velocity lin(3) ang(3)
//   traj point 18 doubles: time_s, state(13), control(4)   (trajectory.hh:9-14)
//   matrices   row-major
// =============================================================================
#include <atomic>
#include <cstdint>
#include <cstring>
#include <thread>

#include "qilqr_oracle.hpp"

using namespace qoracle;

extern "C" {
typedef struct {
  double mass_kg;
  double inertia[9];
  double arm_length_m;
  double torque_to_thrust_ratio_m;
  double g_mpss;
  double Q[144];
  double R[16];
  double dt_s;
  double step_update;
  double desired_reduction_frac;
  double rtol;
  double atol;
  double max_iters;  // a double in the reference too (ilqr_options.hh:14)
  double quu_regularization;
  int32_t ls_max_iters;
  int32_t populate_debug;
  int32_t symmetrize_vxx;
  int32_t model_kind;  // 0: the reference's QuadrotorModel; bit 0: RK4 integrator, bit 1: Coriolis term (QuadrotorModelVariant)
} qoracle_config_t;
}

namespace {
template <class T>
State<T> load_state(const double *s) {
  State<T> x;
  for (int i = 0; i < 3; ++i) x.inertial_from_body.t[i] = T(s[i]);
  x.inertial_from_body.q = Quat<T>{T(s[3]), T(s[4]), T(s[5]), T(s[6])};
  for (int i = 0; i < 6; ++i) x.body_velocity[i] = T(s[7 + i]);
  return x;
}
template <class T>
void store_state(const State<T> &x, double *s) {
  for (int i = 0; i < 3; ++i) s[i] = to_double(x.inertial_from_body.t[i]);
  s[3] = to_double(x.inertial_from_body.q.x);
  s[4] = to_double(x.inertial_from_body.q.y);
  s[5] = to_double(x.inertial_from_body.q.z);
  s[6] = to_double(x.inertial_from_body.q.w);
  for (int i = 0; i < 6; ++i) s[7 + i] = to_double(x.body_velocity[i]);
}
template <class T>
SE3<T> load_se3(const double *s) {
  SE3<T> X;
  for (int i = 0; i < 3; ++i) X.t[i] = T(s[i]);
  X.q = Quat<T>{T(s[3]), T(s[4]), T(s[5]), T(s[6])};
  return X;
}
template <class T>
void store_se3(const SE3<T> &X, double *s) {
  for (int i = 0; i < 3; ++i) s[i] = to_double(X.t[i]);
  s[3] = to_double(X.q.x); s[4] = to_double(X.q.y); s[5] = to_double(X.q.z); s[6] = to_double(X.q.w);
}
template <class T, int R, int C>
Mat<T, R, C> load_mat(const double *p) {
  Mat<T, R, C> m;
  for (int i = 0; i < R * C; ++i) m.a[i] = T(p[i]);
  return m;
}
template <class T, int R, int C>
void store_mat(const Mat<T, R, C> &m, double *p) {
  if (!p) return;
  for (int i = 0; i < R * C; ++i) p[i] = to_double(m.a[i]);
}
template <class T>
Trajectory<T> load_traj(const double *p, int n) {
  Trajectory<T> tr(n);
  for (int i = 0; i < n; ++i) {
    tr[i].time_s = T(p[18 * i]);
    tr[i].state = load_state<T>(p + 18 * i + 1);
    for (int j = 0; j < 4; ++j) tr[i].control[j] = T(p[18 * i + 14 + j]);
  }
  return tr;
}
template <class T>
void store_traj(const Trajectory<T> &tr, double *p) {
  for (size_t i = 0; i < tr.size(); ++i) {
    p[18 * i] = to_double(tr[i].time_s);
    store_state(tr[i].state, p + 18 * i + 1);
    for (int j = 0; j < 4; ++j) p[18 * i + 14 + j] = to_double(tr[i].control[j]);
  }
}
template <class T>
QuadrotorModel<T> make_model(const qoracle_config_t *c) {
  return QuadrotorModel<T>(T(c->mass_kg), load_mat<T, 3, 3>(c->inertia), T(c->arm_length_m),
                           T(c->torque_to_thrust_ratio_m), T(c->g_mpss));
}
ILQROptions make_options(const qoracle_config_t *c) {
  ILQROptions o;
  o.line_search_params = {c->step_update, c->desired_reduction_frac, c->ls_max_iters};
  o.convergence_criteria = {c->rtol, c->atol, c->max_iters};
  o.populate_debug = c->populate_debug != 0;
  o.symmetrize_vxx = c->symmetrize_vxx != 0;
  o.quu_regularization = c->quu_regularization;
  return o;
}
template <class T>
ILQR<T> make_ilqr(const qoracle_config_t *c, const double *desired, int n) {
  return ILQR<T>{make_model<T>(c),
                 CostFunction<T>{load_mat<T, 12, 12>(c->Q), load_mat<T, 4, 4>(c->R),
                                 load_traj<T>(desired, n)},
                 T(c->dt_s), make_options(c)};
}
template <class T>
QuadrotorModelVariant<T> make_variant(const qoracle_config_t *c) {
  return QuadrotorModelVariant<T>{make_model<T>(c), c->model_kind & 1, (c->model_kind & 2) != 0};
}
// f(solver) with the ILQR<ModelT> instantiation the configuration asks for
template <class T, class F>
void with_ilqr(const qoracle_config_t *c, const double *desired, int n, F &&f) {
  if (c->model_kind == 0) {
    f(make_ilqr<T>(c, desired, n));
  } else {
    f(ILQR<T, QuadrotorModelVariant<T>>{make_variant<T>(c),
                                        CostFunction<T>{load_mat<T, 12, 12>(c->Q), load_mat<T, 4, 4>(c->R),
                                                        load_traj<T>(desired, n)},
                                        T(c->dt_s), make_options(c)});
  }
}
template <class T, class F>
void with_model(const qoracle_config_t *c, F &&f) {
  if (c->model_kind == 0) f(make_model<T>(c));
  else f(make_variant<T>(c));
}
template <class T>
ControlUpdateTrajectory<T> load_update(const double *k, const double *K, int n) {
  ControlUpdateTrajectory<T> u(n);
  for (int i = 0; i < n; ++i) {
    u[i].ff_update = load_mat<T, 4, 1>(k + 4 * i);
    u[i].feedback = load_mat<T, 4, 12>(K + 48 * i);
  }
  return u;
}
template <class T>
void store_update(const ControlUpdateTrajectory<T> &u, double *k, double *K) {
  for (size_t i = 0; i < u.size(); ++i) {
    if (k) store_mat(u[i].ff_update, k + 4 * i);
    if (K) store_mat(u[i].feedback, K + 48 * i);
  }
}
}  // namespace

extern "C" {

// ---- Lie-group primitives (for scipy cross-checks) --------------------------
void qoracle_se3_exp(const double *tau, double *X) { store_se3(se3_exp(load_mat<double, 6, 1>(tau)), X); }
void qoracle_se3_log(const double *X, double *tau) { store_mat(se3_log(load_se3<double>(X)), tau); }
void qoracle_se3_compose(const double *A, const double *B, double *X) {
  store_se3(se3_compose(load_se3<double>(A), load_se3<double>(B)), X);
}
void qoracle_se3_inverse(const double *A, double *X) { store_se3(se3_inverse(load_se3<double>(A)), X); }
void qoracle_se3_adj(const double *A, double *Ad) { store_mat(se3_adj(load_se3<double>(A)), Ad); }
void qoracle_se3_rjac(const double *tau, double *J) { store_mat(se3_rjac(load_mat<double, 6, 1>(tau)), J); }
void qoracle_se3_ljac(const double *tau, double *J) { store_mat(se3_ljac(load_mat<double, 6, 1>(tau)), J); }
void qoracle_se3_rjacinv(const double *tau, double *J) { store_mat(se3_rjacinv(load_mat<double, 6, 1>(tau)), J); }
void qoracle_se3_ljacinv(const double *tau, double *J) { store_mat(se3_ljacinv(load_mat<double, 6, 1>(tau)), J); }
void qoracle_rotation_matrix(const double *q_xyzw, double *R) {
  store_mat(rotation_matrix(Quat<double>{q_xyzw[0], q_xyzw[1], q_xyzw[2], q_xyzw[3]}), R);
}
void qoracle_se3_plus(const double *X, const double *tau, double *out, double *J_X, double *J_tau) {
  Mat6<double> a, b;
  store_se3(se3_plus(load_se3<double>(X), load_mat<double, 6, 1>(tau), &a, &b), out);
  store_mat(a, J_X);
  store_mat(b, J_tau);
}
void qoracle_se3_minus(const double *A, const double *B, double *t, double *J_A, double *J_B) {
  Mat6<double> a, b;
  store_mat(se3_minus(load_se3<double>(A), load_se3<double>(B), &a, &b), t);
  store_mat(a, J_A);
  store_mat(b, J_B);
}
void qoracle_ldlt4_solve(const double *A, const double *b, int ncols, double *x) {
  LDLT4<double> f;
  f.compute(load_mat<double, 4, 4>(A));
  for (int c = 0; c < ncols; ++c) {
    Mat<double, 4, 1> rhs;
    for (int i = 0; i < 4; ++i) rhs[i] = b[i * ncols + c];
    const auto s = f.solve(rhs);
    for (int i = 0; i < 4; ++i) x[i * ncols + c] = s[i];
  }
}

// ---- model (quadrotor_model.cc) ---------------------------------------------
// returns 0, or 1 if the inertia matrix is rejected (quadrotor_model.cc:21-24)
int qoracle_check_model(const qoracle_config_t *c) {
  try { make_model<double>(c); } catch (const std::runtime_error &) { return 1; }
  return 0;
}
int qoracle_continuous_dynamics(const qoracle_config_t *c, const double *x, const double *u,
                                double *xdot, double *J_x, double *J_u) {
  try {
    with_model<double>(c, [&](const auto &m) {
      DynamicsDifferentials<double> d;
      const auto r = m.continuous_dynamics(load_state<double>(x), load_mat<double, 4, 1>(u),
                                           (J_x || J_u) ? &d : nullptr);
      store_mat(r.coeffs(), xdot);
      if (J_x) store_mat(d.J_x, J_x);
      if (J_u) store_mat(d.J_u, J_u);
    });
  } catch (const std::runtime_error &) { return 1; }
  return 0;
}
int qoracle_discrete_dynamics(const qoracle_config_t *c, const double *x, const double *u,
                              double dt_s, double *x_next, double *J_x, double *J_u) {
  try {
    with_model<double>(c, [&](const auto &m) {
      DynamicsDifferentials<double> d;
      const auto r = m.discrete_dynamics(load_state<double>(x), load_mat<double, 4, 1>(u), dt_s,
                                         (J_x || J_u) ? &d : nullptr);
      store_state(r, x_next);
      if (J_x) store_mat(d.J_x, J_x);
      if (J_u) store_mat(d.J_u, J_u);
    });
  } catch (const std::runtime_error &) { return 1; }
  return 0;
}
void qoracle_state_add(const double *x, const double *tangent, double *out, double *J_lhs, double *J_rhs) {
  BinaryStateFuncDiffs<double> d;
  const auto t = StateTangent<double>::from_coeffs(load_mat<double, 12, 1>(tangent));
  store_state(add(load_state<double>(x), t, (J_lhs || J_rhs) ? &d : nullptr), out);
  if (J_lhs) store_mat(d.J_x_lhs, J_lhs);
  if (J_rhs) store_mat(d.J_x_rhs, J_rhs);
}
void qoracle_state_minus(const double *lhs, const double *rhs, double *out, double *J_lhs, double *J_rhs) {
  BinaryStateFuncDiffs<double> d;
  const auto r = (J_lhs || J_rhs) ? minus(load_state<double>(lhs), load_state<double>(rhs), &d)
                                  : minus(load_state<double>(lhs), load_state<double>(rhs));
  store_mat(r.coeffs(), out);
  if (J_lhs) store_mat(d.J_x_lhs, J_lhs);
  if (J_rhs) store_mat(d.J_x_rhs, J_rhs);
}
void qoracle_euler_step(const double *x, const double *xdot, double dt_s, double *out, double *J_lhs,
                        double *J_rhs) {
  BinaryStateFuncDiffs<double> d;
  const auto t = StateTangent<double>::from_coeffs(load_mat<double, 12, 1>(xdot));
  store_state(euler_step(load_state<double>(x), t, dt_s, (J_lhs || J_rhs) ? &d : nullptr), out);
  if (J_lhs) store_mat(d.J_x_lhs, J_lhs);
  if (J_rhs) store_mat(d.J_x_rhs, J_rhs);
}

// ---- cost (cost.hh:36-61) -----------------------------------------------------
double qoracle_cost(const qoracle_config_t *c, const double *x, const double *u, const double *x_d,
                    const double *u_d, double *Cx, double *Cu, double *Cxx, double *Cuu, double *Cxu) {
  Trajectory<double> des(1);
  des[0].time_s = 0;
  des[0].state = load_state<double>(x_d);
  des[0].control = load_mat<double, 4, 1>(u_d);
  CostFunction<double> f{load_mat<double, 12, 12>(c->Q), load_mat<double, 4, 4>(c->R), des};
  CostDifferentials<double> d;
  const bool want = Cx || Cu || Cxx || Cuu || Cxu;
  const double cost = f(load_state<double>(x), load_mat<double, 4, 1>(u), 0, want ? &d : nullptr);
  if (want) {
    store_mat(d.x, Cx); store_mat(d.u, Cu); store_mat(d.xx, Cxx); store_mat(d.uu, Cuu); store_mat(d.xu, Cxu);
  }
  return cost;
}

// ---- solver pieces (ilqr.hh) --------------------------------------------------
// All take one problem: desired trajectory [n][18]; return 0 on success.
int qoracle_forward_sim(const qoracle_config_t *c, int n, const double *desired, const double *cur,
                        const double *k, const double *K, double alpha, double *out) {
  try {
    with_ilqr<double>(c, desired, n, [&](const auto &s) {
      store_traj(s.forward_sim(load_traj<double>(cur, n), load_update<double>(k, K, n), alpha), out);
    });
  } catch (const std::exception &) { return 1; }
  return 0;
}
int qoracle_cost_trajectory(const qoracle_config_t *c, int n_desired, const double *desired, int n,
                            const double *traj, double *cost) {
  try {
    const auto s = make_ilqr<double>(c, desired, n_desired);
    *cost = s.cost_trajectory(load_traj<double>(traj, n));
  } catch (const std::out_of_range &) { return 2; }  // cost.hh:39-40
  catch (const std::exception &) { return 1; }
  return 0;
}
int qoracle_backwards_pass(const qoracle_config_t *c, int n, const double *desired, const double *traj,
                           double *k, double *K, double *QuTk, double *kTQuuk) {
  try {
    with_ilqr<double>(c, desired, n, [&](const auto &s) {
      auto [upd, terms] = s.backwards_pass(load_traj<double>(traj, n));
      store_update(upd, k, K);
      *QuTk = terms.QuTk;
      *kTQuuk = terms.kTQuuk;
    });
  } catch (const std::exception &) { return 1; }
  return 0;
}
// returns 0 ok, 4 line search exhausted (ilqr.hh:191-193)
int qoracle_line_search(const qoracle_config_t *c, int n, const double *desired, const double *cur,
                        double cur_cost, const double *k, const double *K, double QuTk, double kTQuuk,
                        double *out, double *new_cost, double *step) {
  try {
    with_ilqr<double>(c, desired, n, [&](const auto &s) {
      CostReductionTerms t{QuTk, kTQuuk};
      auto r = s.line_search(load_traj<double>(cur, n), cur_cost, load_update<double>(k, K, n), t);
      store_traj(r.traj, out);
      *new_cost = r.cost;
      *step = r.step;
    });
  } catch (const LineSearchFailure &) { return 4; }
  catch (const std::exception &) { return 1; }
  return 0;
}

typedef struct {
  int32_t status;           // qoracle::Status
  int32_t backward_passes;  // number of backwards_pass calls
  int32_t rollouts;         // number of forward_sim calls
  int32_t num_debug;        // completed iterations (= ILQRDebug entries if populate_debug)
  double final_cost;
} qoracle_result_t;

// One solve.  cost_hist/step_hist: capacity ceil(max_iters) doubles (may be NULL);
// debug_traj: capacity ceil(max_iters)*n*18 (may be NULL; filled iff populate_debug);
// out_k/out_K: gains of the last backward pass (may be NULL).
int qoracle_solve(const qoracle_config_t *c, int n, const double *desired, const double *initial,
                  double *out_traj, double *out_k, double *out_K, double *cost_hist, double *step_hist,
                  double *debug_traj, qoracle_result_t *res) {
  try {
    with_ilqr<double>(c, desired, n, [&](const auto &s) {
      const auto r = s.solve(load_traj<double>(initial, n));
      store_traj(r.traj, out_traj);
      store_update(r.last_update, out_k, out_K);
      for (size_t i = 0; i < r.cost_history.size(); ++i) {
        if (cost_hist) cost_hist[i] = r.cost_history[i];
        if (step_hist) step_hist[i] = r.step_history[i];
      }
      if (debug_traj)
        for (size_t i = 0; i < r.debug.size(); ++i) store_traj(r.debug[i].trajectory, debug_traj + i * n * 18);
      res->status = r.status;
      res->backward_passes = r.backward_passes;
      res->rollouts = r.rollouts;
      res->num_debug = int(r.cost_history.size());
      res->final_cost = r.final_cost;
    });
  } catch (const std::exception &) { return 1; }
  return 0;
}

// Batch of independent solves over `nthreads` host threads (one problem per task).
// desired_stride = 0: one shared desired trajectory; n*18: one per problem.
// hist_cap: per-problem capacity of cost_hist (0/NULL to skip).
int qoracle_solve_batch(const qoracle_config_t *c, int batch, int n, const double *desired,
                        long desired_stride, const double *initial, double *out_traj, double *out_k,
                        double *out_K, double *cost_hist, int hist_cap, qoracle_result_t *res,
                        int nthreads) {
  if (nthreads < 1) nthreads = 1;
  std::atomic<int> next{0}, failed{0};
  auto work = [&]() {
    std::vector<double> ch(size_t(c->max_iters) + 2), sh(size_t(c->max_iters) + 2);
    qoracle_config_t cc = *c;
    cc.populate_debug = 0;
    for (;;) {
      const int b = next.fetch_add(1);
      if (b >= batch) break;
      const int rc = qoracle_solve(&cc, n, desired + size_t(b) * desired_stride,
                                   initial + size_t(b) * n * 18, out_traj + size_t(b) * n * 18,
                                   out_k ? out_k + size_t(b) * n * 4 : nullptr,
                                   out_K ? out_K + size_t(b) * n * 48 : nullptr, ch.data(), sh.data(),
                                   nullptr, &res[b]);
      if (rc) failed.fetch_add(1);
      if (cost_hist && hist_cap > 0)
        for (int i = 0; i < hist_cap; ++i)
          cost_hist[size_t(b) * hist_cap + i] = i < res[b].num_debug ? ch[i] : 0.0;
    }
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < nthreads; ++t) pool.emplace_back(work);
  work();
  for (auto &t : pool) t.join();
  return failed.load() ? 1 : 0;
}

// Dense as-written FLOP counts per knot:
// out[0] = backwards_pass per knot, out[1] = forward_sim per knot,
// out[2] = cost_trajectory per knot (as written, incl. the unused Jacobians),
// out[3] = cost per knot without the unused minus() Jacobians (cost.hh:42-43).
int qoracle_count_flops(const qoracle_config_t *c, int n, const double *desired, const double *traj,
                        double *out) {
  try {
    with_ilqr<CountedDouble>(c, desired, n, [&](const auto &s) {
    const auto tr = load_traj<CountedDouble>(traj, n);
    FlopCounter::reset();
    auto [upd, terms] = s.backwards_pass(tr);
    out[0] = double(FlopCounter::total()) / n;
    FlopCounter::reset();
    const auto nt = s.forward_sim(tr, upd, 1.0);
    out[1] = double(FlopCounter::total()) / n;
    FlopCounter::reset();
    (void)s.cost_trajectory(nt);
    out[2] = double(FlopCounter::total()) / n;
    FlopCounter::reset();
    for (int i = 0; i < n; ++i) {
      const auto dx = minus(nt[i].state, s.cost_function_.desired_trajectory_[i].state).coeffs();
      const auto du = nt[i].control - s.cost_function_.desired_trajectory_[i].control;
      (void)(((dx.transpose() * s.cost_function_.Q_) * dx)(0, 0) +
             ((du.transpose() * s.cost_function_.R_) * du)(0, 0));
    }
    out[3] = double(FlopCounter::total()) / n + 1;  // + the running sum
    FlopCounter::reset();
    });
  } catch (const std::exception &) { return 1; }
  return 0;
}

int qoracle_hardware_threads(void) { return int(std::thread::hardware_concurrency()); }

}  // extern "C"
