// =============================================================================
// quadrotor_ilqr.hpp -- header-only C++17 host mirror of the reference's C++ interface for the
// hot path, implemented entirely on top of the C ABI (include/qilqr.h -> libqilqr_b200.so):
//
//   reference (namespace src)                         here (namespace qilqr)
//   QuadrotorModel            quadrotor_model.hh:7-67   qilqr::QuadrotorModel
//   add / minus               quadrotor_model.hh:86-103 qilqr::add / qilqr::minus
//   TrajectoryPoint/Trajectory trajectory.hh:9-24       qilqr::TrajectoryPoint / Trajectory
//   ILQROptions & friends     ilqr_options.hh:4-22      qilqr::LineSearchParams / ConvergenceCriteria / ILQROptions
//   ILQRIterDebug / ILQRDebug ilqr_debug.hh:9-22        qilqr::ILQRIterDebug / ILQRDebug
//   CostFunction<ModelT>      cost.hh:10-66             qilqr::CostFunction<ModelT>
//   ILQR<ModelT>              ilqr.hh:25-206            qilqr::ILQR<ModelT>  (+ solve_batch)
//
// Same member names, argument meaning and exceptions (std::runtime_error for a non-PD inertia and
// for line-search exhaustion, std::out_of_range when a trajectory is longer than the desired one).
// Eigen/manif types are replaced by plain std::array aggregates: the arithmetic lives in the CUDA
// library, not here.  Matrices are row-major.
// =============================================================================
#pragma once
#include <array>
#include <cmath>
#include <map>
#include <memory>
#include <ostream>
#include <stdexcept>
#include <tuple>
#include <string>
#include <utility>
#include <vector>

#include "../../include/qilqr.h"

namespace qilqr {

using Vec3 = std::array<double, 3>;
using Vec4 = std::array<double, 4>;
using Vec6 = std::array<double, 6>;
using Vec12 = std::array<double, 12>;
using Mat3 = std::array<double, 9>;
using Mat4 = std::array<double, 16>;
using Mat12 = std::array<double, 144>;
using Mat12x4 = std::array<double, 48>;
using Mat4x12 = std::array<double, 48>;

inline Mat3 Identity3() { return {1, 0, 0, 0, 1, 0, 0, 0, 1}; }
inline Mat4 Identity4() {
  Mat4 m{};
  for (int i = 0; i < 4; ++i) m[5 * i] = 1.0;
  return m;
}
inline Mat12 Identity12() {
  Mat12 m{};
  for (int i = 0; i < 12; ++i) m[13 * i] = 1.0;
  return m;
}

// manif::SE3d stand-in: translation + unit quaternion in (x, y, z, w) order.
struct SE3 {
  Vec3 translation{0, 0, 0};
  Vec4 quaternion{0, 0, 0, 1};
  static SE3 Identity() { return SE3{}; }
};

inline void check(int rc, qilqr_solver_t *h = nullptr) {
  if (rc == QILQR_OK) return;
  std::string msg = qilqr_error_string(rc);
  if (h && *qilqr_last_error_message(h)) msg += std::string(": ") + qilqr_last_error_message(h);
  if (rc == QILQR_ERR_OUT_OF_RANGE) throw std::out_of_range(msg);
  throw std::runtime_error(msg);
}

struct LineSearchParams {  // ilqr_options.hh:4-8
  double step_update;
  double desired_reduction_frac;
  int max_iters;
};
struct ConvergenceCriteria {  // ilqr_options.hh:11-15
  double rtol;
  double atol;
  double max_iters;
};
struct ILQROptions {  // ilqr_options.hh:18-22 (+ extensions, defaults = reference behaviour)
  LineSearchParams line_search_params{0.5, 0.5, 100};
  ConvergenceCriteria convergence_criteria{1e-12, 1e-12, 100};
  bool populate_debug = false;
  bool symmetrize_vxx = false;
  int num_parallel_alphas = 1;
  double quu_regularization = 0.0;
};
inline bool operator==(const LineSearchParams &a, const LineSearchParams &b) {
  return a.step_update == b.step_update && a.desired_reduction_frac == b.desired_reduction_frac && a.max_iters == b.max_iters;
}
inline bool operator==(const ConvergenceCriteria &a, const ConvergenceCriteria &b) {
  return a.rtol == b.rtol && a.atol == b.atol && a.max_iters == b.max_iters;
}
// (the reference's version ends in `lhs.populate_debug && rhs.populate_debug`, ilqr_options.cc:17-21, which
// makes two option sets with populate_debug == false unequal; that slip is not reproduced)
inline bool operator==(const ILQROptions &a, const ILQROptions &b) {
  return a.line_search_params == b.line_search_params && a.convergence_criteria == b.convergence_criteria &&
         a.populate_debug == b.populate_debug && a.symmetrize_vxx == b.symmetrize_vxx &&
         a.num_parallel_alphas == b.num_parallel_alphas && a.quu_regularization == b.quu_regularization;
}
inline qilqr_options_t to_c(const ILQROptions &o) {
  return qilqr_options_t{o.line_search_params.step_update, o.line_search_params.desired_reduction_frac,
                         o.line_search_params.max_iters, o.populate_debug ? 1 : 0, o.convergence_criteria.rtol,
                         o.convergence_criteria.atol, o.convergence_criteria.max_iters, o.symmetrize_vxx ? 1 : 0,
                         o.num_parallel_alphas, o.quu_regularization};
}

// RAII handle on a solver of the CUDA library.
struct Handle {
  qilqr_solver_t *h = nullptr;
  Handle(const qilqr_model_t &m, const Mat12 &Q, const Mat4 &R, double dt_s, const ILQROptions &o, int device = 0,
         int model_flags = 0) {
    const qilqr_options_t co = to_c(o);
    check(qilqr_create(&m, Q.data(), R.data(), dt_s, &co, device, &h));
    if (model_flags) check(qilqr_set_model_variant(h, model_flags));
  }
  ~Handle() { qilqr_destroy(h); }
  Handle(const Handle &) = delete;
  Handle &operator=(const Handle &) = delete;
};

struct QuadrotorModel {
  static constexpr int CONFIG_DIM = 6;
  static constexpr int STATE_DIM = 12;
  static constexpr int CONTROL_DIM = 4;
  struct State {
    SE3 inertial_from_body;
    Vec6 body_velocity{};
  };
  struct StateTangent {  // quadrotor_model.hh:18-28
    Vec6 body_velocity{};
    Vec6 body_acceleration{};
    Vec12 coeffs() const {
      Vec12 c{};
      for (int i = 0; i < 6; ++i) { c[i] = body_velocity[i]; c[6 + i] = body_acceleration[i]; }
      return c;
    }
    double &operator[](int i) { return i < CONFIG_DIM ? body_velocity[i] : body_acceleration[i - CONFIG_DIM]; }
    const double &operator[](int i) const { return i < CONFIG_DIM ? body_velocity[i] : body_acceleration[i - CONFIG_DIM]; }
    static StateTangent Zero() { return StateTangent{}; }
  };
  using StateJacobian = Mat12;
  using Control = Vec4;
  using ControlJacobian = Mat12x4;
  struct DynamicsDifferentials {
    StateJacobian J_x;
    ControlJacobian J_u;
  };
  struct BinaryStateFuncDiffs {
    StateJacobian J_x_lhs;
    StateJacobian J_x_rhs;
  };

  double mass_kg_;
  Mat3 inertia_;
  double arm_length_m_;
  double torque_to_thrust_ratio_m_;
  double g_mpss_;
  int model_flags_ = 0;  // QILQR_MODEL_*: 0 = the reference's model (see QuadrotorModelVariant below)

  QuadrotorModel(double mass_kg, const Mat3 &inertia, double arm_length_m, double torque_to_thrust_ratio_m,
                 double g_mpss = 9.81)
      : mass_kg_(mass_kg), inertia_(inertia), arm_length_m_(arm_length_m),
        torque_to_thrust_ratio_m_(torque_to_thrust_ratio_m), g_mpss_(g_mpss) {
    const qilqr_model_t m = c_model();
    if (qilqr_check_model(&m) != QILQR_OK)
      throw std::runtime_error("Inertia matrix is not positive definite!");  // quadrotor_model.cc:21-24
  }
  // hook run on every solver handle created for this model (see UserModel)
  void configure(qilqr_solver_t *) const {}
  qilqr_model_t c_model() const {
    qilqr_model_t m{};
    m.mass_kg = mass_kg_;
    for (int i = 0; i < 9; ++i) m.inertia[i] = inertia_[i];
    m.arm_length_m = arm_length_m_;
    m.torque_to_thrust_ratio_m = torque_to_thrust_ratio_m_;
    m.g_mpss = g_mpss_;
    return m;
  }

  static void pack(const State &x, double *p) {
    for (int i = 0; i < 3; ++i) p[i] = x.inertial_from_body.translation[i];
    for (int i = 0; i < 4; ++i) p[3 + i] = x.inertial_from_body.quaternion[i];
    for (int i = 0; i < 6; ++i) p[7 + i] = x.body_velocity[i];
  }
  static State unpack(const double *p) {
    State x;
    for (int i = 0; i < 3; ++i) x.inertial_from_body.translation[i] = p[i];
    for (int i = 0; i < 4; ++i) x.inertial_from_body.quaternion[i] = p[3 + i];
    for (int i = 0; i < 6; ++i) x.body_velocity[i] = p[7 + i];
    return x;
  }

  // quadrotor_model.cc:33-49
  State discrete_dynamics(const State &x, const Control &u, double dt_s, DynamicsDifferentials *diffs = nullptr) const {
    double xs[13], xn[13];
    pack(x, xs);
    check(qilqr_discrete_dynamics_host(handle(dt_s), 1, xs, u.data(), xn, diffs ? diffs->J_x.data() : nullptr,
                                       diffs ? diffs->J_u.data() : nullptr));
    return unpack(xn);
  }
  // quadrotor_model.cc:65-122
  StateTangent continuous_dynamics(const State &x, const Control &u, DynamicsDifferentials *diffs = nullptr) const {
    double xs[13], xd[12];
    pack(x, xs);
    check(qilqr_continuous_dynamics_host(handle(0.1), 1, xs, u.data(), xd, diffs ? diffs->J_x.data() : nullptr,
                                         diffs ? diffs->J_u.data() : nullptr));
    StateTangent t;
    for (int i = 0; i < 12; ++i) t[i] = xd[i];
    return t;
  }
  // a solver handle for model-level calls at step dt (identity cost, default options), cached per dt
  qilqr_solver_t *handle(double dt_s) const {
    auto it = handles_.find(dt_s);
    if (it == handles_.end())
      it = handles_.emplace(dt_s, std::make_shared<Handle>(c_model(), Identity12(), Identity4(), dt_s, ILQROptions{}, 0,
                                                           model_flags_)).first;
    return it->second->h;
  }

 private:
  mutable std::map<double, std::shared_ptr<Handle>> handles_;
};

// A second model behind the ModelT concept of ilqr.hh:25-44 (not in the reference): the same state
// manifold with an RK4 integrator (the scheme commented out at quadrotor_model.cc:51-63) and/or the
// Coriolis term -omega x v.  ILQR<QuadrotorModelVariant> runs on the model-agnostic CUDA kernels.
struct QuadrotorModelVariant : QuadrotorModel {
  QuadrotorModelVariant(double mass_kg, const Mat3 &inertia, double arm_length_m, double torque_to_thrust_ratio_m,
                        double g_mpss, bool rk4, bool coriolis)
      : QuadrotorModel(mass_kg, inertia, arm_length_m, torque_to_thrust_ratio_m, g_mpss) {
    model_flags_ = (rk4 ? QILQR_MODEL_RK4 : 0) | (coriolis ? QILQR_MODEL_CORIOLIS : 0);
    if (!model_flags_) model_flags_ = QILQR_MODEL_GENERIC;  // the reference dynamics on the generic kernels
  }
};

// A model SUPPLIED BY THE CALLER behind the ModelT concept (ilqr.hh:25-44): CUDA C++ source defining
// qilqr_user_discrete_dynamics (include/qilqr.h, qilqr_set_user_model), compiled at run time and run on the
// model-agnostic kernels.  `base` provides the State / Control types and the constants the cost's minus() needs;
// its own dynamics are not used by ILQR<UserModel>::solve / forward_sim / backwards_pass.
struct UserModel : QuadrotorModel {
  std::string cuda_source_;
  std::vector<double> params_;
  UserModel(const QuadrotorModel &base, std::string cuda_source, std::vector<double> params)
      : QuadrotorModel(base), cuda_source_(std::move(cuda_source)), params_(std::move(params)) {}
  void configure(qilqr_solver_t *h) const {
    check(qilqr_set_user_model(h, cuda_source_.c_str(), params_.empty() ? nullptr : params_.data(), int(params_.size())), h);
  }
};

namespace detail {
inline qilqr_solver_t *default_handle() {  // for the model-independent Lie operations add / minus
  static QuadrotorModel m(1.0, Identity3(), 1.0, 0.0, 9.81);
  return m.handle(0.1);
}
}  // namespace detail

// add(State, StateTangent, diffs) -- quadrotor_model.cc:174-206
inline QuadrotorModel::State add(const QuadrotorModel::State &x, const QuadrotorModel::StateTangent &tangent,
                                 QuadrotorModel::BinaryStateFuncDiffs *diffs = nullptr) {
  double xs[13], out[13];
  QuadrotorModel::pack(x, xs);
  const Vec12 t = tangent.coeffs();
  check(qilqr_state_add_host(detail::default_handle(), 1, xs, t.data(), out, diffs ? diffs->J_x_lhs.data() : nullptr,
                             diffs ? diffs->J_x_rhs.data() : nullptr));
  return QuadrotorModel::unpack(out);
}
inline QuadrotorModel::State operator+(const QuadrotorModel::State &x, const QuadrotorModel::StateTangent &t) {
  return add(x, t);
}
// minus(State, State, diffs) -- quadrotor_model.cc:215-250
inline QuadrotorModel::StateTangent minus(const QuadrotorModel::State &lhs, const QuadrotorModel::State &rhs,
                                          QuadrotorModel::BinaryStateFuncDiffs *diffs = nullptr) {
  double a[13], b[13], out[12];
  QuadrotorModel::pack(lhs, a);
  QuadrotorModel::pack(rhs, b);
  check(qilqr_state_minus_host(detail::default_handle(), 1, a, b, out, diffs ? diffs->J_x_lhs.data() : nullptr,
                               diffs ? diffs->J_x_rhs.data() : nullptr));
  QuadrotorModel::StateTangent t;
  for (int i = 0; i < 12; ++i) t[i] = out[i];
  return t;
}
inline QuadrotorModel::StateTangent operator-(const QuadrotorModel::State &lhs, const QuadrotorModel::State &rhs) {
  return minus(lhs, rhs);
}

template <class ModelT>
struct TrajectoryPoint {  // trajectory.hh:9-14
  double time_s;
  typename ModelT::State state;
  typename ModelT::Control control;
};
template <class ModelT>
using Trajectory = std::vector<TrajectoryPoint<ModelT>>;

template <class ModelT>
struct ILQRIterDebug {  // ilqr_debug.hh:9-13
  Trajectory<ModelT> trajectory;
  double cost;
};
template <class ModelT>
using ILQRDebug = std::vector<ILQRIterDebug<ModelT>>;

// Equality and printing, as the reference defines them (quadrotor_model.cc:252-263, trajectory.hh:16-44,
// ilqr_debug.hh:15-19).  manif compares group elements and tangents approximately (tolerance
// Constants<double>::eps = 1e-14); here: coefficient-wise within 1e-14 (relative to max(1, |value|)), a
// unit quaternion and its negative being the same rotation.
namespace detail {
inline bool approx(double a, double b) { return std::fabs(a - b) <= 1e-14 * std::fmax(1.0, std::fmax(std::fabs(a), std::fabs(b))); }
template <size_t N>
bool approx(const std::array<double, N> &a, const std::array<double, N> &b, double sign = 1.0) {
  for (size_t i = 0; i < N; ++i)
    if (!approx(a[i], sign * b[i])) return false;
  return true;
}
template <size_t N>
void print(std::ostream &out, const std::array<double, N> &v) {
  for (size_t i = 0; i < N; ++i) out << (i ? " " : "") << v[i];
}
}  // namespace detail
inline bool operator==(const SE3 &a, const SE3 &b) {
  return detail::approx(a.translation, b.translation) &&
         (detail::approx(a.quaternion, b.quaternion) || detail::approx(a.quaternion, b.quaternion, -1.0));
}
inline bool operator==(const QuadrotorModel::State &lhs, const QuadrotorModel::State &rhs) {
  return detail::approx(lhs.body_velocity, rhs.body_velocity) && lhs.inertial_from_body == rhs.inertial_from_body;
}
inline bool operator!=(const QuadrotorModel::State &lhs, const QuadrotorModel::State &rhs) { return !(lhs == rhs); }
inline std::ostream &operator<<(std::ostream &out, const QuadrotorModel::State &state) {
  out << "inertial_from_body: ";
  detail::print(out, state.inertial_from_body.translation);
  out << " ";
  detail::print(out, state.inertial_from_body.quaternion);
  out << ", body velocity: ";
  detail::print(out, state.body_velocity);
  return out;
}
template <class ModelT>
bool operator==(const TrajectoryPoint<ModelT> &lhs, const TrajectoryPoint<ModelT> &rhs) {
  return lhs.time_s == rhs.time_s && lhs.control == rhs.control && lhs.state == rhs.state;
}
template <class ModelT>
bool operator!=(const TrajectoryPoint<ModelT> &lhs, const TrajectoryPoint<ModelT> &rhs) { return !(lhs == rhs); }
template <class ModelT>
std::ostream &operator<<(std::ostream &out, const TrajectoryPoint<ModelT> &pt) {
  out << "{\n\ttime_s: " << pt.time_s << ",\n\tstate: " << pt.state << ",\n\tcontrol:";
  detail::print(out, pt.control);
  return out << "\n}";
}
template <class ModelT>
std::ostream &operator<<(std::ostream &out, const Trajectory<ModelT> &traj) {
  out << "{\n";
  for (const auto &pt : traj) out << pt << ",\n";
  return out << "}";
}
template <class ModelT>
bool operator==(const ILQRIterDebug<ModelT> &lhs, const ILQRIterDebug<ModelT> &rhs) {
  return lhs.trajectory == rhs.trajectory && lhs.cost == rhs.cost;
}
template <class ModelT>
bool operator!=(const ILQRIterDebug<ModelT> &lhs, const ILQRIterDebug<ModelT> &rhs) { return !(lhs == rhs); }

template <class ModelT>
std::vector<double> flatten(const Trajectory<ModelT> &t) {
  std::vector<double> out(t.size() * 18);
  for (size_t i = 0; i < t.size(); ++i) {
    out[18 * i] = t[i].time_s;
    ModelT::pack(t[i].state, &out[18 * i + 1]);
    for (int j = 0; j < 4; ++j) out[18 * i + 14 + j] = t[i].control[j];
  }
  return out;
}
template <class ModelT>
Trajectory<ModelT> unflatten(const double *p, size_t n) {
  Trajectory<ModelT> t(n);
  for (size_t i = 0; i < n; ++i) {
    t[i].time_s = p[18 * i];
    t[i].state = ModelT::unpack(p + 18 * i + 1);
    for (int j = 0; j < 4; ++j) t[i].control[j] = p[18 * i + 14 + j];
  }
  return t;
}

template <class ModelT>
class CostFunction {  // cost.hh:10-66
 public:
  using CostJacobianState = Vec12;
  using CostJacobianControl = Vec4;
  using CostHessianStateState = Mat12;
  using CostHessianControlControl = Mat4;
  using CostHessianStateControl = Mat12x4;
  struct CostDifferentials {
    CostJacobianState x;
    CostJacobianControl u;
    CostHessianStateState xx;
    CostHessianControlControl uu;
    CostHessianStateControl xu;
  };
  CostFunction(CostHessianStateState Q, CostHessianControlControl R, Trajectory<ModelT> desired_trajectory)
      : Q_(Q), R_(R), desired_trajectory_(std::move(desired_trajectory)) {}

  double operator()(const typename ModelT::State &x, const typename ModelT::Control &u, int i,
                    CostDifferentials *diffs = nullptr) const {
    const auto &d = desired_trajectory_.at(i);  // std::out_of_range as cost.hh:39-40
    if (!h_) {
      const ModelT unit(1.0, Identity3(), 1.0, 0.0, 9.81);
      h_ = std::make_shared<Handle>(unit.c_model(), Q_, R_, 0.1, ILQROptions{});
    }
    double xs[13], xd[13], cost = 0.0;
    ModelT::pack(x, xs);
    ModelT::pack(d.state, xd);
    check(qilqr_cost_host(h_->h, 1, xs, u.data(), xd, d.control.data(), &cost, diffs ? diffs->x.data() : nullptr,
                          diffs ? diffs->u.data() : nullptr, diffs ? diffs->xx.data() : nullptr,
                          diffs ? diffs->uu.data() : nullptr, diffs ? diffs->xu.data() : nullptr));
    return cost;
  }
  const CostHessianStateState &Q() const { return Q_; }
  const CostHessianControlControl &R() const { return R_; }
  const Trajectory<ModelT> &desired_trajectory() const { return desired_trajectory_; }

 private:
  CostHessianStateState Q_;
  CostHessianControlControl R_;
  Trajectory<ModelT> desired_trajectory_;
  mutable std::shared_ptr<Handle> h_;
};

namespace detail {
struct CostReductionTerms {  // ilqr.hh:13-16
  double QuTk = 0;
  double kTQuuk = 0;
};
inline double calculate_cost_reduction(const CostReductionTerms &t, const double step = 1.0) {  // ilqr.hh:18-22
  return step * t.QuTk + step * step * t.kTQuuk / 2.0;
}
}  // namespace detail

template <class ModelT>
struct ILQR {  // ilqr.hh:25-206
  using CostFunc = CostFunction<ModelT>;
  using FeedbackGains = Mat4x12;
  struct ControlUpdate {  // ilqr.hh:43-47
    typename ModelT::Control ff_update;
    FeedbackGains feedback;
  };
  using ControlUpdateTrajectory = std::vector<ControlUpdate>;

  ModelT model_;
  CostFunc cost_function_;
  double dt_s_;
  ILQROptions options_;

  ILQR(ModelT model, CostFunc cost_function, double dt_s, ILQROptions options)
      : model_(std::move(model)), cost_function_(std::move(cost_function)), dt_s_(dt_s), options_(options),
        h_(std::make_shared<Handle>(model_.c_model(), cost_function_.Q(), cost_function_.R(), dt_s, options, 0,
                                    model_.model_flags_)),
        desired_(flatten<ModelT>(cost_function_.desired_trajectory())) {
    model_.configure(h_->h);
  }

  // ilqr.hh:53-87
  std::pair<Trajectory<ModelT>, ILQRDebug<ModelT>> solve(const Trajectory<ModelT> &initial_traj) const {
    const size_t n = initial_traj.size();
    require_desired(n);
    const std::vector<double> in = flatten<ModelT>(initial_traj);
    const int cap = int(std::ceil(options_.convergence_criteria.max_iters)) + 1;
    std::vector<double> out(n * 18), hist(cap), dbg(options_.populate_debug ? size_t(cap) * n * 18 : 0);
    qilqr_result_t res{};
    check(qilqr_solve_host(h_->h, 1, int(n), desired_.data(), 1, in.data(), out.data(), nullptr, nullptr, hist.data(),
                           cap, dbg.empty() ? nullptr : dbg.data(), cap, &res), h_->h);
    if (res.status == QILQR_STATUS_LINE_SEARCH_FAILED || res.status == QILQR_STATUS_NONFINITE)  // ilqr.hh:191-193
      throw std::runtime_error("Reached maximum number of line search iterations, " +
                               std::to_string(options_.line_search_params.max_iters) + "\n");
    ILQRDebug<ModelT> debug;
    if (options_.populate_debug)
      for (int i = 0; i < res.num_debug; ++i)
        debug.push_back(ILQRIterDebug<ModelT>{unflatten<ModelT>(&dbg[size_t(i) * n * 18], n), hist[i]});
    return {unflatten<ModelT>(out.data(), n), std::move(debug)};
  }

  // The batched form: many independent problems per call (the reason to use the GPU path).
  // All trajectories have the same length as the desired trajectory's prefix in use.
  std::vector<Trajectory<ModelT>> solve_batch(const std::vector<Trajectory<ModelT>> &initial,
                                              std::vector<qilqr_result_t> *results = nullptr) const {
    if (initial.empty()) return {};
    const size_t n = initial[0].size(), B = initial.size();
    require_desired(n);
    std::vector<double> in(B * n * 18), out(B * n * 18);
    for (size_t b = 0; b < B; ++b) {
      if (initial[b].size() != n) throw std::invalid_argument("solve_batch: trajectories must have equal length");
      const auto f = flatten<ModelT>(initial[b]);
      std::copy(f.begin(), f.end(), in.begin() + b * n * 18);
    }
    std::vector<qilqr_result_t> res(B);
    check(qilqr_solve_host(h_->h, int(B), int(n), desired_.data(), 1, in.data(), out.data(), nullptr, nullptr, nullptr, 0,
                           nullptr, 0, res.data()), h_->h);
    std::vector<Trajectory<ModelT>> sol(B);
    for (size_t b = 0; b < B; ++b) sol[b] = unflatten<ModelT>(&out[b * n * 18], n);
    if (results) *results = std::move(res);
    return sol;
  }

  // ilqr.hh:89-95
  double cost_trajectory(const Trajectory<ModelT> &traj) const {
    const std::vector<double> t = flatten<ModelT>(traj);
    double cost = 0.0;
    check(qilqr_cost_trajectory_host(h_->h, 1, int(traj.size()), desired_.data(), 1, int(desired_.size() / 18), t.data(),
                                     &cost), h_->h);
    return cost;
  }

  // ilqr.hh:97-147
  std::pair<ControlUpdateTrajectory, detail::CostReductionTerms> backwards_pass(const Trajectory<ModelT> &traj) const {
    const size_t n = traj.size();
    require_desired(n);
    const std::vector<double> t = flatten<ModelT>(traj);
    std::vector<double> k(n * 4), K(n * 48);
    double terms[2];
    check(qilqr_backwards_pass_host(h_->h, 1, int(n), desired_.data(), 1, t.data(), k.data(), K.data(), terms), h_->h);
    return {to_updates(k, K, n), detail::CostReductionTerms{terms[0], terms[1]}};
  }

  // ilqr.hh:149-172
  Trajectory<ModelT> forward_sim(const Trajectory<ModelT> &current_traj, const ControlUpdateTrajectory &ctrl_update_traj,
                                 const double line_search_alpha = 1.0) const {
    const size_t n = current_traj.size();
    const std::vector<double> t = flatten<ModelT>(current_traj);
    std::vector<double> k, K, out(n * 18);
    from_updates(ctrl_update_traj, k, K);
    check(qilqr_forward_sim_host(h_->h, 1, int(n), t.data(), k.data(), K.data(), &line_search_alpha, out.data()), h_->h);
    return unflatten<ModelT>(out.data(), n);
  }

  // ilqr.hh:174-194
  std::tuple<Trajectory<ModelT>, double, double> line_search(const Trajectory<ModelT> &current_traj,
                                                             const double current_cost,
                                                             const ControlUpdateTrajectory &ctrl_update_traj,
                                                             const detail::CostReductionTerms &terms) const {
    const size_t n = current_traj.size();
    require_desired(n);
    const std::vector<double> t = flatten<ModelT>(current_traj);
    std::vector<double> k, K, out(n * 18);
    from_updates(ctrl_update_traj, k, K);
    const double tr[2] = {terms.QuTk, terms.kTQuuk};
    double new_cost = 0, step = 0;
    int32_t status = 0;
    check(qilqr_line_search_host(h_->h, 1, int(n), desired_.data(), 1, t.data(), &current_cost, k.data(), K.data(), tr,
                                 out.data(), &new_cost, &step, &status), h_->h);
    if (status != 0)
      throw std::runtime_error("Reached maximum number of line search iterations, " +
                               std::to_string(options_.line_search_params.max_iters) + "\n");
    return {unflatten<ModelT>(out.data(), n), new_cost, step};
  }

  // ilqr.hh:196-205
  bool is_converged(const double cost, const double new_cost) const {
    if (std::abs(cost - new_cost) / std::abs(cost) < options_.convergence_criteria.rtol) return true;
    if (std::abs(cost - new_cost) < options_.convergence_criteria.atol) return true;
    return false;
  }

 private:
  void require_desired(size_t n) const {
    if (n * 18 > desired_.size()) throw std::out_of_range("vector::_M_range_check: trajectory longer than desired");
  }
  static ControlUpdateTrajectory to_updates(const std::vector<double> &k, const std::vector<double> &K, size_t n) {
    ControlUpdateTrajectory u(n);
    for (size_t i = 0; i < n; ++i) {
      for (int j = 0; j < 4; ++j) u[i].ff_update[j] = k[4 * i + j];
      for (int j = 0; j < 48; ++j) u[i].feedback[j] = K[48 * i + j];
    }
    return u;
  }
  static void from_updates(const ControlUpdateTrajectory &u, std::vector<double> &k, std::vector<double> &K) {
    k.resize(u.size() * 4);
    K.resize(u.size() * 48);
    for (size_t i = 0; i < u.size(); ++i) {
      for (int j = 0; j < 4; ++j) k[4 * i + j] = u[i].ff_update[j];
      for (int j = 0; j < 48; ++j) K[48 * i + j] = u[i].feedback[j];
    }
  }
  std::shared_ptr<Handle> h_;
  std::vector<double> desired_;
};

}  // namespace qilqr
