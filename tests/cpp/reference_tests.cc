// The reference's own C++ unit tests for the hot path, ported onto the C++ host mirror
// (quadrotorilqr_b200/cpp/quadrotor_ilqr.hpp), i.e. running through the C ABI on the GPU:
//   ilqr_test.cc:102-190 (all six ILQRFixture tests), quadrotor_model_test.cc:94-143 and :348-397,
//   cost_test.cc:27-39, plus the three exceptions the reference can throw on this path.
// gtest is not in the image, so a 20-line harness stands in for it.
#include <cmath>
#include <cstdio>
#include <functional>
#include <sstream>
#include <string>
#include <vector>

#include "../../quadrotorilqr_b200/cpp/quadrotor_ilqr.hpp"

using namespace qilqr;
using State = QuadrotorModel::State;
using Control = QuadrotorModel::Control;
using ILQRSolver = ILQR<QuadrotorModel>;
using CostFunc = CostFunction<QuadrotorModel>;

static int g_failed = 0, g_checks = 0;
#define EXPECT_TRUE(c) do { ++g_checks; if (!(c)) { ++g_failed; std::printf("  FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); } } while (0)
#define EXPECT_EQ(a, b) EXPECT_TRUE((a) == (b))
#define EXPECT_LT(a, b) EXPECT_TRUE((a) < (b))
#define EXPECT_NEAR(a, b, tol) EXPECT_TRUE(std::fabs((a) - (b)) <= (tol))

static State create_identity_state() { return State{}; }
static Trajectory<QuadrotorModel> create_identity_traj(int num_pts, double dt_s) {  // ilqr_test.cc:23-36
  Trajectory<QuadrotorModel> traj;
  double time_s = 0.0;
  for (int i = 0; i < num_pts; ++i) {
    traj.push_back({time_s, create_identity_state(), Control{0, 0, 0, 0}});
    time_s += dt_s;
  }
  return traj;
}
static double norm12(const QuadrotorModel::StateTangent &t) {
  double s = 0;
  for (int i = 0; i < 12; ++i) s += t[i] * t[i];
  return std::sqrt(s);
}
static bool approx_state_eq(const State &lhs, const State &rhs, double tol) {  // ilqr_test.cc:38-48
  return norm12(rhs - lhs) < tol;  // log(lhs^-1 rhs) and the velocity difference
}
static void check_approx_traj_eq(const Trajectory<QuadrotorModel> &lhs, const Trajectory<QuadrotorModel> &rhs, double tol) {
  EXPECT_EQ(lhs.size(), rhs.size());
  for (size_t i = 0; i < lhs.size() && i < rhs.size(); ++i) {
    EXPECT_TRUE(approx_state_eq(lhs[i].state, rhs[i].state, tol));
    for (int j = 0; j < 4; ++j) EXPECT_NEAR(lhs[i].control[j], rhs[i].control[j], tol * std::fmax(1.0, std::fabs(rhs[i].control[j])));
  }
}

constexpr double mass_kg = 1.0;

struct ILQRFixture {  // ilqr_test.cc:68-100
  size_t N_ = 3;
  double dt_s_ = 0.1;
  ILQRSolver::ControlUpdateTrajectory ctrl_update_traj_;
  Trajectory<QuadrotorModel> current_traj_ = create_identity_traj(3, 0.1);
  ILQRSolver ilqr_{QuadrotorModel{mass_kg, Identity3(), 1.0, 1.0, 0.0},
                   CostFunc{Identity12(), Identity4(), create_identity_traj(3, 0.1)}, 0.1,
                   ILQROptions{LineSearchParams{0.5, 0.5, 10}, ConvergenceCriteria{1e-12, 1e-12, 100}}};
  ILQRFixture() {
    ILQRSolver::ControlUpdate u{};
    u.ff_update = {1, 1, 1, 1};
    ctrl_update_traj_.assign(N_, u);
  }
};

static void ForwardSimGeneratesCorrectTrajectory() {  // ilqr_test.cc:102-126
  ILQRFixture f;
  const Control u{1, 1, 1, 1};
  const double accel = 4.0 / mass_kg;
  State s0 = create_identity_state(), s1 = s0, s2 = s0;
  s1.body_velocity[2] = f.dt_s_ * accel;
  QuadrotorModel::StateTangent off{};
  off.body_velocity[2] = f.dt_s_ * f.dt_s_ * accel;
  s2 = s2 + off;
  s2.body_velocity[2] = 2.0 * f.dt_s_ * accel;
  Trajectory<QuadrotorModel> expected{{0.0, s0, u}, {f.dt_s_, s1, u}, {2 * f.dt_s_, s2, u}};
  const auto new_traj = f.ilqr_.forward_sim(f.current_traj_, f.ctrl_update_traj_);
  check_approx_traj_eq(new_traj, expected, 1e-6);
}
static void CostTrajectoryCalculatesCorrectCost() {  // ilqr_test.cc:128-141
  ILQRFixture f;
  const auto new_traj = f.ilqr_.forward_sim(f.current_traj_, f.ctrl_update_traj_);
  const double cost = f.ilqr_.cost_trajectory(new_traj);
  const double a = 4.0;
  const double expected = std::pow(f.dt_s_ * a, 2.0) + std::pow(f.dt_s_ * f.dt_s_ * a, 2.0) + std::pow(2.0 * f.dt_s_ * a, 2.0) + 3 * 4;
  EXPECT_TRUE(std::fabs(cost - expected) <= 4 * (std::nextafter(expected, 1e9) - expected));  // EXPECT_DOUBLE_EQ
}
static void BackwardPassReturnsZeroUpdateIfZeroGradient() {  // ilqr_test.cc:143-153
  ILQRFixture f;
  const auto [upd, terms] = f.ilqr_.backwards_pass(f.current_traj_);
  EXPECT_EQ(upd.size(), f.N_);
  EXPECT_EQ(terms.QuTk, 0.0);
  EXPECT_EQ(terms.kTQuuk, 0.0);
  for (const auto &c : upd) EXPECT_TRUE(c.ff_update == (Control{0, 0, 0, 0}));
}
static void BackwardsPassExpectedValueReductionIsNegativeIfReductionPossible() {  // ilqr_test.cc:155-164
  ILQRFixture f;
  const auto new_traj = f.ilqr_.forward_sim(f.current_traj_, f.ctrl_update_traj_);
  EXPECT_LT(f.ilqr_.backwards_pass(new_traj).second.QuTk, 0.0);
}
static void LineSearchFindsStepSizeThatReducesCost() {  // ilqr_test.cc:166-177
  ILQRFixture f;
  const auto traj = f.ilqr_.forward_sim(f.current_traj_, f.ctrl_update_traj_);
  const auto cost = f.ilqr_.cost_trajectory(traj);
  const auto [upd, terms] = f.ilqr_.backwards_pass(traj);
  const auto [new_traj, new_cost, step] = f.ilqr_.line_search(traj, cost, upd, terms);
  EXPECT_LT(new_cost - cost, f.ilqr_.options_.line_search_params.desired_reduction_frac * detail::calculate_cost_reduction(terms, step));
  bool threw = false;
  try { f.ilqr_.line_search(traj, -1e30, upd, terms); } catch (const std::runtime_error &) { threw = true; }  // ilqr.hh:191-193
  EXPECT_TRUE(threw);
}
static void SolveFindsOptimalTrajectory() {  // ilqr_test.cc:179-190
  ILQRFixture f;
  for (auto &pt : f.ctrl_update_traj_) { pt.ff_update[0] *= 100; pt.ff_update[2] *= 100; }
  const auto initial = f.ilqr_.forward_sim(f.current_traj_, f.ctrl_update_traj_);
  const auto [opt, debug] = f.ilqr_.solve(initial);
  check_approx_traj_eq(f.current_traj_, opt, 1e-6);
  // batched form of the same call
  std::vector<qilqr_result_t> res;
  const auto sols = f.ilqr_.solve_batch({initial, f.current_traj_, initial}, &res);
  EXPECT_EQ(sols.size(), size_t(3));
  for (const auto &s : sols) check_approx_traj_eq(f.current_traj_, s, 1e-6);
  EXPECT_EQ(res[0].backward_passes, res[2].backward_passes);
}
// ILQR<ModelT> with a second ModelT: the fixture's problem (ilqr_test.cc:68-100, :179-190)
// solved with the RK4 variant and with the reference dynamics on the model-agnostic kernels.
static void SecondModelSolveFindsOptimalTrajectory() {
  using VSolver = ILQR<QuadrotorModelVariant>;
  using VCost = CostFunction<QuadrotorModelVariant>;
  for (int variant = 0; variant < 3; ++variant) {
    const bool rk4 = variant >= 1, coriolis = variant == 2;
    Trajectory<QuadrotorModelVariant> identity;
    for (const auto &pt : create_identity_traj(3, 0.1)) identity.push_back({pt.time_s, pt.state, pt.control});
    VSolver ilqr{QuadrotorModelVariant{mass_kg, Identity3(), 1.0, 1.0, 0.0, rk4, coriolis},
                 VCost{Identity12(), Identity4(), identity}, 0.1,
                 ILQROptions{LineSearchParams{0.5, 0.5, 10}, ConvergenceCriteria{1e-12, 1e-12, 100}}};
    VSolver::ControlUpdate u{};
    u.ff_update = {100, 1, 100, 1};
    const VSolver::ControlUpdateTrajectory upd(3, u);
    const auto initial = ilqr.forward_sim(identity, upd);
    EXPECT_TRUE(ilqr.cost_trajectory(initial) > 1.0);
    const auto [opt, debug] = ilqr.solve(initial);
    for (size_t i = 0; i < opt.size(); ++i) {
      const auto d = (opt[i].state - identity[i].state).coeffs();
      for (int j = 0; j < 12; ++j) EXPECT_NEAR(d[j], 0.0, 1e-6);
      for (int j = 0; j < 4; ++j) EXPECT_NEAR(opt[i].control[j], 0.0, 1e-6);
    }
  }
}
// ILQR<ModelT> with a model SUPPLIED BY THE CALLER (examples/user_model_drag.cu, compiled at run time): the same
// fixture problem; with c_d = 0 the optimum is the identity trajectory as for the reference model.
static void UserModelSolveFindsOptimalTrajectory() {
  std::string path = std::string(QILQR_REPO_ROOT) + "/examples/user_model_drag.cu", src;
  if (FILE *f = std::fopen(path.c_str(), "rb")) {
    char buf[4096];
    size_t n;
    while ((n = std::fread(buf, 1, sizeof(buf), f)) > 0) src.append(buf, n);
    std::fclose(f);
  }
  EXPECT_TRUE(!src.empty());
  using USolver = ILQR<UserModel>;
  using UCost = CostFunction<UserModel>;
  const QuadrotorModel base{mass_kg, Identity3(), 1.0, 1.0, 0.0};
  const UserModel model{base, src, {mass_kg, 1.0, 1.0, 1.0, 1.0, 1.0, 0.0, 0.0}};  // g = 0 as in the fixture
  Trajectory<UserModel> identity;
  for (const auto &pt : create_identity_traj(3, 0.1)) identity.push_back({pt.time_s, pt.state, pt.control});
  USolver ilqr{model, UCost{Identity12(), Identity4(), identity}, 0.1,
               ILQROptions{LineSearchParams{0.5, 0.5, 10}, ConvergenceCriteria{1e-12, 1e-12, 100}}};
  USolver::ControlUpdate u{};
  u.ff_update = {100, 1, 100, 1};
  const USolver::ControlUpdateTrajectory upd(3, u);
  const auto initial = ilqr.forward_sim(identity, upd);
  EXPECT_TRUE(ilqr.cost_trajectory(initial) > 1.0);
  const auto [opt, debug] = ilqr.solve(initial);
  for (size_t i = 0; i < opt.size(); ++i) {
    const auto d = (opt[i].state - identity[i].state).coeffs();
    for (int j = 0; j < 12; ++j) EXPECT_NEAR(d[j], 0.0, 1e-6);
    for (int j = 0; j < 4; ++j) EXPECT_NEAR(opt[i].control[j], 0.0, 1e-6);
  }
}
static void DiscreteDynamicsKnownAnswers() {  // quadrotor_model_test.cc:94-143
  QuadrotorModel quad{mass_kg, Identity3(), 1.0, 1.0};
  State x = create_identity_state();
  x.body_velocity = {1.0, 2.0, 3.0, 0, 0, 0};
  auto xn = quad.discrete_dynamics(x, Control{1, 1, 1, 1}, 0.1);
  EXPECT_NEAR(xn.inertial_from_body.translation[0], 0.1, 1e-6);
  EXPECT_NEAR(xn.inertial_from_body.translation[1], 0.2, 1e-6);
  EXPECT_NEAR(xn.inertial_from_body.translation[2], 0.3, 1e-6);
  EXPECT_NEAR(xn.body_velocity[2], 3.0 + (4.0 - 9.81) * 0.1, 1e-6);
  State y = create_identity_state();
  y.body_velocity[3] = 1.2;
  auto yn = quad.discrete_dynamics(y, Control{0.0, -1.0, 0.0, 1.0}, 0.1);
  EXPECT_NEAR(yn.body_velocity[3], 1.2 + 2.0 * 0.1, 1e-6);
  EXPECT_NEAR(yn.inertial_from_body.quaternion[0], std::sin(0.06), 1e-9);
  // analytic Jacobian vs central differences on the control (quadrotor_model_test.cc:57-79,172-197)
  QuadrotorModel::DynamicsDifferentials d;
  const Control u{1.0, 2.0, 3.0, 4.0};
  quad.discrete_dynamics(x, u, 0.1, &d);
  for (int j = 0; j < 4; ++j) {
    Control up = u, um = u;
    up[j] += 1e-6; um[j] -= 1e-6;
    const auto diff = quad.discrete_dynamics(x, up, 0.1) - quad.discrete_dynamics(x, um, 0.1);
    for (int i = 0; i < 12; ++i) EXPECT_NEAR(diff[i] / 2e-6, d.J_u[4 * i + j], 1e-5);
  }
}
static void StateTangentAndExceptions() {  // quadrotor_model_test.cc:348-369, quadrotor_model.cc:21-24, cost.hh:39-40
  QuadrotorModel::StateTangent t;
  for (int i = 0; i < 12; ++i) t[i] = i;
  EXPECT_EQ(t.body_velocity[5], 5.0);
  EXPECT_EQ(t.body_acceleration[0], 6.0);
  bool threw = false;
  try { QuadrotorModel bad{1.0, Mat3{1, 0, 0, 0, -1, 0, 0, 0, 1}, 1.0, 1.0}; } catch (const std::runtime_error &) { threw = true; }
  EXPECT_TRUE(threw);
  ILQRFixture f;
  threw = false;
  try { f.ilqr_.cost_trajectory(create_identity_traj(4, 0.1)); } catch (const std::out_of_range &) { threw = true; }
  EXPECT_TRUE(threw);
}
static void CostZeroAtZeroError() {  // cost_test.cc:27-39
  State x = create_identity_state();
  QuadrotorModel::StateTangent d{};
  d.body_velocity = {0.3, -0.2, 0.5, 0.4, 0.1, -0.7};
  x = x + d;
  x.body_velocity = {1, 2, 3, 4, 5, 6};
  const Control u{0.1, 0.2, 0.3, 0.4};
  CostFunc c{Identity12(), Identity4(), {{0.0, x, u}}};
  EXPECT_EQ(c(x, u, 0), 0.0);
  CostFunc::CostDifferentials diffs;
  State x2 = x + d;
  EXPECT_TRUE(c(x2, u, 0, &diffs) > 0.0);
  EXPECT_EQ(diffs.uu[0], 2.0);
  EXPECT_EQ(diffs.xu[7], 0.0);
}

// trajectory_test.cc:20-48, ilqr_debug_test.cc, ilqr_options_test.cc: the equality operators (host only)
static void EqualityOperators() {
  const TrajectoryPoint<QuadrotorModel> pt0{0.0, create_identity_state(), Control{0, 0, 0, 0}};
  EXPECT_TRUE(pt0 == pt0);
  auto pt1 = pt0;
  pt1.time_s = 1.0;
  EXPECT_TRUE(pt0 != pt1);
  pt1 = pt0;
  QuadrotorModel::StateTangent off{};
  off.body_velocity[2] = 1.0;
  pt1.state = pt1.state + off;
  EXPECT_TRUE(pt0 != pt1);
  pt1 = pt0;
  pt1.control = {1, 1, 1, 1};
  EXPECT_TRUE(pt0 != pt1);
  auto pt2 = pt0;  // q and -q are the same rotation
  for (auto &c : pt2.state.inertial_from_body.quaternion) c = -c;
  EXPECT_TRUE(pt0 == pt2);
  const ILQRIterDebug<QuadrotorModel> d0{{pt0}, 23.3};
  auto d1 = d0;
  EXPECT_TRUE(d0 == d1);
  d1.cost = 5;
  EXPECT_TRUE(d0 != d1);
  d1 = d0;
  d1.trajectory.front().time_s = 1.0;
  EXPECT_TRUE(d0 != d1);
  EXPECT_TRUE((LineSearchParams{1.0, 2.0, 3} == LineSearchParams{1.0, 2.0, 3}));
  EXPECT_TRUE(!(LineSearchParams{0.0, 2.0, 3} == LineSearchParams{1.0, 2.0, 3}));
  std::ostringstream os;
  os << Trajectory<QuadrotorModel>{pt0};
  EXPECT_TRUE(os.str().find("time_s: 0") != std::string::npos && os.str().find("body velocity") != std::string::npos);
}

int main() {
  const std::vector<std::pair<std::string, std::function<void()>>> tests = {
      {"ILQRFixture.ForwardSimGeneratesCorrectTrajectory", ForwardSimGeneratesCorrectTrajectory},
      {"ILQRFixture.CostTrajectoryCalculatesCorrectCost", CostTrajectoryCalculatesCorrectCost},
      {"ILQRFixture.BackwardPassReturnsZeroUpdateIfZeroGradient", BackwardPassReturnsZeroUpdateIfZeroGradient},
      {"ILQRFixture.BackwardsPassExpectedValueReductionIsNegativeIfReductionPossible",
       BackwardsPassExpectedValueReductionIsNegativeIfReductionPossible},
      {"ILQRFixture.LineSearchFindsStepSizeThatReducesCost", LineSearchFindsStepSizeThatReducesCost},
      {"ILQRFixture.SolveFindsOptimalTrajectory", SolveFindsOptimalTrajectory},
      {"SecondModel.SolveFindsOptimalTrajectory", SecondModelSolveFindsOptimalTrajectory},
      {"UserModel.SolveFindsOptimalTrajectory", UserModelSolveFindsOptimalTrajectory},
      {"QuadrotorModelTest.DiscreteDynamicsKnownAnswers", DiscreteDynamicsKnownAnswers},
      {"StateTangentAndExceptions", StateTangentAndExceptions},
      {"EqualityOperators", EqualityOperators},
      {"ComputeCost.ReturnsZeroCostWhenZeroError", CostZeroAtZeroError},
  };
  for (const auto &t : tests) {
    const int before = g_failed;
    try { t.second(); } catch (const std::exception &e) { ++g_failed; std::printf("  EXCEPTION %s\n", e.what()); }
    std::printf("[%s] %s\n", g_failed == before ? "  OK  " : "FAILED", t.first.c_str());
  }
  std::printf("%d checks, %d failed\n", g_checks, g_failed);
  return g_failed ? 1 : 0;
}
