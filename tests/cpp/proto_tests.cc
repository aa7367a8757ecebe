// Tests of the hand-written proto3 codec (quadrotorilqr_b200/cpp/qilqr_proto.hpp).  CPU only: no solver
// call is made.  Modes:
//   proto_tests roundtrip              the reference's trajectory_to_proto_test.cc, ilqr_options_to_proto_test.cc
//                                      and ilqr_debug_to_proto_test.cc (from_proto(to_proto(x)) == x), through bytes
//   proto_tests reencode KIND IN OUT   parse IN as message KIND, serialise it again into OUT
//   proto_tests emit KIND OUT          serialise a fixed message of KIND (the Python test builds the same one)
//   proto_tests solve OPTS DESIRED INITIAL OUT_TRAJ OUT_DEBUG   (GPU) the reference's default problem
//                                      (quadrotor_ilqr.py:256-306) through qilqr::QuadrotorILQR, bytes in / bytes out
#include <cstdio>
#include <fstream>
#include <iterator>
#include <sstream>

#include "../../quadrotorilqr_b200/cpp/qilqr_proto.hpp"

using namespace qilqr;
namespace pb = qilqr::proto;

static int g_failed = 0, g_checks = 0;
#define EXPECT_TRUE(c)                                                   \
  do {                                                                   \
    ++g_checks;                                                          \
    if (!(c)) { ++g_failed; std::printf("  %s:%d: %s\n", __FILE__, __LINE__, #c); } \
  } while (0)

static bool same(const QuadrotorModel::State &a, const QuadrotorModel::State &b) {
  return a.inertial_from_body.translation == b.inertial_from_body.translation &&
         a.inertial_from_body.quaternion == b.inertial_from_body.quaternion && a.body_velocity == b.body_velocity;
}
static bool same(const Trajectory<QuadrotorModel> &a, const Trajectory<QuadrotorModel> &b) {
  if (a.size() != b.size()) return false;
  for (size_t i = 0; i < a.size(); ++i)
    if (a[i].time_s != b[i].time_s || !same(a[i].state, b[i].state) || a[i].control != b[i].control) return false;
  return true;
}
static Trajectory<QuadrotorModel> sample_traj() {  // trajectory_to_proto_test.cc:12-34 (poses written out instead of Exp())
  QuadrotorModel::State x0, x1;
  x0.inertial_from_body.translation = {1.0, 2.0, 3.0};
  x0.inertial_from_body.quaternion = {0.5, -0.5, 0.5, 0.5};
  x0.body_velocity = {2.0, 3.0, 4.0, 5.0, 6.0, 7.0};
  x1.inertial_from_body.translation = {2.0, 3.0, 4.0};
  x1.inertial_from_body.quaternion = {0.0, 0.6, 0.0, 0.8};
  x1.body_velocity = {3.0, 4.0, 5.0, 6.0, 7.0, 8.0};
  return {{1.0, x0, {3.0, 4.0, 5.0, 6.0}}, {2.0, x1, {4.0, 5.0, 6.0, 7.0}}};
}
static ILQROptions sample_options() {  // ilqr_options_to_proto_test.cc:7-12
  ILQROptions o;
  o.line_search_params = {1.0, 2.0, 3};
  o.convergence_criteria = {4.0, 5.0, 6};
  o.populate_debug = true;
  return o;
}
static ILQRDebug<QuadrotorModel> sample_debug() {  // ilqr_debug_to_proto_test.cc:14-37
  ILQRIterDebug<QuadrotorModel> d0{{{0.0, QuadrotorModel::State{}, {0, 0, 0, 0}}}, 23.3};
  ILQRIterDebug<QuadrotorModel> d1 = d0;
  d1.trajectory.front().time_s = 1.0;
  d1.cost = 5;
  return {d0, d1};
}

template <class M>
static M through_bytes(const M &m) {
  M out;
  EXPECT_TRUE(pb::ParseFromString(pb::SerializeAsString(m), &out));
  return out;
}

static int roundtrip() {
  {
    const auto traj = sample_traj();
    EXPECT_TRUE(same(traj, pb::from_proto(pb::to_proto(traj))));
    EXPECT_TRUE(same(traj, pb::from_proto(through_bytes(pb::to_proto(traj)))));
    // the message stores the quaternion w-first (trajectory_to_proto.cc:67-83)
    const auto p = pb::to_proto(traj);
    EXPECT_TRUE(p.points[1].state->inertial_from_body->rotation->quaternion->c0 == 0.8);
    EXPECT_TRUE(p.points[1].state->inertial_from_body->rotation->quaternion->c2 == 0.6);
  }
  {
    const auto o = sample_options();
    EXPECT_TRUE(o == pb::from_proto(pb::to_proto(o)));
    EXPECT_TRUE(o == pb::from_proto(through_bytes(pb::to_proto(o))));
    ILQROptions neg = o;
    neg.line_search_params.max_iters = -7;  // negative int32: ten-byte varint
    EXPECT_TRUE(neg == pb::from_proto(through_bytes(pb::to_proto(neg))));
  }
  {
    const auto d = sample_debug();
    const auto rt = pb::from_proto(through_bytes(pb::to_proto(d)));
    EXPECT_TRUE(rt.size() == d.size());
    for (size_t i = 0; i < d.size() && i < rt.size(); ++i)
      EXPECT_TRUE(rt[i].cost == d[i].cost && same(rt[i].trajectory, d[i].trajectory));
  }
  {  // zero-valued scalars are not written; an all-default message is empty; presence of sub-messages is kept
    EXPECT_TRUE(pb::SerializeAsString(pb::Vec6{}).empty());
    EXPECT_TRUE(pb::SerializeAsString(pb::to_proto(QuadrotorModel::State{})).size() > 0);
    pb::QuadrotorTrajectoryPoint pt;
    EXPECT_TRUE(pb::SerializeAsString(pt).empty());
    pt.control = pb::Vec4{};
    EXPECT_TRUE(pb::SerializeAsString(pt) == std::string("\x1a\x00", 2));
    EXPECT_TRUE(through_bytes(pt).control.has_value() && !through_bytes(pt).state.has_value());
  }
  {  // malformed input is rejected, unknown fields are skipped
    pb::QuadrotorTrajectory t;
    EXPECT_TRUE(!pb::ParseFromString(std::string("\x0a\x05\x09", 3), &t));            // truncated
    pb::Vec3 v;
    EXPECT_TRUE(pb::ParseFromString(std::string("\x78\x05\x11\0\0\0\0\0\0\xf0\x3f", 11), &v));  // field 15 varint, then c1 = 1.0
    EXPECT_TRUE(v.c1 == 1.0 && v.c0 == 0.0);
  }
  std::printf("%d checks, %d failed\n", g_checks, g_failed);
  return g_failed ? 1 : 0;
}

static std::string slurp(const char *path) {
  std::ifstream f(path, std::ios::binary);
  return std::string(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
}
static void spit(const char *path, const std::string &bytes) {
  std::ofstream f(path, std::ios::binary);
  f.write(bytes.data(), std::streamsize(bytes.size()));
}
template <class M>
static int reencode(const char *in, const char *out) {
  M m;
  if (!pb::ParseFromString(slurp(in), &m)) return 2;
  spit(out, pb::SerializeAsString(m));
  return 0;
}

int main(int argc, char **argv) {
  const std::string mode = argc > 1 ? argv[1] : "roundtrip";
  if (mode == "roundtrip") return roundtrip();
  const std::string kind = argc > 2 ? argv[2] : "";
  if (mode == "reencode" && argc == 5) {
    if (kind == "trajectory") return reencode<pb::QuadrotorTrajectory>(argv[3], argv[4]);
    if (kind == "options") return reencode<pb::ILQROptions>(argv[3], argv[4]);
    if (kind == "debug") return reencode<pb::QuadrotorILQRDebug>(argv[3], argv[4]);
  }
  if (mode == "emit" && argc == 4) {
    if (kind == "trajectory") { spit(argv[3], pb::SerializeAsString(pb::to_proto(sample_traj()))); return 0; }
    if (kind == "options") { spit(argv[3], pb::SerializeAsString(pb::to_proto(sample_options()))); return 0; }
    if (kind == "debug") { spit(argv[3], pb::SerializeAsString(pb::to_proto(sample_debug()))); return 0; }
  }
  if (mode == "solve" && argc == 7) {
    Mat12 Q{};
    for (int i = 0; i < 12; ++i) Q[13 * i] = i < 6 ? 100.0 : 1.0;
    try {
      QuadrotorILQR ilqr(1.0, Identity3(), 1.0, 0.0, 9.81, Q, Identity4(), slurp(argv[3]), 0.1, slurp(argv[2]));
      const auto [traj_bytes, debug_bytes] = ilqr.solve(slurp(argv[4]));
      spit(argv[5], traj_bytes);
      spit(argv[6], debug_bytes);
    } catch (const std::exception &e) {
      std::fprintf(stderr, "solve failed: %s\n", e.what());
      return 3;
    }
    return 0;
  }
  std::fprintf(stderr, "usage: proto_tests roundtrip | reencode KIND IN OUT | emit KIND OUT\n");
  return 64;
}
