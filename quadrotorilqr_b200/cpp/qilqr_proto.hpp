// =============================================================================
// qilqr_proto.hpp -- the reference's protobuf messages (src/trajectory.proto,
// src/ilqr_options.proto, src/ilqr_debug.proto; package src.proto) as plain C++ structs with a
// hand-written proto3 wire codec, and the converters of trajectory_to_proto.{hh,cc},
// ilqr_options_to_proto.{hh,cc} and ilqr_debug_to_proto.{hh,cc} on top of them.
//
// Why hand-written: the image has neither protoc nor libprotobuf, and the
// eleven messages use only three wire types.  Bytes written here parse with any protobuf runtime
// built from the reference's .proto files and vice versa (tests/test_proto_codec.py checks both
// directions, byte for byte, against the Python runtime).
//
// proto3 rules implemented:
//   * field key = (number << 3) | wire type; double = type 1 (8 bytes little endian), int32/bool =
//     type 0 (varint; negative int32 sign-extended to 10 bytes), message = type 2 (length-delimited)
//   * scalars without presence are omitted when zero -- decided on the bit pattern, so -0.0 is written
//     (what upb and current protobuf C++ do; protobuf 3.19's generated code would drop it)
//   * message fields have presence (std::optional); the reference's converters always set them
//   * fields are written in field-number order; on parse any order is accepted, unknown fields are
//     skipped, a repeated occurrence of a singular message field merges into the previous one
// =============================================================================
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <optional>
#include <string>
#include <vector>

#include "quadrotor_ilqr.hpp"

namespace qilqr {
namespace proto {

// ---- messages (field numbers = declaration order, as in the .proto files) ---------------------
struct Vec3 { double c0 = 0, c1 = 0, c2 = 0; };
struct Vec4 { double c0 = 0, c1 = 0, c2 = 0, c3 = 0; };
struct Vec6 { double c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0, c5 = 0; };
struct SO3 { std::optional<Vec4> quaternion; };  // coefficients w, x, y, z (trajectory.proto:27-30)
struct SE3 { std::optional<Vec3> translation; std::optional<SO3> rotation; };
struct QuadrotorState { std::optional<SE3> inertial_from_body; std::optional<Vec6> body_velocity; };
struct QuadrotorTrajectoryPoint { double time_s = 0; std::optional<QuadrotorState> state; std::optional<Vec4> control; };
struct QuadrotorTrajectory { std::vector<QuadrotorTrajectoryPoint> points; };
struct LineSearchParams { double step_update = 0, desired_reduction_frac = 0; int32_t max_iters = 0; };
struct ConvergenceCriteria { double rtol = 0, atol = 0, max_iters = 0; };
struct ILQROptions {
  std::optional<LineSearchParams> line_search_params;
  std::optional<ConvergenceCriteria> convergence_criteria;
  bool populate_debug = false;
};
struct QuadrotorILQRIterDebug { std::optional<QuadrotorTrajectory> trajectory; double cost = 0; };
struct QuadrotorILQRDebug { std::vector<QuadrotorILQRIterDebug> iter_debugs; };

// ---- wire primitives -----------------------------------------------------------------------
namespace wire {
inline void put_varint(std::string &out, uint64_t v) {
  while (v >= 0x80) { out.push_back(char((v & 0x7f) | 0x80)); v >>= 7; }
  out.push_back(char(v));
}
inline void put_key(std::string &out, int field, int type) { put_varint(out, uint64_t(field) << 3 | uint64_t(type)); }
inline void put_double(std::string &out, int field, double v) {
  uint64_t bits;
  std::memcpy(&bits, &v, 8);
  if (bits == 0) return;
  put_key(out, field, 1);
  for (int i = 0; i < 8; ++i) out.push_back(char((bits >> (8 * i)) & 0xff));
}
inline void put_int32(std::string &out, int field, int32_t v) {
  if (v == 0) return;
  put_key(out, field, 0);
  put_varint(out, uint64_t(int64_t(v)));
}
inline void put_bool(std::string &out, int field, bool v) {
  if (!v) return;
  put_key(out, field, 0);
  out.push_back(char(1));
}
struct Reader {
  const uint8_t *p, *end;
  bool ok = true;
  bool done() const { return p >= end; }
  uint64_t varint() {
    uint64_t v = 0;
    for (int shift = 0; shift < 70; shift += 7) {
      if (p >= end) { ok = false; return 0; }
      const uint8_t b = *p++;
      if (shift < 64) v |= uint64_t(b & 0x7f) << shift;
      if (!(b & 0x80)) return v;
    }
    ok = false;
    return 0;
  }
  double fixed64() {
    if (end - p < 8) { ok = false; p = end; return 0; }
    uint64_t bits = 0;
    for (int i = 0; i < 8; ++i) bits |= uint64_t(p[i]) << (8 * i);
    p += 8;
    double v;
    std::memcpy(&v, &bits, 8);
    return v;
  }
  Reader sub() {  // a length-delimited payload
    const uint64_t n = varint();
    if (!ok || n > uint64_t(end - p)) { ok = false; return Reader{end, end, false}; }
    Reader r{p, p + n};
    p += n;
    return r;
  }
  void skip(int type) {
    switch (type) {
      case 0: varint(); break;
      case 1: if (end - p < 8) ok = false; else p += 8; break;
      case 2: sub(); break;
      case 5: if (end - p < 4) ok = false; else p += 4; break;
      default: ok = false;  // groups (3, 4) do not occur in proto3 files
    }
  }
};
}  // namespace wire

// ---- encode / decode, one overload pair per message ------------------------------------------
inline void encode(const Vec3 &m, std::string &out) { wire::put_double(out, 1, m.c0); wire::put_double(out, 2, m.c1); wire::put_double(out, 3, m.c2); }
inline void encode(const Vec4 &m, std::string &out) {
  wire::put_double(out, 1, m.c0); wire::put_double(out, 2, m.c1); wire::put_double(out, 3, m.c2); wire::put_double(out, 4, m.c3);
}
inline void encode(const Vec6 &m, std::string &out) {
  wire::put_double(out, 1, m.c0); wire::put_double(out, 2, m.c1); wire::put_double(out, 3, m.c2);
  wire::put_double(out, 4, m.c3); wire::put_double(out, 5, m.c4); wire::put_double(out, 6, m.c5);
}
template <class M>
void put_message(std::string &out, int field, const M &m);
inline void encode(const SO3 &m, std::string &out) { if (m.quaternion) put_message(out, 1, *m.quaternion); }
inline void encode(const SE3 &m, std::string &out) {
  if (m.translation) put_message(out, 1, *m.translation);
  if (m.rotation) put_message(out, 2, *m.rotation);
}
inline void encode(const QuadrotorState &m, std::string &out) {
  if (m.inertial_from_body) put_message(out, 1, *m.inertial_from_body);
  if (m.body_velocity) put_message(out, 2, *m.body_velocity);
}
inline void encode(const QuadrotorTrajectoryPoint &m, std::string &out) {
  wire::put_double(out, 1, m.time_s);
  if (m.state) put_message(out, 2, *m.state);
  if (m.control) put_message(out, 3, *m.control);
}
inline void encode(const QuadrotorTrajectory &m, std::string &out) { for (const auto &p : m.points) put_message(out, 1, p); }
inline void encode(const LineSearchParams &m, std::string &out) {
  wire::put_double(out, 1, m.step_update); wire::put_double(out, 2, m.desired_reduction_frac); wire::put_int32(out, 3, m.max_iters);
}
inline void encode(const ConvergenceCriteria &m, std::string &out) {
  wire::put_double(out, 1, m.rtol); wire::put_double(out, 2, m.atol); wire::put_double(out, 3, m.max_iters);
}
inline void encode(const ILQROptions &m, std::string &out) {
  if (m.line_search_params) put_message(out, 1, *m.line_search_params);
  if (m.convergence_criteria) put_message(out, 2, *m.convergence_criteria);
  wire::put_bool(out, 3, m.populate_debug);
}
inline void encode(const QuadrotorILQRIterDebug &m, std::string &out) {
  if (m.trajectory) put_message(out, 1, *m.trajectory);
  wire::put_double(out, 2, m.cost);
}
inline void encode(const QuadrotorILQRDebug &m, std::string &out) { for (const auto &d : m.iter_debugs) put_message(out, 1, d); }
template <class M>
void put_message(std::string &out, int field, const M &m) {
  std::string body;
  encode(m, body);
  wire::put_key(out, field, 2);
  wire::put_varint(out, body.size());
  out += body;
}

// decode(reader, message): merges the payload into `m`; false on malformed input
template <class M, class F>
bool decode_fields(wire::Reader &r, M &m, F &&field) {
  while (r.ok && !r.done()) {
    const uint64_t key = r.varint();
    if (!r.ok) return false;
    const int number = int(key >> 3), type = int(key & 7);
    if (!field(number, type, m)) r.skip(type);
  }
  return r.ok;
}
template <class M>
bool decode_sub(wire::Reader &r, std::optional<M> &dst);
inline bool dbl(wire::Reader &r, int type, double &dst) { if (type != 1) return false; dst = r.fixed64(); return true; }

inline bool decode(wire::Reader &r, Vec3 &m) {
  return decode_fields(r, m, [&](int n, int t, Vec3 &v) {
    return n == 1 ? dbl(r, t, v.c0) : n == 2 ? dbl(r, t, v.c1) : n == 3 ? dbl(r, t, v.c2) : false; });
}
inline bool decode(wire::Reader &r, Vec4 &m) {
  return decode_fields(r, m, [&](int n, int t, Vec4 &v) {
    return n == 1 ? dbl(r, t, v.c0) : n == 2 ? dbl(r, t, v.c1) : n == 3 ? dbl(r, t, v.c2) : n == 4 ? dbl(r, t, v.c3) : false; });
}
inline bool decode(wire::Reader &r, Vec6 &m) {
  return decode_fields(r, m, [&](int n, int t, Vec6 &v) {
    double *c[6] = {&v.c0, &v.c1, &v.c2, &v.c3, &v.c4, &v.c5};
    return (n >= 1 && n <= 6) ? dbl(r, t, *c[n - 1]) : false; });
}
inline bool decode(wire::Reader &r, SO3 &m) {
  return decode_fields(r, m, [&](int n, int t, SO3 &v) { return (n == 1 && t == 2) ? decode_sub(r, v.quaternion) : false; });
}
inline bool decode(wire::Reader &r, SE3 &m) {
  return decode_fields(r, m, [&](int n, int t, SE3 &v) {
    if (t != 2) return false;
    return n == 1 ? decode_sub(r, v.translation) : n == 2 ? decode_sub(r, v.rotation) : false; });
}
inline bool decode(wire::Reader &r, QuadrotorState &m) {
  return decode_fields(r, m, [&](int n, int t, QuadrotorState &v) {
    if (t != 2) return false;
    return n == 1 ? decode_sub(r, v.inertial_from_body) : n == 2 ? decode_sub(r, v.body_velocity) : false; });
}
inline bool decode(wire::Reader &r, QuadrotorTrajectoryPoint &m) {
  return decode_fields(r, m, [&](int n, int t, QuadrotorTrajectoryPoint &v) {
    if (n == 1) return dbl(r, t, v.time_s);
    if (t != 2) return false;
    return n == 2 ? decode_sub(r, v.state) : n == 3 ? decode_sub(r, v.control) : false; });
}
inline bool decode(wire::Reader &r, QuadrotorTrajectory &m) {
  return decode_fields(r, m, [&](int n, int t, QuadrotorTrajectory &v) {
    if (n != 1 || t != 2) return false;
    wire::Reader s = r.sub();
    v.points.emplace_back();
    if (!s.ok || !decode(s, v.points.back())) r.ok = false;
    return true; });
}
inline bool decode(wire::Reader &r, LineSearchParams &m) {
  return decode_fields(r, m, [&](int n, int t, LineSearchParams &v) {
    if (n == 3 && t == 0) { v.max_iters = int32_t(uint32_t(r.varint())); return true; }
    return n == 1 ? dbl(r, t, v.step_update) : n == 2 ? dbl(r, t, v.desired_reduction_frac) : false; });
}
inline bool decode(wire::Reader &r, ConvergenceCriteria &m) {
  return decode_fields(r, m, [&](int n, int t, ConvergenceCriteria &v) {
    return n == 1 ? dbl(r, t, v.rtol) : n == 2 ? dbl(r, t, v.atol) : n == 3 ? dbl(r, t, v.max_iters) : false; });
}
inline bool decode(wire::Reader &r, ILQROptions &m) {
  return decode_fields(r, m, [&](int n, int t, ILQROptions &v) {
    if (n == 3 && t == 0) { v.populate_debug = r.varint() != 0; return true; }
    if (t != 2) return false;
    return n == 1 ? decode_sub(r, v.line_search_params) : n == 2 ? decode_sub(r, v.convergence_criteria) : false; });
}
inline bool decode(wire::Reader &r, QuadrotorILQRIterDebug &m) {
  return decode_fields(r, m, [&](int n, int t, QuadrotorILQRIterDebug &v) {
    if (n == 2) return dbl(r, t, v.cost);
    return (n == 1 && t == 2) ? decode_sub(r, v.trajectory) : false; });
}
inline bool decode(wire::Reader &r, QuadrotorILQRDebug &m) {
  return decode_fields(r, m, [&](int n, int t, QuadrotorILQRDebug &v) {
    if (n != 1 || t != 2) return false;
    wire::Reader s = r.sub();
    v.iter_debugs.emplace_back();
    if (!s.ok || !decode(s, v.iter_debugs.back())) r.ok = false;
    return true; });
}
template <class M>
bool decode_sub(wire::Reader &r, std::optional<M> &dst) {
  wire::Reader s = r.sub();
  if (!dst) dst.emplace();
  if (!s.ok || !decode(s, *dst)) r.ok = false;
  return true;
}

// The two calls of the protobuf API the reference uses at its boundary
template <class M>
std::string SerializeAsString(const M &m) {
  std::string out;
  encode(m, out);
  return out;
}
template <class M>
bool ParseFromString(const std::string &bytes, M *m) {
  *m = M{};
  wire::Reader r{reinterpret_cast<const uint8_t *>(bytes.data()), reinterpret_cast<const uint8_t *>(bytes.data()) + bytes.size()};
  return decode(r, *m);
}

// ---- converters: trajectory_to_proto.cc, ilqr_options_to_proto.cc, ilqr_debug_to_proto.cc --------
inline Vec3 to_proto(const qilqr::Vec3 &v) { return {v[0], v[1], v[2]}; }
inline qilqr::Vec3 from_proto(const Vec3 &p) { return {p.c0, p.c1, p.c2}; }
inline Vec4 to_proto(const qilqr::Vec4 &v) { return {v[0], v[1], v[2], v[3]}; }
inline qilqr::Vec4 from_proto(const Vec4 &p) { return {p.c0, p.c1, p.c2, p.c3}; }
inline Vec6 to_proto(const qilqr::Vec6 &v) { return {v[0], v[1], v[2], v[3], v[4], v[5]}; }
inline qilqr::Vec6 from_proto(const Vec6 &p) { return {p.c0, p.c1, p.c2, p.c3, p.c4, p.c5}; }
// trajectory_to_proto.cc:67-96: the message holds w, x, y, z; qilqr::SE3 holds x, y, z, w
inline SE3 to_proto(const qilqr::SE3 &X) {
  SE3 p;
  p.translation = to_proto(X.translation);
  p.rotation = SO3{Vec4{X.quaternion[3], X.quaternion[0], X.quaternion[1], X.quaternion[2]}};
  return p;
}
inline qilqr::SE3 from_proto(const SE3 &p) {
  qilqr::SE3 X;
  X.translation = from_proto(p.translation.value_or(Vec3{}));
  const Vec4 q = p.rotation.value_or(SO3{}).quaternion.value_or(Vec4{});
  // the reference builds a manif::SO3d here (trajectory_to_proto.cc:76-83), which rejects a quaternion that is not
  // normalised -- e.g. the all-zero one of a message without a rotation.  (manif's own tolerance is its eps on
  // | |q|^2 - 1 |; this one is deliberately looser so that accumulated rounding is not rejected.)
  const double n2 = q.c0 * q.c0 + q.c1 * q.c1 + q.c2 * q.c2 + q.c3 * q.c3;
  if (!(std::fabs(n2 - 1.0) <= 1e-9)) throw std::invalid_argument("SO3 assigned data not normalized !");
  X.quaternion = {q.c1, q.c2, q.c3, q.c0};
  return X;
}
inline QuadrotorState to_proto(const QuadrotorModel::State &s) {
  QuadrotorState p;
  p.inertial_from_body = to_proto(s.inertial_from_body);
  p.body_velocity = to_proto(s.body_velocity);
  return p;
}
inline QuadrotorModel::State from_proto(const QuadrotorState &p) {
  QuadrotorModel::State s;
  s.inertial_from_body = from_proto(p.inertial_from_body.value_or(SE3{}));
  s.body_velocity = from_proto(p.body_velocity.value_or(Vec6{}));
  return s;
}
inline QuadrotorTrajectoryPoint to_proto(const TrajectoryPoint<QuadrotorModel> &pt) {
  QuadrotorTrajectoryPoint p;
  p.time_s = pt.time_s;
  p.state = to_proto(pt.state);
  p.control = to_proto(pt.control);
  return p;
}
inline TrajectoryPoint<QuadrotorModel> from_proto(const QuadrotorTrajectoryPoint &p) {
  return {p.time_s, from_proto(p.state.value_or(QuadrotorState{})), from_proto(p.control.value_or(Vec4{}))};
}
inline QuadrotorTrajectory to_proto(const Trajectory<QuadrotorModel> &traj) {
  QuadrotorTrajectory p;
  for (const auto &pt : traj) p.points.push_back(to_proto(pt));
  return p;
}
inline Trajectory<QuadrotorModel> from_proto(const QuadrotorTrajectory &p) {
  Trajectory<QuadrotorModel> t;
  for (const auto &pt : p.points) t.push_back(from_proto(pt));
  return t;
}
inline LineSearchParams to_proto(const qilqr::LineSearchParams &v) { return {v.step_update, v.desired_reduction_frac, v.max_iters}; }
inline qilqr::LineSearchParams from_proto(const LineSearchParams &p) { return {p.step_update, p.desired_reduction_frac, p.max_iters}; }
inline ConvergenceCriteria to_proto(const qilqr::ConvergenceCriteria &v) { return {v.rtol, v.atol, v.max_iters}; }
inline qilqr::ConvergenceCriteria from_proto(const ConvergenceCriteria &p) { return {p.rtol, p.atol, p.max_iters}; }
inline ILQROptions to_proto(const qilqr::ILQROptions &o) {
  ILQROptions p;
  p.line_search_params = to_proto(o.line_search_params);
  p.convergence_criteria = to_proto(o.convergence_criteria);
  p.populate_debug = o.populate_debug;
  return p;
}
inline qilqr::ILQROptions from_proto(const ILQROptions &p) {
  qilqr::ILQROptions o;
  o.line_search_params = from_proto(p.line_search_params.value_or(LineSearchParams{}));
  o.convergence_criteria = from_proto(p.convergence_criteria.value_or(ConvergenceCriteria{}));
  o.populate_debug = p.populate_debug;
  return o;
}
inline QuadrotorILQRIterDebug to_proto(const ILQRIterDebug<QuadrotorModel> &d) {
  QuadrotorILQRIterDebug p;
  p.trajectory = to_proto(d.trajectory);
  p.cost = d.cost;
  return p;
}
inline ILQRIterDebug<QuadrotorModel> from_proto(const QuadrotorILQRIterDebug &p) {
  return {from_proto(p.trajectory.value_or(QuadrotorTrajectory{})), p.cost};
}
inline QuadrotorILQRDebug to_proto(const ILQRDebug<QuadrotorModel> &debug) {
  QuadrotorILQRDebug p;
  for (const auto &d : debug) p.iter_debugs.push_back(to_proto(d));
  return p;
}
inline ILQRDebug<QuadrotorModel> from_proto(const QuadrotorILQRDebug &p) {
  ILQRDebug<QuadrotorModel> debug;
  for (const auto &d : p.iter_debugs) debug.push_back(from_proto(d));
  return debug;
}

}  // namespace proto

// The reference's pybind entry point with its wire-level signature (quadrotor_ilqr_binding.cc:20-49 hands
// protobuf messages across the language boundary; serialised bytes are the portable form of that):
// ILQROptions / QuadrotorTrajectory bytes in, (QuadrotorTrajectory, QuadrotorILQRDebug) bytes out.
class QuadrotorILQR {
 public:
  QuadrotorILQR(double mass_kg, const Mat3 &inertia, double arm_length_m, double torque_to_thrust_ratio_m, double g_mpss,
                const Mat12 &Q, const Mat4 &R, const std::string &desired_traj_bytes, double dt_s,
                const std::string &options_bytes)
      : ilqr_(QuadrotorModel{mass_kg, inertia, arm_length_m, torque_to_thrust_ratio_m, g_mpss},
              CostFunction<QuadrotorModel>{Q, R, parse_traj(desired_traj_bytes)}, dt_s, parse_options(options_bytes)) {}
  std::pair<std::string, std::string> solve(const std::string &initial_traj_bytes) const {
    const auto [traj, debug] = ilqr_.solve(parse_traj(initial_traj_bytes));
    return {proto::SerializeAsString(proto::to_proto(traj)), proto::SerializeAsString(proto::to_proto(debug))};
  }

 private:
  static Trajectory<QuadrotorModel> parse_traj(const std::string &bytes) {
    proto::QuadrotorTrajectory p;
    if (!proto::ParseFromString(bytes, &p)) throw std::invalid_argument("malformed QuadrotorTrajectory");
    return proto::from_proto(p);
  }
  static ILQROptions parse_options(const std::string &bytes) {
    proto::ILQROptions p;
    if (!proto::ParseFromString(bytes, &p)) throw std::invalid_argument("malformed ILQROptions");
    return proto::from_proto(p);
  }
  ILQR<QuadrotorModel> ilqr_;
};

}  // namespace qilqr
