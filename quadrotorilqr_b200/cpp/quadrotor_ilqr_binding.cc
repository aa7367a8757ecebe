// =============================================================================
// quadrotor_ilqr_binding.cc -- the reference's pybind11 module (src/quadrotor_ilqr_binding.cc:20-49),
// same module name, class name, constructor arguments and solve() signature, over the CUDA library:
//
//   QuadrotorILQR(mass_kg, inertia 3x3, arm_length_m, torque_to_thrust_ratio_m, g_mpss, Q 12x12, R 4x4,
//                 desired_traj: QuadrotorTrajectory, dt_s, options: ILQROptions)
//   .solve(initial_traj: QuadrotorTrajectory) -> (QuadrotorTrajectory, QuadrotorILQRDebug)
//
// The reference moves protobuf messages across the language boundary with pybind11_protobuf's native
// casters; that library (and libprotobuf) is not in this image, so a message crosses as its serialised
// bytes -- msg.SerializeToString() on the way in, Class.FromString() on the way out -- and the C++ side
// uses the codec of qilqr_proto.hpp.  The returned messages are instances of the classes that belong
// to the caller's own descriptor pool (the reference's generated *_pb2 modules or
// quadrotorilqr_b200.protos).  Unlike the reference (binding.cc:34-41) the GIL is released while the
// GPU solves.  solve_batch() is the batched form (not in the reference).
// =============================================================================
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include "qilqr_proto.hpp"

namespace py = pybind11;
using namespace qilqr;
using DenseArray = py::array_t<double, py::array::c_style | py::array::forcecast>;

namespace {
template <size_t N>
std::array<double, N> to_array(const DenseArray &a, const char *what) {
  if (size_t(a.size()) != N) throw std::invalid_argument(std::string(what) + ": wrong number of elements");
  std::array<double, N> out;
  std::copy(a.data(), a.data() + N, out.begin());
  return out;
}
std::string bytes_of(const py::handle &msg) { return msg.attr("SerializeToString")().cast<std::string>(); }
template <class M>
M parse(const py::handle &msg, const char *what) {
  M m;
  if (!proto::ParseFromString(bytes_of(msg), &m)) throw std::invalid_argument(std::string("malformed ") + what);
  return m;
}
// the Python class of message `full_name` in the descriptor pool `like` comes from
py::object message_class(const py::handle &like, const char *full_name) {
  py::object pool = like.attr("DESCRIPTOR").attr("file").attr("pool");
  py::object factory = py::module_::import("google.protobuf.message_factory");
  try {
    return factory.attr("GetMessageClass")(pool.attr("FindMessageTypeByName")(full_name));
  } catch (py::error_already_set &) {  // the caller's pool does not hold that file: use the package's own classes
    py::object protos = py::module_::import("quadrotorilqr_b200.protos");
    return protos.attr("ilqr_debug_pb2").attr("QuadrotorILQRDebug");
  }
}

class PyQuadrotorILQR {
 public:
  PyQuadrotorILQR(double mass_kg, const DenseArray &inertia, double arm_length_m, double torque_to_thrust_ratio_m,
                  double g_mpss, const DenseArray &Q, const DenseArray &R, const py::object &desired_traj, double dt_s,
                  const py::object &options)
      : ilqr_(QuadrotorModel{mass_kg, to_array<9>(inertia, "inertia"), arm_length_m, torque_to_thrust_ratio_m, g_mpss},
              CostFunction<QuadrotorModel>{to_array<144>(Q, "Q"), to_array<16>(R, "R"),
                                           proto::from_proto(parse<proto::QuadrotorTrajectory>(desired_traj, "QuadrotorTrajectory"))},
              dt_s, proto::from_proto(parse<proto::ILQROptions>(options, "ILQROptions"))) {}

  py::tuple solve(const py::object &initial_traj) const {
    const auto initial = proto::from_proto(parse<proto::QuadrotorTrajectory>(initial_traj, "QuadrotorTrajectory"));
    if (initial.empty()) throw std::invalid_argument("empty trajectory");  // undefined behaviour in the reference (ilqr.hh:156)
    std::string traj_bytes, debug_bytes;
    {
      py::gil_scoped_release release;
      const auto [traj, debug] = ilqr_.solve(initial);
      traj_bytes = proto::SerializeAsString(proto::to_proto(traj));
      debug_bytes = proto::SerializeAsString(proto::to_proto(debug));
    }
    py::object traj_cls = py::type::of(initial_traj);
    py::object debug_cls = message_class(initial_traj, "src.proto.QuadrotorILQRDebug");
    return py::make_tuple(traj_cls.attr("FromString")(py::bytes(traj_bytes)),
                          debug_cls.attr("FromString")(py::bytes(debug_bytes)));
  }

  // Many problems per call: a list of QuadrotorTrajectory messages of equal length ->
  // (list of QuadrotorTrajectory, list of (status, backward_passes, rollouts, final_cost))
  py::tuple solve_batch(const py::list &initial_trajs) const {
    std::vector<Trajectory<QuadrotorModel>> initial;
    for (const auto &t : initial_trajs)
      initial.push_back(proto::from_proto(parse<proto::QuadrotorTrajectory>(t, "QuadrotorTrajectory")));
    std::vector<qilqr_result_t> res;
    std::vector<std::string> out_bytes;
    {
      py::gil_scoped_release release;
      const auto sols = ilqr_.solve_batch(initial, &res);
      for (const auto &s : sols) out_bytes.push_back(proto::SerializeAsString(proto::to_proto(s)));
    }
    py::list trajs, results;
    for (size_t b = 0; b < out_bytes.size(); ++b) {
      trajs.append(py::type::of(initial_trajs[b]).attr("FromString")(py::bytes(out_bytes[b])));
      results.append(py::make_tuple(res[b].status, res[b].backward_passes, res[b].rollouts, res[b].final_cost));
    }
    return py::make_tuple(trajs, results);
  }

 private:
  ILQR<QuadrotorModel> ilqr_;
};
}  // namespace

PYBIND11_MODULE(quadrotor_ilqr_binding, m) {
  m.doc() = "Drop-in for the reference's src.quadrotor_ilqr_binding, backed by libqilqr_b200.so (sm_100a)";
  py::class_<PyQuadrotorILQR>(m, "QuadrotorILQR")
      .def(py::init<double, const DenseArray &, double, double, double, const DenseArray &, const DenseArray &,
                    const py::object &, double, const py::object &>())
      .def("solve", &PyQuadrotorILQR::solve)
      .def("solve_batch", &PyQuadrotorILQR::solve_batch);
}
