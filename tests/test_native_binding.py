"""The pybind11 build of the reference's ``quadrotor_ilqr_binding`` module
(quadrotorilqr_b200/cpp/quadrotor_ilqr_binding.cc mirroring src/quadrotor_ilqr_binding.cc:20-49)."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def qb():
    from quadrotorilqr_b200 import native

    native.build()
    from quadrotorilqr_b200.native import quadrotor_ilqr_binding

    return quadrotor_ilqr_binding


def default_args():
    from quadrotorilqr_b200 import problems, protos

    m, opts = problems.default_model(), problems.default_options(True)
    desired = protos.trajectory_to_proto(problems.default_desired_trajectory())
    return m, desired, protos.options_to_proto(opts)


def make(qb, m, desired, options):
    return qb.QuadrotorILQR(m["mass_kg"], m["inertia"], m["arm_length_m"], m["torque_to_thrust_ratio_m"], m["g_mpss"],
                            m["Q"], m["R"], desired, m["dt_s"], options)


def test_module_surface_and_argument_checks(qb):
    assert hasattr(qb, "QuadrotorILQR") and hasattr(qb.QuadrotorILQR, "solve")
    m, desired, options = default_args()
    with pytest.raises(ValueError):  # inertia must have 9 elements
        qb.QuadrotorILQR(1.0, np.eye(2), 1.0, 0.0, 9.81, m["Q"], m["R"], desired, 0.1, options)
    with pytest.raises(RuntimeError, match="[Ii]nertia"):  # quadrotor_model.cc:21-24
        qb.QuadrotorILQR(1.0, -np.eye(3), 1.0, 0.0, 9.81, m["Q"], m["R"], desired, 0.1, options)


@pytest.mark.gpu
def test_native_binding_equals_ctypes_binding_on_the_default_problem(qb):
    """quadrotor_ilqr.py:256-306 through both routes: identical protobuf messages out."""
    from quadrotorilqr_b200.quadrotor_ilqr_binding import QuadrotorILQR

    m, desired, options = default_args()
    traj_n, debug_n = make(qb, m, desired, options).solve(desired)
    traj_c, debug_c = QuadrotorILQR(m["mass_kg"], m["inertia"], m["arm_length_m"], m["torque_to_thrust_ratio_m"],
                                    m["g_mpss"], m["Q"], m["R"], desired, m["dt_s"], options).solve(desired)
    assert type(traj_n) is type(desired)
    assert traj_n == traj_c and debug_n == debug_c
    assert len(debug_n.iter_debugs) == 76 and abs(debug_n.iter_debugs[-1].cost - 22556.50259198) < 1e-5


@pytest.mark.gpu
def test_native_binding_errors_and_batch(qb):
    from quadrotorilqr_b200 import protos

    m, desired, options = default_args()
    ilqr = make(qb, m, desired, options)
    longer = protos.trajectory_pb2.QuadrotorTrajectory()
    longer.CopyFrom(desired)
    longer.points.add().CopyFrom(desired.points[-1])
    with pytest.raises(IndexError):  # std::out_of_range, cost.hh:39-40
        ilqr.solve(longer)
    with pytest.raises(ValueError):
        ilqr.solve(protos.trajectory_pb2.QuadrotorTrajectory())
    single, _ = ilqr.solve(desired)
    trajs, results = ilqr.solve_batch([desired, desired, desired])
    assert len(trajs) == 3 and all(t == single for t in trajs)
    assert all(r[0] in (1, 2) and r[1] == 77 for r in results)
