"""Proto layer: wire compatibility with the reference's .proto files and the round trips
the reference tests pin (trajectory_to_proto_test.cc:13-37, ilqr_options_to_proto_test.cc:7-18,
ilqr_debug_to_proto_test.cc:30-40)."""
import struct

import numpy as np
import pytest


def test_wire_format_matches_proto3_rules():
    from quadrotorilqr_b200 import protos

    tp, op = protos.trajectory_pb2, protos.ilqr_options_pb2
    # double = fixed64 with tag (field << 3) | 1; zero-valued scalars are omitted
    assert tp.Vec3(c0=1.0).SerializeToString() == b"\x09" + struct.pack("<d", 1.0)
    assert tp.Vec3(c1=-2.5, c2=0.0).SerializeToString() == b"\x11" + struct.pack("<d", -2.5)
    # nested message = length-delimited; int32 / bool = varint
    ls = op.LineSearchParams(step_update=0.5, desired_reduction_frac=0.5, max_iters=100)
    assert ls.SerializeToString() == b"\x09" + struct.pack("<d", 0.5) + b"\x11" + struct.pack("<d", 0.5) + b"\x18\x64"
    o = op.ILQROptions(line_search_params=ls, populate_debug=True)
    raw = o.SerializeToString()
    assert raw[:2] == b"\x0a" + bytes([len(ls.SerializeToString())]) and raw.endswith(b"\x18\x01")
    # descriptors carry the reference's package and field numbers
    d = tp.QuadrotorTrajectoryPoint.DESCRIPTOR
    assert d.full_name == "src.proto.QuadrotorTrajectoryPoint"
    assert [(f.name, f.number) for f in d.fields] == [("time_s", 1), ("state", 2), ("control", 3)]
    assert protos.ilqr_debug_pb2.QuadrotorILQRDebug.DESCRIPTOR.fields[0].name == "iter_debugs"


def test_trajectory_round_trip_and_quaternion_order():
    from quadrotorilqr_b200 import problems, protos

    t = problems.default_desired_trajectory()
    t[:, 8:14] = np.random.default_rng(0).uniform(-1, 1, (40, 6))
    t[:, 14:] = np.random.default_rng(1).uniform(-1, 1, (40, 4))
    msg = protos.trajectory_to_proto(t)
    # proto quaternion is w-first (trajectory.proto:27-30), the C-ABI layout is (x, y, z, w)
    q = msg.points[15].state.inertial_from_body.rotation.quaternion
    assert (q.c0, q.c1) == (t[15, 7], t[15, 4])
    back = protos.trajectory_from_proto(protos.trajectory_pb2.QuadrotorTrajectory.FromString(msg.SerializeToString()))
    assert np.array_equal(back, t)
    # and the reference script's own 18-column layout (quadrotor_ilqr.py:19-65) is a permutation of ours
    assert np.array_equal(problems.from_idx_layout(problems.to_idx_layout(t)), t)
    assert problems.to_idx_layout(t)[15, problems.IDX.quaternion_w] == t[15, 7]


def test_options_and_debug_round_trip():
    from quadrotorilqr_b200 import problems, protos

    o = problems.default_options(True)
    assert protos.options_from_proto(protos.options_to_proto(o)) == o
    trajs = [problems.default_desired_trajectory() * s for s in (1.0, 0.5)]
    trajs[1][:, 4:8] = trajs[0][:, 4:8]  # (quaternions stay normalised: the converters reject others, as manif does)
    msg = protos.debug_to_proto(trajs, [3.0, 1.5])
    tr, costs = protos.debug_from_proto(protos.ilqr_debug_pb2.QuadrotorILQRDebug.FromString(msg.SerializeToString()))
    assert costs == [3.0, 1.5] and all(np.array_equal(a, b) for a, b in zip(tr, trajs))


@pytest.mark.gpu
def test_reference_demo_flow_through_the_binding(O):
    """quadrotor_ilqr.py:256-306 (main()) with the drop-in QuadrotorILQR: same protos in, same protos out."""
    from conftest import oracle_config
    from quadrotorilqr_b200 import problems, protos
    from quadrotorilqr_b200.quadrotor_ilqr_binding import QuadrotorILQR

    traj, opts = protos.trajectory_pb2, protos.ilqr_options_pb2
    desired = problems.default_desired_trajectory()
    desired_traj = protos.trajectory_to_proto(desired)
    options = opts.ILQROptions(
        line_search_params=opts.LineSearchParams(step_update=0.5, desired_reduction_frac=0.5, max_iters=100),
        convergence_criteria=opts.ConvergenceCriteria(rtol=1e-12, atol=1e-12, max_iters=100), populate_debug=True)
    Q = np.diag(np.concatenate((100 * np.ones(6), 1 * np.ones(6))))
    ilqr = QuadrotorILQR(1.0, np.eye(3), 1.0, 0.0, 9.81, Q, np.eye(4), desired_traj, 0.1, options)
    opt_traj, debug = ilqr.solve(desired_traj)
    assert isinstance(opt_traj, traj.QuadrotorTrajectory) and len(opt_traj.points) == 40
    costs = [d.cost for d in debug.iter_debugs]  # quadrotor_ilqr.py:312
    o = O.solve(oracle_config(O, problems.default_model(), problems.default_options(True)), desired, desired)
    assert len(costs) == 76 == o["num_debug"]
    assert np.allclose(costs, o["cost_history"], rtol=1e-9, atol=0)
    got = protos.trajectory_from_proto(opt_traj)
    assert np.max(np.abs(got - o["traj"])) <= 1e-9 * np.max(np.abs(o["traj"]))
    last = protos.trajectory_from_proto(debug.iter_debugs[-1].trajectory)
    assert np.max(np.abs(last - o["debug"][-1])) <= 1e-9 * np.max(np.abs(o["debug"][-1]))
    with pytest.raises(IndexError):  # cost.hh:39-40
        QuadrotorILQR(1.0, np.eye(3), 1.0, 0.0, 9.81, Q, np.eye(4), protos.trajectory_to_proto(desired[:10]), 0.1,
                      options).solve(desired_traj)


def test_unnormalised_quaternions_are_rejected():
    """from_proto builds a manif::SO3d in the reference (trajectory_to_proto.cc:76-83), which rejects a quaternion
    that is not normalised -- e.g. the all-zero one of a message without a rotation."""
    from quadrotorilqr_b200 import problems, protos

    good = protos.trajectory_to_proto(problems.default_desired_trajectory())
    protos.trajectory_from_proto(good)
    bad = protos.trajectory_pb2.QuadrotorTrajectory()
    bad.points.add().time_s = 1.0  # no state at all: quaternion (0, 0, 0, 0)
    with pytest.raises(ValueError, match="not normalized"):
        protos.trajectory_from_proto(bad)
    scaled = protos.trajectory_pb2.QuadrotorTrajectory()
    scaled.CopyFrom(good)
    scaled.points[3].state.inertial_from_body.rotation.quaternion.c0 *= 1.001
    with pytest.raises(ValueError, match="not normalized"):
        protos.trajectory_from_proto(scaled)
