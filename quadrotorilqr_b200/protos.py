"""Wire-compatible protobuf messages of the reference (``src/trajectory.proto``,
``src/ilqr_options.proto``, ``src/ilqr_debug.proto``; package ``src.proto``), built from programmatic
descriptors because the image has the Python protobuf runtime but no ``protoc``.

Field names and numbers are the reference's, so bytes serialised by either side parse on the other.
``trajectory_pb2``, ``ilqr_options_pb2`` and ``ilqr_debug_pb2`` stand in for the generated modules the
reference script imports (``src/quadrotor_ilqr.py:14-15``).
"""
from __future__ import annotations

import types

import numpy as np
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

_D = descriptor_pb2.FieldDescriptorProto
_POOL = descriptor_pool.DescriptorPool()


def _msg(file, name, fields):
    m = file.message_type.add()
    m.name = name
    for number, (fname, ftype, label, type_name) in enumerate(fields, start=1):
        f = m.field.add()
        f.name, f.number, f.type, f.label = fname, number, ftype, label
        if type_name:
            f.type_name = ".src.proto." + type_name
    return m


def _dbl(name):
    return (name, _D.TYPE_DOUBLE, _D.LABEL_OPTIONAL, None)


def _sub(name, type_name, repeated=False):
    return (name, _D.TYPE_MESSAGE, _D.LABEL_REPEATED if repeated else _D.LABEL_OPTIONAL, type_name)


def _build():
    traj = descriptor_pb2.FileDescriptorProto(name="src/trajectory.proto", package="src.proto", syntax="proto3")
    _msg(traj, "Vec3", [_dbl("c0"), _dbl("c1"), _dbl("c2")])
    _msg(traj, "Vec4", [_dbl("c0"), _dbl("c1"), _dbl("c2"), _dbl("c3")])
    _msg(traj, "Vec6", [_dbl(f"c{i}") for i in range(6)])
    _msg(traj, "SO3", [_sub("quaternion", "Vec4")])  # coefficients w, x, y, z (trajectory.proto:27-30)
    _msg(traj, "SE3", [_sub("translation", "Vec3"), _sub("rotation", "SO3")])
    _msg(traj, "QuadrotorState", [_sub("inertial_from_body", "SE3"), _sub("body_velocity", "Vec6")])
    _msg(traj, "QuadrotorTrajectoryPoint", [_dbl("time_s"), _sub("state", "QuadrotorState"), _sub("control", "Vec4")])
    _msg(traj, "QuadrotorTrajectory", [_sub("points", "QuadrotorTrajectoryPoint", repeated=True)])

    opts = descriptor_pb2.FileDescriptorProto(name="src/ilqr_options.proto", package="src.proto", syntax="proto3")
    _msg(opts, "LineSearchParams", [_dbl("step_update"), _dbl("desired_reduction_frac"),
                                    ("max_iters", _D.TYPE_INT32, _D.LABEL_OPTIONAL, None)])
    _msg(opts, "ConvergenceCriteria", [_dbl("rtol"), _dbl("atol"), _dbl("max_iters")])
    _msg(opts, "ILQROptions", [_sub("line_search_params", "LineSearchParams"),
                               _sub("convergence_criteria", "ConvergenceCriteria"),
                               ("populate_debug", _D.TYPE_BOOL, _D.LABEL_OPTIONAL, None)])

    dbg = descriptor_pb2.FileDescriptorProto(name="src/ilqr_debug.proto", package="src.proto", syntax="proto3",
                                             dependency=["src/trajectory.proto"])
    _msg(dbg, "QuadrotorILQRIterDebug", [_sub("trajectory", "QuadrotorTrajectory"), _dbl("cost")])
    _msg(dbg, "QuadrotorILQRDebug", [_sub("iter_debugs", "QuadrotorILQRIterDebug", repeated=True)])

    for f in (traj, opts, dbg):
        _POOL.Add(f)

    def module(file, names):
        ns = types.SimpleNamespace()
        for n in names:
            setattr(ns, n, message_factory.GetMessageClass(_POOL.FindMessageTypeByName("src.proto." + n)))
        return ns

    return (module(traj, ["Vec3", "Vec4", "Vec6", "SO3", "SE3", "QuadrotorState", "QuadrotorTrajectoryPoint",
                          "QuadrotorTrajectory"]),
            module(opts, ["LineSearchParams", "ConvergenceCriteria", "ILQROptions"]),
            module(dbg, ["QuadrotorILQRIterDebug", "QuadrotorILQRDebug"]))


trajectory_pb2, ilqr_options_pb2, ilqr_debug_pb2 = _build()


# ---- converters (the roles of trajectory_to_proto.cc / ilqr_options_to_proto.cc / ilqr_debug_to_proto.cc) ----
def trajectory_from_proto(msg) -> np.ndarray:
    """QuadrotorTrajectory -> [N, 18] in the C-ABI layout (quaternion x,y,z,w).
    The proto stores the quaternion as c0..c3 = w,x,y,z (trajectory_to_proto.cc:67-83)."""
    out = np.zeros((len(msg.points), 18))
    for i, pt in enumerate(msg.points):
        t, q, v, u = (pt.state.inertial_from_body.translation, pt.state.inertial_from_body.rotation.quaternion,
                      pt.state.body_velocity, pt.control)
        # the reference builds a manif::SO3d here (trajectory_to_proto.cc:76-83), which rejects a quaternion that is
        # not normalised -- e.g. the all-zero one of a point without a rotation (tolerance looser than manif's eps)
        if not abs(q.c0 * q.c0 + q.c1 * q.c1 + q.c2 * q.c2 + q.c3 * q.c3 - 1.0) <= 1e-9:
            raise ValueError(f"SO3 assigned data not normalized ! (trajectory point {i})")
        out[i] = [pt.time_s, t.c0, t.c1, t.c2, q.c1, q.c2, q.c3, q.c0, v.c0, v.c1, v.c2, v.c3, v.c4, v.c5,
                  u.c0, u.c1, u.c2, u.c3]
    return out


def trajectory_to_proto(arr):
    tp = trajectory_pb2
    msg = tp.QuadrotorTrajectory()
    for r in np.asarray(arr, dtype=np.float64):
        pt = msg.points.add()
        pt.time_s = r[0]
        pt.state.inertial_from_body.translation.CopyFrom(tp.Vec3(c0=r[1], c1=r[2], c2=r[3]))
        pt.state.inertial_from_body.rotation.quaternion.CopyFrom(tp.Vec4(c0=r[7], c1=r[4], c2=r[5], c3=r[6]))
        pt.state.body_velocity.CopyFrom(tp.Vec6(c0=r[8], c1=r[9], c2=r[10], c3=r[11], c4=r[12], c5=r[13]))
        pt.control.CopyFrom(tp.Vec4(c0=r[14], c1=r[15], c2=r[16], c3=r[17]))
    return msg


def options_from_proto(msg):
    from .options import ConvergenceCriteria, ILQROptions, LineSearchParams

    ls, cc = msg.line_search_params, msg.convergence_criteria
    return ILQROptions(LineSearchParams(ls.step_update, ls.desired_reduction_frac, ls.max_iters),
                       ConvergenceCriteria(cc.rtol, cc.atol, cc.max_iters), populate_debug=msg.populate_debug)


def options_to_proto(o):
    op = ilqr_options_pb2
    return op.ILQROptions(
        line_search_params=op.LineSearchParams(step_update=o.line_search_params.step_update,
                                               desired_reduction_frac=o.line_search_params.desired_reduction_frac,
                                               max_iters=o.line_search_params.max_iters),
        convergence_criteria=op.ConvergenceCriteria(rtol=o.convergence_criteria.rtol, atol=o.convergence_criteria.atol,
                                                    max_iters=o.convergence_criteria.max_iters),
        populate_debug=o.populate_debug)


def debug_to_proto(trajs, costs):
    msg = ilqr_debug_pb2.QuadrotorILQRDebug()
    for tr, c in zip(trajs, costs):
        it = msg.iter_debugs.add()
        it.trajectory.CopyFrom(trajectory_to_proto(tr))
        it.cost = float(c)
    return msg


def debug_from_proto(msg):
    return [trajectory_from_proto(it.trajectory) for it in msg.iter_debugs], [it.cost for it in msg.iter_debugs]
