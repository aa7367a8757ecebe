"""The hand-written C++ proto3 codec (quadrotorilqr_b200/cpp/qilqr_proto.hpp) against the protobuf runtime.

The reference's boundary carries protobuf messages (src/*.proto; quadrotor_ilqr_binding.cc:20-49).  The
image has no protoc / libprotobuf, so the C++ side has its own codec; these CPU tests check that it is
wire-compatible, byte for byte and in both directions, with the Python protobuf runtime working from the
same message definitions (quadrotorilqr_b200/protos.py), and port the reference's three round-trip tests
(trajectory_to_proto_test.cc, ilqr_options_to_proto_test.cc, ilqr_debug_to_proto_test.cc).
"""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "proto_tests.cc")
EXE = os.path.join(ROOT, "tests", "cpp", "proto_tests")


@pytest.fixture(scope="module")
def exe():
    from quadrotorilqr_b200 import _capi

    _capi.build()
    lib_dir = os.path.dirname(_capi.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-o", EXE, SRC, "-L" + lib_dir, "-lqilqr_b200",
                           "-Wl,-rpath," + lib_dir, "-L/usr/local/cuda/lib64", "-lcudart"])
    return EXE


def test_reference_round_trip_tests(exe):
    out = subprocess.run([exe, "roundtrip"], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stdout + out.stderr
    assert " 0 failed" in out.stdout


def sample_messages():
    """The same three messages tests/cpp/proto_tests.cc `emit`s."""
    from quadrotorilqr_b200 import protos
    from quadrotorilqr_b200.options import ConvergenceCriteria, ILQROptions, LineSearchParams

    traj = np.array([[1.0, 1, 2, 3, 0.5, -0.5, 0.5, 0.5, 2, 3, 4, 5, 6, 7, 3, 4, 5, 6],
                     [2.0, 2, 3, 4, 0.0, 0.6, 0.0, 0.8, 3, 4, 5, 6, 7, 8, 4, 5, 6, 7]])
    opts = ILQROptions(LineSearchParams(1.0, 2.0, 3), ConvergenceCriteria(4.0, 5.0, 6.0), populate_debug=True)
    ident = np.zeros((1, 18))
    ident[0, 7] = 1.0
    ident1 = ident.copy()
    ident1[0, 0] = 1.0
    return dict(trajectory=protos.trajectory_to_proto(traj), options=protos.options_to_proto(opts),
                debug=protos.debug_to_proto([ident, ident1], [23.3, 5.0]))


@pytest.mark.parametrize("kind", ["trajectory", "options", "debug"])
def test_cpp_bytes_equal_protobuf_runtime_bytes(exe, kind, tmp_path):
    out = tmp_path / "cpp.bin"
    assert subprocess.run([exe, "emit", kind, str(out)]).returncode == 0
    msg = sample_messages()[kind]
    assert out.read_bytes() == msg.SerializeToString(deterministic=True)
    parsed = type(msg)()
    parsed.ParseFromString(out.read_bytes())
    assert parsed == msg


def test_cpp_parses_and_reproduces_runtime_bytes(exe, tmp_path):
    """Random trajectories / debug streams (incl. zeros, -0.0, inf, denormals) written by the protobuf runtime
    parse in C++ and serialise back to the same bytes."""
    from quadrotorilqr_b200 import protos

    rng = np.random.default_rng(0)
    traj = rng.normal(size=(50, 18))
    traj[3] = 0.0
    traj[4, 2:6] = [-0.0, np.inf, 5e-324, -1e308]
    dbg = protos.debug_to_proto([traj[:7], traj[7:9], traj[:0]], [1.5, 0.0, -2.0])
    neg = protos.ilqr_options_pb2.ILQROptions(
        line_search_params=protos.ilqr_options_pb2.LineSearchParams(step_update=0.5, max_iters=-3))
    for kind, msg in (("trajectory", protos.trajectory_to_proto(traj)), ("debug", dbg), ("options", neg),
                      ("trajectory", protos.trajectory_pb2.QuadrotorTrajectory())):
        src, dst = tmp_path / "in.bin", tmp_path / "out.bin"
        src.write_bytes(msg.SerializeToString(deterministic=True))
        assert subprocess.run([exe, "reencode", kind, str(src), str(dst)]).returncode == 0
        assert dst.read_bytes() == src.read_bytes()


def test_cpp_rejects_malformed_bytes(exe, tmp_path):
    src, dst = tmp_path / "in.bin", tmp_path / "out.bin"
    src.write_bytes(b"\x0a\xff\x01\x00")  # length prefix runs past the end
    assert subprocess.run([exe, "reencode", "trajectory", str(src), str(dst)]).returncode == 2


@pytest.mark.gpu
def test_cpp_binding_solves_the_default_problem_from_proto_bytes(exe, tmp_path):
    """qilqr::QuadrotorILQR (C++, bytes in / bytes out) == the Python drop-in binding on quadrotor_ilqr.py's problem."""
    from quadrotorilqr_b200 import problems, protos
    from quadrotorilqr_b200.quadrotor_ilqr_binding import QuadrotorILQR

    m, opts = problems.default_model(), problems.default_options(True)
    desired = protos.trajectory_to_proto(problems.default_desired_trajectory())
    options = protos.options_to_proto(opts)
    files = {k: tmp_path / f"{k}.bin" for k in ("opts", "desired", "initial", "traj", "debug")}
    files["opts"].write_bytes(options.SerializeToString())
    files["desired"].write_bytes(desired.SerializeToString())
    files["initial"].write_bytes(desired.SerializeToString())
    rc = subprocess.run([exe, "solve"] + [str(files[k]) for k in ("opts", "desired", "initial", "traj", "debug")],
                        capture_output=True, text=True, timeout=300)
    assert rc.returncode == 0, rc.stderr
    traj_cpp = protos.trajectory_pb2.QuadrotorTrajectory()
    traj_cpp.ParseFromString(files["traj"].read_bytes())
    debug_cpp = protos.ilqr_debug_pb2.QuadrotorILQRDebug()
    debug_cpp.ParseFromString(files["debug"].read_bytes())
    py = QuadrotorILQR(m["mass_kg"], m["inertia"], m["arm_length_m"], m["torque_to_thrust_ratio_m"], m["g_mpss"],
                       m["Q"], m["R"], desired, m["dt_s"], options)
    traj_py, debug_py = py.solve(desired)
    assert traj_cpp == traj_py
    assert debug_cpp == debug_py
    assert len(debug_cpp.iter_debugs) == 76   # 76 completed iterations on the default problem


def test_cpp_parser_agrees_with_the_runtime_on_fuzzed_input(exe, tmp_path):
    """Random bytes and mutated / truncated valid messages: the hand-written parser never crashes or hangs, it
    accepts exactly the inputs the protobuf runtime accepts, and what it re-encodes parses to the same message."""
    from google.protobuf.message import DecodeError

    from quadrotorilqr_b200 import protos

    rng = np.random.default_rng(7)
    valid = sample_messages()["debug"].SerializeToString()
    cases = [bytes(rng.integers(0, 256, rng.integers(0, 64), dtype=np.uint8)) for _ in range(150)]
    for _ in range(250):
        b = bytearray(valid)
        for _ in range(rng.integers(1, 4)):
            b[rng.integers(0, len(b))] = rng.integers(0, 256)
        cases.append(bytes(b[: rng.integers(1, len(b) + 1)]))
    src, dst = tmp_path / "in.bin", tmp_path / "out.bin"
    accepted = 0
    for raw in cases:
        src.write_bytes(raw)
        rc = subprocess.run([exe, "reencode", "debug", str(src), str(dst)], timeout=20).returncode
        assert rc in (0, 2), (rc, raw.hex())
        try:
            msg = protos.ilqr_debug_pb2.QuadrotorILQRDebug.FromString(raw)
        except DecodeError:
            msg = None
        assert (rc == 0) == (msg is not None), raw.hex()
        if msg is not None:
            accepted += 1
            msg.DiscardUnknownFields()
            assert protos.ilqr_debug_pb2.QuadrotorILQRDebug.FromString(dst.read_bytes()) == msg
    assert accepted > 0
