"""ctypes declarations for ``libqilqr_b200.so`` (the C ABI of ``include/qilqr.h``).

The library is built in-tree by ``quadrotorilqr_b200/csrc/Makefile`` (see
``__graft_entry__.build``).  There is no CPU fallback: if the shared library is
missing, or no sm_100 device is present, the product raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QILQR_LIB") or os.path.join(_HERE, "libqilqr_b200.so")  # QILQR_LIB: tuning builds

OK = 0
ERR_INVALID_ARGUMENT = 1
ERR_INERTIA_NOT_PD = 2
ERR_OUT_OF_RANGE = 3
ERR_NO_DEVICE = 4
ERR_CUDA = 5
ERR_OUT_OF_MEMORY = 6
ERR_LINE_SEARCH = 7

STATUS_NOT_RUN = 0
STATUS_CONVERGED_EXPECTED = 1
STATUS_CONVERGED_ACTUAL = 2
STATUS_MAX_ITERS = 3
STATUS_LINE_SEARCH_FAILED = 4
STATUS_NONFINITE = 5


class Model(C.Structure):
    _fields_ = [
        ("mass_kg", C.c_double),
        ("inertia", C.c_double * 9),
        ("arm_length_m", C.c_double),
        ("torque_to_thrust_ratio_m", C.c_double),
        ("g_mpss", C.c_double),
    ]


class Options(C.Structure):
    _fields_ = [
        ("step_update", C.c_double),
        ("desired_reduction_frac", C.c_double),
        ("line_search_max_iters", C.c_int32),
        ("populate_debug", C.c_int32),
        ("rtol", C.c_double),
        ("atol", C.c_double),
        ("max_iters", C.c_double),
        ("symmetrize_vxx", C.c_int32),
        ("num_parallel_alphas", C.c_int32),
        ("quu_regularization", C.c_double),
    ]


class Result(C.Structure):
    _fields_ = [
        ("status", C.c_int32),
        ("backward_passes", C.c_int32),
        ("rollouts", C.c_int32),
        ("num_debug", C.c_int32),
        ("final_cost", C.c_double),
    ]


class SolveStats(C.Structure):
    _fields_ = [
        ("solver_iterations", C.c_int64),
        ("problem_iterations", C.c_int64),
        ("problem_rollouts", C.c_int64),
        ("kernel_launches", C.c_int64),
        ("backward_ms", C.c_double),
        ("rollout_ms", C.c_double),
        ("backward_problem_knots", C.c_int64),
        ("rollout_problem_knots", C.c_int64),
        ("bulk_wall_ms", C.c_double),
        ("tail_wall_ms", C.c_double),
        ("backward_ms_bulk", C.c_double),
        ("rollout_ms_bulk", C.c_double),
        ("backward_problem_knots_bulk", C.c_int64),
        ("backward_launches_bulk", C.c_int64),
    ]


# every symbol include/qilqr.h declares
EXPORTED_SYMBOLS = [
    "qilqr_create", "qilqr_destroy", "qilqr_set_options", "qilqr_error_string",
    "qilqr_last_error_message", "qilqr_kernel_launch_count", "qilqr_stream",
    "qilqr_solve_host", "qilqr_forward_sim_host", "qilqr_cost_trajectory_host",
    "qilqr_backwards_pass_host", "qilqr_line_search_host", "qilqr_discrete_dynamics_host",
    "qilqr_continuous_dynamics_host", "qilqr_state_minus_host", "qilqr_state_add_host",
    "qilqr_cost_host", "qilqr_solve_device", "qilqr_pack_trajectory_device",
    "qilqr_unpack_trajectory_device", "qilqr_rollout_constant_control_device",
    "qilqr_last_solve_stats", "qilqr_set_profiling", "qilqr_measure_fp64_peak",
    "qilqr_mpc_advance_device", "qilqr_mpc_run_device", "qilqr_check_model",
    "qilqr_set_model_variant", "qilqr_build_info", "qilqr_solve_host_begin", "qilqr_solve_host_finish",
    "qilqr_solve_device_begin", "qilqr_solve_device_finish", "qilqr_set_debug_sampling",
    "qilqr_read_debug_samples_host", "qilqr_last_cost_history_host", "qilqr_solve_from_controls_host",
    "qilqr_solve_from_controls_host_begin", "qilqr_set_user_model",
]

MODEL_REFERENCE = 0
MODEL_RK4 = 1
MODEL_CORIOLIS = 2
MODEL_GENERIC = 4


STRICT_LIB_PATH = os.path.join(_HERE, "libqilqr_b200_strict.so")  # -DQILQR_STRICT build (csrc/Makefile)
PLIBM_LIB_PATH = os.path.join(_HERE, "libqilqr_b200_strict_plibm.so")  # ... + portable sin/cos/atan2


def build(force: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ... (cross-compiles without a GPU).

    Builds the production library and the opt-in STRICT build (no fused multiply-adds, true divisions;
    load it with ``QILQR_LIB=quadrotorilqr_b200/libqilqr_b200_strict.so``)."""
    csrc = os.path.join(_HERE, "csrc")
    srcs = [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cu", ".cuh"))]
    srcs.append(os.path.join(_HERE, "..", "include", "qilqr.h"))
    targets = [os.path.join(_HERE, "libqilqr_b200.so"), STRICT_LIB_PATH, PLIBM_LIB_PATH]
    stale = force or any(not os.path.exists(t) or any(os.path.getmtime(s) > os.path.getmtime(t) for s in srcs)
                         for t in targets)
    if stale:
        r = subprocess.run(["make", "-C", csrc, "-B", "-j3", "all"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                           text=True)
        if r.returncode != 0:
            raise RuntimeError(f"building the CUDA library failed (make -C {csrc}, exit code {r.returncode}):\n"
                               + r.stdout[-4000:])
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    """Load the CUDA library; fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.qilqr_error_string.restype = C.c_char_p
        L.qilqr_build_info.restype = C.c_char_p
        L.qilqr_last_error_message.restype = C.c_char_p
        L.qilqr_kernel_launch_count.restype = C.c_int64
        L.qilqr_stream.restype = C.c_void_p
        for name in EXPORTED_SYMBOLS:
            fn = getattr(L, name)
            if name not in ("qilqr_error_string", "qilqr_last_error_message", "qilqr_build_info",
                            "qilqr_kernel_launch_count", "qilqr_stream", "qilqr_destroy"):
                fn.restype = C.c_int
        L.qilqr_destroy.restype = None
        _lib = L
    return _lib
