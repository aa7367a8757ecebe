"""CPU oracle for the batched quadrotor iLQR hot path -- TEST INFRASTRUCTURE ONLY.

ctypes front-end of ``oracle/libqilqr_oracle.so`` (built from
``qilqr_oracle_capi.cc`` / ``qilqr_oracle.hpp`` by ``oracle/Makefile``).  The C++
restates the reference's ``src/ilqr.hh``, ``src/quadrotor_model.{hh,cc}`` and
``src/cost.hh`` plus the manif / Eigen routines they call; see the header of
``qilqr_oracle.hpp`` for provenance and the parity status ("pinned" only by the
reference's toy known-answer tests; otherwise parity unpinned).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product
(``quadrotorilqr_b200``) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# QORACLE_LIB=plibm selects the build whose sin / cos / atan2 come from the portable implementation shared with
# the STRICT CUDA build (bit-for-bit comparisons only; oracle/Makefile)
_LIB_NAME = "libqilqr_oracle_plibm.so" if os.environ.get("QORACLE_LIB", "") == "plibm" else "libqilqr_oracle.so"
_LIB_PATH = os.path.join(_HERE, _LIB_NAME)

STATUS_CONVERGED_EXPECTED = 1
STATUS_CONVERGED_ACTUAL = 2
STATUS_MAX_ITERS = 3
STATUS_LINE_SEARCH_FAILED = 4
STATUS_NONFINITE = 5


class Config(C.Structure):
    """Mirror of qoracle_config_t."""

    _fields_ = [
        ("mass_kg", C.c_double),
        ("inertia", C.c_double * 9),
        ("arm_length_m", C.c_double),
        ("torque_to_thrust_ratio_m", C.c_double),
        ("g_mpss", C.c_double),
        ("Q", C.c_double * 144),
        ("R", C.c_double * 16),
        ("dt_s", C.c_double),
        ("step_update", C.c_double),
        ("desired_reduction_frac", C.c_double),
        ("rtol", C.c_double),
        ("atol", C.c_double),
        ("max_iters", C.c_double),
        ("quu_regularization", C.c_double),
        ("ls_max_iters", C.c_int32),
        ("populate_debug", C.c_int32),
        ("symmetrize_vxx", C.c_int32),
        ("model_kind", C.c_int32),
    ]


class Result(C.Structure):
    _fields_ = [
        ("status", C.c_int32),
        ("backward_passes", C.c_int32),
        ("rollouts", C.c_int32),
        ("num_debug", C.c_int32),
        ("final_cost", C.c_double),
    ]


RESULT_DTYPE = np.dtype(
    [("status", "<i4"), ("backward_passes", "<i4"), ("rollouts", "<i4"), ("num_debug", "<i4"),
     ("final_cost", "<f8")]
)


def build(force: bool = False) -> str:
    """Compile the oracle (g++ -O2 -ffp-contract=off).  Building is not using."""
    src = [os.path.join(_HERE, f) for f in ("qilqr_oracle_capi.cc", "qilqr_oracle.hpp")]
    src.append(os.path.join(_HERE, "..", "quadrotorilqr_b200", "csrc", "qilqr_portable_libm.h"))
    targets = [os.path.join(_HERE, n) for n in ("libqilqr_oracle.so", "libqilqr_oracle_plibm.so")]
    stale = force or any(not os.path.exists(t) or any(os.path.getmtime(s) > os.path.getmtime(t) for s in src)
                         for t in targets)
    if stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "-j2", "all"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.qoracle_cost.restype = C.c_double
    return _lib


def _p(a):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _in(a, shape=None):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    if shape is not None:
        a = a.reshape(shape)
    return a


def make_config(mass_kg=1.0, inertia=None, arm_length_m=1.0, torque_to_thrust_ratio_m=0.0,
                g_mpss=9.81, Q=None, R=None, dt_s=0.1, step_update=0.5, desired_reduction_frac=0.5,
                ls_max_iters=100, rtol=1e-12, atol=1e-12, max_iters=100.0, populate_debug=False,
                symmetrize_vxx=False, quu_regularization=0.0, model_kind=0) -> Config:
    c = Config()
    c.mass_kg = mass_kg
    c.inertia[:] = _in(np.eye(3) if inertia is None else inertia, (9,)).tolist()
    c.arm_length_m = arm_length_m
    c.torque_to_thrust_ratio_m = torque_to_thrust_ratio_m
    c.g_mpss = g_mpss
    c.Q[:] = _in(np.eye(12) if Q is None else Q, (144,)).tolist()
    c.R[:] = _in(np.eye(4) if R is None else R, (16,)).tolist()
    c.dt_s = dt_s
    c.step_update = step_update
    c.desired_reduction_frac = desired_reduction_frac
    c.ls_max_iters = int(ls_max_iters)
    c.rtol = rtol
    c.atol = atol
    c.max_iters = float(max_iters)
    c.populate_debug = int(bool(populate_debug))
    c.symmetrize_vxx = int(bool(symmetrize_vxx))
    c.quu_regularization = quu_regularization
    c.model_kind = int(model_kind)  # bit 0: RK4 integrator, bit 1: Coriolis term (QuadrotorModelVariant)
    return c


# ---- Lie primitives ---------------------------------------------------------------
def _unary(name, x, out_shape):
    x = _in(x)
    out = np.zeros(out_shape)
    getattr(lib(), name)(_p(x), _p(out))
    return out


def se3_exp(tau): return _unary("qoracle_se3_exp", tau, 7)
def se3_log(X): return _unary("qoracle_se3_log", X, 6)
def se3_inverse(X): return _unary("qoracle_se3_inverse", X, 7)
def se3_adj(X): return _unary("qoracle_se3_adj", X, (6, 6))
def se3_rjac(tau): return _unary("qoracle_se3_rjac", tau, (6, 6))
def se3_ljac(tau): return _unary("qoracle_se3_ljac", tau, (6, 6))
def se3_rjacinv(tau): return _unary("qoracle_se3_rjacinv", tau, (6, 6))
def se3_ljacinv(tau): return _unary("qoracle_se3_ljacinv", tau, (6, 6))
def rotation_matrix(q_xyzw): return _unary("qoracle_rotation_matrix", q_xyzw, (3, 3))


def se3_compose(A, B):
    A, B = _in(A), _in(B)
    out = np.zeros(7)
    lib().qoracle_se3_compose(_p(A), _p(B), _p(out))
    return out


def se3_plus(X, tau):
    X, tau = _in(X), _in(tau)
    out, JX, Jt = np.zeros(7), np.zeros((6, 6)), np.zeros((6, 6))
    lib().qoracle_se3_plus(_p(X), _p(tau), _p(out), _p(JX), _p(Jt))
    return out, JX, Jt


def se3_minus(A, B):
    A, B = _in(A), _in(B)
    t, JA, JB = np.zeros(6), np.zeros((6, 6)), np.zeros((6, 6))
    lib().qoracle_se3_minus(_p(A), _p(B), _p(t), _p(JA), _p(JB))
    return t, JA, JB


def ldlt4_solve(A, b):
    A = _in(A, (4, 4))
    b = _in(b)
    b2 = b.reshape(4, -1)
    x = np.zeros_like(b2)
    lib().qoracle_ldlt4_solve(_p(A), _p(np.ascontiguousarray(b2)), C.c_int(b2.shape[1]), _p(x))
    return x.reshape(b.shape)


# ---- model / cost -------------------------------------------------------------------
def check_model(cfg) -> int:
    return lib().qoracle_check_model(C.byref(cfg))


def continuous_dynamics(cfg, x, u, diffs=False):
    x, u = _in(x, (13,)), _in(u, (4,))
    xdot = np.zeros(12)
    Jx = np.zeros((12, 12)) if diffs else None
    Ju = np.zeros((12, 4)) if diffs else None
    rc = lib().qoracle_continuous_dynamics(C.byref(cfg), _p(x), _p(u), _p(xdot), _p(Jx), _p(Ju))
    assert rc == 0
    return (xdot, Jx, Ju) if diffs else xdot


def discrete_dynamics(cfg, x, u, dt_s=None, diffs=False):
    x, u = _in(x, (13,)), _in(u, (4,))
    out = np.zeros(13)
    Jx = np.zeros((12, 12)) if diffs else None
    Ju = np.zeros((12, 4)) if diffs else None
    rc = lib().qoracle_discrete_dynamics(C.byref(cfg), _p(x), _p(u),
                                         C.c_double(cfg.dt_s if dt_s is None else dt_s), _p(out),
                                         _p(Jx), _p(Ju))
    assert rc == 0
    return (out, Jx, Ju) if diffs else out


def state_add(x, tangent, diffs=False):
    x, tangent = _in(x, (13,)), _in(tangent, (12,))
    out = np.zeros(13)
    Jl = np.zeros((12, 12)) if diffs else None
    Jr = np.zeros((12, 12)) if diffs else None
    lib().qoracle_state_add(_p(x), _p(tangent), _p(out), _p(Jl), _p(Jr))
    return (out, Jl, Jr) if diffs else out


def state_minus(lhs, rhs, diffs=False):
    lhs, rhs = _in(lhs, (13,)), _in(rhs, (13,))
    out = np.zeros(12)
    Jl = np.zeros((12, 12)) if diffs else None
    Jr = np.zeros((12, 12)) if diffs else None
    lib().qoracle_state_minus(_p(lhs), _p(rhs), _p(out), _p(Jl), _p(Jr))
    return (out, Jl, Jr) if diffs else out


def euler_step(x, xdot, dt_s, diffs=False):
    x, xdot = _in(x, (13,)), _in(xdot, (12,))
    out = np.zeros(13)
    Jl = np.zeros((12, 12)) if diffs else None
    Jr = np.zeros((12, 12)) if diffs else None
    lib().qoracle_euler_step(_p(x), _p(xdot), C.c_double(dt_s), _p(out), _p(Jl), _p(Jr))
    return (out, Jl, Jr) if diffs else out


def cost(cfg, x, u, x_d, u_d, diffs=False):
    x, u, x_d, u_d = _in(x, (13,)), _in(u, (4,)), _in(x_d, (13,)), _in(u_d, (4,))
    if not diffs:
        return lib().qoracle_cost(C.byref(cfg), _p(x), _p(u), _p(x_d), _p(u_d), None, None, None,
                                  None, None)
    Cx, Cu, Cxx, Cuu, Cxu = (np.zeros(12), np.zeros(4), np.zeros((12, 12)), np.zeros((4, 4)),
                             np.zeros((12, 4)))
    c = lib().qoracle_cost(C.byref(cfg), _p(x), _p(u), _p(x_d), _p(u_d), _p(Cx), _p(Cu), _p(Cxx),
                           _p(Cuu), _p(Cxu))
    return c, Cx, Cu, Cxx, Cuu, Cxu


# ---- solver pieces (single problem; trajectories are [n, 18]) -------------------------
def forward_sim(cfg, desired, cur, k, K, alpha=1.0):
    desired, cur = _in(desired), _in(cur)
    n = cur.shape[0]
    k, K = _in(k, (n, 4)), _in(K, (n, 4, 12))
    out = np.zeros((n, 18))
    rc = lib().qoracle_forward_sim(C.byref(cfg), C.c_int(n), _p(desired), _p(cur), _p(k), _p(K),
                                   C.c_double(alpha), _p(out))
    assert rc == 0
    return out


def cost_trajectory(cfg, desired, traj):
    desired, traj = _in(desired), _in(traj)
    c = C.c_double(0)
    rc = lib().qoracle_cost_trajectory(C.byref(cfg), C.c_int(desired.shape[0]), _p(desired),
                                       C.c_int(traj.shape[0]), _p(traj), C.byref(c))
    if rc == 2:
        raise IndexError("trajectory longer than desired trajectory (cost.hh:39-40)")
    assert rc == 0
    return c.value


def backwards_pass(cfg, desired, traj):
    desired, traj = _in(desired), _in(traj)
    n = traj.shape[0]
    k, K = np.zeros((n, 4)), np.zeros((n, 4, 12))
    a, b = C.c_double(0), C.c_double(0)
    rc = lib().qoracle_backwards_pass(C.byref(cfg), C.c_int(n), _p(desired), _p(traj), _p(k), _p(K),
                                      C.byref(a), C.byref(b))
    assert rc == 0
    return k, K, a.value, b.value


def line_search(cfg, desired, cur, cur_cost, k, K, QuTk, kTQuuk):
    desired, cur = _in(desired), _in(cur)
    n = cur.shape[0]
    k, K = _in(k, (n, 4)), _in(K, (n, 4, 12))
    out = np.zeros((n, 18))
    nc, st = C.c_double(0), C.c_double(0)
    rc = lib().qoracle_line_search(C.byref(cfg), C.c_int(n), _p(desired), _p(cur),
                                   C.c_double(cur_cost), _p(k), _p(K), C.c_double(QuTk),
                                   C.c_double(kTQuuk), _p(out), C.byref(nc), C.byref(st))
    if rc == 4:
        raise RuntimeError("Reached maximum number of line search iterations")
    assert rc == 0
    return out, nc.value, st.value


def solve(cfg, desired, initial):
    """One solve -> dict(traj, k, K, cost_history, step_history, debug, status, ...)."""
    desired, initial = _in(desired), _in(initial)
    n = initial.shape[0]
    cap = int(np.ceil(cfg.max_iters)) + 1
    out = np.zeros((n, 18))
    k, K = np.zeros((n, 4)), np.zeros((n, 4, 12))
    ch, sh = np.zeros(cap), np.zeros(cap)
    dbg = np.zeros((cap, n, 18)) if cfg.populate_debug else None
    res = Result()
    rc = lib().qoracle_solve(C.byref(cfg), C.c_int(n), _p(desired), _p(initial), _p(out), _p(k), _p(K),
                             _p(ch), _p(sh), _p(dbg), C.byref(res))
    assert rc == 0
    nd = res.num_debug
    return dict(traj=out, k=k, K=K, cost_history=ch[:nd].copy(), step_history=sh[:nd].copy(),
                debug=None if dbg is None else dbg[:nd].copy(), status=res.status,
                backward_passes=res.backward_passes, rollouts=res.rollouts, num_debug=nd,
                final_cost=res.final_cost)


def solve_batch(cfg, desired, initial, nthreads=None, want_gains=False, hist_cap=0):
    """Batch of independent solves on `nthreads` host threads.

    initial [B, n, 18]; desired [n, 18] (shared) or [B, n, 18].
    """
    initial = _in(initial)
    B, n, _ = initial.shape
    desired = _in(desired)
    stride = 0 if desired.ndim == 2 else n * 18
    out = np.zeros((B, n, 18))
    k = np.zeros((B, n, 4)) if want_gains else None
    K = np.zeros((B, n, 4, 12)) if want_gains else None
    ch = np.zeros((B, hist_cap)) if hist_cap else None
    res = np.zeros(B, dtype=RESULT_DTYPE)
    if nthreads is None:
        nthreads = hardware_threads()
    rc = lib().qoracle_solve_batch(C.byref(cfg), C.c_int(B), C.c_int(n), _p(desired),
                                   C.c_long(stride), _p(initial), _p(out), _p(k), _p(K), _p(ch),
                                   C.c_int(hist_cap), res.ctypes.data_as(C.c_void_p),
                                   C.c_int(nthreads))
    assert rc == 0
    return dict(traj=out, k=k, K=K, cost_history=ch, status=res["status"].copy(),
                backward_passes=res["backward_passes"].copy(), rollouts=res["rollouts"].copy(),
                num_debug=res["num_debug"].copy(), final_cost=res["final_cost"].copy())


def count_flops(cfg, desired, traj):
    desired, traj = _in(desired), _in(traj)
    out = np.zeros(4)
    rc = lib().qoracle_count_flops(C.byref(cfg), C.c_int(traj.shape[0]), _p(desired), _p(traj), _p(out))
    assert rc == 0
    return dict(backward_per_knot=out[0], rollout_per_knot=out[1], cost_per_knot=out[2],
                cost_per_knot_no_dead_jacobians=out[3])


def hardware_threads() -> int:
    return max(1, lib().qoracle_hardware_threads())
