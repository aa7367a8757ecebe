"""Host-side mirror of ``ILQR<QuadrotorModel>`` (``src/ilqr.hh:25-206``) for batches.

:class:`BatchILQR` is constructed like the reference's pybind ``QuadrotorILQR``
(``src/quadrotor_ilqr_binding.cc:20-32``: mass, inertia, arm length,
torque-to-thrust ratio, g, Q, R, dt, options) and exposes the reference's public
methods -- ``solve``, ``forward_sim``, ``cost_trajectory``, ``backwards_pass``,
``line_search`` -- plus the model/cost functions under them, each over a batch of
independent problems.  All arithmetic happens in the CUDA library behind the C ABI
(``include/qilqr.h``); this file only marshals numpy / torch buffers.

Array conventions (see ``include/qilqr.h``): state = 13 doubles ``t, q(x,y,z,w),
v(6)``; trajectory point = 18 doubles ``time_s, state, control``; host trajectories
are ``[batch, knots, 18]``; device-resident trajectories are ``[knots, 17, batch]``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from .options import ILQROptions

RESULT_DTYPE = np.dtype(
    [("status", "<i4"), ("backward_passes", "<i4"), ("rollouts", "<i4"), ("num_debug", "<i4"),
     ("final_cost", "<f8")]
)


class QilqrError(RuntimeError):
    def __init__(self, code, message=""):
        self.code = code
        text = _capi.lib().qilqr_error_string(code).decode()
        super().__init__(f"{text}{(': ' + message) if message else ''}")


def _f64(a, shape=None):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    if shape is not None:
        a = a.reshape(shape)
    return a


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return C.c_void_p(a.ctypes.data)
    # torch tensor
    assert a.is_contiguous()
    return C.c_void_p(a.data_ptr())


def _options_struct(o: ILQROptions) -> _capi.Options:
    s = _capi.Options()
    s.step_update = o.line_search_params.step_update
    s.desired_reduction_frac = o.line_search_params.desired_reduction_frac
    s.line_search_max_iters = int(o.line_search_params.max_iters)
    s.populate_debug = int(bool(o.populate_debug))
    s.rtol = o.convergence_criteria.rtol
    s.atol = o.convergence_criteria.atol
    s.max_iters = float(o.convergence_criteria.max_iters)
    s.symmetrize_vxx = int(bool(o.symmetrize_vxx))
    s.num_parallel_alphas = int(o.num_parallel_alphas)
    s.quu_regularization = float(o.quu_regularization)
    return s


class BatchILQR:
    """``ILQR<QuadrotorModel>`` over a batch, on one B200."""

    def __init__(self, mass_kg, inertia, arm_length_m, torque_to_thrust_ratio_m, g_mpss, Q, R, dt_s,
                 options: ILQROptions | None = None, device: int = 0, model_flags: int = 0):
        """``model_flags``: 0 = the reference's ``QuadrotorModel``; ``_capi.MODEL_RK4`` /
        ``MODEL_CORIOLIS`` select a second dynamics function behind the ModelT concept
        (``ilqr.hh:25-44``), ``MODEL_GENERIC`` the model-agnostic kernels for the reference model."""
        self._h = None
        self._dbg = None
        self.options = options if options is not None else ILQROptions()
        self.dt_s = float(dt_s)
        m = _capi.Model()
        m.mass_kg = float(mass_kg)
        m.inertia[:] = _f64(inertia, (9,)).tolist()
        m.arm_length_m = float(arm_length_m)
        m.torque_to_thrust_ratio_m = float(torque_to_thrust_ratio_m)
        m.g_mpss = float(g_mpss)
        self._Q = _f64(Q, (144,))
        self._R = _f64(R, (16,))
        self.device = int(device)
        h = C.c_void_p()
        opts = _options_struct(self.options)
        rc = _capi.lib().qilqr_create(C.byref(m), _ptr(self._Q), _ptr(self._R), C.c_double(self.dt_s),
                                      C.byref(opts), C.c_int(self.device), C.byref(h))
        if rc != _capi.OK:
            raise QilqrError(rc)
        self._h = h
        self.model_flags = int(model_flags)
        if self.model_flags:
            self._check(_capi.lib().qilqr_set_model_variant(self._h, C.c_int(self.model_flags)))

    # ---- lifetime -----------------------------------------------------------------
    def close(self):
        if self._h is not None:
            _capi.lib().qilqr_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != _capi.OK:
            raise QilqrError(rc, _capi.lib().qilqr_last_error_message(self._h).decode())

    def set_options(self, options: ILQROptions):
        self.options = options
        o = _options_struct(options)
        self._check(_capi.lib().qilqr_set_options(self._h, C.byref(o)))

    def set_user_model(self, cuda_source: str, params=()):
        """``qilqr_set_user_model``: run every later call of this solver with the caller's ``discrete_dynamics``
        (CUDA C++ defining ``qilqr_user_discrete_dynamics``, compiled at run time) on the model-agnostic kernels."""
        par = _f64(np.asarray(params, dtype=np.float64).ravel())
        rc = _capi.lib().qilqr_set_user_model(self._h, C.c_char_p(cuda_source.encode()), _ptr(par) if par.size else None,
                                              C.c_int(par.size))
        if rc:
            raise QilqrError(rc, _capi.lib().qilqr_last_error_message(self._h).decode())

    def set_profiling(self, enabled: bool):
        self._check(_capi.lib().qilqr_set_profiling(self._h, C.c_int(int(enabled))))

    @property
    def kernel_launch_count(self) -> int:
        return int(_capi.lib().qilqr_kernel_launch_count(self._h))

    @property
    def stream_handle(self) -> int:
        return int(_capi.lib().qilqr_stream(self._h) or 0)

    def last_solve_stats(self) -> dict:
        s = _capi.SolveStats()
        self._check(_capi.lib().qilqr_last_solve_stats(self._h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in s._fields_}

    # ---- ILQR::solve (ilqr.hh:53-87) -------------------------------------------------
    def solve(self, initial, desired, want_gains=False, hist_cap=0, want_debug=False, out=None):
        """Solve ``batch`` problems.  initial ``[B, N, 18]``; desired ``[N, 18]`` (shared) or ``[B, N, 18]``.

        Returns a dict with ``traj [B,N,18]``, ``results`` (structured array: status,
        backward_passes, rollouts, num_debug, final_cost) and, on request, ``k [B,N,4]``,
        ``K [B,N,4,12]``, ``cost_history [B,hist_cap]``, ``debug [B,cap,N,18]``.
        """
        initial = _f64(initial)
        if initial.ndim == 2:
            initial = initial[None]
        B, N, _ = initial.shape
        desired = _f64(desired)
        Bd = 1 if desired.ndim == 2 else desired.shape[0]
        if out is not None and not (isinstance(out, np.ndarray) and out.dtype == np.float64
                                    and out.shape == (B, N, 18) and out.flags["C_CONTIGUOUS"]):
            raise ValueError(f"out must be a C-contiguous float64 array of shape {(B, N, 18)}")
        traj = out if out is not None else np.empty((B, N, 18))
        k = np.empty((B, N, 4)) if want_gains else None
        K = np.empty((B, N, 4, 12)) if want_gains else None
        hist = np.zeros((B, hist_cap)) if hist_cap else None
        max_iters = self.options.convergence_criteria.max_iters
        dbg_cap = 0
        if want_debug:  # ILQRDebug of every problem: one trajectory per completed iteration (ilqr_debug.hh:9-22)
            dbg_cap = int(np.ceil(max_iters)) if np.isfinite(max_iters) and max_iters > 0 else 0
            if B * dbg_cap * N * 18 * 8 > self.MAX_DEBUG_BYTES:
                raise ValueError(
                    f"populate_debug for {B} problems x {dbg_cap} iterations x {N} knots needs "
                    f"{B * dbg_cap * N * 144 / 2**30:.1f} GiB; sample problems / iterations with solve_sampled_debug(), "
                    "solve a smaller batch, or lower convergence_criteria.max_iters")
        dbg = np.zeros((B, dbg_cap, N, 18)) if dbg_cap else None
        res = np.zeros(B, dtype=RESULT_DTYPE)
        rc = _capi.lib().qilqr_solve_host(self._h, C.c_int(B), C.c_int(N), _ptr(desired), C.c_int(Bd),
                                          _ptr(initial), _ptr(traj), _ptr(k), _ptr(K), _ptr(hist),
                                          C.c_int(hist_cap), _ptr(dbg), C.c_int(dbg_cap), _ptr(res))
        self._check(rc)
        return dict(traj=traj, results=res, k=k, K=K, cost_history=hist, debug=dbg)

    MAX_DEBUG_BYTES = 8 << 30  # refuse to stage more than this for a full ILQRDebug capture

    # ---- ILQRDebug at batch scale (include/qilqr.h: qilqr_set_debug_sampling) ---------------------------
    def set_debug_sampling(self, problems=None, every=1, ring=None):
        """Capture ILQRDebug trajectories (``ilqr.hh:78-80``) for a sample only: ``problems`` = indices into the
        batch, iterations ``i % every == 0``, the last ``ring`` sampled iterations per problem (default: enough for
        every iteration up to ``max_iters``).  ``problems=None`` switches sampling off.  Needs
        ``options.populate_debug``; read the rings back with :meth:`read_debug_samples` after a solve."""
        if problems is None or len(problems) == 0:
            self._check(_capi.lib().qilqr_set_debug_sampling(self._h, C.c_int(0), None, C.c_int(0), C.c_int(0)))
            self._dbg = None
            return
        idx = np.ascontiguousarray(np.asarray(problems, dtype=np.int32))
        if ring is None:
            mi = self.options.convergence_criteria.max_iters
            ring = max(1, int(np.ceil(min(mi, 4096.0) / every))) if np.isfinite(mi) and mi > 0 else 1
        self._check(_capi.lib().qilqr_set_debug_sampling(self._h, C.c_int(int(every)), _ptr(idx),
                                                         C.c_int(idx.size), C.c_int(int(ring))))
        self._dbg = (idx, int(every), int(ring))

    def read_debug_samples(self, n_knots, out=None):
        """Rings of the last solve: dict with ``problems [S]``, ``traj [S, ring, N, 18]``, ``iters [S, ring]`` (-1 =
        empty slot), ``costs [S, ring]``, ``counts [S]``; :meth:`debug_of` orders one problem's entries."""
        idx, every, ring = self._dbg
        S = idx.size
        traj = out if out is not None else np.empty((S, ring, n_knots, 18))
        iters = np.empty((S, ring), dtype=np.int32)
        costs = np.empty((S, ring))
        counts = np.empty(S, dtype=np.int32)
        self._check(_capi.lib().qilqr_read_debug_samples_host(self._h, _ptr(traj), _ptr(iters), _ptr(costs),
                                                              _ptr(counts)))
        return dict(problems=idx, traj=traj, iters=iters, costs=costs, counts=counts, every=every, ring=ring)

    @staticmethod
    def debug_of(samples, s):
        """(iteration indices, trajectories, costs) of sampled problem number ``s``, oldest first."""
        it = samples["iters"][s]
        order = np.argsort(it[it >= 0], kind="stable")
        sel = np.where(it >= 0)[0][order]
        return it[sel], samples["traj"][s, sel], samples["costs"][s, sel]

    def last_cost_history(self, first=0, count=None, cap=None):
        """Per-iteration costs of the last solve, kept on the device for every problem (always on):
        ``[count, cap]``, zero-padded; a problem's valid entries are ``results['num_debug']``."""
        stored = C.c_int(0)
        self._check(_capi.lib().qilqr_last_cost_history_host(self._h, C.c_int(0), C.c_int(0), None, C.c_int(0),
                                                             C.byref(stored)))
        cap = int(cap or stored.value)
        if count is None:
            raise ValueError("count (number of problems) is required")
        out = np.zeros((count, max(cap, 1)))
        self._check(_capi.lib().qilqr_last_cost_history_host(self._h, C.c_int(first), C.c_int(count), _ptr(out),
                                                             C.c_int(out.shape[1]), C.byref(stored)))
        return out

    def solve_host_begin(self, initial, desired, out_traj, results):
        """First half of :meth:`solve_host_buffers` (``qilqr_solve_host_begin``): returns once the device finishes the
        batch on its own; call :meth:`solve_host_finish` before touching ``out_traj`` / ``results``."""
        B, N = int(initial.shape[0]), int(initial.shape[1])
        Bd = 1 if desired.ndim == 2 else int(desired.shape[0])
        self._check(_capi.lib().qilqr_solve_host_begin(self._h, C.c_int(B), C.c_int(N), _ptr(desired), C.c_int(Bd),
                                                       _ptr(initial), _ptr(out_traj), _ptr(results)))

    def solve_host_finish(self):
        self._check(_capi.lib().qilqr_solve_host_finish(self._h))

    def solve_device_begin(self, traj_soa, desired_soa, results=None, k=None, K=None, cost_hist=None):
        """First half of :meth:`solve_device` (``qilqr_solve_device_begin``)."""
        N, rows, B = traj_soa.shape
        assert rows == 17 and desired_soa.shape[0] == N and desired_soa.shape[1] == 17
        Bd = int(desired_soa.shape[2])
        hist_cap = int(cost_hist.shape[0]) if cost_hist is not None else 0
        self._check(_capi.lib().qilqr_solve_device_begin(self._h, C.c_int(B), C.c_int(N), _ptr(desired_soa),
                                                         C.c_int(Bd), _ptr(traj_soa), _ptr(k), _ptr(K),
                                                         _ptr(cost_hist), C.c_int(hist_cap), _ptr(results)))

    def solve_device_finish(self):
        self._check(_capi.lib().qilqr_solve_device_finish(self._h))

    def solve_from_controls(self, x0, controls, desired, out_traj=None, out_controls=None, results=None,
                            want_traj=True, want_controls=False, begin_only=False):
        """``qilqr_solve_from_controls_host``: the initial guess is a control sequence (``controls [N, 4]`` shared or
        ``[B, N, 4]``), the initial trajectory its open-loop rollout from ``x0 [B, 13]`` on the device."""
        x0 = x0 if not isinstance(x0, np.ndarray) else _f64(x0)
        B = int(x0.shape[0])
        controls = controls if not isinstance(controls, np.ndarray) else _f64(controls)
        Bc = 1 if controls.ndim == 2 else int(controls.shape[0])
        N = int(controls.shape[-2])
        Bd = 1 if desired.ndim == 2 else int(desired.shape[0])
        if out_traj is None and want_traj:
            out_traj = np.empty((B, N, 18))
        if out_controls is None and want_controls:
            out_controls = np.empty((B, N, 4))
        res = results if results is not None else np.zeros(B, dtype=RESULT_DTYPE)
        fn = (_capi.lib().qilqr_solve_from_controls_host_begin if begin_only
              else _capi.lib().qilqr_solve_from_controls_host)
        self._check(fn(self._h, C.c_int(B), C.c_int(N), _ptr(desired), C.c_int(Bd), _ptr(x0), _ptr(controls),
                       C.c_int(Bc), _ptr(out_traj), _ptr(out_controls), _ptr(res)))
        return dict(traj=out_traj, controls=out_controls, results=res)

    def solve_host_buffers(self, initial, desired, out_traj, results):
        """Zero-allocation variant for (pinned) host buffers: numpy arrays or CPU torch tensors."""
        B, N = int(initial.shape[0]), int(initial.shape[1])
        Bd = 1 if desired.ndim == 2 else int(desired.shape[0])
        rc = _capi.lib().qilqr_solve_host(self._h, C.c_int(B), C.c_int(N), _ptr(desired), C.c_int(Bd),
                                          _ptr(initial), _ptr(out_traj), None, None, None, C.c_int(0), None,
                                          C.c_int(0), _ptr(results))
        self._check(rc)

    # ---- device-resident path ----------------------------------------------------------
    def solve_device(self, traj_soa, desired_soa, results=None, k=None, K=None, cost_hist=None):
        """In-place solve of device-resident data (torch CUDA float64 tensors).

        traj_soa ``[N, 17, B]`` (in: initial trajectory, out: solution), desired_soa
        ``[N, 17, 1]`` or ``[N, 17, B]``; results: uint8 tensor of ``B*24`` bytes or None.
        """
        N, rows, B = traj_soa.shape
        assert rows == 17 and desired_soa.shape[0] == N and desired_soa.shape[1] == 17
        Bd = int(desired_soa.shape[2])
        hist_cap = int(cost_hist.shape[0]) if cost_hist is not None else 0
        rc = _capi.lib().qilqr_solve_device(self._h, C.c_int(B), C.c_int(N), _ptr(desired_soa), C.c_int(Bd),
                                            _ptr(traj_soa), _ptr(k), _ptr(K), _ptr(cost_hist),
                                            C.c_int(hist_cap), _ptr(results))
        self._check(rc)

    def pack_trajectory_device(self, aos, soa):
        B, N, _ = aos.shape
        self._check(_capi.lib().qilqr_pack_trajectory_device(self._h, C.c_int(B), C.c_int(N), _ptr(aos), _ptr(soa)))

    def unpack_trajectory_device(self, soa, aos, time_src=None):
        N, _, B = soa.shape
        self._check(_capi.lib().qilqr_unpack_trajectory_device(self._h, C.c_int(B), C.c_int(N), _ptr(soa),
                                                               _ptr(time_src), _ptr(aos)))

    def rollout_constant_control_device(self, x0_soa, u, traj_soa):
        N, _, B = traj_soa.shape
        u = _f64(u, (4,))
        self._check(_capi.lib().qilqr_rollout_constant_control_device(self._h, C.c_int(B), C.c_int(N),
                                                                      _ptr(x0_soa), _ptr(u), _ptr(traj_soa)))

    def mpc_advance_device(self, traj_soa, plant_soa, disturbance=None, applied_u=None):
        N, _, B = traj_soa.shape
        self._check(_capi.lib().qilqr_mpc_advance_device(self._h, C.c_int(B), C.c_int(N), _ptr(traj_soa),
                                                         _ptr(plant_soa), _ptr(disturbance), _ptr(applied_u)))

    def mpc_run_device(self, steps, traj_soa, desired_soa, plant_soa, disturbance=None, state_log=None,
                       control_log=None):
        """Receding-horizon loop on the device (BASELINE config 5): `steps` x (solve from the warm start,
        apply u_0 to the plant, shift).  Returns dict(backward_passes, rollouts, not_converged, resolves)."""
        N, _, B = traj_soa.shape
        Bd = int(desired_soa.shape[2])
        totals = (C.c_int64 * 4)()
        self._check(_capi.lib().qilqr_mpc_run_device(self._h, C.c_int(int(steps)), C.c_int(B), C.c_int(N),
                                                     _ptr(desired_soa), C.c_int(Bd), _ptr(traj_soa), _ptr(plant_soa),
                                                     _ptr(disturbance), _ptr(state_log), _ptr(control_log), totals))
        return dict(backward_passes=int(totals[0]), rollouts=int(totals[1]), not_converged=int(totals[2]),
                    resolves=int(totals[3]))

    # ---- ILQR::forward_sim / cost_trajectory / backwards_pass / line_search ------------------
    def forward_sim(self, current, k, K, alpha=1.0):
        current = _f64(current)
        squeeze = current.ndim == 2
        if squeeze:
            current = current[None]
        B, N, _ = current.shape
        k = _f64(k, (B, N, 4))
        K = _f64(K, (B, N, 48))
        alpha = _f64(np.broadcast_to(np.asarray(alpha, dtype=np.float64), (B,)))
        out = np.empty((B, N, 18))
        self._check(_capi.lib().qilqr_forward_sim_host(self._h, C.c_int(B), C.c_int(N), _ptr(current), _ptr(k),
                                                       _ptr(K), _ptr(alpha), _ptr(out)))
        return out[0] if squeeze else out

    def cost_trajectory(self, traj, desired):
        traj = _f64(traj)
        squeeze = traj.ndim == 2
        if squeeze:
            traj = traj[None]
        B, N, _ = traj.shape
        desired = _f64(desired)
        Bd = 1 if desired.ndim == 2 else desired.shape[0]
        nd = desired.shape[-2]
        cost = np.empty(B)
        self._check(_capi.lib().qilqr_cost_trajectory_host(self._h, C.c_int(B), C.c_int(N), _ptr(desired),
                                                           C.c_int(Bd), C.c_int(nd), _ptr(traj), _ptr(cost)))
        return float(cost[0]) if squeeze else cost

    def backwards_pass(self, traj, desired):
        traj = _f64(traj)
        squeeze = traj.ndim == 2
        if squeeze:
            traj = traj[None]
        B, N, _ = traj.shape
        desired = _f64(desired)
        Bd = 1 if desired.ndim == 2 else desired.shape[0]
        k, K, terms = np.empty((B, N, 4)), np.empty((B, N, 4, 12)), np.empty((B, 2))
        self._check(_capi.lib().qilqr_backwards_pass_host(self._h, C.c_int(B), C.c_int(N), _ptr(desired),
                                                          C.c_int(Bd), _ptr(traj), _ptr(k), _ptr(K), _ptr(terms)))
        if squeeze:
            return k[0], K[0], float(terms[0, 0]), float(terms[0, 1])
        return k, K, terms[:, 0].copy(), terms[:, 1].copy()

    def line_search(self, current, desired, current_cost, k, K, QuTk, kTQuuk):
        current = _f64(current)
        squeeze = current.ndim == 2
        if squeeze:
            current = current[None]
        B, N, _ = current.shape
        desired = _f64(desired)
        Bd = 1 if desired.ndim == 2 else desired.shape[0]
        k, K = _f64(k, (B, N, 4)), _f64(K, (B, N, 48))
        cc = _f64(np.broadcast_to(np.asarray(current_cost, dtype=np.float64), (B,)))
        terms = _f64(np.stack([np.broadcast_to(np.asarray(QuTk, dtype=np.float64), (B,)),
                               np.broadcast_to(np.asarray(kTQuuk, dtype=np.float64), (B,))], axis=1))
        out, nc, step = np.empty((B, N, 18)), np.empty(B), np.empty(B)
        status = np.zeros(B, dtype=np.int32)
        self._check(_capi.lib().qilqr_line_search_host(self._h, C.c_int(B), C.c_int(N), _ptr(desired),
                                                       C.c_int(Bd), _ptr(current), _ptr(cc), _ptr(k), _ptr(K),
                                                       _ptr(terms), _ptr(out), _ptr(nc), _ptr(step),
                                                       _ptr(status)))
        if squeeze:
            if status[0] != 0:
                raise QilqrError(int(status[0]))  # ilqr.hh:191-193
            return out[0], float(nc[0]), float(step[0])
        return out, nc, step, status

    # ---- QuadrotorModel / CostFunction -----------------------------------------------------------
    def _model_call(self, fn, a, na, b, nb, nout, diffs, jshapes):
        a = _f64(a)
        squeeze = a.ndim == 1
        a = a.reshape(-1, na)
        B = a.shape[0]
        b = _f64(b).reshape(B, nb)
        out = np.empty((B, nout))
        J1 = np.empty((B,) + jshapes[0]) if diffs else None
        J2 = np.empty((B,) + jshapes[1]) if diffs else None
        self._check(fn(self._h, C.c_int(B), _ptr(a), _ptr(b), _ptr(out), _ptr(J1), _ptr(J2)))
        if squeeze:
            return (out[0], J1[0], J2[0]) if diffs else out[0]
        return (out, J1, J2) if diffs else out

    def discrete_dynamics(self, x, u, diffs=False):
        return self._model_call(_capi.lib().qilqr_discrete_dynamics_host, x, 13, u, 4, 13, diffs, ((12, 12), (12, 4)))

    def continuous_dynamics(self, x, u, diffs=False):
        return self._model_call(_capi.lib().qilqr_continuous_dynamics_host, x, 13, u, 4, 12, diffs, ((12, 12), (12, 4)))

    def state_minus(self, lhs, rhs, diffs=False):
        return self._model_call(_capi.lib().qilqr_state_minus_host, lhs, 13, rhs, 13, 12, diffs, ((12, 12), (12, 12)))

    def state_add(self, x, tangent, diffs=False):
        return self._model_call(_capi.lib().qilqr_state_add_host, x, 13, tangent, 12, 13, diffs, ((12, 12), (12, 12)))

    def cost(self, x, u, x_d, u_d, diffs=False):
        x = _f64(x)
        squeeze = x.ndim == 1
        x = x.reshape(-1, 13)
        B = x.shape[0]
        u, x_d, u_d = _f64(u).reshape(B, 4), _f64(x_d).reshape(B, 13), _f64(u_d).reshape(B, 4)
        cost = np.empty(B)
        if diffs:
            Cx, Cu, Cxx, Cuu, Cxu = (np.empty((B, 12)), np.empty((B, 4)), np.empty((B, 12, 12)),
                                     np.empty((B, 4, 4)), np.empty((B, 12, 4)))
        else:
            Cx = Cu = Cxx = Cuu = Cxu = None
        self._check(_capi.lib().qilqr_cost_host(self._h, C.c_int(B), _ptr(x), _ptr(u), _ptr(x_d), _ptr(u_d),
                                                _ptr(cost), _ptr(Cx), _ptr(Cu), _ptr(Cxx), _ptr(Cuu), _ptr(Cxu)))
        if not diffs:
            return float(cost[0]) if squeeze else cost
        if squeeze:
            return float(cost[0]), Cx[0], Cu[0], Cxx[0], Cuu[0], Cxu[0]
        return cost, Cx, Cu, Cxx, Cuu, Cxu
