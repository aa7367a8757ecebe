"""Problem definitions: the reference's default problem and the synthetic batches of
``BASELINE.json``.  Pure numpy -- inputs only, no solver arithmetic.

Trajectory arrays use the C-ABI point layout (``include/qilqr.h``): 18 doubles
``time_s, t(3), q(x,y,z,w), v_lin(3), v_ang(3), u(4)``.  :class:`IDX` /
:func:`to_idx_layout` give the reference script's own 18-column layout, which has the
quaternion w first (``src/quadrotor_ilqr.py:19-37``).
"""
from __future__ import annotations

from enum import IntEnum

import numpy as np

from .options import ConvergenceCriteria, ILQROptions, LineSearchParams


class IDX(IntEnum):  # src/quadrotor_ilqr.py:19-37
    time_s = 0
    translation_x_m = 1
    translation_y_m = 2
    translation_z_m = 3
    quaternion_w = 4
    quaternion_x = 5
    quaternion_y = 6
    quaternion_z = 7
    vel_translational_x_mps = 8
    vel_translational_y_mps = 9
    vel_translational_z_mps = 10
    vel_rotational_x_radps = 11
    vel_rotational_y_radps = 12
    vel_rotational_z_radps = 13
    control_0 = 14
    control_1 = 15
    control_2 = 16
    control_3 = 17


_TO_IDX = [0, 1, 2, 3, 7, 4, 5, 6] + list(range(8, 18))  # (x,y,z,w) -> (w,x,y,z)
_FROM_IDX = [0, 1, 2, 3, 5, 6, 7, 4] + list(range(8, 18))


def to_idx_layout(traj):
    """C-ABI layout -> the reference script's ``extract_traj_array`` layout (w-first quaternion)."""
    return np.asarray(traj)[..., _TO_IDX]


def from_idx_layout(arr):
    return np.asarray(arr)[..., _FROM_IDX]


# ---------------------------------------------------------------------------------
# C1: the reference's default problem (src/quadrotor_ilqr.py:256-306)
# ---------------------------------------------------------------------------------
def _roll_quat_xyzw(roll_rad):
    try:  # the reference uses scipy (quadrotor_ilqr.py:70); identical for a pure roll
        from scipy.spatial.transform import Rotation

        return Rotation.from_euler("xyz", [roll_rad, 0.0, 0.0]).as_quat()
    except Exception:  # pragma: no cover
        return np.array([np.sin(roll_rad / 2.0), 0.0, 0.0, np.cos(roll_rad / 2.0)])


def make_traj_pt(t_s, vel_mps, horizon_s):
    """State of the 4-segment reference path (quadrotor_ilqr.py:83-106) -> 13 doubles."""
    qh = horizon_s / 4.0
    if t_s < qh:
        x, y, z, roll = vel_mps * t_s, 0.0, 0.0, 0.0
    elif t_s < 2.0 * qh:
        x, y, z, roll = vel_mps * qh, vel_mps * (t_s - qh), 10.0 / 3.0, 1 * np.pi / 3.0
    elif t_s < 3.0 * qh:
        x, y, z, roll = vel_mps * (3.0 * qh - t_s), vel_mps * qh, 20.0 / 3.0, 2.0 * np.pi / 3.0
    else:
        x, y, z, roll = 0.0, vel_mps * (4.0 * qh - t_s), 10.0, np.pi
    q = _roll_quat_xyzw(roll)
    return np.array([x, y, z, q[0], q[1], q[2], q[3], 0, 0, 0, 0, 0, 0], dtype=np.float64)


def default_model():
    """Model / cost / dt of quadrotor_ilqr.py:257-292."""
    return dict(mass_kg=1.0, inertia=np.eye(3), arm_length_m=1.0, torque_to_thrust_ratio_m=0.0,
                g_mpss=9.81, Q=np.diag(np.concatenate((100 * np.ones(6), 1 * np.ones(6)))),
                R=np.eye(4), dt_s=0.1)


def default_options(populate_debug=True):
    """quadrotor_ilqr.py:272-284."""
    return ILQROptions(LineSearchParams(0.5, 0.5, 100), ConvergenceCriteria(1e-12, 1e-12, 100.0),
                       populate_debug=populate_debug)


def default_desired_trajectory():
    """quadrotor_ilqr.py:257-270: N=40, dt=0.1, 10 m/s; zero controls.  Also the initial trajectory (:306)."""
    dt_s, horizon_s, vel = 0.1, 4.0, 10
    time_s = np.arange(0, horizon_s, dt_s)
    traj = np.zeros((len(time_s), 18))
    for i, t in enumerate(time_s):
        traj[i, 0] = t
        traj[i, 1:14] = make_traj_pt(t, vel, horizon_s)
    return traj


# ---------------------------------------------------------------------------------
# C2 / C3: random initial SE(3) states -> hover goal
# ---------------------------------------------------------------------------------
HOVER_DRAWS_PER_PROBLEM = 12


def hover_model():
    """Common model/cost of the synthetic configs: as C1 but torque_to_thrust_ratio = 0.1."""
    m = default_model()
    m["torque_to_thrust_ratio_m"] = 0.1
    return m


def hover_initial_states(batch, seed=0, first=0, pos=1.0, theta_max=0.5, vel=0.25):
    """x0 [batch, 13] for problems ``first .. first+batch-1`` of the stream ``seed``.

    Counter-based (Philox) so that problem b gets the same numbers whatever the batch
    size or the rank that generates it: 12 uniforms per problem --
    position U[-pos,pos]^3, rotation Exp(axis*theta) with axis uniform on S^2 and
    theta U[0,theta_max], body velocity U[-vel,vel]^6.
    """
    bitgen = np.random.Philox(key=seed)
    # 4 doubles per Philox counter step; 12 per problem = 3 steps
    bitgen.advance(int(first) * (HOVER_DRAWS_PER_PROBLEM // 4))
    u = np.random.Generator(bitgen).random((batch, HOVER_DRAWS_PER_PROBLEM))
    x0 = np.zeros((batch, 13))
    x0[:, 0:3] = (2.0 * u[:, 0:3] - 1.0) * pos
    z = 2.0 * u[:, 3] - 1.0
    phi = 2.0 * np.pi * u[:, 4]
    s = np.sqrt(np.maximum(0.0, 1.0 - z * z))
    axis = np.stack([s * np.cos(phi), s * np.sin(phi), z], axis=1)
    theta = theta_max * u[:, 5]
    x0[:, 3:6] = axis * np.sin(theta / 2.0)[:, None]
    x0[:, 6] = np.cos(theta / 2.0)
    x0[:, 7:13] = (2.0 * u[:, 6:12] - 1.0) * vel
    return x0


def hover_desired_trajectory(n_knots=40, dt_s=0.1, mass_kg=1.0, g_mpss=9.81):
    """Identity pose, zero velocity, u_d = m g / 4 per rotor (the hover fixed point)."""
    d = np.zeros((n_knots, 18))
    d[:, 0] = np.arange(n_knots) * dt_s
    d[:, 7] = 1.0  # qw
    d[:, 14:18] = mass_kg * g_mpss / 4.0
    return d


def constant_state_trajectory(x0, n_knots, dt_s, u):
    """[B, N, 18] with every knot = x0 and control u: seed for an open-loop rollout
    (forward_sim with zero gains uses only the first state and the controls)."""
    x0 = np.atleast_2d(x0)
    B = x0.shape[0]
    t = np.zeros((B, n_knots, 18))
    t[:, :, 0] = np.arange(n_knots) * dt_s
    t[:, :, 1:14] = x0[:, None, :]
    t[:, :, 14:18] = np.asarray(u, dtype=np.float64)
    return t


def figure_eight_desired(n_knots=1000, dt_s=0.02, amp_m=2.0, period_s=20.0, mass_kg=1.0, g_mpss=9.81):
    """C4: p(t) = (A sin wt, A sin wt cos wt, 1), identity attitude, zero desired velocity."""
    t = np.arange(n_knots) * dt_s
    w = 2.0 * np.pi / period_s
    d = np.zeros((n_knots, 18))
    d[:, 0] = t
    d[:, 1] = amp_m * np.sin(w * t)
    d[:, 2] = amp_m * np.sin(w * t) * np.cos(w * t)
    d[:, 3] = 1.0
    d[:, 7] = 1.0
    d[:, 14:18] = mass_kg * g_mpss / 4.0
    return d


# ---------------------------------------------------------------------------------
# C3 "waypoint" variant and C4 initial states
# ---------------------------------------------------------------------------------
def waypoint_desired_trajectories(batch, n_knots=40, dt_s=0.1, mass_kg=1.0, g_mpss=9.81, seed=3, first=0,
                                  segments=4):
    """[batch, n_knots, 18]: every problem tracks its own piecewise-constant position path of `segments`
    waypoints drawn U[-1,1]^3 (identity attitude, zero velocity, hover thrust) -- one desired trajectory
    per problem (``desired_count = batch``).  Counter-based like :func:`hover_initial_states`: problem b
    gets the same waypoints whatever the batch or rank (12 uniforms = 3 Philox steps per problem)."""
    assert segments == 4, "12 uniforms per problem"
    bitgen = np.random.Philox(key=seed)
    bitgen.advance(int(first) * 3)
    way = 2.0 * np.random.Generator(bitgen).random((batch, segments, 3)) - 1.0
    base = hover_desired_trajectory(n_knots, dt_s, mass_kg, g_mpss)
    desired = np.repeat(base[None], batch, axis=0)
    for seg in range(segments):
        desired[:, seg * n_knots // segments:(seg + 1) * n_knots // segments, 1:4] = way[:, seg][:, None, :]
    return desired


def figure_eight_initial_states(batch, desired, seed=4, first=0, pos=0.3):
    """C4: x0 = first desired state with the position perturbed by U[-pos,pos]^3 (counter-based, 4 draws
    = one Philox step per problem, the fourth unused)."""
    bitgen = np.random.Philox(key=seed)
    bitgen.advance(int(first))
    u = np.random.Generator(bitgen).random((batch, 4))
    x0 = np.tile(np.asarray(desired)[0, 1:14], (batch, 1))
    x0[:, 0:3] += (2.0 * u[:, 0:3] - 1.0) * pos
    return x0
