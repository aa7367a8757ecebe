"""A user-supplied model behind the ModelT concept (ilqr.hh:25-44): CUDA source compiled at run time by
qilqr_set_user_model and run on the model-agnostic kernels (examples/user_model_drag.cu)."""
import os

import numpy as np
import pytest

from conftest import make_solver

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SOURCE = open(os.path.join(ROOT, "examples", "user_model_drag.cu")).read()


def params(model, cd):
    I = np.asarray(model["inertia"])
    return [model["mass_kg"], I[0, 0], I[1, 1], I[2, 2], model["arm_length_m"], model["torque_to_thrust_ratio_m"],
            model["g_mpss"], cd]


def hover(s, B, N, seed):
    from quadrotorilqr_b200 import problems

    m = problems.hover_model()
    desired = problems.hover_desired_trajectory(N, m["dt_s"], m["mass_kg"], m["g_mpss"])
    x0 = problems.hover_initial_states(B, seed=seed)
    seedtraj = problems.constant_state_trajectory(x0, N, m["dt_s"], desired[0, 14:18])
    return desired, seedtraj


def test_user_model_without_drag_reproduces_the_reference_model():
    """c_d = 0: the same dynamics as QuadrotorModel, written by a 'user' -- the solver must take the same decisions
    and reach the same trajectories (1e-9; the summation orders differ) as with the built-in model."""
    from quadrotorilqr_b200 import problems

    model, opts = problems.hover_model(), problems.default_options(False)
    ref, usr = make_solver(model, opts), make_solver(model, opts)
    usr.set_user_model(SOURCE, params(model, 0.0))
    desired, seedtraj = hover(ref, 64, 40, seed=3)
    zeros = (np.zeros((64, 40, 4)), np.zeros((64, 40, 48)))
    init_ref, init_usr = ref.forward_sim(seedtraj, *zeros), usr.forward_sim(seedtraj, *zeros)
    assert np.max(np.abs(init_ref - init_usr)) <= 1e-12
    a, b = ref.solve(init_ref, desired, want_gains=True), usr.solve(init_ref, desired, want_gains=True)
    assert np.array_equal(a["results"]["status"], b["results"]["status"])
    assert np.array_equal(a["results"]["backward_passes"], b["results"]["backward_passes"])
    for key in ("traj", "k", "K"):
        scale = max(1.0, float(np.max(np.abs(a[key]))))
        assert np.max(np.abs(a[key] - b[key])) <= 1e-9 * scale, key


def test_user_model_with_drag():
    from quadrotorilqr_b200 import problems

    model, opts = problems.hover_model(), problems.default_options(False)
    usr, ref = make_solver(model, opts), make_solver(model, opts)
    usr.set_user_model(SOURCE, params(model, 0.4))
    B, N = 48, 40
    desired, seedtraj = hover(usr, B, N, seed=5)
    zeros = (np.zeros((B, N, 4)), np.zeros((B, N, 48)))
    initial = usr.forward_sim(seedtraj, *zeros)
    assert np.max(np.abs(initial - ref.forward_sim(seedtraj, *zeros))) > 1e-3      # the drag is felt
    # its dynamics, checked independently on the host (numpy restatement of the velocity update)
    v0, v1 = initial[:, 0, 8:11], initial[:, 1, 8:11]
    q = initial[:, 0, 4:8]
    Rz = np.stack([2 * (q[:, 0] * q[:, 2] - q[:, 1] * q[:, 3]), 2 * (q[:, 1] * q[:, 2] + q[:, 0] * q[:, 3]),
                   1 - 2 * (q[:, 0] ** 2 + q[:, 1] ** 2)], axis=1)                  # third row of R(q) = R^T e_z
    acc = -model["g_mpss"] * Rz - 0.4 * v0
    acc[:, 2] += initial[:, 0, 14:18].sum(axis=1) / model["mass_kg"]
    assert np.max(np.abs(v1 - (v0 + model["dt_s"] * acc))) <= 1e-12
    r = usr.solve(initial, desired, hist_cap=100)
    res = r["results"]
    assert np.all(np.isin(res["status"], [1, 2])) and res["backward_passes"].max() <= 40   # the Jacobians are right
    nd = res["num_debug"]
    for b in range(B):
        h = r["cost_history"][b, :nd[b]]
        assert np.all(np.diff(h[1:]) < 0)                                           # Armijo after the first step
    # the solution is a trajectory of the user's dynamics: rolling its controls out reproduces its states
    again = usr.forward_sim(r["traj"], *zeros)
    assert np.array_equal(again[:, :, 1:14], r["traj"][:, :, 1:14])
    # and not of the built-in model's
    assert np.max(np.abs(ref.forward_sim(r["traj"], *zeros) - r["traj"])) > 1e-4


def test_compile_errors_are_reported():
    from quadrotorilqr_b200 import problems
    from quadrotorilqr_b200.solver import QilqrError

    s = make_solver(problems.hover_model(), problems.default_options(False))
    with pytest.raises(QilqrError) as e:
        s.set_user_model("this is not CUDA", [])
    assert "user_model.cu" in str(e.value) and "error" in str(e.value)
    # the handle still works with its built-in model
    desired, seedtraj = hover(s, 4, 40, seed=1)
    s.solve(s.forward_sim(seedtraj, np.zeros((4, 40, 4)), np.zeros((4, 40, 48))), desired)
