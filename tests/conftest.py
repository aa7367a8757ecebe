import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def O():
    """The CPU oracle (test infrastructure)."""
    import oracle

    oracle.build()
    return oracle


def make_solver(model=None, options=None, model_flags=0, **overrides):
    """BatchILQR for a model dict (quadrotorilqr_b200.problems.default_model style)."""
    from quadrotorilqr_b200 import BatchILQR, problems

    m = dict(problems.default_model() if model is None else model)
    m.update(overrides)
    return BatchILQR(m["mass_kg"], m["inertia"], m["arm_length_m"], m["torque_to_thrust_ratio_m"],
                     m["g_mpss"], m["Q"], m["R"], m["dt_s"], options, model_flags=model_flags)


def oracle_config(O, model, options, model_kind=0):
    return O.make_config(
        model_kind=model_kind,
        mass_kg=model["mass_kg"], inertia=model["inertia"], arm_length_m=model["arm_length_m"],
        torque_to_thrust_ratio_m=model["torque_to_thrust_ratio_m"], g_mpss=model["g_mpss"],
        Q=model["Q"], R=model["R"], dt_s=model["dt_s"],
        step_update=options.line_search_params.step_update,
        desired_reduction_frac=options.line_search_params.desired_reduction_frac,
        ls_max_iters=options.line_search_params.max_iters,
        rtol=options.convergence_criteria.rtol, atol=options.convergence_criteria.atol,
        max_iters=options.convergence_criteria.max_iters, populate_debug=options.populate_debug,
        symmetrize_vxx=options.symmetrize_vxx, quu_regularization=options.quu_regularization)


def random_spd_inertia(seed=0):
    """Stand-in for make_random_inertia_matrix (quadrotor_model_test.cc:22-28): A A^T + 3 I."""
    A = np.random.default_rng(seed).uniform(-1, 1, (3, 3))
    return A @ A.T + 3 * np.eye(3)


def identity_traj(n, dt):
    """create_identity_traj (ilqr_test.cc:23-36)."""
    t = np.zeros((n, 18))
    t[:, 0] = np.arange(n) * dt
    t[:, 7] = 1.0
    return t
