"""C-ABI contract details: argument validation, solver reuse across batch sizes / horizons / options,
statistics, and that concurrent handles (the pipelining mode of bench.py) do not interfere."""
import ctypes as C
import threading

import numpy as np
import pytest

from conftest import make_solver

pytestmark = pytest.mark.gpu


def hover(s, B, N, seed):
    from quadrotorilqr_b200 import problems

    m = problems.hover_model()
    d = problems.hover_desired_trajectory(N, m["dt_s"], m["mass_kg"], m["g_mpss"])
    x0 = problems.hover_initial_states(B, seed=seed)
    init = s.forward_sim(problems.constant_state_trajectory(x0, N, m["dt_s"], d[0, 14:18]), np.zeros((B, N, 4)),
                         np.zeros((B, N, 48)))
    return d, init


def test_invalid_arguments_are_rejected():
    from quadrotorilqr_b200 import QilqrError, _capi, problems

    s = make_solver(problems.hover_model(), problems.default_options(False))
    d, init = hover(s, 4, 8, 0)
    lib = _capi.lib()
    out = np.empty_like(init)
    res = np.zeros(4, dtype=np.dtype([("a", "<i4", 4), ("c", "<f8")]))
    p = lambda a: C.c_void_p(a.ctypes.data)
    bad = [
        (0, 8, 1),    # empty batch
        (4, 0, 1),    # empty trajectory (undefined behaviour in the reference, ilqr.hh:156)
        (4, 8, 3),    # desired_count must be 1 or batch
    ]
    for B, N, Bd in bad:
        rc = lib.qilqr_solve_host(s._h, C.c_int(B), C.c_int(N), p(d), C.c_int(Bd), p(init), p(out), None, None, None,
                                  C.c_int(0), None, C.c_int(0), p(res))
        assert rc == _capi.ERR_INVALID_ARGUMENT
    assert lib.qilqr_solve_host(s._h, C.c_int(4), C.c_int(8), None, C.c_int(1), p(init), p(out), None, None, None,
                                C.c_int(0), None, C.c_int(0), p(res)) == _capi.ERR_INVALID_ARGUMENT
    assert lib.qilqr_create(None, None, None, C.c_double(0.1), None, C.c_int(0), None) == _capi.ERR_INVALID_ARGUMENT
    with pytest.raises(QilqrError):
        s.cost_trajectory(init, d[:4])  # shorter desired than trajectory


def test_solver_reuse_and_option_changes(O):
    import dataclasses
    from conftest import oracle_config
    from quadrotorilqr_b200 import problems

    model, opts = problems.hover_model(), problems.default_options(False)
    s = make_solver(model, opts)
    first = None
    for B, N in ((50, 40), (7, 12), (300, 40), (50, 40)):  # grow, shrink, grow: the workspace is reused
        d, init = hover(s, B, N, seed=B)
        r = s.solve(init, d)
        if (B, N) == (50, 40):
            if first is None:
                first = r
            else:
                assert np.array_equal(r["traj"], first["traj"]) and np.array_equal(r["results"], first["results"])
    # options can be changed on a live solver and take effect
    loose = dataclasses.replace(opts)
    loose.convergence_criteria = dataclasses.replace(opts.convergence_criteria, rtol=1e-4, atol=1e-4)
    d, init = hover(s, 50, 40, seed=50)
    s.set_options(loose)
    r2 = s.solve(init, d)
    assert r2["results"]["backward_passes"].sum() < first["results"]["backward_passes"].sum()
    cfg = oracle_config(O, model, loose)
    o = O.solve_batch(cfg, d, init)
    assert np.array_equal(r2["results"]["backward_passes"], o["backward_passes"])
    s.set_options(opts)
    r3 = s.solve(init, d)
    assert np.array_equal(r3["traj"], first["traj"])
    st = s.last_solve_stats()
    assert st["problem_iterations"] == int(r3["results"]["backward_passes"].sum())
    assert st["problem_rollouts"] == int(r3["results"]["rollouts"].sum())
    # host-sequenced super-steps: none beyond the slowest problem's rollouts (the persistent tail kernel, which takes
    # over small batches from the start, does not count)
    assert 0 <= st["solver_iterations"] <= int(r3["results"]["rollouts"].max()) + 1


def test_concurrent_handles_do_not_interfere():
    from quadrotorilqr_b200 import problems

    model, opts = problems.hover_model(), problems.default_options(False)
    solvers = [make_solver(model, opts) for _ in range(3)]
    data = [hover(solvers[j], 3000 + 17 * j, 40, seed=100 + j) for j in range(3)]
    ref = [solvers[j].solve(data[j][1], data[j][0]) for j in range(3)]
    out = [None] * 3

    def work(j):
        for _ in range(3):
            out[j] = solvers[j].solve(data[j][1], data[j][0])

    ths = [threading.Thread(target=work, args=(j,)) for j in range(3)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    for j in range(3):
        assert np.array_equal(out[j]["traj"], ref[j]["traj"]) and np.array_equal(out[j]["results"], ref[j]["results"])


def test_solve_from_controls_equals_rollout_then_solve():
    """qilqr_solve_from_controls_host: x0 + nominal controls in, the initial trajectory rolled out on the device --
    bit-identical to forward_sim (zero gains) followed by solve; shared and per-problem control sequences;
    trajectories and / or controls out."""
    from quadrotorilqr_b200 import problems

    model, opts = problems.hover_model(), problems.default_options(False)
    s = make_solver(model, opts)
    B, N = 130, 40
    desired = problems.hover_desired_trajectory(N, model["dt_s"], model["mass_kg"], model["g_mpss"])
    x0 = problems.hover_initial_states(B, seed=9)
    rng = np.random.default_rng(1)
    for controls in (np.tile(desired[0, 14:18], (N, 1)), desired[0, 14:18] + rng.uniform(-0.2, 0.2, (B, N, 4))):
        seedtraj = problems.constant_state_trajectory(x0, N, model["dt_s"], desired[0, 14:18])
        seedtraj[:, :, 14:18] = controls
        initial = s.forward_sim(seedtraj, np.zeros((B, N, 4)), np.zeros((B, N, 48)))
        want = s.solve(initial, desired)
        got = s.solve_from_controls(x0, controls, desired, want_controls=True)
        assert np.array_equal(got["results"], want["results"])
        assert np.array_equal(got["traj"], want["traj"])            # time_s = knot * dt_s as well
        assert np.array_equal(got["controls"], want["traj"][:, :, 14:18])
        only_u = s.solve_from_controls(x0, controls, desired, want_traj=False, want_controls=True)
        assert only_u["traj"] is None and np.array_equal(only_u["controls"], got["controls"])
