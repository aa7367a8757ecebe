"""ILQRDebug at batch scale (ilqr.hh:78-80, ilqr_debug.hh:9-22): the always-on per-iteration cost
history and the sampled trajectory rings, against the full capture and against the oracle."""
import dataclasses

import numpy as np
import pytest

from conftest import make_solver, oracle_config

pytestmark = pytest.mark.gpu


def hover(s, B, N, seed):
    from quadrotorilqr_b200 import problems

    m = problems.hover_model()
    desired = problems.hover_desired_trajectory(N, m["dt_s"], m["mass_kg"], m["g_mpss"])
    x0 = problems.hover_initial_states(B, seed=seed)
    initial = s.forward_sim(problems.constant_state_trajectory(x0, N, m["dt_s"], desired[0, 14:18]),
                            np.zeros((B, N, 4)), np.zeros((B, N, 48)))
    return desired, initial


def test_sampled_rings_equal_the_full_capture():
    from quadrotorilqr_b200 import problems

    model, opts = problems.hover_model(), problems.default_options(True)
    s = make_solver(model, opts)
    B, N = 96, 40
    desired, initial = hover(s, B, N, seed=21)
    full = s.solve(initial, desired, want_debug=True, hist_cap=100)
    sample = np.array([5, 90, 17, 64, 0], dtype=np.int32)  # any order
    # every iteration, ring large enough: the rings hold exactly the ILQRDebug entries of those problems
    s.set_debug_sampling(sample)
    r = s.solve(initial, desired)
    assert np.array_equal(r["traj"], full["traj"])
    d = s.read_debug_samples(N)
    for j, b in enumerate(sample):
        nd = int(full["results"]["num_debug"][b])
        it, tr, co = s.debug_of(d, j)
        assert int(d["counts"][j]) == nd and np.array_equal(it, np.arange(nd))
        assert np.array_equal(tr, full["debug"][b, :nd])                 # same trajectories, time_s column included
        assert np.array_equal(co, full["cost_history"][b, :nd])          # ILQRIterDebug.cost
    # every 3rd iteration in a ring of 2: the last two sampled iterations survive
    s.set_debug_sampling(sample, every=3, ring=2)
    s.solve(initial, desired)
    d = s.read_debug_samples(N)
    for j, b in enumerate(sample):
        nd = int(full["results"]["num_debug"][b])
        want = np.arange(0, nd, 3)
        it, tr, co = s.debug_of(d, j)
        assert int(d["counts"][j]) == want.size and np.array_equal(it, want[-2:])
        assert np.array_equal(tr, full["debug"][b, want[-2:]])
    # sampling off again: populate_debug without a destination captures nothing, results unchanged
    s.set_debug_sampling(None)
    assert np.array_equal(s.solve(initial, desired)["traj"], full["traj"])


def test_always_on_cost_history():
    from quadrotorilqr_b200 import problems

    model, opts = problems.hover_model(), problems.default_options(False)
    s = make_solver(model, opts)
    B, N = 300, 40
    desired, initial = hover(s, B, N, seed=22)
    want = s.solve(initial, desired, hist_cap=100)
    r = s.solve(initial, desired)                       # no history requested ...
    h = s.last_cost_history(first=10, count=200)        # ... it is there anyway
    assert h.shape[1] == 100
    assert np.array_equal(h, want["cost_history"][10:210])
    assert np.array_equal(r["results"], want["results"])


def test_one_percent_of_the_full_batch_sampled(O):
    """B = 65536 with 1 % of the problems sampled, every 2nd iteration, ring of 8 (46 MB instead of 35 GB)."""
    from quadrotorilqr_b200 import problems

    model = problems.hover_model()
    opts = problems.default_options(True)
    s = make_solver(model, opts)
    B, N = 65536, 40
    desired, initial = hover(s, B, N, seed=2026)
    sample = np.sort(np.random.default_rng(0).choice(B, B // 100, replace=False)).astype(np.int32)
    s.set_debug_sampling(sample, every=2, ring=8)
    r = s.solve(initial, desired)
    d = s.read_debug_samples(N)
    res = r["results"]
    # counts: sampled iterations = ceil(num_debug / 2)
    assert np.array_equal(d["counts"], (res["num_debug"][sample] + 1) // 2)
    # the newest slot of every sampled problem whose last completed iteration is even is its returned trajectory
    for j in range(0, sample.size, 37):
        b, nd = int(sample[j]), int(res["num_debug"][sample[j]])
        it, tr, co = s.debug_of(d, j)
        assert np.array_equal(it, np.arange(0, nd, 2)[-8:])
        if (nd - 1) % 2 == 0:
            assert np.array_equal(tr[-1], r["traj"][b]) and co[-1] == res["final_cost"][b]
    # against the oracle's ILQRDebug for a few of them
    cfg = oracle_config(O, model, dataclasses.replace(opts, populate_debug=True))
    for j in (0, 100, 333):
        b = int(sample[j])
        o = O.solve(cfg, desired, initial[b])
        it, tr, co = s.debug_of(d, j)
        assert o["debug"].shape[0] == int(res["num_debug"][b])
        for i, t_, c_ in zip(it, tr, co):
            ref, ref_cost = o["debug"][int(i)], o["cost_history"][int(i)]
            assert np.max(np.abs(t_ - ref)) <= 1e-9 * max(1.0, np.max(np.abs(ref)))
            assert abs(c_ - ref_cost) <= 1e-9 * max(1.0, abs(ref_cost))
