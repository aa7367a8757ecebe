"""BASELINE.json's full size (batch 65536, N = 40) checked through size-independent properties, plus an
oracle spot check on a random sample: the headline workload is correct, not just fast."""
import numpy as np
import pytest

from conftest import make_solver, oracle_config

pytestmark = pytest.mark.gpu

B, N = 65536, 40


@pytest.fixture(scope="module")
def full_batch():
    from quadrotorilqr_b200 import problems

    model, opts = problems.hover_model(), problems.default_options(False)
    s = make_solver(model, opts)
    desired = problems.hover_desired_trajectory(N, model["dt_s"], model["mass_kg"], model["g_mpss"])
    x0 = problems.hover_initial_states(B, seed=2026)
    seedtraj = problems.constant_state_trajectory(x0, N, model["dt_s"], desired[0, 14:18])
    initial = s.forward_sim(seedtraj, np.zeros((B, N, 4)), np.zeros((B, N, 48)))
    r = s.solve(initial, desired, hist_cap=100)
    return s, model, opts, desired, initial, r


def test_statuses_counts_and_cost_histories(full_batch):
    s, model, opts, desired, initial, r = full_batch
    res, hist = r["results"], r["cost_history"]
    assert np.all(np.isin(res["status"], [1, 2, 3]))                     # no line-search failures on this workload
    assert np.mean(np.isin(res["status"], [1, 2])) > 0.999               # converged fraction
    assert res["backward_passes"].min() >= 2 and res["backward_passes"].max() <= 100
    # solve() bookkeeping invariants (ilqr.hh:53-87): exit A -> one more backward pass than completed
    # iterations; exits B and C -> equal; rollouts >= completed iterations
    nd, bp = res["num_debug"], res["backward_passes"]
    assert np.all(bp[res["status"] == 1] == nd[res["status"] == 1] + 1)
    assert np.all(bp[res["status"] != 1] == nd[res["status"] != 1])
    assert np.all(res["rollouts"] >= nd)
    # Armijo: after the unconditional first step every accepted step decreases the cost
    idx = np.arange(100)[None, :]
    valid = (idx >= 1) & (idx < nd[:, None])
    dec = hist[:, 1:] - hist[:, :-1]
    assert np.all(dec[valid[:, 1:]] < 0.0)
    # final_cost is the last history entry, and it is the cost of the returned trajectory
    last = hist[np.arange(B), np.maximum(nd - 1, 0)]
    assert np.array_equal(last, res["final_cost"])
    sample = np.random.default_rng(0).choice(B, 512, replace=False)
    c = s.cost_trajectory(r["traj"][sample], desired)
    assert np.array_equal(c, res["final_cost"][sample])


def test_returned_trajectories_satisfy_the_dynamics(full_batch):
    s, model, opts, desired, initial, r = full_batch
    traj = r["traj"]
    assert np.array_equal(traj[:, 0, 1:14], initial[:, 0, 1:14])        # x0 is kept (ilqr.hh:156)
    assert np.array_equal(traj[:, :, 0], initial[:, :, 0])              # time_s copied through (ilqr.hh:164)
    sample = np.random.default_rng(1).choice(B, 256, replace=False)
    x = traj[sample][:, :-1, 1:14].reshape(-1, 13)
    u = traj[sample][:, :-1, 14:18].reshape(-1, 4)
    xn = s.discrete_dynamics(x, u).reshape(len(sample), N - 1, 13)
    assert np.array_equal(xn, traj[sample][:, 1:, 1:14])                # bit-exact: same device function
    q = traj[:, :, 4:8]
    assert np.max(np.abs(np.sum(q * q, axis=2) - 1.0)) < 1e-12          # quaternions stay normalised


def test_batch_composition_invariance_at_full_size(full_batch):
    s, model, opts, desired, initial, r = full_batch
    sample = np.sort(np.random.default_rng(2).choice(B, 300, replace=False))
    sub = s.solve(initial[sample], desired, hist_cap=100)
    assert np.array_equal(sub["traj"], r["traj"][sample])
    assert np.array_equal(sub["results"], r["results"][sample])
    assert np.array_equal(sub["cost_history"], r["cost_history"][sample])


def test_oracle_spot_check_at_full_size(O, full_batch):
    s, model, opts, desired, initial, r = full_batch
    cfg = oracle_config(O, model, opts)
    res = r["results"]
    # the slowest problems (including the ones that hit max_iters) and a random sample
    hard = np.argsort(-res["backward_passes"])[:8]
    sample = np.unique(np.concatenate([hard, np.random.default_rng(3).choice(B, 40, replace=False)]))
    o = O.solve_batch(cfg, desired, initial[sample], hist_cap=100)
    assert np.array_equal(res["status"][sample], o["status"])
    assert np.array_equal(res["backward_passes"][sample], o["backward_passes"])
    assert np.array_equal(res["rollouts"][sample], o["rollouts"])
    for j, b in enumerate(sample):
        err = np.max(np.abs(r["traj"][b] - o["traj"][j]))
        assert err <= 1e-9 * max(1.0, np.max(np.abs(o["traj"][j]))), (b, err)
        herr = np.max(np.abs(r["cost_history"][b] - o["cost_history"][j]))
        assert herr <= 1e-9 * max(1.0, np.max(np.abs(o["cost_history"][j]))), (b, herr)


def test_every_problem_of_the_full_batch_against_the_oracle(O, full_batch):
    """All 65536 problems against the CPU oracle (about 15 s on 16 host cores).

    The north star asks for IDENTICAL iteration counts and convergence flags.  What is achievable, and what this
    test holds the production build to (measured: profiles/r2/r2_parity_c3_*.json, four Philox seeds):

    * every discrete decision (flag, backward passes, rollouts) is identical except for a handful of problems --
      6 to 14 of 65536 -- whose LAST iterations sit on a threshold of the reference: the rtol = 1e-12 test of
      ilqr.hh:196-205 with a relative cost step between 0.9995 and 1.0005 x rtol, or the Armijo test of
      ilqr.hh:186 on a final step that changes the cost by 1e-13 of itself.  Any two implementations whose
      roundings differ anywhere decide these differently: the STRICT build (no FMA, true divisions) differs from
      the oracle on 6-14 problems as well, through the CUDA math library alone, and the STRICT + portable-libm
      build, which shares sin / cos / atan2 with the oracle, on NONE (test_strict_build_* below);
    * bar: at most 32 such problems; each at most one iteration and one rollout apart, converged on both sides,
      the relative cost step of its last iteration below 10 x rtol on both sides (it is at the convergence
      threshold), final costs equal to 2e-12;
    * everywhere else: trajectories, cost histories and final costs within 1e-9 (measured 1.3e-11 / 2e-15)."""
    s, model, opts, desired, initial, r = full_batch
    cfg = oracle_config(O, model, opts)
    o = O.solve_batch(cfg, desired, initial, hist_cap=100)
    res = r["results"]
    same = ((res["status"] == o["status"]) & (res["backward_passes"] == o["backward_passes"])
            & (res["rollouts"] == o["rollouts"]))
    diff = np.where(~same)[0]
    assert diff.size <= 32, diff.size
    scale = np.maximum(1.0, np.max(np.abs(o["traj"]), axis=(1, 2)))
    err = np.max(np.abs(r["traj"] - o["traj"]), axis=(1, 2)) / scale
    cerr = np.abs(res["final_cost"] - o["final_cost"]) / np.maximum(1.0, np.abs(o["final_cost"]))
    herr = np.max(np.abs(r["cost_history"] - o["cost_history"]) / np.maximum(1.0, np.abs(o["cost_history"])), axis=1)
    assert err[same].max() <= 1e-9, (int(err.argmax()), float(err.max()))
    assert herr[same].max() <= 1e-9 and cerr[same].max() <= 1e-9
    rtol = opts.convergence_criteria.rtol

    def last_rel_step(hist, nd):
        return abs(hist[nd - 2] - hist[nd - 1]) / abs(hist[nd - 2]) / rtol

    for b in diff:
        assert abs(int(res["backward_passes"][b]) - int(o["backward_passes"][b])) <= 1, b
        assert abs(int(res["rollouts"][b]) - int(o["rollouts"][b])) <= 1, b
        assert res["status"][b] in (1, 2) and o["status"][b] in (1, 2), b
        assert cerr[b] <= 2e-12, (b, cerr[b])
        nd_o = int(np.count_nonzero(o["cost_history"][b]))
        assert last_rel_step(r["cost_history"][b], int(res["num_debug"][b])) <= 10.0, b
        assert last_rel_step(o["cost_history"][b], nd_o) <= 10.0, b
    print(f"identical decisions: {int(same.sum())}/{same.size}; threshold cases: {diff.tolist()}; max rel traj err where "
          f"identical {err[same].max():.2e}; max rel final-cost err overall {cerr.max():.2e}")


def _parity_tool(config, lib, oracle_lib=None, batch=None):
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, CONFIG=config, QILQR_LIB=os.path.join(root, "quadrotorilqr_b200", lib))
    if oracle_lib:
        env["QORACLE_LIB"] = oracle_lib
    if batch:
        env["B"] = str(batch)
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "full_batch_parity.py")], env=env,
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    return json.loads(out.stdout.strip().splitlines()[-1])


def test_strict_build_with_the_shared_libm_is_bit_identical_to_the_oracle_on_the_full_batch():
    """The STRICT + portable-libm build (no FMA, true divisions, sin / cos / atan2 from the source the oracle is
    built with too) against that oracle on ALL 65536 problems: identical decisions everywhere and every double of
    every trajectory and cost history equal -- the CUDA implementation executes the reference's arithmetic, operation
    for operation.  (The production build differs from it only by the explicit fused multiply-adds, the reciprocal
    multiplies and the CUDA math library.)"""
    d = _parity_tool("c3", "libqilqr_b200_strict_plibm.so", oracle_lib="plibm")
    assert d["build"].startswith("strict+portable-libm") and d["batch"] == B
    assert d["different_decisions"] == 0
    assert d["bit_identical_trajectories"] == B and d["bit_identical_cost_histories"] == B


def test_strict_build_is_bit_identical_where_no_libm_call_is_made():
    """STRICT build (CUDA math library) on problems whose attitude stays the identity, so that no sin / cos / atan2
    is ever evaluated: bit-identical to the oracle.  Whatever the STRICT build differs by on the hover batch
    (6-14 threshold decisions of 65536) comes from the math library alone."""
    d = _parity_tool("libmfree", "libqilqr_b200_strict.so")
    assert d["build"].startswith("strict:") and d["different_decisions"] == 0
    assert d["bit_identical_trajectories"] == d["batch"] and d["initial_rollout_bit_identical"]
