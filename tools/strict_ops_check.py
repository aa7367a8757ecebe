#!/usr/bin/env python
"""Op-by-op comparison of a build of the CUDA library with the CPU oracle in units of ulp.

With QILQR_LIB=quadrotorilqr_b200/libqilqr_b200_strict_plibm.so and QORACLE_LIB=plibm both sides execute
the same IEEE operations in the same order (no FMA, true divisions, shared sin / cos / atan2), so every line
must read `mismatching elements 0`.  With the production build the numbers show how far the explicit fused
multiply-adds, the reciprocal multiplies and the CUDA math library move each result.  One JSON line."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O  # noqa: E402
from quadrotorilqr_b200 import BatchILQR, ILQROptions, _capi, problems  # noqa: E402


def ulps(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    scale = np.maximum(np.spacing(np.maximum(np.abs(b), 1e-300)), 0.0)
    return np.abs(a - b) / scale


def report(out, name, a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    u = ulps(a, b)
    out[name] = {"mismatching_elements": int(np.sum(a != b)), "elements": int(a.size),
                 "max_ulp": float(u.max()) if u.size else 0.0,
                 "max_abs": float(np.max(np.abs(a - b))) if a.size else 0.0}


def random_states(n, seed, big=False):
    rng = np.random.default_rng(seed)
    x = np.zeros((n, 13))
    for i in range(n):
        s = 3.0 if big else 0.7
        tau = rng.uniform(-s, s, 6)
        if i % 7 == 3:
            tau[3:] *= 1e-9
        if i % 11 == 5:
            tau[3:] = 0.0
        x[i, :7] = O.se3_exp(tau)
        x[i, 7:] = rng.uniform(-2, 2, 6)
    return x


def main():
    rng = np.random.default_rng(0)
    A = rng.uniform(-1, 1, (3, 3))
    inertia = A @ A.T + 2.0 * np.eye(3)
    out = {"build": _capi.lib().qilqr_build_info().decode(), "oracle": os.environ.get("QORACLE_LIB", "libm")}
    for mname, model in (("default_model", problems.hover_model()),
                         ("random_inertia", dict(problems.hover_model(), inertia=inertia, mass_kg=1.3, arm_length_m=0.4))):
        s = BatchILQR(model["mass_kg"], model["inertia"], model["arm_length_m"], model["torque_to_thrust_ratio_m"],
                      model["g_mpss"], model["Q"], model["R"], model["dt_s"], ILQROptions())
        cfg = O.make_config(mass_kg=model["mass_kg"], inertia=model["inertia"], arm_length_m=model["arm_length_m"],
                            torque_to_thrust_ratio_m=model["torque_to_thrust_ratio_m"], g_mpss=model["g_mpss"],
                            Q=model["Q"], R=model["R"], dt_s=model["dt_s"])
        n = 256
        x, xd = random_states(n, 1), random_states(n, 2)
        xb, yb = random_states(n, 3, big=True), random_states(n, 4, big=True)
        u, ud = rng.uniform(-3, 6, (n, 4)), rng.uniform(-3, 6, (n, 4))
        tangent = rng.uniform(-2, 2, (n, 12))
        o = {}
        xn, Jx, Ju = s.discrete_dynamics(x, u, diffs=True)
        xc, Jxc, Juc = s.continuous_dynamics(x, u, diffs=True)
        d, Jl, Jr = s.state_minus(xb, yb, diffs=True)
        y, Al, Ar = s.state_add(xb, tangent, diffs=True)
        c, Cx, Cu, Cxx, Cuu, Cxu = s.cost(x, u, xd, ud, diffs=True)
        ref = {k: [] for k in ("xn Jx Ju xc Jxc Juc d Jl Jr y Al Ar c Cx Cu Cxx").split()}
        for i in range(n):
            r = O.discrete_dynamics(cfg, x[i], u[i], diffs=True)
            ref["xn"].append(r[0]); ref["Jx"].append(r[1]); ref["Ju"].append(r[2])
            r = O.continuous_dynamics(cfg, x[i], u[i], diffs=True)
            ref["xc"].append(r[0]); ref["Jxc"].append(r[1]); ref["Juc"].append(r[2])
            r = O.state_minus(xb[i], yb[i], diffs=True)
            ref["d"].append(r[0]); ref["Jl"].append(r[1]); ref["Jr"].append(r[2])
            r = O.state_add(xb[i], tangent[i], diffs=True)
            ref["y"].append(r[0]); ref["Al"].append(r[1]); ref["Ar"].append(r[2])
            r = O.cost(cfg, x[i], u[i], xd[i], ud[i], diffs=True)
            ref["c"].append(r[0]); ref["Cx"].append(r[1]); ref["Cu"].append(r[2]); ref["Cxx"].append(r[3])
        got = dict(xn=xn, Jx=Jx, Ju=Ju, xc=xc, Jxc=Jxc, Juc=Juc, d=d, Jl=Jl, Jr=Jr, y=y, Al=Al, Ar=Ar, c=c, Cx=Cx,
                   Cu=Cu, Cxx=Cxx)
        names = dict(xn="discrete_dynamics x+", Jx="discrete_dynamics J_x", Ju="discrete_dynamics J_u",
                     xc="continuous_dynamics xdot", Jxc="continuous_dynamics J_x", Juc="continuous_dynamics J_u",
                     d="minus", Jl="minus J_lhs", Jr="minus J_rhs", y="add", Al="add J_lhs", Ar="add J_rhs",
                     c="cost", Cx="cost C.x", Cu="cost C.u", Cxx="cost C.xx")
        for k in got:
            report(o, names[k], np.asarray(got[k]).reshape(n, -1), np.asarray(ref[k]).reshape(n, -1))
        # one backward pass and one closed-loop rollout on hover trajectories
        N, B = 40, 64
        desired = problems.hover_desired_trajectory(N, model["dt_s"], model["mass_kg"], model["g_mpss"])
        x0 = problems.hover_initial_states(B, seed=7)
        init = s.forward_sim(problems.constant_state_trajectory(x0, N, model["dt_s"], desired[0, 14:18]),
                             np.zeros((B, N, 4)), np.zeros((B, N, 48)))
        k, K, qutk, ktq = s.backwards_pass(init, desired)
        ok, oK, ot, oroll, ocost = [], [], [], [], []
        for b in range(B):
            r = O.backwards_pass(cfg, desired, init[b])
            ok.append(r[0]); oK.append(r[1]); ot.append([r[2], r[3]])
            rr = O.forward_sim(cfg, desired, init[b], r[0], r[1])
            oroll.append(rr)
            ocost.append(O.cost_trajectory(cfg, desired, rr))
        report(o, "backwards_pass k", k, np.asarray(ok))
        report(o, "backwards_pass K", K.reshape(B, N, 48), np.asarray(oK).reshape(B, N, 48))
        report(o, "backwards_pass terms", np.stack([qutk, ktq], axis=1), np.asarray(ot))
        report(o, "forward_sim (oracle gains on both sides)",
               s.forward_sim(init, np.asarray(ok), np.asarray(oK).reshape(B, N, 48), 1.0), np.asarray(oroll))
        report(o, "cost_trajectory", s.cost_trajectory(np.asarray(oroll), desired), np.asarray(ocost))
        out[mname] = o
    print(json.dumps(out))


if __name__ == "__main__":
    main()
