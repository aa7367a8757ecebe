"""quadrotorilqr_b200/csrc/qilqr_portable_libm.h -- the sin / cos / atan2 shared by the STRICT + portable-libm CUDA
build and the `plibm` oracle build for bit-for-bit comparisons: accuracy against long-double references (<= 1 ulp on
the argument ranges of this code), and the plibm oracle itself against the reference's known answers."""
import os
import re
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_portable_sin_cos_atan2_are_accurate_to_one_ulp(tmp_path):
    exe = str(tmp_path / "plibm_check")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-o", exe,
                           os.path.join(ROOT, "tests", "cpp", "portable_libm_check.cc")])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120).stdout
    m = re.search(r"sin ([\d.]+) cos ([\d.]+) atan2 ([\d.]+)", out)
    assert m, out
    assert all(float(v) <= 1.0 for v in m.groups()), out
    # special values of atan2 as the C library returns them
    assert "0x1.921fb54442d18p+1 0x1.921fb54442d18p+0 -0x0p+0" in out


def test_plibm_oracle_agrees_with_the_libm_oracle():
    """Same algorithm, sin / cos / atan2 one ulp apart at most: one solve of the default problem agrees to 1e-9 and
    takes the same number of iterations (it is not a threshold case)."""
    code = ("import json, numpy as np, oracle as O\n"
            "from quadrotorilqr_b200 import problems\n"
            "m = problems.default_model(); d = problems.default_desired_trajectory()\n"
            "cfg = O.make_config(mass_kg=m['mass_kg'], inertia=m['inertia'], arm_length_m=m['arm_length_m'],"
            " torque_to_thrust_ratio_m=m['torque_to_thrust_ratio_m'], g_mpss=m['g_mpss'], Q=m['Q'], R=m['R'], dt_s=m['dt_s'])\n"
            "r = O.solve(cfg, d, d)\n"
            "print(json.dumps(dict(bp=int(r['backward_passes']), cost=float(r['final_cost']), traj=r['traj'].tolist())))\n")
    outs = []
    for lib in ("", "plibm"):
        env = dict(os.environ, QORACLE_LIB=lib, PYTHONPATH=ROOT)
        p = subprocess.run(["python", "-c", code], env=env, capture_output=True, text=True, timeout=300, cwd=ROOT)
        assert p.returncode == 0, p.stderr[-2000:]
        import json

        outs.append(json.loads(p.stdout.strip().splitlines()[-1]))
    a, b = outs
    assert a["bp"] == b["bp"] == 77
    assert abs(a["cost"] - b["cost"]) <= 1e-9 * abs(a["cost"])
    ta, tb = np.array(a["traj"]), np.array(b["traj"])
    assert np.max(np.abs(ta - tb)) <= 1e-9 * max(1.0, np.max(np.abs(ta)))
