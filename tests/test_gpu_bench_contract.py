"""bench.py on the GPU at a reduced batch: the JSON line carries every key of the measurement contract."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def test_bench_line_has_the_contract_keys():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--batch", "4096", "--steps", "4", "--warmup", "3",
                          "--pipeline", "2", "--cpu-sample-per-core", "8"], capture_output=True, text=True, timeout=900,
                         cwd=ROOT)
    assert out.returncode == 0, out.stderr[-3000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert key in j, key
    assert j["unit"] == "solves/s" and j["scaling"] == "weak" and j["dtype"] == "f64" and j["vs_baseline"] is None
    assert j["value"] > 0 and j["gpu_launches"] > 0 and j["steps"] == 4 and j["n_gpus"] == 1
    assert j["converged_fraction"] > 0.99
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(j["e2e"])
    assert j["e2e"]["value"] > 0 and j["e2e"]["h2d_bytes_per_step"] > 0 and j["e2e"]["d2h_bytes_per_step"] > 0
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(j["roofline"])
    assert 0 < j["roofline"]["frac"] < 2.0 and j["roofline"]["peak"] > 10  # dense-count fraction: > 1 is possible
    assert 0 < j["roofline"]["executed_frac"] < 1.0
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(j["cpu_baseline"])
    assert set(("sm_mhz", "sm_max_mhz", "reasons")) <= set(j["clocks"])
    assert j["clocks"]["samples"] >= 1 and j["clocks"]["sm_mhz"] > 0     # sampled inside the timed region
    # the compact entry points give the same results as the full-trajectory call
    for key in ("e2e_from_controls", "e2e_controls_in_controls_out"):
        assert j[key]["value"] > 0 and j[key]["same_results_as_e2e"] is True, key
    assert j["e2e_controls_in_controls_out"]["d2h_bytes_per_step"] < j["e2e"]["d2h_bytes_per_step"] / 4
    assert "workload" in j["config"]
