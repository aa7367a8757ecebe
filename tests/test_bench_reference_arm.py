"""bench.py --impl reference (the CPU arm the driver runs beside the GPU arm): one JSON line with the contract's
keys; under torchrun with two ranks, rank 0 alone works and prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ARGS = ["--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-sample-per-core", "4"]


def check_line(out):
    lines = [l for l in out.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "solves/s" and j["higher_is_better"] is True
    assert j["value"] > 0 and j["steps"] == 1 and j["dtype"] == "f64" and j["data"] == "synthetic"
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in j["config"]
    return j


def test_reference_arm_single_process():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + ARGS, capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr
    assert check_line(out.stdout)["n_gpus"] == 1


def test_reference_arm_under_torchrun_prints_once():
    port = 29600 + os.getpid() % 300
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "bench.py"),
                          "--gpus", "2"] + ARGS, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    check_line(out.stdout)
