#!/usr/bin/env python
"""Per-kernel hardware counters of an `ncu --set full` capture (read on the CPU box) as JSON: executed FP64
instructions (DFMA / DMUL / DADD thread-instructions, predicated on), DRAM bytes, duration, pipe utilisation,
shared-memory wavefronts.  bench.py reads the committed copy (profiles/r2/r2_kernel_counters.json) for
`roofline.traffic` and `roofline.executed_*`.

usage: python tools/ncu_counters.py gpurun_out/prof.ncu-rep PROBLEM_KNOTS > profiles/r2/r2_kernel_counters.json
PROBLEM_KNOTS = problem-knots every captured launch processed (65536 x 40 for the full-batch capture)."""
import csv
import io
import json
import subprocess
import sys

rep, knots = sys.argv[1], float(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, data = rows[0], rows[2:]


def col(row, name, default=None):
    if name not in hdr:
        return default
    v = row[hdr.index(name)].replace(",", "")
    try:
        return float(v)
    except ValueError:
        return default


UNIT = {}
for i, h in enumerate(hdr):
    UNIT[h] = rows[1][i]


def bytes_of(row, name):
    v = col(row, name, 0.0)
    u = UNIT.get(name, "byte")
    return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)


def ms_of(row, name):
    v = col(row, name, 0.0)
    u = UNIT.get(name, "ms")
    return v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1.0)


out = {"source": f"ncu --set full --clock-control none ({rep.split('/')[-1]}); one launch each over {int(knots)} problem-knots",
       "problem_knots_per_launch": knots, "kernels": {}}
for row in data:
    name = row[hdr.index("Kernel Name")]
    short = name.split("(")[0].replace("void ", "").replace("qilqr::", "")
    dfma = col(row, "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", 0.0)
    dmul = col(row, "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", 0.0)
    dadd = col(row, "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", 0.0)
    flops = 2.0 * dfma + dmul + dadd
    dur = ms_of(row, "gpu__time_duration.sum")
    rd, wr = bytes_of(row, "dram__bytes_read.sum"), bytes_of(row, "dram__bytes_write.sum")
    entry = {
        "grid": row[hdr.index("Grid Size")], "block": row[hdr.index("Block Size")], "duration_ms": dur,
        "registers_per_thread": col(row, "launch__registers_per_thread"),
        "dfma": dfma, "dmul": dmul, "dadd": dadd, "executed_flops": flops,
        "executed_flops_per_problem_knot": flops / knots,
        "executed_tflops_in_this_capture": flops / (dur * 1e-3) / 1e12 if dur else None,
        "dram_read_bytes": rd, "dram_write_bytes": wr, "dram_bytes_per_problem_knot": (rd + wr) / knots,
        "fp64_pipe_pct": col(row, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
        "tensor_pipe_pct": col(row, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        "dram_throughput_pct": col(row, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "warps_active_pct": col(row, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        "issue_active_pct": col(row, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "shared_bank_conflicts": col(row, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
        "shared_loads": col(row, "sass__inst_executed_shared_loads"),
        "shared_stores": col(row, "sass__inst_executed_shared_stores"),
        "local_loads": col(row, "sass__inst_executed_local_loads"),
        "local_stores": col(row, "sass__inst_executed_local_stores"),
    }
    key = short
    n = 2
    while key in out["kernels"]:
        key = f"{short}#{n}"
        n += 1
    out["kernels"][key] = entry
k = out["kernels"]
lin = next((v for n_, v in k.items() if n_.startswith("k_linearise")), None)
ric = next((v for n_, v in k.items() if n_.startswith("k_riccati_g4")), None)
if lin and ric:
    out["backward_pass"] = {
        "executed_flops_per_problem_knot": lin["executed_flops_per_problem_knot"] + ric["executed_flops_per_problem_knot"],
        "dram_bytes_per_problem_knot": lin["dram_bytes_per_problem_knot"] + ric["dram_bytes_per_problem_knot"],
        "algorithmic_bytes_per_problem_knot": 552.0,
        "dense_reference_flops_per_problem_knot": 30231.3,
    }
print(json.dumps(out, indent=1))
