#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): headline metrics, stall reasons, hottest source lines.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--top 25] > profiles/<name>.txt"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
KEYS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.sum",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "sass__inst_executed_shared_loads", "sass__inst_executed_shared_stores",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg"]
for ki, row in enumerate(data):
    print(f"=== launch {ki} ===")
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k:72s} {row[i]:>22s} {units[i]}")
    print("-- warp stall reasons (smsp__average_warps_issue_stalled_*_per_issue_active, ratio) --")
    st = []
    for i, h in enumerate(hdr):
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            try:
                st.append((float(row[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    for v, n in sorted(st, reverse=True)[:10]:
        print(f"   {n:40s} {v:8.3f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
try:
    rows = list(csv.reader(io.StringIO(src)))
    # find header row
    hi = next(i for i, r in enumerate(rows) if "Source" in r and any("Sampling" in c for c in r))
    h = rows[hi]
    si = h.index("Source")
    samp = next(i for i, c in enumerate(h) if c.startswith("# Samples") or c == "Warp Stall Sampling (All Samples)" or "Sampling (All" in c)
    agg = defaultdict(float)
    for r in rows[hi + 1:]:
        if len(r) <= max(si, samp):
            continue
        try:
            agg[r[si].strip()] += float(r[samp])
        except ValueError:
            pass
    tot = sum(agg.values()) or 1.0
    print(f"-- hottest lines by warp-stall samples ({h[samp]}), first launch --")
    for line, v in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
        print(f"{100*v/tot:6.2f}%  {line[:150]}")
except Exception as e:  # source page layout differs between ncu versions
    print("source page not parsed:", e)
