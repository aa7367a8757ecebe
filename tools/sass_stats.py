#!/usr/bin/env python
"""Static instruction statistics per kernel from `cuobjdump -sass` of the built library (runs on the CPU box):
FP64 arithmetic (DFMA / DMUL / DADD), TMA bulk copies (UBLKCP) and mbarrier operations (SYNCS), shared / local
memory instructions, CTA barriers.   usage: python tools/sass_stats.py > profiles/r1_sass_static.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sass = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "quadrotorilqr_b200", "libqilqr_b200.so")],
                      capture_output=True, text=True).stdout
cur, stats = None, collections.defaultdict(collections.Counter)
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        stats[cur][m.group(1).split(".")[0]] += 1


def demangle(n):
    return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().split("(")[0]


cols = ("DFMA", "DMUL", "DADD", "UBLKCP", "SYNCS", "LDS", "STS", "LDL", "STL", "BAR", "HMMA", "UTCHMMA")
print("cuobjdump -sass quadrotorilqr_b200/libqilqr_b200.so (sm_100a): static instruction counts per kernel")
print(f"{'kernel':42s} {'all':>6s} " + " ".join(f"{c:>7s}" for c in cols))
for f, c in sorted(stats.items(), key=lambda kv: -sum(kv[1].values())):
    print(f"{demangle(f)[:42]:42s} {sum(c.values()):6d} " + " ".join(f"{c[k]:7d}" for k in cols))
print("\nUBLKCP = cp.async.bulk (TMA bulk copy, record tiles of the Riccati kernels); SYNCS = mbarrier operations; "
      "no tensor-core instruction (HMMA / UTC*MMA) anywhere: the per-problem 12x12 / 4x4 FP64 algebra is not a "
      "contraction at scale (DESIGN.md section 4).")
