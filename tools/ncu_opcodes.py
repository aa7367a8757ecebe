#!/usr/bin/env python
"""Executed-instruction histogram by opcode of one kernel of an `ncu --set full --import-source on` capture (read
on the CPU box): warp-instructions executed and stall samples per opcode -- which non-FP64 work a kernel spends
its issue slots on.   usage: python tools/ncu_opcodes.py REP KERNEL_REGEX [TOP]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name",
                      f"regex:{kern}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = next(r for r in rows if r and r[0] == "Address")
iS, iE, iW = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
execd, stall = collections.Counter(), collections.Counter()
first = True
for r in rows:
    if not r or not r[0].startswith("0x"):
        if r and r[0] == "Kernel Name":
            if not first:
                break  # first matching launch only
            first = False
            print(r[1])
        continue
    m = re.match(r"\s*(?:@!?U?P\w+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", r[iS])
    if not m:
        continue
    op = m.group(1)
    base = op.split(".")[0]
    key = op if base in ("IMAD", "LDS", "STS", "LDG", "STG", "SHFL", "MUFU") else base
    execd[key] += int(r[iE] or 0)
    stall[key] += int(r[iW] or 0)
te, ts = sum(execd.values()), sum(stall.values())
print(f"{'opcode':18s} {'warp-inst':>12s} {'%':>6s} {'stall samples':>14s} {'%':>6s}")
for k, v in execd.most_common(top):
    print(f"{k:18s} {v:12d} {100.0 * v / te:6.2f} {stall[k]:14d} {100.0 * stall[k] / max(ts, 1):6.2f}")
print(f"{'total':18s} {te:12d} {100.0:6.2f} {ts:14d}")
