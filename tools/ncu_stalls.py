#!/usr/bin/env python
"""Stall reasons and executed instructions per opcode from the source page of an .ncu-rep (read here, no GPU).
Usage: python tools/ncu_stalls.py file.ncu-rep"""
import csv
import io
import subprocess
import sys
from collections import Counter

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
if rows[0][0] == "Kernel Name":
    print("kernel:", rows[0][1])
    rows = rows[1:]
h = rows[0]
col = {n: i for i, n in enumerate(h)}
stalls = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
tot = Counter()
by_op = Counter()
samples_op = Counter()
n_inst = 0
for r in rows[1:]:
    if len(r) < len(h):
        continue
    for s in stalls:
        tot[s] += float(r[col[s]] or 0)
    parts = r[col["Source"]].split()
    op = parts[1] if parts and parts[0].startswith("@") else (parts[0] if parts else "?")
    op = op.split(".")[0] + ("." + op.split(".")[1] if op.startswith("MUFU") else "")
    by_op[op] += float(r[col["Instructions Executed"]] or 0)
    samples_op[op] += float(r[col["Warp Stall Sampling (All Samples)"]] or 0)
    n_inst += float(r[col["Instructions Executed"]] or 0)
all_s = sum(tot.values())
print(f"\nstall samples (all): {all_s:.0f}")
for s, v in tot.most_common():
    if v:
        print(f"  {s:28s} {v:8.0f}  {100 * v / all_s:5.1f} %")
print(f"\nwarp instructions executed: {n_inst:.0f}")
for op, v in by_op.most_common(28):
    print(f"  {op:12s} {v:12.0f}  {100 * v / n_inst:5.1f} %   stall samples {100 * samples_op[op] / all_s:5.1f} %")
