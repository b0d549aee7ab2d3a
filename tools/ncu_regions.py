#!/usr/bin/env python
"""Stall samples and executed instructions of an .ncu-rep bucketed by code region (every `step` SASS instructions) and by
source line of the .cu file (needs -lineinfo + --import-source on).  Usage: python tools/ncu_regions.py file.ncu-rep [step]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
step = int(sys.argv[2]) if len(sys.argv) > 2 else 100
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
if rows[0][0] == "Kernel Name":
    print("kernel:", rows[0][1])
    rows = rows[1:]
h = rows[0]
col = {n: i for i, n in enumerate(h)}
data = []
for r in rows[1:]:
    if len(r) < len(h):
        continue
    stalls = {c[6:]: float(r[col[c]] or 0) for c in h if c.startswith("stall_") and "Not Issued" not in c}
    data.append((r[col["Source"]].strip(), float(r[col["Warp Stall Sampling (All Samples)"]] or 0),
                 float(r[col["Instructions Executed"]] or 0), stalls))
tot = sum(d[1] for d in data) or 1
ti = sum(d[2] for d in data) or 1
print(f"total samples {tot:.0f}, warp instructions {ti:.0f}")
for k in range(0, len(data), step):
    ch = data[k:k + step]
    s = sum(d[1] for d in ch)
    e = sum(d[2] for d in ch)
    if e == 0 and s == 0:
        continue
    agg = {}
    for d in ch:
        for n, v in d[3].items():
            agg[n] = agg.get(n, 0) + v
    top = sorted(agg.items(), key=lambda x: -x[1])[:3]
    print(f"{k:5d}  samples {100 * s / tot:5.1f}%  instr {100 * e / ti:5.1f}% ({e / 1e6:5.2f}M)  top stalls: "
          + ", ".join(f"{n} {100 * v / max(s, 1):.0f}%" for n, v in top) + f"   | {ch[0][0][:48]}")
