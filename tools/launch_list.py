#!/usr/bin/env python
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total time and share per kernel.
Usage: python tools/launch_list.py gpurun_out/launches.csv "<command that was profiled>" [> profiles/x.md]"""
import csv
import sys
from collections import OrderedDict

path, command = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = OrderedDict()
for r in rows:
    name = r[ik]
    n, t = agg.get(name, (0, 0.0))
    agg[name] = (n + 1, t + float(r[iv].replace(",", "")) / 1e3)


def table(items, title):
    tot = sum(t for _, (_, t) in items)
    print(f"\n{title}\n\n| kernel | launches | total us | share | avg us |\n|---|---|---|---|---|")
    for name, (n, t) in sorted(items, key=lambda kv: -kv[1][1]):
        print(f"| `{name[:90]}` | {n} | {t:.1f} | {100 * t / tot:.1f}% | {t / n:.2f} |")


print(f"# ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none -c {len(rows)}`)\n")
print(f"Command: `{command}`.  Per-launch times are cold-cache and serialised; what must agree with the bench is the SHARE.")
print(f"\nProfiled launches: {len(rows)}")
table(list(agg.items()), "All launches")
table([(k, v) for k, v in agg.items() if "probe" not in k], "Without the FFMA peak probe (it runs before and after the timed region, outside it)")
