#!/usr/bin/env python
"""Throughput of the three kernels over the batch size (one launch timed alone, CUDA events, 20 repetitions after
3 warm-ups): where each kernel leaves the latency regime.  Usage (GPU box): python tools/sweep.py > profiles/r1_size_sweep.md"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

pkg = graft.load_package()
from kcbs_b200 import workloads  # noqa: E402

stream = torch.cuda.Stream()


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3  # us


print("# Size sweep, round 1 (B200; one launch at a time, device-resident buffers, CUDA events)\n")
print("## K1 k_propagate — Empty32x32, one shared parent, durations U{1..10}\n")
print("| control samples | us per launch | propagated states/s | algorithmic TFLOP/s |\n|---|---|---|---|")
scene = "Empty32x32_10robots"
ctx = pkg.Context(pkg.scenes.world(scene))
for n in (4096, 16384, 65536, 262144, 1048576, 4194304):
    p, c, s = workloads.expansion_batch(scene, n, seed=3, shared_parent=True, interior=1.5)
    dp, dc, ds = torch.from_numpy(p).cuda(), torch.from_numpy(c).cuda(), torch.from_numpy(s).cuda()
    seg = torch.empty((5, 10, n), device="cuda")
    fin = torch.empty((5, n), device="cuda")
    val = torch.zeros(n, dtype=torch.int32, device="cuda")
    us = timed(lambda: ctx.propagate_batch_dev(0, n, 1, dp, dc, ds, None, 10, seg, fin, val, stream))
    units = float(torch.minimum(ds, val + 1).sum().item())
    print(f"| {n} | {us:.1f} | {units / us * 1e6:.3e} | {980 * units / us * 1e6 / 1e12:.1f} |")
    del seg, fin, val
ctx.close()

print("\n## K2 k_validity — Congested10x10 (9 obstacles) and Empty32x32\n")
print("| states | Congested us | Congested GB/s | Empty us | Empty GB/s |\n|---|---|---|---|---|")
cc, ce = pkg.Context(pkg.scenes.world("Congested10x10_7robots")), pkg.Context(pkg.scenes.world("Empty32x32_10robots"))
for n in (1 << 14, 1 << 16, 1 << 18, 1 << 20, 1 << 22, 1 << 24, 1 << 26):
    row = []
    for ctx, sc in ((cc, "Congested10x10_7robots"), (ce, "Empty32x32_10robots")):
        st = torch.from_numpy(workloads.validity_states(sc, n, seed=1)).cuda()
        mask = torch.empty((n + 31) // 32, dtype=torch.int32, device="cuda")
        us = timed(lambda: ctx.validity_batch_dev(0, n, st, mask, stream))
        row += [f"{us:.1f}", f"{20.125 * n / us * 1e6 / 1e9:.0f}"]
        del st, mask
    print(f"| {n} | " + " | ".join(row) + " |")
cc.close()
ce.close()

print("\n## K3 k_conflict_keys — Empty32x32, 30 robots (435 pairs), 10,000 steps per plan\n")
print("| plans | us per launch | state-pairs/s | GB/s (12 B per trajectory state) |\n|---|---|---|---|")
scene = "Benchmark_Empty32x32_30robots"
ctx = pkg.Context(pkg.scenes.world(scene))
R, T = 30, 10000
for P in (1, 4, 16, 64, 256):
    traj = workloads.cell_plans(scene, P, T, seed=1, xp=torch)
    res = torch.empty((P, 4), dtype=torch.int32, device="cuda")
    us = timed(lambda: ctx.first_conflict_dev(P, R, T, traj, None, res, stream))
    print(f"| {P} | {us:.1f} | {P * 435 * T / us * 1e6:.3e} | {12.0 * P * R * T / us * 1e6 / 1e9:.0f} |")
    del traj, res
ctx.close()
