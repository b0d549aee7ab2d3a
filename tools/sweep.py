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
    """Microseconds per call: `reps` calls captured into one CUDA graph (the C-ABI launches themselves; no Python between
    them, a ctypes call costs ~9 us) and replayed; CUDA events on the replay stream."""
    with torch.cuda.stream(stream):
        for _ in range(3):
            fn()
        stream.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            for _ in range(reps):
                fn()
        g.replay()
        stream.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(5):
            g.replay()
        e1.record(stream)
        stream.synchronize()
    return e0.elapsed_time(e1) / (5 * reps) * 1e3  # us


print("# Size sweep, round 2 (B200; back-to-back launches of ONE stream replayed from a CUDA graph, device-resident buffers, CUDA events)\n")
print("## K1 k_propagate — Empty32x32, one shared parent, durations U{1..10}\n")
print("| control samples | us per launch | propagated states/s | algorithmic TFLOP/s |\n|---|---|---|---|")
scene = "Empty32x32_10robots"
ctx = pkg.Context(pkg.scenes.world(scene))
print("(columns: launch shape forced to one pair per thread / two pairs per thread / chosen from n; dense and packed segments)\n")
print("| control samples | dense, 1 pair us | dense, 2 pairs us | dense, auto us | packed, auto us | final only, auto us | states/s (dense auto) | TFLOP/s |\n|---|---|---|---|---|---|---|---|")
for n in (4096, 16384, 65536, 131072, 262144, 1048576, 4194304):
    p, c, s = workloads.expansion_batch(scene, n, seed=3, shared_parent=True, interior=1.5)
    dp, dc, ds = torch.from_numpy(p).cuda(), torch.from_numpy(c).cuda(), torch.from_numpy(s).cuda()
    seg = torch.empty((5, 10, n), device="cuda")
    fin = torch.empty((5, n), device="cuda")
    val = torch.zeros(n, dtype=torch.int32, device="cuda")
    off = torch.zeros(n + 1, dtype=torch.int32, device="cuda")
    row = []
    for mode in (1, 2, 0):
        ctx.set_expansion_mode(mode)
        row.append(timed(lambda: ctx.propagate_batch_dev(0, n, 1, dp, dc, ds, None, 10, seg, fin, val, stream)))
    us = row[2]
    units = float(torch.minimum(ds, val + 1).sum().item())
    row.append(timed(lambda: ctx.propagate_packed_dev(0, n, 1, dp, dc, ds, None, 10, seg, 10 * n, off, fin, val, stream)))
    row.append(timed(lambda: ctx.propagate_batch_dev(0, n, 1, dp, dc, ds, None, 10, None, fin, val, stream)))
    print(f"| {n} | " + " | ".join(f"{x:.1f}" for x in row) + f" | {units / us * 1e6:.3e} | {980 * units / us * 1e6 / 1e12:.1f} |")
    del seg, fin, val, off
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
