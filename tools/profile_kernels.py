#!/usr/bin/env python
"""Small profiling target: a handful of launches of each hot-path kernel at the benchmark shapes, with as
little setup as possible, so `ncu` runs stay short.  Usage (on the GPU box):
    python tools/profile_kernels.py [propagate|validity|conflict|all]
Prints CUDA-event timings; under ncu, filter with -k regex:k_propagate etc."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

pkg = graft.load_package()
from kcbs_b200 import workloads  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
stream = torch.cuda.Stream()


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


if what in ("probe", "all"):
    ctx = pkg.Context(pkg.scenes.world("Empty32x32_10robots"))
    print(f"FFMA probe {ctx.measure_fp32_peak():.1f} TFLOP/s; packed fma.rn.f32x2 probe {ctx.measure_fp32x2_peak():.1f} TFLOP/s")
    ctx.close()

if what in ("propagate", "all"):
    for scene, fixed in (("Empty32x32_10robots", None), ("Empty32x32_10robots", 10), ("Congested10x10_7robots", None)):
        ctx = pkg.Context(pkg.scenes.world(scene))
        for n, shared in ((65536, True), (655360, False)):
            p, c, s = workloads.expansion_batch(scene, n, seed=3, shared_parent=shared, interior=1.5, fixed_steps=fixed)
            dp, dc, ds = torch.from_numpy(p).cuda(), torch.from_numpy(c).cuda(), torch.from_numpy(s).cuda()
            seg = torch.empty((5, 10, n), device="cuda")
            fin = torch.empty((5, n), device="cuda")
            val = torch.zeros(n, dtype=torch.int32, device="cuda")
            ms = timed(lambda: ctx.propagate_batch_dev(0, n, p.shape[1], dp, dc, ds, None, 10, seg, fin, val, stream), 5)
            units = float(torch.minimum(ds, val + 1).sum().item())
            print(f"propagate {scene} n={n} steps={'U' if fixed is None else fixed}: {ms * 1e3:.1f} us  {units / ms / 1e-3:.3e} states/s "
                  f"({980 * units / ms / 1e-3 / 1e12:.1f} TFLOP/s algorithmic)")
            off = torch.zeros(n + 1, dtype=torch.int32, device="cuda")
            ms = timed(lambda: ctx.propagate_packed_dev(0, n, p.shape[1], dp, dc, ds, None, 10, seg, 10 * n, off, fin, val, stream), 5)
            print(f"propagate PACKED {scene} n={n} steps={'U' if fixed is None else fixed}: {ms * 1e3:.1f} us  {units / ms / 1e-3:.3e} states/s "
                  f"({980 * units / ms / 1e-3 / 1e12:.1f} TFLOP/s algorithmic); rows {int(off[n])}")
        ctx.close()

if what in ("validity", "all"):
    for scene in ("Congested10x10_7robots", "Empty32x32_10robots"):
        ctx = pkg.Context(pkg.scenes.world(scene))
        n = 1 << 24
        st = torch.from_numpy(workloads.validity_states(scene, n, seed=1)).cuda()
        mask = torch.empty((n + 31) // 32, dtype=torch.int32, device="cuda")
        ms = timed(lambda: ctx.validity_batch_dev(0, n, st, mask, stream), 5)
        print(f"validity {scene} n={n}: {ms * 1e3:.1f} us  {n / ms / 1e-3:.3e} states/s  {20.125 * n / ms / 1e-3 / 1e9:.0f} GB/s")
        ctx.close()

if what in ("conflict", "all"):
    scene = "Benchmark_Empty32x32_30robots"
    ctx = pkg.Context(pkg.scenes.world(scene))
    for P in (256, 1):
        R, T = 30, 10000
        traj = workloads.cell_plans(scene, P, T, seed=1, xp=torch)
        res = torch.empty((P, 4), dtype=torch.int32, device="cuda")
        ms = timed(lambda: ctx.first_conflict_dev(P, R, T, traj, None, res, stream), 5 if P > 1 else 50)
        pairs = P * 435 * T
        print(f"conflict P={P}: {ms * 1e3:.1f} us  {pairs / ms / 1e-3:.3e} pairs/s  {12.0 * P * R * T / ms / 1e-3 / 1e9:.0f} GB/s algorithmic; "
              f"conflicts found: {int((res[:, 0] >= 0).sum())}")
    ctx.close()
