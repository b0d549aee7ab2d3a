#!/usr/bin/env python
"""Wall time of kcbs_rrt_plan_many for 1..30 concurrent low-level plans (how well the planner lanes overlap on the GPU).
Usage: python tools/profile_lanes.py [scene]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

pkg = graft.load_package()
scene = sys.argv[1] if len(sys.argv) > 1 else "Benchmark_Empty32x32_30robots"
with pkg.Context(pkg.scenes.world(scene), device=0) as ctx:
    starts, goals = pkg.scenes.start_states(scene), pkg.scenes.goals(scene)
    R = starts.shape[1]
    for n in (1, 1, 2, 4, 8, 16, 30, 30, 16, 8, 4, 2, 1):
        n = min(n, R)
        robots = np.arange(n, dtype=np.int32)
        l0 = ctx.launch_count()
        t0 = time.perf_counter()
        trajs, lens, stats = ctx.rrt_plan_many(robots, np.ascontiguousarray(starts[:, :n]), goals[:n], ctx.rrt_params(seed=3),
                                         seeds=np.arange(n, dtype=np.uint64) + 11)
        dt = time.perf_counter() - t0
        its = [s.iterations for s in stats]
        print(f"{n:3d} plans: {dt * 1e3:7.2f} ms wall, {dt * 1e3 / n:6.2f} ms per plan, iterations {min(its)}-{max(its)}, "
              f"launches {ctx.launch_count() - l0}, slowest plan {max(s.seconds for s in stats) * 1e3:.2f} ms")
