#!/usr/bin/env python
"""One low-level plan (kcbs_rrt_plan) after a warm-up plan: wall clock here, per-kernel times under
`ncu --metrics gpu__time_duration.sum`.  Usage: python tools/profile_rrt.py [scene] [robot] [mode]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

pkg = graft.load_package()
scene = sys.argv[1] if len(sys.argv) > 1 else "Benchmark_Empty32x32_30robots"
robot = int(sys.argv[2]) if len(sys.argv) > 2 else 0
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 0
with pkg.Context(pkg.scenes.world(scene), device=0) as ctx:
    ctx.set_expansion_mode(mode)
    starts, goals = pkg.scenes.start_states(scene), pkg.scenes.goals(scene)
    ctx.rrt_plan(robot, starts[:, robot], goals[robot], ctx.rrt_params(seed=99))
    for seed in (1, 2, 3):
        l0 = ctx.launch_count()
        t0 = time.perf_counter()
        traj, rs = ctx.rrt_plan(robot, starts[:, robot], goals[robot], ctx.rrt_params(seed=seed))
        dt = time.perf_counter() - t0
        print(f"{scene} robot {robot} seed {seed} mode {mode}: {dt * 1e3:.3f} ms wall, {rs.seconds * 1e3:.3f} ms inside, iterations {rs.iterations}, "
              f"nodes {rs.nodes}, steps {rs.path_steps}, launches {ctx.launch_count() - l0}, {dt * 1e6 / max(rs.iterations, 1):.1f} us per iteration")
