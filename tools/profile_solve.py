#!/usr/bin/env python
"""One K-CBS solve per scene (for `ncu --metrics gpu__time_duration.sum` launch lists and wall-clock breakdowns).
Usage: python tools/profile_solve.py [scene] [seed] [repeats]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

pkg = graft.load_package()
scene = sys.argv[1] if len(sys.argv) > 1 else "Benchmark_Empty32x32_30robots"
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
max_nodes = int(sys.argv[4]) if len(sys.argv) > 4 and int(sys.argv[4]) > 0 else None   # low-level tree size (default of kcbs_rrt_defaults when absent / 0)
extra = {}
if len(sys.argv) > 5: extra["n_targets"] = int(sys.argv[5])      # targets per RRT round (default 1024)
if len(sys.argv) > 6: extra["n_controls"] = int(sys.argv[6])     # controls per target (default 32)
if len(sys.argv) > 7: extra["goal_bias"] = float(sys.argv[7])
with pkg.Context(pkg.scenes.world(scene), device=0) as ctx:
    starts, goals = pkg.scenes.start_states(scene), pkg.scenes.goals(scene)
    for rep in range(reps):
        l0 = ctx.launch_count()
        t0 = time.perf_counter()
        _, lengths, st = ctx.kcbs_solve(starts, goals, ctx.kcbs_params(low_level=ctx.rrt_params(seed=seed + rep, **extra, **({} if max_nodes is None else {'max_nodes': max_nodes})), time_limit_s=60.0))
        dt = time.perf_counter() - t0
        print(f"{scene} seed {seed + rep}: solved={st.solved} {dt * 1e3:.1f} ms, root {st.root_seconds * 1e3:.1f} ms, nodes {st.nodes_expanded}, "
              f"low-level plans {st.low_level_calls}, validations {st.conflict_checks}, merges {st.merges}, "
              f"launches {ctx.launch_count() - l0}, max path {int(lengths.max())}")
    # one low-level plan alone, for its iteration count
    t0 = time.perf_counter()
    traj, rs = ctx.rrt_plan(0, starts[:, 0], goals[0], ctx.rrt_params(seed=seed, **extra))
    print(f"single RRT robot 0: {(time.perf_counter() - t0) * 1e3:.2f} ms, iterations {rs.iterations}, nodes {rs.nodes}, steps {rs.path_steps}")
