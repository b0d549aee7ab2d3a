#!/usr/bin/env python
"""K-CBS solve time of the reference's demo scenes with the GPU planner (kcbs_kcbs_solve): the
"K-CBS solve time (Empty32x32)" part of BASELINE.json's metric.  Usage: python tools/solve_scenes.py [scene ...]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

pkg = graft.load_package()
names = sys.argv[1:] or ["Corridor10x10_4robots", "Congested10x10_7robots", "Empty32x32_10robots", "Empty32x32_15robots",
                         "Empty32x32_20robots", "Benchmark_Empty32x32_30robots"]
for scene in names:
    with pkg.Context(pkg.scenes.world(scene)) as ctx:
        starts, goals = pkg.scenes.start_states(scene), pkg.scenes.goals(scene)
        rows = []
        for seed in (1, 2, 3):
            t0 = time.perf_counter()
            plan, lengths, st = ctx.kcbs_solve(starts, goals, ctx.kcbs_params(low_level=ctx.rrt_params(seed=seed), time_limit_s=180.0))
            rows.append({"seed": seed, "solved": st.solved, "seconds": round(time.perf_counter() - t0, 4), "root_s": round(st.root_seconds, 4),
                         "nodes_expanded": st.nodes_expanded, "low_level_calls": st.low_level_calls, "approx": st.approximate_solutions,
                         "max_len": int(lengths.max())})
        print(json.dumps({"scene": scene, "robots": starts.shape[1], "runs": rows}))
