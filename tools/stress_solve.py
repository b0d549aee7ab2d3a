#!/usr/bin/env python
"""Determinism / race check of kcbs_kcbs_solve: many seeds per scene, a digest of every solution.  Run it twice (or with
KCBS_RRT_ROUNDS=stream) and diff the output: a plan depends on the seed only.  Usage: python tools/stress_solve.py scene n_seeds [merge_bound]"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

pkg = graft.load_package()
scene = sys.argv[1] if len(sys.argv) > 1 else "Congested10x10_7robots"
n_seeds = int(sys.argv[2]) if len(sys.argv) > 2 else 12
merge_bound = int(sys.argv[3]) if len(sys.argv) > 3 else None
with pkg.Context(pkg.scenes.world(scene)) as ctx:
    starts, goals = pkg.scenes.start_states(scene), pkg.scenes.goals(scene)
    for seed in range(1, n_seeds + 1):
        plan, lengths, st = ctx.kcbs_solve(starts, goals, ctx.kcbs_params(low_level=ctx.rrt_params(seed=seed), time_limit_s=60.0, **({} if merge_bound is None else {"merge_bound": merge_bound})))
        h = hashlib.sha1(np.ascontiguousarray(plan).tobytes() + np.ascontiguousarray(lengths).tobytes()).hexdigest()[:12]
        print(f"{scene} seed {seed}: solved={st.solved} nodes={st.nodes_expanded} plans={st.low_level_calls} approx={st.approximate_solutions} "
              f"merges={st.merges} max_len={int(lengths.max())} digest={h}")
