#!/usr/bin/env python
"""Solves every demo scene for several seeds and checks every plan with the CPU oracle: the reference's full conflict scan over
the whole plan, and dynamic feasibility / validity / goal of some robots' paths (the checks of tests/test_gpu_planner.py, on
more seeds).  Usage: python tools/validate_solves.py [n_seeds]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as graft  # noqa: E402

pkg = graft.load_package()
orc = graft.load_oracle()
orc.build()
from test_gpu_planner import check_path  # noqa: E402

n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 10
bad = 0
for scene, bound in (("Corridor10x10_4robots", 10), ("Congested10x10_7robots", 20), ("Empty32x32_10robots", None),
                     ("Empty32x32_20robots", None), ("Benchmark_Empty32x32_30robots", None)):
    oracle = orc.Oracle(pkg.scenes.world(scene))
    starts, goals = pkg.scenes.start_states(scene), pkg.scenes.goals(scene)
    R = starts.shape[1]
    with pkg.Context(pkg.scenes.world(scene)) as ctx:
        for seed in range(1, n_seeds + 1):
            kw = {} if bound is None else {"merge_bound": bound}
            plan, lengths, st = ctx.kcbs_solve(starts, goals, ctx.kcbs_params(low_level=ctx.rrt_params(seed=seed), time_limit_s=60.0, **kw), max_T=2048)
            if not st.solved:
                print(f"{scene} seed {seed}: NOT solved ({st.nodes_expanded} nodes)")
                bad += 1
                continue
            T = int(lengths.max())
            traj = np.zeros((R, 3, T), np.float32)
            for r in range(R):
                traj[r, :, :lengths[r]] = plan[r][[0, 1, 4], :lengths[r]]
            res, _, _ = oracle.first_conflict(traj, lengths=lengths)
            ok = res[0] == -1
            # (single-car feasibility only when nothing was merged: the cars of a merged pair follow the 10-d propagator with
            # its freeze-at-goal rule, which tests/test_gpu_planner.py checks with the pair oracle)
            for r in ([] if st.merges else sorted(set([0, R // 2, R - 1, seed % R]))):
                try:
                    check_path(oracle, plan[r, :, :lengths[r]], starts[:, r], goals[r], r)
                except AssertionError as e:
                    ok = False
                    print(f"{scene} seed {seed} robot {r}: {e}")
            bad += not ok
            print(f"{scene} seed {seed}: {'ok' if ok else 'BAD'} nodes={st.nodes_expanded} merges={st.merges} conflict={res} {st.seconds * 1e3:.1f} ms")
print("bad:", bad)
sys.exit(1 if bad else 0)
