#!/usr/bin/env python
"""Error statistics of kcbs_propagate_batch (K1) against the CPU oracle on seeded expansion batches: largest absolute
error per component, largest error relative to the parity tolerance (1e-5 + 1e-4 |ref|), valid-step agreement.
Usage (on the GPU box): python tools/prop_error.py [n]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

pkg = graft.load_package()
orc = graft.load_oracle()
orc.build()
from kcbs_b200 import workloads  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 16
for scene, shared in (("Empty32x32_10robots", False), ("Empty32x32_10robots", True), ("Congested10x10_7robots", False)):
    world = pkg.scenes.world(scene)
    oracle = orc.Oracle(world)
    with pkg.Context(world) as ctx:
        for seed in (1, 2):
            p, c, s = workloads.expansion_batch(scene, n, seed=seed, shared_parent=shared, interior=1.5 if "Empty" in scene else 0.0)
            seg, fin, valid = ctx.propagate_batch(0, p, c, s)
            oseg, ofin, ovalid, _ = oracle.propagate_batch(0, p, c, s, n_threads=os.cpu_count() or 1)
            agree = valid == ovalid
            j = np.arange(10)[:, None] < np.minimum(valid, ovalid)[None, :]
            d = np.abs(seg - oseg)
            d[4] = np.minimum(d[4], 2 * np.pi - d[4])
            d = np.where(j[None], d, 0.0)
            tol = 1e-5 + 1e-4 * np.abs(oseg)
            print(f"{scene} shared={int(shared)} seed={seed}: valid agree {agree.mean():.5f}  max abs err x,y,v,phi,th = "
                  + " ".join(f"{d[k].max():.2e}" for k in range(5)) + f"  max err/tol = {(d / tol).max():.3f}")
