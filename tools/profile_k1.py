#!/usr/bin/env python
"""ncu target for K1 alone: a warm-up launch and `reps` timed launches of k_propagate on one batch.
Usage: python tools/profile_k1.py [n] [fixed_steps|0] [dense|packed|final] [scene] [reps] [expansion mode 0|1|2]
Under ncu: -k regex:k_propagate -s 1 -c 1"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

pkg = graft.load_package()
from kcbs_b200 import workloads  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 655360
fixed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
layout = sys.argv[3] if len(sys.argv) > 3 else "dense"
scene = sys.argv[4] if len(sys.argv) > 4 else "Empty32x32_10robots"
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 5
mode = int(sys.argv[6]) if len(sys.argv) > 6 else 0
stream = torch.cuda.Stream()
ctx = pkg.Context(pkg.scenes.world(scene))
ctx.set_expansion_mode(mode)
p, c, s = workloads.expansion_batch(scene, n, seed=3, shared_parent=(n <= 65536), interior=1.5, fixed_steps=fixed or None)
dp, dc, ds = torch.from_numpy(p).cuda(), torch.from_numpy(c).cuda(), torch.from_numpy(s).cuda()
seg = torch.empty((5, 10, n), device="cuda")
fin = torch.empty((5, n), device="cuda")
val = torch.zeros(n, dtype=torch.int32, device="cuda")
off = torch.zeros(n + 1, dtype=torch.int32, device="cuda")


def launch():
    if layout == "packed":
        ctx.propagate_packed_dev(0, n, p.shape[1], dp, dc, ds, None, 10, seg, 10 * n, off, fin, val, stream)
    elif layout == "final":
        ctx.propagate_batch_dev(0, n, p.shape[1], dp, dc, ds, None, 10, None, fin, val, stream)
    else:
        ctx.propagate_batch_dev(0, n, p.shape[1], dp, dc, ds, None, 10, seg, fin, val, stream)


launch()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(stream):
    e0.record()
    for _ in range(reps):
        launch()
    e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / reps * 1e3
units = float(torch.minimum(ds, val + 1).sum().item())
print(f"k_propagate {layout} {scene} n={n} steps={'U' if not fixed else fixed}: {us:.1f} us  {units / us * 1e6:.3e} states/s "
      f"({980 * units / us * 1e6 / 1e12:.1f} TFLOP/s algorithmic)")
ctx.close()
