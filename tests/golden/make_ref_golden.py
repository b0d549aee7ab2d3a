#!/usr/bin/env python
"""Generates tests/golden/ref_partial.npz from the REFERENCE's own code: oracle/_ref/libkcbs_ref_partial.so is built by
oracle/Makefile from /root/reference/includes (StatePropagatorDatabase.h, StateSpaceDatabase.h, ControlSpaceDatabase.h,
GoalRegionDatabase.h, included where they lie) against this repository's OMPL API shim.  The fixture carries those
outputs to machines without /root/reference.  Run here: python tests/golden/make_ref_golden.py"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

orc = graft.load_oracle()
orc.build()
L = orc.ref_partial()
assert L is not None, "oracle/_ref/libkcbs_ref_partial.so is missing: needs /root/reference (make -C oracle)"
dp = C.POINTER(C.c_double)
p = lambda a: a.ctypes.data_as(dp)  # noqa: E731

g = np.random.Generator(np.random.Philox(key=[2026, 1]))
n = 1024
q = np.empty((n, 5))
q[:, 0] = g.uniform(0, 32, n)
q[:, 1] = g.uniform(0, 32, n)
q[:, 2] = g.uniform(-1.2, 1.2, n)
q[:, 3] = g.uniform(-1.2, 1.2, n)
q[:, 4] = g.uniform(-4.0, 4.0, n)
u = g.uniform(-1, 1, (n, 2))
# special points: rest, straight, hard lock, headings on the axes
q[:8] = [[1, 2, 0, 0, 0], [1, 2, 1, 0, 0], [1, 2, -1, 0, np.pi], [1, 2, 1, np.pi / 3, np.pi / 2], [1, 2, 1, -np.pi / 3, -np.pi / 2],
         [0, 0, 0.5, 0.25, 3.0], [32, 32, -0.5, -0.25, -3.0], [16, 16, 1, 1e-9, 1e-9]]
qdot = np.empty((n, 5))
for i in range(n):
    L.ref_second_order_car_ode(p(q[i]), p(u[i]), p(qdot[i]))
q2 = np.concatenate([q, q[::-1]], axis=1).copy()
u2 = np.concatenate([u, u[::-1]], axis=1).copy()
qdot2 = np.empty((n, 10))
for i in range(n):
    L.ref_two_second_order_cars_ode(p(q2[i]), p(u2[i]), p(qdot2[i]))
bounds = {}
for size in (10, 32):
    lo, hi = np.empty(4), np.empty(4)
    nsub = L.ref_state_bounds(size, size, p(lo), p(hi))
    bounds[size] = (lo, hi, nsub)
clo, chi = np.empty(2), np.empty(2)
L.ref_control_bounds(p(clo), p(chi))
pts = g.uniform(0, 32, (256, 4))
thr = C.c_double()
gd = np.array([L.ref_goal_distance(a, b, c, d, C.byref(thr)) for a, b, c, d in pts])
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_partial.npz"), q=q, u=u, qdot=qdot, q2=q2, u2=u2, qdot2=qdot2,
                    lo10=bounds[10][0], hi10=bounds[10][1], lo32=bounds[32][0], hi32=bounds[32][1],
                    n_subspaces=np.int64(bounds[10][2]), control_lo=clo, control_hi=chi, goal_pts=pts, goal_dist=gd,
                    goal_threshold=np.float64(thr.value))
print("wrote tests/golden/ref_partial.npz:", n, "right-hand sides, bounds", bounds[32][0], bounds[32][1], "controls", clo, chi,
      "goal threshold", thr.value)
