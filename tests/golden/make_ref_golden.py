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
# ---- the reference's validity checkers, pair propagator, two-car goal and merger (compiled on the OMPL shim + the
# Boost.Geometry stand-in, oracle/ref_partial.cpp) ------------------------------------------------------------------
pkg = graft.load_package()
ip = C.POINTER(C.c_int)
pi_ = lambda a: a.ctypes.data_as(ip)  # noqa: E731
extra = {}
for key, scene in (("congested", "Congested10x10_7robots"), ("corridor", "Corridor10x10_4robots")):
    w = pkg.scenes.world(scene)
    obs = np.ascontiguousarray(np.asarray(w.obstacles, np.float64).reshape(-1, 4))
    L_, W_ = w.robot_dims[0]
    m = 6000
    st = np.empty((m, 5))
    st[:, 0] = g.uniform(-0.3, w.x_max + 0.3, m)
    st[:, 1] = g.uniform(-0.3, w.y_max + 0.3, m)
    st[:, 2] = g.uniform(-1.05, 1.05, m)
    st[:, 3] = g.uniform(-1.1, 1.1, m)
    st[:, 4] = g.uniform(-3.3, 3.3, m)
    st[:8, 2:] = 0.0          # exact bound values of x / y
    st[:4, 0] = [0.0, w.x_max, 0.5, 0.5]
    st[:4, 1] = [0.5, 0.5, 0.0, w.y_max]
    out = np.zeros(m, np.int32)
    L.ref_is_valid(int(w.x_max), int(w.y_max), len(obs), p(obs), L_, W_, m, p(st), pi_(out))
    extra[f"valid_{key}_states"], extra[f"valid_{key}"] = st, out
# robot-robot predicate: pairs at distances around the bounding circles, equal and mixed footprints
m = 6000
for key, (d1, d2) in (("same", ((0.7, 0.5), (0.7, 0.5))), ("mixed", ((1.0, 1.0), (0.7, 0.5)))):
    a = np.empty((m, 5))
    b = np.empty((m, 5))
    a[:, :2] = g.uniform(4, 28, (m, 2))
    ang, dist = g.uniform(-np.pi, np.pi, m), g.uniform(0.0, 1.6, m)
    b[:, 0], b[:, 1] = a[:, 0] + dist * np.cos(ang), a[:, 1] + dist * np.sin(ang)
    a[:, 2:4] = b[:, 2:4] = 0.0
    a[:, 4], b[:, 4] = g.uniform(-np.pi, np.pi, m), g.uniform(-np.pi, np.pi, m)
    out = np.zeros(m, np.int32)
    L.ref_are_states_valid(d1[0], d1[1], d2[0], d2[1], m, p(a), p(b), pi_(out))
    extra[f"pairs_{key}_a"], extra[f"pairs_{key}_b"], extra[f"pairs_{key}"] = a, b, out
# merged pair: isValid of 10-d states in the congested scene, areStatesValid against a third robot
w = pkg.scenes.world("Congested10x10_7robots")
obs = np.ascontiguousarray(np.asarray(w.obstacles, np.float64).reshape(-1, 4))
m = 6000
q10 = np.empty((m, 10))
for c in (0, 5):
    q10[:, c] = g.uniform(-0.2, 10.2, m)
    q10[:, c + 1] = g.uniform(-0.2, 10.2, m)
    q10[:, c + 2] = g.uniform(-1.02, 1.02, m)
    q10[:, c + 3] = g.uniform(-1.07, 1.07, m)
    q10[:, c + 4] = g.uniform(-3.2, 3.2, m)
near = g.random(m) < 0.5   # half of the pairs close to each other
q10[near, 5] = q10[near, 0] + g.uniform(-1.5, 1.5, near.sum())
q10[near, 6] = q10[near, 1] + g.uniform(-1.5, 1.5, near.sum())
q3 = np.empty((m, 5))
q3[:, 0] = q10[:, 5] + g.uniform(-1.6, 1.6, m)
q3[:, 1] = q10[:, 6] + g.uniform(-1.6, 1.6, m)
q3[:, 2:4] = 0
q3[:, 4] = g.uniform(-np.pi, np.pi, m)
o1, o2 = np.zeros(m, np.int32), np.zeros(m, np.int32)
L.ref_pair_is_valid(10, 10, len(obs), p(obs), 1.0, 1.0, 1.0, 1.0, m, p(q10), pi_(o1), 1.0, 1.0, p(q3), pi_(o2))
extra.update(pair_states=q10, pair_valid=o1, pair_third=q3, pair_third_valid=o2)
# propagators through the RK4 stand-in: single car (10 steps), merged pair with the freeze rule (12 steps)
m = 256
cq = np.empty((m, 5))
cq[:, :2] = g.uniform(2, 30, (m, 2))
cq[:, 2], cq[:, 3], cq[:, 4] = g.uniform(-1, 1, m), g.uniform(-1, 1, m), g.uniform(-3.1, 3.1, m)
cu = g.uniform(-1, 1, (m, 2))
cout = np.zeros((m, 10, 5))
for i in range(m):
    L.ref_car_propagate(p(cq[i]), p(cu[i]), 10, p(cout[i]))
pq = np.concatenate([cq, cq[::-1]], axis=1).copy()
pu = np.concatenate([cu, cu[::-1]], axis=1).copy()
pg = np.empty((m, 4))
pg[:, :2] = pq[:, :2] + g.uniform(-2.0, 2.0, (m, 2))      # goals near car 1: some start inside the 1.5 hold radius, some enter it
pg[:, 2:] = pq[:, 5:7] + g.uniform(-2.0, 2.0, (m, 2))
pout = np.zeros((m, 12, 10))
for i in range(m):
    L.ref_pair_propagate(p(pq[i]), p(pu[i]), p(pg[i]), 12, p(pout[i]))
extra.update(car_q=cq, car_u=cu, car_out=cout, pairp_q=pq, pairp_u=pu, pairp_goals=pg, pairp_out=pout)
# two-car goal
tg = np.zeros(m, np.int32)
td = np.zeros(m)
for i in range(m):
    d = C.c_double()
    tg[i] = L.ref_two_car_goal(p(pq[i]), p(pg[i]), C.byref(d))
    td[i] = d.value
extra.update(goal2_flags=tg, goal2_dist=td)
# shape constants
sc = np.zeros(4)
L.ref_shape_constants(0.5, 1.8, 9.0, 1.0, 0.7, 0.5, p(sc))
extra["shape_constants"] = sc
# SystemMerger::merge wiring on the corridor problem (4 robots, bound 10, merge(0, 1))
starts = np.array([[1, 0.5], [1, 3.5], [9, 0.5], [9, 3.5]], np.float64)
goals = np.array([[9, 0.5], [9, 9], [1, 0.5], [1, 9]], np.float64)
names = C.create_string_buffer(512)
dims = np.zeros(8, np.int32)
mlo, mhi, mclo, mchi, mstart = np.zeros(8), np.zeros(8), np.zeros(4), np.zeros(4), np.zeros(10)
msteps = np.zeros(2, np.int32)
mstep = C.c_double()
mre = C.c_int()
n_ind = L.ref_merge(4, p(starts), p(goals), 10, 0, 1, names, 512, pi_(dims), p(mlo), p(mhi), p(mclo), p(mchi), p(mstart), pi_(msteps),
                    C.byref(mstep), C.byref(mre))
extra.update(merge_n=np.int64(n_ind), merge_names=np.array(names.value.decode()), merge_dims=dims[:n_ind], merge_lo=mlo, merge_hi=mhi,
             merge_clo=mclo, merge_chi=mchi, merge_start=mstart, merge_steps=msteps, merge_step_size=np.float64(mstep.value),
             merge_remerge=np.int64(mre.value))
print("merge:", n_ind, names.value.decode(), dims[:n_ind], mstart, msteps, mstep.value, "remerge", mre.value)
print("valid fractions:", {k: float(v.mean()) for k, v in extra.items() if v.dtype == np.int32 and v.ndim == 1 and v.size > 100})

np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_partial.npz"), **extra, q=q, u=u, qdot=qdot, q2=q2, u2=u2, qdot2=qdot2,
                    lo10=bounds[10][0], hi10=bounds[10][1], lo32=bounds[32][0], hi32=bounds[32][1],
                    n_subspaces=np.int64(bounds[10][2]), control_lo=clo, control_hi=chi, goal_pts=pts, goal_dist=gd,
                    goal_threshold=np.float64(thr.value))
print("wrote tests/golden/ref_partial.npz:", n, "right-hand sides, bounds", bounds[32][0], bounds[32][1], "controls", clo, chi,
      "goal threshold", thr.value)
