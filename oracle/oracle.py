"""ctypes wrapper of the CPU oracle (oracle/kcbs_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs, never by the product package.  PARITY UNPINNED (see kcbs_oracle.h)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
MAX_OBS, MAX_ROB = 64, 256
WRAP_HALF_OPEN, WRAP_LEGACY = 0, 1


class OrcWorld(C.Structure):
    _fields_ = [
        ("lo", C.c_double * 4), ("hi", C.c_double * 4),
        ("step_size", C.c_double), ("integration_step", C.c_double), ("car_length", C.c_double),
        ("wrap_mode", C.c_int32), ("n_obstacles", C.c_int32),
        ("obstacles", (C.c_double * 4) * MAX_OBS),
        ("n_robots", C.c_int32),
        ("robot_dims", (C.c_double * 2) * MAX_ROB),
    ]


class OrcConflict(C.Structure):
    _fields_ = [("r1", C.c_int32), ("r2", C.c_int32), ("k_start", C.c_int32), ("k_end", C.c_int32)]


CONFLICT_DTYPE = np.dtype([("r1", "<i4"), ("r2", "<i4"), ("k_start", "<i4"), ("k_end", "<i4")])


def build(force: bool = False) -> str:
    path = os.path.join(_HERE, "libkcbs_oracle.so")
    src = os.path.join(_HERE, "kcbs_oracle.c")
    if force or not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return path


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libkcbs_oracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        dp, vp, i32, i64, dbl = C.POINTER(C.c_double), C.c_void_p, C.c_int, C.c_int64, C.c_double
        W = C.POINTER(OrcWorld)
        L.orc_car_ode.argtypes = [dp, dp, dbl, dp]
        L.orc_rk4_step.argtypes = [dp, dp, dbl, dbl]
        L.orc_integrate_const.argtypes = [dp, dp, dbl, dbl, dbl]
        L.orc_wrap_angle.restype = dbl
        L.orc_wrap_angle.argtypes = [dbl, i32]
        L.orc_propagate.argtypes = [W, dp, dp, dbl, dp]
        L.orc_satisfies_bounds.argtypes = [W, dp, dp]
        L.orc_is_valid.argtypes = [W, i32, dp, dp]
        L.orc_pair_valid.argtypes = [W, i32, dp, i32, dp, dp]
        L.orc_propagate_while_valid.argtypes = [W, i32, dp, dp, i32, dp, dp]
        L.orc_robot_corners.argtypes = [dbl, dbl, dbl, dbl, dbl, dp]
        L.orc_polygons_disjoint.argtypes = [i32, dp, i32, dp]
        L.orc_sat_gap_robot_robot.restype = dbl
        L.orc_sat_gap_robot_robot.argtypes = [dbl, dbl, dp, dbl, dbl, dp]
        L.orc_sat_gap_robot_obstacle.restype = dbl
        L.orc_sat_gap_robot_obstacle.argtypes = [dbl, dbl, dp, dp]
        L.orc_propagate_batch.restype = i64
        L.orc_propagate_batch.argtypes = [W, i32, i64, i64, vp, vp, vp, vp, i32, vp, vp, vp, i32]
        L.orc_validity_batch.restype = i64
        L.orc_validity_batch.argtypes = [W, i32, i64, vp, vp, vp, i32]
        L.orc_first_conflict.restype = i64
        L.orc_first_conflict.argtypes = [W, i32, i32, vp, vp, C.POINTER(OrcConflict), dp]
        L.orc_first_conflict_batch.restype = i64
        L.orc_first_conflict_batch.argtypes = [W, i32, i32, i32, vp, vp, vp, i32, i32]
        L.orc_pair_is_valid.argtypes = [W, i32, i32, dp, dp]
        L.orc_pair_propagate.argtypes = [W, dp, dp, dbl, dp, dbl, dp]
        L.orc_pair_propagate_while_valid.argtypes = [W, i32, i32, dp, dp, i32, dp, dbl, dp, dp]
        L.orc_pair_propagate_batch.restype = i64
        L.orc_pair_propagate_batch.argtypes = [W, i32, i32, i64, i64, vp, vp, vp, vp, i32, dp, dbl, vp, vp, vp, vp, i32]
        L.orc_pair_validity_batch.restype = i64
        L.orc_pair_validity_batch.argtypes = [W, i32, i32, i64, vp, vp, vp, i32]
        L.orc_first_conflict_groups.restype = i64
        L.orc_first_conflict_groups.argtypes = [W, i32, i32, vp, vp, vp, C.POINTER(OrcConflict), dp]
        _LIB = L
    return _LIB


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


def _p(a):
    return None if a is None else a.ctypes.data


class Oracle:
    """Double-precision CPU restatement of the hot path for one scene."""

    def __init__(self, world, wrap_mode: int = WRAP_HALF_OPEN):
        """world: anything with x_max, y_max, robot_dims, obstacles, bounds(), step_size,
        integration_step, car_length (the product's World dataclass duck-types)."""
        self.L = lib()
        w = OrcWorld()
        lo, hi = world.bounds()
        w.lo[:] = lo
        w.hi[:] = hi
        w.step_size, w.integration_step, w.car_length = world.step_size, world.integration_step, world.car_length
        w.wrap_mode = wrap_mode
        obs = np.asarray(world.obstacles, dtype=np.float64).reshape(-1, 4)
        dims = np.asarray(world.robot_dims, dtype=np.float64).reshape(-1, 2)
        w.n_obstacles, w.n_robots = len(obs), len(dims)
        for i, o in enumerate(obs):
            w.obstacles[i][:] = list(o)
        for i, d in enumerate(dims):
            w.robot_dims[i][:] = list(d)
        self.w = w
        self.n_robots = len(dims)
        self.dims = dims

    # ---- scalar -------------------------------------------------------------------------
    def ode(self, q, u):
        qa, qp = _d(q)
        ua, up = _d(u)
        out = np.zeros(5)
        self.L.orc_car_ode(qp, up, self.w.car_length, out.ctypes.data_as(C.POINTER(C.c_double)))
        return out

    def rk4_step(self, q, u, h=None):
        qa, qp = _d(np.array(q, dtype=np.float64))
        ua, up = _d(u)
        self.L.orc_rk4_step(qp, up, self.w.integration_step if h is None else h, self.w.car_length)
        return qa

    def integrate_const(self, q, u, duration, h=None):
        qa, qp = _d(np.array(q, dtype=np.float64))
        ua, up = _d(u)
        n = self.L.orc_integrate_const(qp, up, duration, self.w.integration_step if h is None else h,
                                       self.w.car_length)
        return qa, n

    def wrap(self, theta, mode=None):
        return self.L.orc_wrap_angle(float(theta), self.w.wrap_mode if mode is None else mode)

    def propagate(self, q, u, duration=None):
        qa, qp = _d(q)
        ua, up = _d(u)
        out = np.zeros(5)
        self.L.orc_propagate(C.byref(self.w), qp, up, self.w.step_size if duration is None else duration,
                             out.ctypes.data_as(C.POINTER(C.c_double)))
        return out

    def satisfies_bounds(self, q):
        qa, qp = _d(q)
        s = C.c_double()
        ok = self.L.orc_satisfies_bounds(C.byref(self.w), qp, C.byref(s))
        return bool(ok), s.value

    def is_valid(self, q, robot=0):
        qa, qp = _d(q)
        m = C.c_double()
        ok = self.L.orc_is_valid(C.byref(self.w), robot, qp, C.byref(m))
        return bool(ok), m.value

    def pair_valid(self, r1, s1, r2, s2):
        a, ap = _d(s1)
        b, bp = _d(s2)
        m = C.c_double()
        ok = self.L.orc_pair_valid(C.byref(self.w), r1, ap, r2, bp, C.byref(m))
        return bool(ok), m.value

    def propagate_while_valid(self, q0, u, steps, robot=0):
        qa, qp = _d(q0)
        ua, up = _d(u)
        seg = np.zeros((max(steps, 1), 5))
        fin = np.zeros(5)
        n = self.L.orc_propagate_while_valid(C.byref(self.w), robot, qp, up, steps,
                                             seg.ctypes.data_as(C.POINTER(C.c_double)),
                                             fin.ctypes.data_as(C.POINTER(C.c_double)))
        return n, seg[:n], fin

    def robot_corners(self, length, width, x, y, theta):
        out = np.zeros((4, 2))
        self.L.orc_robot_corners(length, width, x, y, theta, out.ctypes.data_as(C.POINTER(C.c_double)))
        return out

    def polygons_disjoint(self, p1, p2):
        a, ap = _d(p1)
        b, bp = _d(p2)
        return bool(self.L.orc_polygons_disjoint(len(a), ap, len(b), bp))

    def sat_gap_robot_robot(self, la, wa, sa, lb, wb, sb):
        a, ap = _d(sa)
        b, bp = _d(sb)
        return self.L.orc_sat_gap_robot_robot(la, wa, ap, lb, wb, bp)

    def sat_gap_robot_obstacle(self, l, w, s, obs):
        a, ap = _d(s)
        b, bp = _d(obs)
        return self.L.orc_sat_gap_robot_obstacle(l, w, ap, bp)

    # ---- batched (float32 in, double out) ----------------------------------------------------
    def propagate_batch(self, robot, parents, controls, steps=None, parent_idx=None, max_steps=10, want_seg=True,
                        n_threads=1):
        parents = np.ascontiguousarray(parents, dtype=np.float32)
        controls = np.ascontiguousarray(controls, dtype=np.float32)
        n, n_parents = controls.shape[1], parents.shape[1]
        steps = None if steps is None else np.ascontiguousarray(steps, dtype=np.int32)
        parent_idx = None if parent_idx is None else np.ascontiguousarray(parent_idx, dtype=np.int32)
        seg = np.zeros((5, max_steps, n)) if want_seg else None
        fin = np.zeros((5, n))
        valid = np.zeros(n, np.int32)
        calls = self.L.orc_propagate_batch(C.byref(self.w), robot, n, n_parents, _p(parents), _p(parent_idx),
                                           _p(controls), _p(steps), max_steps, _p(seg), _p(fin), _p(valid), n_threads)
        return seg, fin, valid, int(calls)

    def validity_batch(self, robot, states, want_margin=True, n_threads=1):
        states = np.ascontiguousarray(states, dtype=np.float32)
        n = states.shape[1]
        valid = np.zeros(n, np.uint8)
        margin = np.zeros(n) if want_margin else None
        self.L.orc_validity_batch(C.byref(self.w), robot, n, _p(states), _p(valid), _p(margin), n_threads)
        return valid.astype(bool), margin

    def first_conflict(self, traj, lengths=None, want_margin=True):
        traj = np.ascontiguousarray(traj, dtype=np.float32)
        R, _, T = traj.shape
        lengths = None if lengths is None else np.ascontiguousarray(lengths, dtype=np.int32)
        out = OrcConflict()
        mm = C.c_double()
        evals = self.L.orc_first_conflict(C.byref(self.w), R, T, _p(traj), _p(lengths), C.byref(out),
                                          C.byref(mm) if want_margin else None)
        return (out.r1, out.r2, out.k_start, out.k_end), mm.value, int(evals)

    def first_conflict_batch(self, traj, lengths=None, full_scan=False, n_threads=1):
        traj = np.ascontiguousarray(traj, dtype=np.float32)
        P, R, _, T = traj.shape
        lengths = None if lengths is None else np.ascontiguousarray(lengths, dtype=np.int32)
        out = np.zeros(P, CONFLICT_DTYPE)
        evals = self.L.orc_first_conflict_batch(C.byref(self.w), P, R, T, _p(traj), _p(lengths), _p(out),
                                                int(full_scan), n_threads)
        return out, int(evals)

    # ---- merged pair -------------------------------------------------------------------------------
    def pair_is_valid(self, q, r1=0, r2=1):
        qa, qp = _d(q)
        m = C.c_double()
        ok = self.L.orc_pair_is_valid(C.byref(self.w), r1, r2, qp, C.byref(m))
        return bool(ok), m.value

    def pair_propagate(self, q, u, goals, duration=None, hold_radius=1.5):
        qa, qp = _d(q)
        ua, up = _d(u)
        ga, gp = _d(goals)
        out = np.zeros(10)
        self.L.orc_pair_propagate(C.byref(self.w), qp, up, self.w.step_size if duration is None else duration, gp, hold_radius,
                                  out.ctypes.data_as(C.POINTER(C.c_double)))
        return out

    def pair_propagate_batch(self, r1, r2, parents, controls, goals, steps=None, parent_idx=None, max_steps=10, hold_radius=1.5,
                             n_threads=1):
        parents = np.ascontiguousarray(parents, dtype=np.float32)
        controls = np.ascontiguousarray(controls, dtype=np.float32)
        ga, gp = _d(goals)
        n, n_parents = controls.shape[1], parents.shape[1]
        steps = None if steps is None else np.ascontiguousarray(steps, dtype=np.int32)
        parent_idx = None if parent_idx is None else np.ascontiguousarray(parent_idx, dtype=np.int32)
        seg = np.zeros((10, max_steps, n))
        fin = np.zeros((10, n))
        valid = np.zeros(n, np.int32)
        hold = np.zeros(n)
        calls = self.L.orc_pair_propagate_batch(C.byref(self.w), r1, r2, n, n_parents, _p(parents), _p(parent_idx), _p(controls),
                                                _p(steps), max_steps, gp, hold_radius, _p(seg), _p(fin), _p(valid), _p(hold),
                                                n_threads)
        return seg, fin, valid, hold, int(calls)

    def pair_validity_batch(self, r1, r2, states, n_threads=1):
        states = np.ascontiguousarray(states, dtype=np.float32)
        n = states.shape[1]
        valid = np.zeros(n, np.uint8)
        margin = np.zeros(n)
        self.L.orc_pair_validity_batch(C.byref(self.w), r1, r2, n, _p(states), _p(valid), _p(margin), n_threads)
        return valid.astype(bool), margin

    def first_conflict_groups(self, traj, group, lengths=None):
        traj = np.ascontiguousarray(traj, dtype=np.float32)
        R, _, T = traj.shape
        group = np.ascontiguousarray(group, dtype=np.int32)
        lengths = None if lengths is None else np.ascontiguousarray(lengths, dtype=np.int32)
        out = OrcConflict()
        mm = C.c_double()
        self.L.orc_first_conflict_groups(C.byref(self.w), R, T, _p(traj), _p(lengths), _p(group), C.byref(out), C.byref(mm))
        return (out.r1, out.r2, out.k_start, out.k_end), mm.value


# ---- the part of the reference that compiles here (oracle/ref_partial.cpp -> oracle/_ref/) ---------------------------
_REF = None


def ref_partial():
    """ctypes handle of oracle/_ref/libkcbs_ref_partial.so (the reference's own ODE right-hand sides, bounds and goal
    test, compiled from /root/reference/includes against the OMPL API shim), or None when it has not been built
    (the GPU box has no /root/reference; the committed tests/golden/ref_partial.npz carries its outputs there)."""
    global _REF
    if _REF is None:
        path = os.path.join(_HERE, "_ref", "libkcbs_ref_partial.so")
        if not os.path.exists(path):
            return None
        L = C.CDLL(path)
        dp = C.POINTER(C.c_double)
        L.ref_second_order_car_ode.argtypes = [dp, dp, dp]
        L.ref_two_second_order_cars_ode.argtypes = [dp, dp, dp]
        L.ref_state_bounds.argtypes = [C.c_uint, C.c_uint, dp, dp]
        L.ref_control_bounds.argtypes = [dp, dp]
        L.ref_goal_distance.restype = C.c_double
        L.ref_goal_distance.argtypes = [C.c_double, C.c_double, C.c_double, C.c_double, dp]
        ip = C.POINTER(C.c_int)
        L.ref_is_valid.argtypes = [C.c_uint, C.c_uint, C.c_int, dp, C.c_double, C.c_double, C.c_int, dp, ip]
        L.ref_are_states_valid.argtypes = [C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, dp, dp, ip]
        L.ref_pair_is_valid.argtypes = [C.c_uint, C.c_uint, C.c_int, dp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int,
                                        dp, ip, C.c_double, C.c_double, dp, ip]
        L.ref_pair_propagate.argtypes = [dp, dp, dp, C.c_int, dp]
        L.ref_car_propagate.argtypes = [dp, dp, C.c_int, dp]
        L.ref_two_car_goal.argtypes = [dp, dp, dp]
        L.ref_shape_constants.argtypes = [C.c_double] * 6 + [dp]
        L.ref_merge.argtypes = [C.c_int, dp, dp, C.c_uint, C.c_int, C.c_int, C.c_char_p, C.c_int, ip, dp, dp, dp, dp, dp, ip, dp, ip]
        _REF = L
    return _REF
