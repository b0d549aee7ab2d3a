"""Pins the oracle (and the scene constants the product is created with) to the REFERENCE's own code where that code
compiles here: the ODE right-hand sides of StatePropagatorDatabase.h:44-64 / :77-102, the state bounds of
StateSpaceDatabase.h:42-62, the control bounds of ControlSpaceDatabase.h:43-54 and the goal test of
GoalRegionDatabase.h:42-60.  tests/golden/ref_partial.npz holds their outputs (tests/golden/make_ref_golden.py); when
oracle/_ref/libkcbs_ref_partial.so is present (it is wherever /root/reference exists) the same functions are also called
live on fresh inputs.  The RK4 driver, the angle wrap and the Boost.Geometry predicates stay unpinned."""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "ref_partial.npz"))


def test_oracle_rhs_is_the_references_rhs(pkg, orc):
    oracle = orc.Oracle(pkg.scenes.world("Empty32x32_10robots"))
    got = np.array([oracle.ode(q, u) for q, u in zip(GOLD["q"], GOLD["u"])])
    assert np.array_equal(got, GOLD["qdot"]), f"max |diff| {np.abs(got - GOLD['qdot']).max():.3e}"   # bit for bit
    # the two-car ODE (StatePropagatorDatabase.h:77-102) is the one-car ODE applied to each car
    two = np.concatenate([np.array([oracle.ode(q[:5], u[:2]) for q, u in zip(GOLD["q2"], GOLD["u2"])]),
                          np.array([oracle.ode(q[5:], u[2:]) for q, u in zip(GOLD["q2"], GOLD["u2"])])], axis=1)
    assert np.array_equal(two, GOLD["qdot2"])


def test_scene_constants_are_the_references(pkg):
    assert int(GOLD["n_subspaces"]) == 2                      # RealVector(4) + SO2: reals order (x, y, v, phi, theta)
    for scene, lo, hi in (("Empty32x32_10robots", GOLD["lo32"], GOLD["hi32"]), ("Corridor10x10_4robots", GOLD["lo10"], GOLD["hi10"]),
                          ("Congested10x10_7robots", GOLD["lo10"], GOLD["hi10"])):
        wlo, whi = pkg.scenes.world(scene).bounds()
        assert np.array_equal(np.asarray(wlo, np.float64), lo) and np.array_equal(np.asarray(whi, np.float64), hi)
    assert GOLD["control_lo"].tolist() == [-1.0, -1.0] and GOLD["control_hi"].tolist() == [1.0, 1.0]
    # the synthetic control batches are drawn from exactly that box
    from kcbs_b200 import workloads
    _, controls, _ = workloads.expansion_batch("Empty32x32_10robots", 4096, seed=1)
    assert controls.min() >= -1.0 and controls.max() <= 1.0


def test_goal_test_is_the_references(pkg):
    pts, d = GOLD["goal_pts"], GOLD["goal_dist"]
    mine = np.sqrt((pts[:, 0] - pts[:, 2]) ** 2 + (pts[:, 1] - pts[:, 3]) ** 2)
    assert np.allclose(mine, d, rtol=1e-15, atol=0)
    assert float(GOLD["goal_threshold"]) == 0.5
    # the planner's default goal radius is that threshold (kcbs_rrt_defaults)
    from kcbs_b200 import capi
    lib = pkg.load_library()
    p = capi.RrtParams()
    lib.kcbs_rrt_defaults(C.byref(p))
    assert p.goal_radius == 0.5


def test_live_reference_code_when_built(pkg, orc):
    """Fresh random inputs through the compiled reference headers (skipped where /root/reference is absent)."""
    L = orc.ref_partial()
    if L is None:
        pytest.skip("oracle/_ref/libkcbs_ref_partial.so not built (no /root/reference on this machine)")
    dp = C.POINTER(C.c_double)
    oracle = orc.Oracle(pkg.scenes.world("Empty32x32_10robots"))
    g = np.random.default_rng(77)
    for _ in range(2000):
        q = np.concatenate([g.uniform(0, 32, 2), g.uniform(-1.5, 1.5, 2), g.uniform(-7, 7, 1)])
        u = g.uniform(-1, 1, 2)
        out = np.zeros(5)
        L.ref_second_order_car_ode(q.ctypes.data_as(dp), u.ctypes.data_as(dp), out.ctypes.data_as(dp))
        assert np.array_equal(out, oracle.ode(q, u))
    lo, hi = np.zeros(4), np.zeros(4)
    assert L.ref_state_bounds(32, 32, lo.ctypes.data_as(dp), hi.ctypes.data_as(dp)) == 2
    assert np.array_equal(lo, GOLD["lo32"]) and np.array_equal(hi, GOLD["hi32"])    # the fixture is what the code says today
