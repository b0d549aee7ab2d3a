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


# ---- validity checkers, pair propagator, two-car goal, merger: the reference's own control flow ------------------
# (compiled on the OMPL shim + the Boost.Geometry stand-in; outputs in tests/golden/ref_partial.npz)

def _agree_outside_band(ref_bool, orc_bool, margin, band=1e-9):
    """The oracle's margin is the signed distance to its decision boundary: outside a tiny band (the two sides evaluate
    touching in different arithmetic) the booleans must be identical."""
    clear = np.abs(margin) > band
    assert clear.mean() > 0.999
    bad = np.nonzero(clear & (ref_bool != orc_bool))[0]
    assert bad.size == 0, f"{bad.size} disagreements, first at {bad[:5]} (margins {margin[bad[:5]]})"


@pytest.mark.parametrize("key,scene", [("congested", "Congested10x10_7robots"), ("corridor", "Corridor10x10_4robots")])
def test_oracle_is_valid_is_the_references(pkg, orc, key, scene):
    # homogeneous2ndOrderCarSystemSVC::isValid, StateValidityCheckerDatabase.h:54-75 (bounds, quick check, -theta footprint)
    oracle = orc.Oracle(pkg.scenes.world(scene))
    st, ref = GOLD[f"valid_{key}_states"], GOLD[f"valid_{key}"].astype(bool)
    got = np.array([oracle.is_valid(q, 0) for q in st])
    _agree_outside_band(ref, got[:, 0].astype(bool), got[:, 1])
    assert 0.2 < ref.mean() < 0.8


@pytest.mark.parametrize("key,dims", [("same", ((0.7, 0.5), (0.7, 0.5))), ("mixed", ((1.0, 1.0), (0.7, 0.5)))])
def test_oracle_pair_predicate_is_the_references(pkg, orc, key, dims):
    # areStatesValid, non-composed branch, StateValidityCheckerDatabase.h:108-124 + the exact check :144-177
    from kcbs_b200.capi import World
    oracle = orc.Oracle(World(32, 32, robot_dims=[dims[0], dims[1]]))
    a, b, ref = GOLD[f"pairs_{key}_a"], GOLD[f"pairs_{key}_b"], GOLD[f"pairs_{key}"].astype(bool)
    got = np.array([oracle.pair_valid(0, qa[[0, 1, 4]], 1, qb[[0, 1, 4]]) for qa, qb in zip(a, b)])
    _agree_outside_band(ref, got[:, 0].astype(bool), got[:, 1])
    assert 0.2 < ref.mean() < 0.8


def test_oracle_merged_pair_validity_is_the_references(pkg, orc):
    # homogeneousTwo2ndOrderCarSystemSVC::isValid :225-261 and ::areStatesValid :264-287
    oracle = orc.Oracle(pkg.scenes.world("Congested10x10_7robots"))
    q10, ref = GOLD["pair_states"], GOLD["pair_valid"].astype(bool)
    got = np.array([oracle.pair_is_valid(q, 0, 1) for q in q10])
    _agree_outside_band(ref, got[:, 0].astype(bool), got[:, 1])
    # merged pair against a third robot = both of its cars against it (the oracle's single-pair predicate twice)
    q3, ref3 = GOLD["pair_third"], GOLD["pair_third_valid"].astype(bool)
    r = np.array([[*oracle.pair_valid(0, q[[0, 1, 4]], 2, t[[0, 1, 4]]), *oracle.pair_valid(1, q[[5, 6, 9]], 2, t[[0, 1, 4]])]
                  for q, t in zip(q10, q3)])
    both = r[:, 0].astype(bool) & r[:, 2].astype(bool)
    margin = np.minimum(np.abs(r[:, 1]), np.abs(r[:, 3]))
    _agree_outside_band(ref3, both, margin)


def test_oracle_propagators_match_the_reference_wiring(pkg, orc):
    # the reference's ODE functors + post-integration wrap + freeze-at-goal rule (StatePropagatorDatabase.h:67-74, :106-117,
    # :130-153) driven by the RK4 stand-in, against the oracle's restatement: same states to rounding
    oracle = orc.Oracle(pkg.scenes.world("Empty32x32_10robots"))
    for q, u, out in zip(GOLD["car_q"], GOLD["car_u"], GOLD["car_out"]):
        cur = q.copy()
        for j in range(out.shape[0]):
            cur = oracle.propagate(cur, u)
            if abs(cur[3]) > 1.2:
                break      # far outside the steering bound tan(phi) nears its pole and amplifies the last bit
            d = np.abs(cur - out[j])
            d[4] = min(d[4], abs(2 * np.pi - d[4]))
            assert d.max() < 1e-12, (j, d)
    frozen_seen = moved_seen = 0
    for q, u, goals, out in zip(GOLD["pairp_q"], GOLD["pairp_u"], GOLD["pairp_goals"], GOLD["pairp_out"]):
        cur = q.copy()
        for j in range(out.shape[0]):
            nxt = oracle.pair_propagate(cur, u, goals)
            if max(abs(nxt[3]), abs(nxt[8])) > 1.2:
                break
            d = np.abs(nxt - out[j])
            d[[4, 9]] = np.minimum(d[[4, 9]], np.abs(2 * np.pi - d[[4, 9]]))
            assert d.max() < 1e-12, (j, d)
            for c in (0, 5):
                if np.array_equal(nxt[c:c + 5], cur[c:c + 5]):
                    frozen_seen += 1
                else:
                    moved_seen += 1
            cur = nxt
    assert frozen_seen > 100 and moved_seen > 100      # the fixture exercises both sides of the freeze rule


def test_two_car_goal_is_the_references(pkg):
    # GoalRegionTwo2ndOrderCars, GoalRegionDatabase.h:61-108: both cars strictly within 1.5 of their goals; distance = d1 + d2
    q, g = GOLD["pairp_q"], GOLD["pairp_goals"]
    d1 = np.sqrt((q[:, 0] - g[:, 0]) ** 2 + (q[:, 1] - g[:, 1]) ** 2)
    d2 = np.sqrt((q[:, 5] - g[:, 2]) ** 2 + (q[:, 6] - g[:, 3]) ** 2)
    sat = (d1 < 1.5) & (d2 < 1.5)
    assert np.array_equal(GOLD["goal2_flags"], np.where(sat, 3, 0))      # both overloads agree
    assert np.allclose(GOLD["goal2_dist"], d1 + d2, rtol=1e-15, atol=0)
    assert 0 < sat.mean() < 1


def test_shape_constants_are_the_references(pkg, orc):
    # Obstacle.h:82-100 (centre = corner + size / 2, bounding radius = half diagonal), Robot.h:99
    cx, cy, orad, rrad = GOLD["shape_constants"]
    assert (cx, cy) == (0.5 + 9.0 / 2, 1.8 + 1.0 / 2)
    assert orad == np.sqrt((9.0 / 2) ** 2 + (1.0 / 2) ** 2) and rrad == np.sqrt(0.35 ** 2 + 0.25 ** 2)


def test_merge_wiring_fixture():
    # homogeneous2ndOrderCarSystemMerger::merge(0, 1) on the corridor problem, SystemMergerDatabase.h:55-168: the others first,
    # the meta-robot last, 10-d, both cars' bounds, zero velocities in the start state, durations 1..10 of 0.1 s, no re-merge
    assert int(GOLD["merge_n"]) == 3
    assert str(GOLD["merge_names"]) == "Robot 3|Robot 4|Robot 1 and Robot 2"
    assert GOLD["merge_dims"].tolist() == [5, 5, 10]
    assert np.array_equal(GOLD["merge_lo"], np.tile(GOLD["lo10"], 2)) and np.array_equal(GOLD["merge_hi"], np.tile(GOLD["hi10"], 2))
    assert GOLD["merge_clo"].tolist() == [-1.0] * 4 and GOLD["merge_chi"].tolist() == [1.0] * 4
    assert GOLD["merge_start"].tolist() == [1.0, 0.5, 0, 0, 0, 1.0, 3.5, 0, 0, 0]
    assert GOLD["merge_steps"].tolist() == [1, 10] and float(GOLD["merge_step_size"]) == 0.1
    assert int(GOLD["merge_remerge"]) in (1, 2)      # null pair, or robots_.at throwing on the compound name


def test_live_reference_checkers_when_built(pkg, orc):
    """The compiled reference headers on fresh inputs (skipped where /root/reference is absent): the fixture is what the
    code says today, and the oracle agrees on inputs the fixture has never seen."""
    L = orc.ref_partial()
    if L is None:
        pytest.skip("oracle/_ref/libkcbs_ref_partial.so not built (no /root/reference on this machine)")
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    w = pkg.scenes.world("Congested10x10_7robots")
    obs = np.ascontiguousarray(np.asarray(w.obstacles, np.float64).reshape(-1, 4))
    st = GOLD["valid_congested_states"]
    out = np.zeros(len(st), np.int32)
    L.ref_is_valid(10, 10, len(obs), obs.ctypes.data_as(dp), 1.0, 1.0, len(st), st.ctypes.data_as(dp), out.ctypes.data_as(ip))
    assert np.array_equal(out, GOLD["valid_congested"])
    g = np.random.default_rng(2027)
    m = 3000
    fresh = np.column_stack([g.uniform(-0.2, 10.2, m), g.uniform(-0.2, 10.2, m), g.uniform(-1, 1, m), g.uniform(-1, 1, m), g.uniform(-3.1, 3.1, m)])
    fresh = np.ascontiguousarray(fresh)
    out = np.zeros(m, np.int32)
    L.ref_is_valid(10, 10, len(obs), obs.ctypes.data_as(dp), 1.0, 1.0, m, fresh.ctypes.data_as(dp), out.ctypes.data_as(ip))
    oracle = orc.Oracle(w)
    got = np.array([oracle.is_valid(q, 0) for q in fresh])
    _agree_outside_band(out.astype(bool), got[:, 0].astype(bool), got[:, 1])
