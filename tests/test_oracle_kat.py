"""Known-answer tests that pin the CPU oracle to analytic results (the reference
ships no golden vectors, so these are constructed).  Each test names the reference lines whose
semantics it checks."""
import math

import numpy as np
import pytest


@pytest.fixture(scope="module")
def empty(pkg, orc):
    return orc.Oracle(pkg.scenes.world("Empty32x32_10robots"))


@pytest.fixture(scope="module")
def corridor(pkg, orc):
    return orc.Oracle(pkg.scenes.world("Corridor10x10_4robots"))


def test_ode_rhs(empty):
    # StatePropagatorDatabase.h:59-63
    q = [1.0, 2.0, 0.7, 0.3, 1.1]
    d = empty.ode(q, [0.4, -0.2])
    assert d[0] == pytest.approx(0.7 * math.cos(1.1), abs=1e-15)
    assert d[1] == pytest.approx(0.7 * math.sin(1.1), abs=1e-15)
    assert d[2] == 0.4 and d[3] == -0.2
    assert d[4] == pytest.approx(0.7 / 0.5 * math.tan(0.3), abs=1e-15)


def test_integrate_const_step_counts(empty):
    # odeint::integrate_const with dt = 1e-2: 0.1 -> 10 ... 1.0 -> 100 sub-steps
    for j in range(1, 11):
        _, n = empty.integrate_const([0, 0, 0, 0, 0], [0, 0], 0.1 * j)
        assert n == 10 * j


def test_straight_line_exact(empty):
    # phi = 0, phidot = 0: theta constant, v linear, x quadratic in t -> RK4 is exact
    x0, y0, v0, th, a = 3.0, 4.0, 0.2, 0.6, 0.5
    q = empty.propagate([x0, y0, v0, 0.0, th], [a, 0.0], duration=1.0)
    t = 1.0
    s = v0 * t + 0.5 * a * t * t
    assert q[0] == pytest.approx(x0 + math.cos(th) * s, abs=1e-13)
    assert q[1] == pytest.approx(y0 + math.sin(th) * s, abs=1e-13)
    assert q[2] == pytest.approx(v0 + a * t, abs=1e-14)
    assert q[3] == 0.0 and q[4] == pytest.approx(th, abs=1e-15)


def test_constant_curvature(empty):
    # a = 0, phidot = 0: theta(t) = th0 + 2 v tan(phi) t, circular arc
    x0, y0, v, phi, th0, t = 5.0, 6.0, 0.8, 0.4, -0.3, 1.0
    q = empty.propagate([x0, y0, v, phi, th0], [0.0, 0.0], duration=t)
    k = 2.0 * math.tan(phi)
    th1 = th0 + v * k * t
    assert q[4] == pytest.approx(th1, abs=1e-12)
    assert q[0] == pytest.approx(x0 + (math.sin(th1) - math.sin(th0)) / k, abs=1e-9)
    assert q[1] == pytest.approx(y0 - (math.cos(th1) - math.cos(th0)) / k, abs=1e-9)


def test_v_phi_linear(empty):
    q = empty.propagate([1, 1, -0.3, 0.2, 0.1], [0.7, -0.9], duration=0.5)
    assert q[2] == pytest.approx(-0.3 + 0.7 * 0.5, abs=1e-14)
    assert q[3] == pytest.approx(0.2 - 0.9 * 0.5, abs=1e-14)


def test_against_high_order_integrator(empty):
    # independent check of the RK4 driver: scipy DOP853 on the same ODE, rtol 1e-12
    from scipy.integrate import solve_ivp

    rng = np.random.default_rng(0)
    for _ in range(20):
        q0 = [rng.uniform(0, 32), rng.uniform(0, 32), rng.uniform(-1, 1), rng.uniform(-0.6, 0.6), rng.uniform(-3, 3)]
        u = [rng.uniform(-1, 1), rng.uniform(-0.4, 0.4)]   # keeps |phi| <= 1.0, inside the bounds

        def f(t, q):
            return [q[2] * math.cos(q[4]), q[2] * math.sin(q[4]), u[0], u[1], q[2] / 0.5 * math.tan(q[3])]

        ref = solve_ivp(f, (0, 1.0), q0, method="DOP853", rtol=1e-12, atol=1e-14).y[:, -1]
        got, n = empty.integrate_const(q0, u, 1.0)
        assert n == 100
        assert np.allclose(got, ref, atol=1e-8, rtol=0)


def test_wrap_half_open(empty, orc):
    # SO2StateSpace::enforceBounds -> [-pi, pi); the comment at StatePropagatorDatabase.h:66 says
    # [0, 2pi] but the code is the contract
    assert empty.wrap(math.pi) == pytest.approx(-math.pi)
    assert empty.wrap(-math.pi) == -math.pi
    assert empty.wrap(math.pi + 0.25) == pytest.approx(-math.pi + 0.25)
    assert empty.wrap(-math.pi - 0.25) == pytest.approx(math.pi - 0.25)
    assert empty.wrap(7 * math.pi + 0.5) == pytest.approx(-math.pi + 0.5)
    assert empty.wrap(0.3) == 0.3
    # legacy switch: (-pi, pi]
    assert empty.wrap(math.pi, orc.WRAP_LEGACY) == math.pi
    assert empty.wrap(-math.pi, orc.WRAP_LEGACY) == pytest.approx(math.pi)


def test_wrap_applied_per_step(empty):
    # theta crossing +pi during a step lands in [-pi, pi) (StatePropagatorDatabase.h:67-74)
    q = empty.propagate([5, 5, 1.0, 1.0, math.pi - 0.05], [0, 0])
    assert -math.pi <= q[4] < math.pi
    assert q[4] < 0


def test_bounds(empty):
    # StateSpaceDatabase.h:51-58 via satisfiesBounds with DBL_EPSILON slack
    ok = lambda q: empty.is_valid(q)[0]
    assert ok([0.0, 0.0, 0.0, 0.0, 0.0])
    assert ok([32.0, 32.0, 1.0, math.pi / 3, 0.0])
    assert not ok([-1e-3, 1.0, 0.0, 0.0, 0.0])
    assert not ok([1.0, 32.001, 0.0, 0.0, 0.0])
    assert not ok([1.0, 1.0, 1.0 + 1e-3, 0.0, 0.0])
    assert not ok([1.0, 1.0, 0.0, -math.pi / 3 - 1e-6, 0.0])
    assert ok([1.0, 1.0, 1.0 + 1e-17, 0.0, 0.0])
    # the footprint is NOT tested against the workspace boundary (only the centre is), :57
    assert ok([0.1, 0.1, 0.0, 0.0, 0.3])


def test_minus_theta_footprint(empty):
    # StateValidityCheckerDatabase.h:152-155: rotation by -theta
    c = empty.robot_corners(0.7, 0.5, 0.0, 0.0, 0.3)
    # corner (+L/2, +W/2) is the third appended point (Robot.h:92)
    assert c[2][0] == pytest.approx(0.35 * math.cos(0.3) + 0.25 * math.sin(0.3))
    assert c[2][1] == pytest.approx(-0.35 * math.sin(0.3) + 0.25 * math.cos(0.3))
    assert c[2] == pytest.approx([0.40825, 0.13540], abs=1e-5)
    assert not np.allclose(c[2], [0.26049, 0.34227], atol=1e-3)  # what +theta would give


def test_touching_is_collision(empty):
    # boost::geometry::disjoint is false when boundaries touch (:173)
    a = empty.robot_corners(1.0, 1.0, 0.0, 0.0, 0.0)
    b = empty.robot_corners(1.0, 1.0, 1.0, 0.0, 0.0)
    assert not empty.polygons_disjoint(a, b)
    b = empty.robot_corners(1.0, 1.0, 1.0 + 1e-9, 0.0, 0.0)
    assert empty.polygons_disjoint(a, b)
    # corner touch
    b = empty.robot_corners(1.0, 1.0, 1.0, 1.0, 0.0)
    assert not empty.polygons_disjoint(a, b)
    # containment without boundary contact
    b = empty.robot_corners(0.2, 0.2, 0.0, 0.0, 0.4)
    assert not empty.polygons_disjoint(a, b)
    assert empty.sat_gap_robot_robot(1, 1, [0, 0, 0], 1, 1, [1, 0, 0]) == 0.0


def test_sat_matches_polygon_disjoint(empty):
    # SAT gap sign == restated boost disjoint on random rectangle pairs (robot-robot and robot-obstacle)
    rng = np.random.default_rng(1)
    n_close = 0
    for _ in range(20000):
        la, wa, lb, wb = rng.uniform(0.3, 1.5, 4)
        sa = [rng.uniform(0, 3), rng.uniform(0, 3), rng.uniform(-math.pi, math.pi)]
        sb = [rng.uniform(0, 3), rng.uniform(0, 3), rng.uniform(-math.pi, math.pi)]
        gap = empty.sat_gap_robot_robot(la, wa, sa, lb, wb, sb)
        dis = empty.polygons_disjoint(empty.robot_corners(la, wa, *sa), empty.robot_corners(lb, wb, *sb))
        if abs(gap) > 1e-12:
            assert dis == (gap > 0)
            n_close += abs(gap) < 0.05
        obs = [rng.uniform(0, 3), rng.uniform(0, 3), rng.uniform(0.2, 2), rng.uniform(0.2, 2)]
        gap = empty.sat_gap_robot_obstacle(la, wa, sa, obs)
        oc = [[obs[0], obs[1]], [obs[0] + obs[2], obs[1]], [obs[0] + obs[2], obs[1] + obs[3]], [obs[0], obs[1] + obs[3]]]
        dis = empty.polygons_disjoint(empty.robot_corners(la, wa, *sa), oc)
        if abs(gap) > 1e-12:
            assert dis == (gap > 0)
    assert n_close > 100  # the sample does probe the decision boundary


def test_scene_start_states_valid(pkg, orc):
    # implied fixture: every demo start state is valid and pairwise conflict free
    for name in pkg.scenes.names():
        o = orc.Oracle(pkg.scenes.world(name))
        st = pkg.scenes.start_states(name)
        R = st.shape[1]
        for r in range(R):
            assert o.is_valid(st[:, r], r)[0], (name, r)
        for r1 in range(R):
            for r2 in range(r1 + 1, R):
                assert o.pair_valid(r1, st[[0, 1, 4], r1], r2, st[[0, 1, 4], r2])[0], (name, r1, r2)


def test_obstacle_validity(corridor):
    # Corridor obstacle (0.5, 1.8, 9.0, 1.0): robot 1x1 (demo_Corridor...:75,84)
    assert not corridor.is_valid([5.0, 2.3, 0, 0, 0])[0]          # centre inside the obstacle
    assert not corridor.is_valid([5.0, 1.3, 0, 0, 0])[0]          # top edge touches y = 1.8
    assert corridor.is_valid([5.0, 1.3 - 1e-9, 0, 0, 0])[0]
    ok, m = corridor.is_valid([5.0, 1.0, 0, 0, 0])
    assert ok and m == pytest.approx(0.3)
    # rotated by 45 degrees the footprint reaches 0.7071 up
    assert not corridor.is_valid([5.0, 1.2, 0, 0, math.pi / 4])[0]
    assert corridor.is_valid([5.0, 1.05, 0, 0, math.pi / 4])[0]


def test_propagate_while_valid(corridor):
    # stops at the first invalid step and returns the last valid state (propagateWhileValid)
    q0 = [5.0, 0.9, 1.0, 0.0, -math.pi / 2]   # rotation by -theta only matters for the footprint;
    n, seg, fin = corridor.propagate_while_valid(q0, [0.0, 0.0], 10)   # dynamics drive along -y: out of bounds
    assert 0 < n < 10
    assert fin == pytest.approx(seg[n - 1])
    nxt = corridor.propagate(fin, [0.0, 0.0])
    assert not corridor.is_valid(nxt)[0]
    # zero valid steps -> result is the start state
    n, seg, fin = corridor.propagate_while_valid([5.0, 1.29, 1.0, 0.0, math.pi / 2], [0.0, 0.0], 5)
    assert n == 0 and fin == pytest.approx([5.0, 1.29, 1.0, 0.0, math.pi / 2])


def _line_plan(R, T, spacing=3.0):
    traj = np.zeros((R, 3, T), np.float32)
    for r in range(R):
        traj[r, 0] = 2.0 + spacing * r
        traj[r, 1] = 5.0
    return traj


def test_first_conflict_order(pkg, orc):
    o = orc.Oracle(pkg.scenes.world("Empty32x32_10robots"))
    R, T = 6, 50
    traj = _line_plan(R, T)
    assert o.first_conflict(traj)[0] == (-1, -1, -1, -1)
    # two conflicts at the same k: lexicographically smaller pair wins
    t2 = traj.copy()
    t2[5, 0, 20] = t2[4, 0, 20] + 0.1     # pair (4, 5)
    t2[3, 0, 20] = t2[1, 0, 20] + 0.1     # pair (1, 3)
    assert o.first_conflict(t2)[0] == (1, 3, 20, 20)
    # smaller k wins regardless of pair; k_end = last consecutive colliding step of that pair
    t3 = t2.copy()
    t3[5, 0, 10:14] = t3[4, 0, 10:14] + 0.1
    assert o.first_conflict(t3)[0] == (4, 5, 10, 13)
    # conflict running to the end of the plan
    t4 = traj.copy()
    t4[2, 0, 45:] = t4[0, 0, 45:] + 0.2
    assert o.first_conflict(t4)[0] == (0, 2, 45, 49)


def test_first_conflict_padding(pkg, orc):
    # shorter paths are padded with their last state (KCBS::findConflicts)
    o = orc.Oracle(pkg.scenes.world("Empty32x32_10robots"))
    R, T = 3, 40
    traj = _line_plan(R, T)
    # robot 1 drives through where robot 0 parks after its path ends at k = 9
    traj[0, 0, :10] = np.linspace(0.5, 2.0, 10)
    traj[0, 0, 10:] = 30.0     # garbage beyond its length must be ignored
    traj[1, 0, :] = np.linspace(8.0, -2.0, T)
    traj[2, 1, :] = 20.0       # robot 2 out of the way
    lengths = np.array([10, T, T], np.int32)
    res = o.first_conflict(traj, lengths)[0]
    assert res[0] == 0 and res[1] == 1 and res[2] > 10
    assert o.first_conflict(traj, None)[0] == (-1, -1, -1, -1)


def test_merged_pair_oracle(pkg, orc):
    # StatePropagatorDatabase.h:130-153 and StateValidityCheckerDatabase.h:225-261 (merged pair)
    o = orc.Oracle(pkg.scenes.world("Corridor10x10_4robots"))
    goals = [9.0, 0.5, 9.0, 9.0]
    # both cars far from their goals: each moves like a single car
    q = [1.0, 0.5, 0.5, 0.0, 0.0, 1.0, 3.5, 0.3, 0.1, 0.2]
    u = [1.0, 0.0, -0.5, 0.4]
    out = o.pair_propagate(q, u, goals)
    assert np.allclose(out[:5], o.propagate(q[:5], u[:2]), atol=0) and np.allclose(out[5:], o.propagate(q[5:], u[2:]), atol=0)
    # car 1 within 1.5 of its goal before the step keeps its state, car 2 still moves
    q2 = [8.0, 0.5, 0.5, 0.0, 0.0, 1.0, 3.5, 0.3, 0.1, 0.2]
    out = o.pair_propagate(q2, u, goals)
    assert list(out[:5]) == q2[:5] and out[5] != q2[5]
    # exactly at the hold radius is NOT inside (strict <)
    q3 = [7.5, 0.5, 0.5, 0.0, 0.0, 1.0, 3.5, 0.3, 0.1, 0.2]
    assert o.pair_propagate(q3, u, goals)[0] != 7.5
    # validity: both in bounds and obstacle free, not overlapping each other
    assert o.pair_is_valid([1.0, 0.5, 0, 0, 0, 1.0, 3.5, 0, 0, 0])[0]
    assert not o.pair_is_valid([1.0, 0.5, 0, 0, 0, 1.5, 0.5, 0, 0, 0])[0]       # the cars overlap
    assert not o.pair_is_valid([1.0, 0.5, 0, 0, 0, 2.0, 0.5, 0, 0, 0])[0]       # touching
    assert not o.pair_is_valid([1.0, 0.5, 0, 0, 0, 5.0, 2.3, 0, 0, 0])[0]       # second car inside the obstacle
    assert not o.pair_is_valid([1.0, 0.5, 1.2, 0, 0, 1.0, 3.5, 0, 0, 0])[0]     # first car's speed out of bounds


def test_first_conflict_groups(pkg, orc):
    # bodies of one merged individual are never tested against each other
    o = orc.Oracle(pkg.scenes.world("Empty32x32_10robots"))
    R, T = 5, 30
    traj = _line_plan(R, T)
    traj[4, 0, :] = traj[3, 0, :] + 0.2          # bodies 3 and 4 overlap all the time
    traj[2, 0, 12:15] = traj[0, 0, 12:15] + 0.1  # bodies 0 and 2 collide at 12..14
    assert o.first_conflict(traj)[0] == (3, 4, 0, T - 1)
    assert o.first_conflict_groups(traj, [0, 1, 2, 3, 3])[0] == (0, 2, 12, 14)
    assert o.first_conflict_groups(traj, [0, 1, 0, 3, 3])[0] == (-1, -1, -1, -1)
