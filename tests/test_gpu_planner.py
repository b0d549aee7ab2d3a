"""GPU tests of the "next" rows: the batched low-level planner (kcbs_rrt_plan) and the K-CBS
high level (kcbs_kcbs_solve).  Plans are checked with the CPU oracle: every state valid, consecutive states
connected by one reference propagation step under an admissible control, goals reached, and the multi-robot plan
conflict free under the reference's scan."""
import numpy as np
import pytest

from helpers import state_close

pytestmark = pytest.mark.gpu


def check_path(oracle, traj, start, goal, robot=0, radius=0.5):
    """traj [5, T]: dynamically feasible under the reference propagator, valid, start -> goal region."""
    T = traj.shape[1]
    assert T >= 1
    assert (traj[:, 0] == np.asarray(start, np.float32)).all()
    assert np.hypot(traj[0, -1] - goal[0], traj[1, -1] - goal[1]) <= radius + 1e-4
    for k in range(T - 1):
        q, qn = traj[:, k].astype(np.float64), traj[:, k + 1].astype(np.float64)
        u = np.array([(qn[2] - q[2]) / 0.1, (qn[3] - q[3]) / 0.1])
        assert np.abs(u).max() <= 1.0 + 1e-3, f"step {k}: control {u} outside [-1, 1]^2"
        ref = oracle.propagate(q, np.clip(u, -1, 1))
        ok, d = state_close(qn.astype(np.float32)[:, None], ref[:, None])
        assert ok.all(), f"step {k}: {d.ravel()}"
        valid, margin = oracle.is_valid(qn, robot)
        assert valid or abs(margin) < 2e-5, f"state {k + 1} invalid (margin {margin})"


@pytest.mark.parametrize("scene,robot", [("Empty32x32_10robots", 0), ("Empty32x32_10robots", 7), ("Corridor10x10_4robots", 1)])
def test_rrt_reaches_goal(pkg, orc, gpu_ctx_factory, scene, robot):
    ctx = gpu_ctx_factory(scene)
    oracle = orc.Oracle(pkg.scenes.world(scene))
    start = pkg.scenes.start_states(scene)[:, robot]
    goal = pkg.scenes.goals(scene)[robot]
    traj, st = ctx.rrt_plan(robot, start, goal, ctx.rrt_params(seed=3))
    assert st.solved == 1 and st.iterations >= 1 and st.nodes > 1
    assert traj.shape[1] == st.path_steps + 1
    check_path(oracle, traj, start, goal, robot)


def test_rrt_respects_constraints(pkg, orc, gpu_ctx_factory):
    # a robot parked across the straight line to the goal for the whole horizon must be avoided
    scene = "Empty32x32_10robots"
    ctx = gpu_ctx_factory(scene)
    oracle = orc.Oracle(pkg.scenes.world(scene))
    start = np.array([5.0, 16.0, 0, 0, 0], np.float32)
    goal = np.array([12.0, 16.0], np.float32)
    K = 600
    block = np.zeros((3, K), np.float32)
    block[0], block[1], block[2] = 8.5, 16.0, np.pi / 2
    traj, st = ctx.rrt_plan(0, start, goal, ctx.rrt_params(seed=5), constraints=[(1, 0, block)])
    assert st.solved == 1
    check_path(oracle, traj, start, goal)
    for k in range(traj.shape[1]):
        ok, _ = oracle.pair_valid(0, traj[[0, 1, 4], k], 1, block[:, min(k, K - 1)])
        assert ok, f"constraint violated at step {k}"
    # without the constraint the planner is free to drive through that spot
    free, st2 = ctx.rrt_plan(0, start, goal, ctx.rrt_params(seed=5))
    assert st2.solved == 1 and free.shape[1] <= traj.shape[1] + 200


def test_rrt_unreachable_goal_times_out(pkg, gpu_ctx_factory):
    ctx = gpu_ctx_factory("Corridor10x10_4robots")
    # goal inside the long obstacle (0.5, 1.8, 9.0, 1.0): never reachable
    traj, st = ctx.rrt_plan(0, [1.0, 0.5, 0, 0, 0], [5.0, 2.3], ctx.rrt_params(max_iterations=20, goal_radius=0.2))
    assert st.solved == 0 and traj.shape[1] == 0 and st.iterations == 20


@pytest.mark.parametrize("scene", ["Empty32x32_10robots", "Corridor10x10_4robots"])
def test_kcbs_solves_demo_scene(pkg, orc, gpu_ctx_factory, scene):
    ctx = gpu_ctx_factory(scene)
    oracle = orc.Oracle(pkg.scenes.world(scene))
    starts, goals = pkg.scenes.start_states(scene), pkg.scenes.goals(scene)
    R = starts.shape[1]
    plan, lengths, st = ctx.kcbs_solve(starts, goals, ctx.kcbs_params(time_limit_s=120.0))
    assert st.solved == 1, f"K-CBS failed: {st.nodes_expanded} nodes, {st.low_level_calls} low-level calls"
    assert st.nodes_expanded >= 1 and st.root_seconds > 0 and st.seconds >= st.root_seconds
    T = int(lengths.max())
    for r in range(R):
        check_path(oracle, plan[r, :, :lengths[r]], starts[:, r], goals[r], r)
    traj = np.ascontiguousarray(plan[:, [0, 1, 4], :T])
    res, _, _ = oracle.first_conflict(traj, lengths)
    assert res == (-1, -1, -1, -1), f"plan has a conflict {res}"


def test_rrt_plan_many_equals_single_plans(pkg, orc, gpu_ctx_factory):
    # the concurrent batch entry point returns, job by job, exactly what kcbs_rrt_plan returns alone (same seeds):
    # the interleaving of the planner lanes does not leak into the results
    scene = "Empty32x32_20robots"
    ctx = gpu_ctx_factory(scene)
    oracle = orc.Oracle(pkg.scenes.world(scene))
    starts, goals = pkg.scenes.start_states(scene), pkg.scenes.goals(scene)
    robots = [0, 3, 5, 8, 11, 12, 13, 17, 19, 2, 4]
    seeds = np.array([1000 + r for r in robots], np.uint64)
    traj, T, stats = ctx.rrt_plan_many(robots, starts[:, robots], goals[robots], seeds=seeds, max_T=1024)
    assert all(s.solved for s in stats)
    for j, r in enumerate(robots):
        single, st = ctx.rrt_plan(r, starts[:, r], goals[r], ctx.rrt_params(seed=int(seeds[j])), max_T=1024)
        assert single.shape[1] == T[j] and (single == traj[j, :, :T[j]]).all()
        assert (traj[j, :, T[j]:] == 0).all()
    check_path(oracle, traj[0, :, :T[0]], starts[:, robots[0]], goals[robots[0]], robots[0])
    check_path(oracle, traj[7, :, :T[7]], starts[:, robots[7]], goals[robots[7]], robots[7])


def test_sharded_replans_single_rank(pkg, orc, gpu_ctx_factory):
    # config 'per-robot replans sharded across GPUs' on one rank: all 20 replans + merge + validation
    import torch
    from kcbs_b200 import sharding
    scene = "Empty32x32_20robots"
    ctx = gpu_ctx_factory(scene)
    starts, goals = pkg.scenes.start_states(scene), pkg.scenes.goals(scene)
    plan, lengths, ok = sharding.plan_replans_sharded(ctx, starts, goals, 0, 1, 1024, seed=11)
    assert ok and plan.shape == (20, 5, 1024) and int(lengths.min()) > 1
    oracle = orc.Oracle(pkg.scenes.world(scene))
    p = plan.cpu().numpy()
    for r in (0, 9, 19):
        check_path(oracle, p[r, :, :int(lengths[r])], starts[:, r], goals[r], r)
    assert torch.cuda.is_available()


@pytest.mark.parametrize("scene,bound", [("Corridor10x10_4robots", 0), ("Congested10x10_7robots", 1), ("Empty32x32_10robots", 0)])
def test_kcbs_merges_when_the_bound_is_exceeded(pkg, orc, gpu_ctx_factory, scene, bound):
    # KCBS::setMergeBound: a pair with more conflicts than the bound becomes one individual (planned together, its
    # bodies never tested against each other by the high level); the plan that comes out must still be conflict free
    # under the reference's FULL scan (no groups), every path feasible and valid
    ctx = gpu_ctx_factory(scene)
    oracle = orc.Oracle(pkg.scenes.world(scene))
    starts, goals = pkg.scenes.start_states(scene), pkg.scenes.goals(scene)
    params = ctx.kcbs_params(low_level=ctx.rrt_params(seed=2, time_limit_s=5.0), time_limit_s=120.0, merge_bound=bound)
    plan, lengths, st = ctx.kcbs_solve(starts, goals, params, max_T=1024)
    assert st.solved == 1
    R = starts.shape[1]
    T = int(lengths.max())
    traj = np.zeros((R, 3, T), np.float32)
    for r in range(R):
        L = int(lengths[r])
        traj[r, :, :L] = plan[r][[0, 1, 4], :L]
    res, _, _ = oracle.first_conflict(traj, lengths=lengths)
    assert res[0] == -1, f"conflict {res} in a merged plan"
    for r in range(R):
        # (a car of a merged pair stops within 1.5 of its goal and stands still from then on: its path is checked state by
        # state here, the joint dynamics in test_pair_rrt_plans_two_cars_as_one_system)
        L = int(lengths[r])
        assert np.hypot(plan[r][0, L - 1] - goals[r][0], plan[r][1, L - 1] - goals[r][1]) < 1.5 + 1e-4
        for k in range(0, L, 7):
            valid, margin = oracle.is_valid(plan[r][:, k].astype(np.float64), r)
            assert valid or abs(margin) < 2e-5
    if scene != "Empty32x32_10robots":
        assert st.merges >= 1     # these scenes do conflict at the root


def test_rrt_never_returns_a_truncated_path(pkg, gpu_ctx_factory):
    # a trajectory buffer too short for any path to the goal: the plan must come back unsolved, not cut off
    scene = "Empty32x32_10robots"
    ctx = gpu_ctx_factory(scene)
    start = np.array([3.0, 3.0, 0, 0, 0], np.float32)
    goal = np.array([29.0, 29.0], np.float32)       # > 36 m away at |v| <= 1: more than 360 steps
    traj, st = ctx.rrt_plan(0, start, goal, ctx.rrt_params(seed=3, max_iterations=200), max_T=128)
    assert st.solved == 0 and traj.shape[1] == 0
    full, st2 = ctx.rrt_plan(0, start, goal, ctx.rrt_params(seed=3), max_T=4096)
    assert st2.solved == 1 and full.shape[1] > 128
    assert np.hypot(full[0, -1] - goal[0], full[1, -1] - goal[1]) <= 0.5 + 1e-4
    # exactly fitting buffer: solved, complete
    fit, st3 = ctx.rrt_plan(0, start, goal, ctx.rrt_params(seed=3), max_T=full.shape[1])
    assert st3.solved == 1 and (fit == full).all()


def test_rrt_persistent_constraint(pkg, orc, gpu_ctx_factory):
    # a NEGATIVE constraint length marks a persistent constraint: the other robot drives for 40 steps and then stays
    # parked across the way for ever; the planned path (much longer than 40 steps) must avoid the parked state too
    scene = "Empty32x32_10robots"
    ctx = gpu_ctx_factory(scene)
    oracle = orc.Oracle(pkg.scenes.world(scene))
    start = np.array([5.0, 16.0, 0, 0, 0], np.float32)
    goal = np.array([12.0, 16.0], np.float32)
    K = 40
    other = np.zeros((3, K), np.float32)
    other[0] = np.linspace(8.5, 8.5, K)
    other[1] = np.linspace(12.0, 16.0, K)      # arrives on the straight line start -> goal and parks there
    other[2] = np.pi / 2
    traj, st = ctx.rrt_plan(0, start, goal, ctx.rrt_params(seed=5), constraints=[(1, 0, other)], persistent=[True])
    assert st.solved == 1 and traj.shape[1] > K
    check_path(oracle, traj, start, goal)
    for k in range(traj.shape[1]):
        ok, _ = oracle.pair_valid(0, traj[[0, 1, 4], k], 1, other[:, min(k, K - 1)])
        assert ok, f"persistent constraint violated at step {k}"
    # the same constraint without persistence stops binding after K steps: some seed drives through the parked spot
    # (or at least the plans differ), so the flag is what made the difference
    free, st2 = ctx.rrt_plan(0, start, goal, ctx.rrt_params(seed=5), constraints=[(1, 0, other)])
    assert st2.solved == 1


def check_pair_path(oracle, traj, start10, goals, r1=0, r2=1):
    """traj [10, T]: every transition is one step of the reference's merged propagator (freeze rule included) under
    admissible controls, every state valid for the merged system, both cars end within 1.5 of their goals."""
    T = traj.shape[1]
    assert (traj[:, 0] == np.asarray(start10, np.float32)).all()
    for c, g in ((0, goals[:2]), (5, goals[2:])):
        assert np.hypot(traj[c, -1] - g[0], traj[c + 1, -1] - g[1]) < 1.5 + 1e-4
    for k in range(T - 1):
        q, qn = traj[:, k].astype(np.float64), traj[:, k + 1].astype(np.float64)
        u = np.array([(qn[2] - q[2]) / 0.1, (qn[3] - q[3]) / 0.1, (qn[7] - q[7]) / 0.1, (qn[8] - q[8]) / 0.1])
        assert np.abs(u).max() <= 1.0 + 1e-3
        ref = oracle.pair_propagate(q, np.clip(u, -1, 1), goals)
        for c in (0, 5):
            ok, d = state_close(qn[c:c + 5].astype(np.float32)[:, None], ref[c:c + 5][:, None])
            assert ok.all(), f"step {k} car {c // 5}: {d.ravel()}"
        valid, margin = oracle.pair_is_valid(qn, r1, r2)
        assert valid or abs(margin) < 2e-5, f"state {k + 1} invalid (margin {margin})"


def test_pair_rrt_plans_two_cars_as_one_system(pkg, orc, gpu_ctx_factory):
    # two cars that have to cross each other's straight line, planned jointly in the 10-d space
    scene = "Empty32x32_10robots"
    ctx = gpu_ctx_factory(scene)
    oracle = orc.Oracle(pkg.scenes.world(scene))
    start = np.array([10, 16, 0, 0, 0, 16, 10, 0, 0, np.pi / 2], np.float32)
    goals = np.array([22, 16, 16, 22], np.float32)
    traj, st = ctx.rrt_plan_pair(0, 1, start, goals, ctx.rrt_params(seed=4, time_limit_s=30.0))
    assert st.solved == 1 and traj.shape[0] == 10 and traj.shape[1] == st.path_steps + 1
    check_pair_path(oracle, traj, start, goals)
    # a third robot parked for ever on car 1's straight line: both cars of the pair avoid it at all times
    K = 5
    block = np.zeros((3, K), np.float32)
    block[0], block[1], block[2] = 16.0, 16.0, 0.3
    traj2, st2 = ctx.rrt_plan_pair(0, 1, start, goals, ctx.rrt_params(seed=4, time_limit_s=30.0), constraints=[(2, 0, block)],
                                   persistent=[True])
    assert st2.solved == 1
    check_pair_path(oracle, traj2, start, goals)
    for k in range(traj2.shape[1]):
        for c in (0, 5):
            ok, _ = oracle.pair_valid(c // 5, traj2[[c, c + 1, c + 4], k], 2, block[:, -1])
            assert ok, f"constraint violated by car {c // 5} at step {k}"


def test_kcbs_plans_the_merged_pair_jointly(pkg, orc, gpu_ctx_factory):
    # merge bound 0: the first conflicting pair is merged at once and from then on planned as ONE 10-d system (no car-by-car
    # fallback, no split); Robot 1 and Robot 3 of the corridor scene have to swap places in a passage two robots wide
    scene = "Corridor10x10_4robots"
    ctx = gpu_ctx_factory(scene)
    oracle = orc.Oracle(pkg.scenes.world(scene))
    starts, goals = pkg.scenes.start_states(scene), pkg.scenes.goals(scene)
    params = ctx.kcbs_params(low_level=ctx.rrt_params(seed=3, time_limit_s=20.0), time_limit_s=170.0, merge_bound=0)
    plan, lengths, st = ctx.kcbs_solve(starts, goals, params, max_T=2048)
    assert st.solved == 1 and st.merges >= 1
    R = starts.shape[1]
    T = int(lengths.max())
    traj = np.zeros((R, 3, T), np.float32)
    for r in range(R):
        L = int(lengths[r])
        traj[r, :, :L] = plan[r][[0, 1, 4], :L]
    res, _, _ = oracle.first_conflict(traj, lengths=lengths)
    assert res[0] == -1, f"conflict {res} in the merged plan"
    # merged cars stop within 1.5 of their goals (the merged goal region), the others within 0.5
    within = [float(np.hypot(plan[r][0, lengths[r] - 1] - goals[r][0], plan[r][1, lengths[r] - 1] - goals[r][1])) for r in range(R)]
    assert sum(d <= 0.5 + 1e-4 for d in within) >= R - 2 * st.merges and max(within) < 1.5 + 1e-4


def test_concurrent_solves_and_device_wide_work(pkg, gpu_ctx_factory):
    # Host threads of one process: two K-CBS solves on two contexts (their planner lanes capture the round-loop graphs
    # while they run) beside a thread whose plan scans keep growing -- every growth re-allocates after a device-wide
    # synchronisation, which must never meet a capture in progress (capture_gate, csrc/ctx.h).  Results = the same calls
    # made one after the other: a plan depends on its seed only.
    import threading

    from kcbs_b200 import workloads

    scenes = ["Corridor10x10_4robots", "Empty32x32_10robots"]
    seeds = [11, 12, 13]

    def solve_all(ctx, scene):
        starts, goals = pkg.scenes.start_states(scene), pkg.scenes.goals(scene)
        out = []
        for seed in seeds:
            plan, lengths, st = ctx.kcbs_solve(starts, goals, ctx.kcbs_params(low_level=ctx.rrt_params(seed=seed), time_limit_s=60.0),
                                               max_T=1024)
            out.append((st.solved, lengths.copy(), plan.copy()))
        return out

    def fresh(scene):   # new contexts: the lanes, their workspaces and graphs are built while the other threads run
        return pkg.Context(pkg.scenes.world(scene), device=0)

    results, errors = {}, []

    def worker(name, fn):
        try:
            results[name] = fn()
        except Exception as e:  # noqa: BLE001
            errors.append((name, repr(e)))

    def scans():
        scene = "Empty32x32_10robots"
        n = 0
        with fresh(scene) as ctx:
            for P in (1, 3, 9, 40, 130, 300, 700):
                res = ctx.first_conflict(workloads.cell_plans(scene, P, 256, seed=P))
                assert (res["r1"] == -1).all()
                n += P
        return n

    ctxs = [fresh(s) for s in scenes]
    try:
        threads = [threading.Thread(target=worker, args=(s, lambda c=c, s=s: solve_all(c, s))) for c, s in zip(ctxs, scenes)]
        threads.append(threading.Thread(target=worker, args=("scans", scans)))
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        assert not errors, errors
        assert results["scans"] == 1 + 3 + 9 + 40 + 130 + 300 + 700
    finally:
        for c in ctxs:
            c.close()
    for scene in scenes:
        with fresh(scene) as ctx:
            alone = solve_all(ctx, scene)
        for (s1, l1, p1), (s2, l2, p2) in zip(results[scene], alone):
            assert s1 == 1 and s2 == 1
            assert (l1 == l2).all() and (p1 == p2).all(), f"{scene}: a concurrent solve differs from the same solve alone"


def test_rrt_many_constraints_and_equivalent_splits(pkg, orc, gpu_ctx_factory):
    # (a) more constraints than the 32 the expansion kernel's per-sample mask covers: a wall of 40 parked robots across the
    # straight line, each present for its own time window, with a gap that opens late; every state of the plan respects
    # every constraint.  (b) one long constraint and the same states cut into 7 pieces are the same constraint set: the
    # mask (bounding circle and time window per piece) must not change the plan.
    scene = "Empty32x32_10robots"
    ctx = gpu_ctx_factory(scene)
    oracle = orc.Oracle(pkg.scenes.world(scene))
    start = np.array([4.0, 16.0, 0, 0, 0], np.float32)
    goal = np.array([14.0, 16.0], np.float32)
    K = 400
    cons = []
    for q in range(40):
        y = 6.0 + 0.5 * q                       # a wall at x = 9 from y = 6 to 25.5
        k0 = 3 * q
        block = np.zeros((3, K - k0), np.float32)
        block[0], block[1], block[2] = 9.0, y, 0.0
        if 19 <= q <= 21:                       # the part in front of the robot goes away at step 150
            block = block[:, :150 - k0]
        cons.append((1 + q % 9, k0, block))
    traj, st = ctx.rrt_plan(0, start, goal, ctx.rrt_params(seed=8), constraints=cons, max_T=1024)
    assert st.solved == 1
    check_path(oracle, traj, start, goal)
    for k in range(traj.shape[1]):
        for other, k0, block in cons:
            if k0 <= k < k0 + block.shape[1]:
                ok, _ = oracle.pair_valid(0, traj[[0, 1, 4], k], other, block[:, k - k0])
                assert ok, f"constraint of robot {other} from step {k0} violated at step {k}"
    # (b)
    long = np.zeros((3, 300), np.float32)
    long[0] = np.linspace(6.0, 12.0, 300)
    long[1] = 16.0 + 0.8 * np.sin(np.linspace(0, 6, 300))
    long[2] = 0.2
    whole, st1 = ctx.rrt_plan(0, start, goal, ctx.rrt_params(seed=9), constraints=[(2, 20, long)], max_T=1024)
    cuts = [0, 11, 50, 51, 140, 200, 299, 300]
    pieces = [(2, 20 + a, long[:, a:b]) for a, b in zip(cuts[:-1], cuts[1:])]
    split, st2 = ctx.rrt_plan(0, start, goal, ctx.rrt_params(seed=9), constraints=pieces, max_T=1024)
    assert st1.solved == 1 and st2.solved == 1
    assert whole.shape == split.shape and (whole == split).all()
