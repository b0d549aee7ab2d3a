"""GPU parity: kcbs_first_conflict (K3) against the oracle's findConflicts restatement: the first
(robot pair, time index) must be bit-exact whenever no pair-time predicate up to the conflict lies
within 1e-6 of touching (north_star)."""
import numpy as np
import pytest

from helpers import BAND, rollout_plan

pytestmark = pytest.mark.gpu


def _t(res):
    return [tuple(int(v) for v in r) for r in res]


@pytest.mark.parametrize("scene,T", [("Corridor10x10_4robots", 600), ("Congested10x10_7robots", 500),
                                     ("Empty32x32_10robots", 1500), ("Empty32x32_20robots", 1000),
                                     ("Benchmark_Empty32x32_30robots", 1000)])
def test_natural_conflicts(pkg, orc, gpu_ctx_factory, scene, T):
    ctx = gpu_ctx_factory(scene)
    oracle = orc.Oracle(pkg.scenes.world(scene))
    starts = pkg.scenes.start_states(scene)
    plans = np.stack([rollout_plan(oracle, starts, T, seed) for seed in (1, 2, 3)])
    got = _t(ctx.first_conflict(plans))
    n_conf = 0
    for p in range(plans.shape[0]):
        ref, min_margin, _ = oracle.first_conflict(plans[p])
        assert min_margin > BAND, "test plan has a predicate inside the excluded band; change the seed"
        assert got[p] == ref, f"plan {p}: gpu {got[p]} oracle {ref}"
        n_conf += ref[0] >= 0
    if "Empty" in scene and len(starts[0]) >= 20:
        assert n_conf > 0   # crowded random driving does collide


def test_planted_conflicts_and_tie_break(pkg, orc, workloads, gpu_ctx_factory):
    scene = "Benchmark_Empty32x32_30robots"
    ctx = gpu_ctx_factory(scene)
    oracle = orc.Oracle(pkg.scenes.world(scene))
    T = 1000
    traj = workloads.cell_plans(scene, 8, T, seed=3)
    assert _t(ctx.first_conflict(traj)) == [(-1, -1, -1, -1)] * 8
    workloads.plant_conflict(traj, 1, 4, 9, 500, length=7)
    workloads.plant_conflict(traj, 2, 20, 29, 640)                 # same k, two pairs: (3, 17) wins
    workloads.plant_conflict(traj, 2, 3, 17, 640, length=3)
    workloads.plant_conflict(traj, 3, 28, 29, 0, length=2)         # first time index
    workloads.plant_conflict(traj, 4, 0, 1, T - 1)                 # last time index
    workloads.plant_conflict(traj, 5, 10, 11, 300, length=700)     # runs to the end of the plan
    workloads.plant_conflict(traj, 6, 5, 6, 129, length=140)       # extension crosses tile borders
    workloads.plant_conflict(traj, 6, 1, 2, 400)                   # later conflict of a smaller pair loses
    workloads.plant_conflict(traj, 7, 7, 8, 127)
    workloads.plant_conflict(traj, 7, 2, 3, 128)
    got = _t(ctx.first_conflict(traj))
    ref, _ = oracle.first_conflict_batch(traj, n_threads=8)
    assert got == _t(ref)
    assert got[1] == (4, 9, 500, 506) and got[2] == (3, 17, 640, 642) and got[3] == (28, 29, 0, 1)
    assert got[4] == (0, 1, T - 1, T - 1) and got[5] == (10, 11, 300, T - 1) and got[6] == (5, 6, 129, 268)
    assert got[7] == (7, 8, 127, 127)


@pytest.mark.parametrize("T", [1, 2, 127, 128, 129, 1000])
def test_ragged_lengths_and_padding(pkg, orc, workloads, gpu_ctx_factory, T):
    scene = "Empty32x32_10robots"
    ctx = gpu_ctx_factory(scene)
    oracle = orc.Oracle(pkg.scenes.world(scene))
    rng = np.random.default_rng(T)
    P, R = 6, 10
    traj = workloads.cell_plans(scene, P, T, seed=T)
    lengths = rng.integers(1, T + 1, (P, R)).astype(np.int32)
    lengths[0] = T
    # garbage beyond each path's length must be ignored (padding uses the last state)
    for p in range(P):
        for r in range(R):
            traj[p, r, :, lengths[p, r]:] = 1.0
    if T > 4:
        # a parked robot (short path) gets run over later by robot 1
        lengths[1, 0] = 3
        traj[1, 1, 0, T // 2:] = traj[1, 0, 0, 2]
        traj[1, 1, 1, T // 2:] = traj[1, 0, 1, 2]
        lengths[1, 1] = T
    got = _t(ctx.first_conflict(traj, lengths))
    ref, _ = oracle.first_conflict_batch(traj, lengths, n_threads=8)
    assert got == _t(ref)
    if T > 4:
        assert got[1][:3] == (0, 1, T // 2)


def test_sharded_keys_reduce_to_full_scan(pkg, orc, workloads, gpu_ctx_factory):
    # time slabs scanned separately + min over the packed keys == one full scan (multi-GPU path)
    import torch

    scene = "Empty32x32_20robots"
    ctx = gpu_ctx_factory(scene)
    oracle = orc.Oracle(pkg.scenes.world(scene))
    P, R, T = 5, 20, 900
    traj = workloads.cell_plans(scene, P, T, seed=8)
    workloads.plant_conflict(traj, 0, 3, 4, 100)
    workloads.plant_conflict(traj, 0, 1, 2, 700)
    workloads.plant_conflict(traj, 2, 6, 19, 899)
    workloads.plant_conflict(traj, 3, 0, 5, 450, length=300)
    d_traj = torch.from_numpy(traj).cuda()
    shards = 4
    keys = torch.empty((shards, P), dtype=torch.int64, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    for s in range(shards):
        ctx.conflict_keys_dev(P, R, T, s * T // shards, (s + 1) * T // shards, d_traj, None, keys[s], stream)
    # unsigned min: the no-conflict key is all ones (-1 as int64), so compare as uint64 via a bias flip
    red = (keys ^ (1 << 63)).min(dim=0).values ^ (1 << 63)
    out = torch.empty((P, 4), dtype=torch.int32, device="cuda")
    ctx.conflict_finish_dev(P, R, T, d_traj, None, red, out, stream)
    torch.cuda.synchronize()
    got = [tuple(int(v) for v in row) for row in out.cpu().numpy()]
    ref, _ = oracle.first_conflict_batch(traj, n_threads=8)
    assert got == _t(ref)
    assert got[0] == (3, 4, 100, 100) and got[3] == (0, 5, 450, 749)


def test_device_entry_point_matches_host(pkg, workloads, gpu_ctx_factory):
    import torch

    scene = "Benchmark_Empty32x32_30robots"
    ctx = gpu_ctx_factory(scene)
    traj = workloads.cell_plans(scene, 4, 2000, seed=2)
    workloads.plant_conflict(traj, 2, 11, 12, 1234, length=4)
    host = _t(ctx.first_conflict(traj))
    d = torch.from_numpy(traj).cuda()
    out = torch.empty((4, 4), dtype=torch.int32, device="cuda")
    ctx.first_conflict_dev(4, 30, 2000, d, None, out, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert [tuple(int(v) for v in r) for r in out.cpu().numpy()] == host
    assert host[2] == (11, 12, 1234, 1237)


def test_errors(pkg, gpu_ctx_factory):
    ctx = gpu_ctx_factory("Corridor10x10_4robots")
    with pytest.raises(pkg.KcbsError) as e:   # more robots than the world describes
        ctx.first_conflict(np.zeros((1, 9, 3, 10), np.float32))
    assert e.value.status == 5
    assert ctx.first_conflict(np.zeros((0, 4, 3, 10), np.float32)).shape == (0,)


@pytest.mark.parametrize("R", [2, 9, 33, 64])
def test_robot_counts_and_mixed_footprints(pkg, orc, gpu_ctx_factory, R):
    # register tiles of 8 robots: partial last tile (sentinel rows), exactly full tiles, the 64-robot maximum;
    # robots of different sizes use their own rectangles and radii
    rng = np.random.default_rng(R)
    dims = [[0.7, 0.5] if r % 3 else [1.2, 0.4] for r in range(R)]
    w = pkg.World(x_max=32, y_max=32, robot_dims=dims)
    ctx = gpu_ctx_factory(w)
    oracle = orc.Oracle(w)
    P, T = 4, 300
    traj = np.zeros((P, R, 3, T), np.float32)
    for p in range(P):
        pos = rng.uniform(1, 31, (R, 2))
        vel = rng.uniform(-0.05, 0.05, (R, 2))
        for k in range(T):
            vel += rng.normal(0, 0.005, (R, 2))
            pos = np.clip(pos + vel, 0.5, 31.5)
            traj[p, :, 0, k], traj[p, :, 1, k] = pos[:, 0], pos[:, 1]
            traj[p, :, 2, k] = np.arctan2(vel[:, 1], vel[:, 0])
    got = _t(ctx.first_conflict(traj))
    for p in range(P):
        ref, min_margin, _ = oracle.first_conflict(traj[p])
        if min_margin > BAND:
            assert got[p] == ref, f"R={R} plan {p}: gpu {got[p]} oracle {ref}"


@pytest.mark.parametrize("counts", [(1, 100), (128, 256), (1, 66, 102, 300, 5, 301)])
def test_plan_count_grows_on_one_context(pkg, workloads, counts):
    # the keys / tickets workspace of a context is laid out for a plan count; a later call with MORE plans must re-lay
    # it out even when the bytes still fit the (over-allocated) buffer -- otherwise keys alias tickets and plans past
    # the old count decode as a conflict at k = 0, r1 = r2 = 0
    import torch
    scene = "Benchmark_Empty32x32_30robots"
    T = 512
    with pkg.Context(pkg.scenes.world(scene), device=0) as ctx:
        for n in counts:
            traj = workloads.cell_plans(scene, n, T, seed=5)
            want = [(-1, -1, -1, -1)] * n
            if n >= 5:
                workloads.plant_conflict(traj, n - 1, 2, 9, 300, length=4)
                want[n - 1] = (2, 9, 300, 303)
            assert _t(ctx.first_conflict(traj)) == want, f"host entry point, {n} plans after {counts}"
            d = torch.from_numpy(traj).cuda()
            out = torch.zeros((n, 4), dtype=torch.int32, device="cuda")
            for _ in range(2):   # twice: the kernel must also leave keys and tickets clean
                ctx.first_conflict_dev(n, traj.shape[1], T, d, None, out, torch.cuda.current_stream())
                torch.cuda.synchronize()
                assert _t(out.cpu().numpy()) == want, f"device entry point, {n} plans after {counts}"
