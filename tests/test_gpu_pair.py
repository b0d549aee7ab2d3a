"""GPU parity of the merged-pair path: kcbs_propagate_pair_batch, kcbs_validity_pair_batch and
kcbs_first_conflict_groups against the oracle's restatement of TwoSecondOrderCarStatePropagator /
homogeneousTwo2ndOrderCarSystemSVC."""
import numpy as np
import pytest

from helpers import BAND, state_close

pytestmark = pytest.mark.gpu


def _pair_batch(scene_dict, n, seed, interior=0.8):
    g = np.random.Generator(np.random.Philox(key=[seed, 9]))
    p = np.empty((10, n), np.float32)
    for c in range(2):
        p[5 * c + 0] = g.uniform(interior, scene_dict["x_max"] - interior, n)
        p[5 * c + 1] = g.uniform(interior, scene_dict["y_max"] - interior, n)
        p[5 * c + 2] = g.uniform(-1, 1, n)
        p[5 * c + 3] = g.uniform(-1.0, 1.0, n)
        p[5 * c + 4] = g.uniform(-np.pi, np.pi, n)
    u = g.uniform(-1, 1, (4, n)).astype(np.float32)
    steps = g.integers(1, 11, n).astype(np.int32)
    return p, u, steps


@pytest.mark.parametrize("scene", ["Corridor10x10_4robots", "Congested10x10_7robots", "Empty32x32_10robots"])
def test_pair_propagation(pkg, orc, gpu_ctx_factory, scene):
    ctx = gpu_ctx_factory(scene)
    oracle = orc.Oracle(pkg.scenes.world(scene))
    sd = pkg.scenes.SCENES[scene]
    n = 20000
    parents, controls, steps = _pair_batch(sd, n, seed=3)
    goals = np.array([sd["x_max"] * 0.7, sd["y_max"] * 0.3, sd["x_max"] * 0.3, sd["y_max"] * 0.7], np.float32)
    seg, fin, valid = ctx.propagate_pair_batch(0, 1, parents, controls, goals, steps)
    oseg, ofin, ovalid, hold, _ = oracle.pair_propagate_batch(0, 1, parents, controls, goals, steps, n_threads=8)
    clear = hold > 1e-4                      # freeze decisions away from the hold radius
    agree = (valid == ovalid) & clear
    assert (valid[clear] == ovalid[clear]).mean() >= 0.995
    for j in range(10):
        idx = np.nonzero(agree & (valid > j))[0]
        if len(idx) == 0:
            continue
        for c in range(2):
            ok, d = state_close(seg[5 * c:5 * c + 5, j, idx], oseg[5 * c:5 * c + 5, j, idx])
            assert ok.all(), f"step {j} car {c}: {d.max():.3e}"
    idx = np.nonzero(agree)[0]
    for c in range(2):
        ok, d = state_close(fin[5 * c:5 * c + 5, idx], ofin[5 * c:5 * c + 5, idx])
        assert ok.all()
    # the batch exercises freezing, early termination and full segments
    d1 = np.hypot(parents[0] - goals[0], parents[1] - goals[1])
    assert (d1 < 1.5).sum() > 50 and (ovalid < steps).sum() > 50 and (ovalid == steps).sum() > 50
    # frozen cars keep their parent state in every written step
    frozen = np.nonzero((d1 < 1.4) & (valid > 0))[0]
    assert (seg[0:5, 0, frozen] == parents[0:5, frozen]).all()


@pytest.mark.parametrize("scene", ["Corridor10x10_4robots", "Congested10x10_7robots"])
def test_pair_validity(pkg, orc, workloads, gpu_ctx_factory, scene):
    ctx = gpu_ctx_factory(scene)
    oracle = orc.Oracle(pkg.scenes.world(scene))
    n = 1 << 16
    a = workloads.validity_states(scene, n, seed=4)
    b = workloads.validity_states(scene, n, seed=5)
    # a third of the pairs close to each other so that the car-car test matters
    close = np.arange(n) % 3 == 0
    rng = np.random.default_rng(0)
    b[0, close] = a[0, close] + rng.uniform(-1.5, 1.5, close.sum()).astype(np.float32)
    b[1, close] = a[1, close] + rng.uniform(-1.5, 1.5, close.sum()).astype(np.float32)
    states = np.ascontiguousarray(np.concatenate([a, b], axis=0))
    got = ctx.unpack_mask(ctx.validity_pair_batch(0, 1, states), n)
    ref, margin = oracle.pair_validity_batch(0, 1, states, n_threads=8)
    clear = np.abs(margin) > BAND
    assert (got[clear] == ref[clear]).all()
    assert 0.02 < ref.mean() < 0.9


def test_grouped_conflicts(pkg, orc, workloads, gpu_ctx_factory):
    scene = "Empty32x32_10robots"
    ctx = gpu_ctx_factory(scene)
    oracle = orc.Oracle(pkg.scenes.world(scene))
    P, R, T = 4, 10, 700
    traj = workloads.cell_plans(scene, P, T, seed=6)
    group = np.array([0, 1, 2, 3, 4, 5, 6, 7, 8, 8], np.int32)     # bodies 8 and 9 form one merged individual
    workloads.plant_conflict(traj, 0, 8, 9, 50, length=600)         # ignored: same individual
    workloads.plant_conflict(traj, 0, 2, 9, 300, length=3)          # merged individual's second car vs robot 2
    workloads.plant_conflict(traj, 1, 8, 9, 0, length=T)
    workloads.plant_conflict(traj, 2, 0, 8, 10)
    workloads.plant_conflict(traj, 2, 8, 9, 5)
    got = [tuple(int(v) for v in r) for r in ctx.first_conflict_groups(traj, group)]
    for p in range(P):
        ref, _ = oracle.first_conflict_groups(traj[p], group)
        assert got[p] == ref
    assert got[0] == (2, 9, 300, 302) and got[1] == (-1, -1, -1, -1) and got[2] == (0, 8, 10, 10)
    # without groups the same plans report the intra-individual overlaps
    plain = [tuple(int(v) for v in r) for r in ctx.first_conflict(traj)]
    assert plain[0][:3] == (8, 9, 50) and plain[1] == (8, 9, 0, T - 1)
