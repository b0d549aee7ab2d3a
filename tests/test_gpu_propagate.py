"""GPU parity: kcbs_propagate_batch (K1) against the CPU oracle's propagateWhileValid restatement on
identical seeded batches.  Gates (north_star): states within 1e-4 rel / 1e-5 abs over the horizon;
valid-step counts identical except where the deciding state lies within the tolerance band of a
validity boundary."""
import numpy as np
import pytest

from helpers import ATOL, RTOL, state_close

pytestmark = pytest.mark.gpu

SCENES = ["Empty32x32_10robots", "Corridor10x10_4robots", "Congested10x10_7robots"]


def _check(oracle, world_scene, seg, fin, valid, oseg, ofin, ovalid, parents_of, controls, steps):
    n = valid.shape[0]
    agree = valid == ovalid
    # every propagated state both sides consider valid must match within tolerance
    common = np.minimum(valid, ovalid)
    worst = 0.0
    for j in range(seg.shape[1]):
        idx = np.nonzero(common > j)[0]
        if len(idx) == 0:
            continue
        ok, d = state_close(seg[:, j, idx], oseg[:, j, idx])
        worst = max(worst, d.max())
        assert ok.all(), f"step {j}: max |diff| {d.max():.3e} at sample {idx[np.argmax(d.max(axis=0))]}"
    # final state (samples that agree on the count)
    idx = np.nonzero(agree)[0]
    ok, d = state_close(fin[:, idx], ofin[:, idx])
    assert ok.all(), f"final states: max |diff| {d.max():.3e}"
    # count mismatches are only allowed when the oracle's deciding state sits on a boundary
    bad = np.nonzero(~agree)[0]
    assert len(bad) <= 1e-3 * n + 1, f"{len(bad)} of {n} valid-step counts differ"
    for i in bad:
        q = parents_of(i).astype(np.float64)
        u = controls[:, i].astype(np.float64)
        lo = min(valid[i], ovalid[i])
        for _ in range(lo):
            q = oracle.propagate(q, u)
        nxt = oracle.propagate(q, u)
        _, margin = oracle.is_valid(nxt)
        assert abs(margin) < 2e-5 + 1e-4 * 32, f"sample {i}: counts {valid[i]} vs {ovalid[i]}, margin {margin:.3e}"
    return agree.mean(), worst


@pytest.mark.parametrize("scene", SCENES)
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_independent_parents(pkg, orc, workloads, gpu_ctx_factory, scene, seed):
    ctx = gpu_ctx_factory(scene)
    oracle = orc.Oracle(pkg.scenes.world(scene))
    n = 20000
    parents, controls, steps = workloads.expansion_batch(scene, n, seed=seed)
    seg, fin, valid = ctx.propagate_batch(0, parents, controls, steps)
    oseg, ofin, ovalid, _ = oracle.propagate_batch(0, parents, controls, steps, n_threads=8)
    frac, worst = _check(oracle, scene, seg, fin, valid, oseg, ofin, ovalid, lambda i: parents[:, i], controls, steps)
    assert frac >= 0.999
    # states past the valid count are never written (host path zero-fills them)
    for j in range(10):
        assert (seg[:, j, valid <= j] == 0).all()


@pytest.mark.parametrize("scene", SCENES[:2])
def test_shared_parent_is_an_expansion(pkg, orc, workloads, gpu_ctx_factory, scene):
    # one tree node, 65,536 sampled controls (config 'batched propagation 64K controls/expansion')
    ctx = gpu_ctx_factory(scene)
    oracle = orc.Oracle(pkg.scenes.world(scene))
    n = 65536
    parents, controls, steps = workloads.expansion_batch(scene, n, seed=5, shared_parent=True, interior=1.0)
    seg, fin, valid = ctx.propagate_batch(0, parents, controls, steps)
    oseg, ofin, ovalid, _ = oracle.propagate_batch(0, parents, controls, steps, n_threads=8)
    frac, _ = _check(oracle, scene, seg, fin, valid, oseg, ofin, ovalid, lambda i: parents[:, 0], controls, steps)
    assert frac >= 0.999


def test_parent_index_gather(pkg, orc, workloads, gpu_ctx_factory):
    scene = "Empty32x32_10robots"
    ctx = gpu_ctx_factory(scene)
    oracle = orc.Oracle(pkg.scenes.world(scene))
    n, m = 5000, 37
    nodes, _, _ = workloads.expansion_batch(scene, m, seed=9)
    _, controls, steps = workloads.expansion_batch(scene, n, seed=10)
    idx = np.random.default_rng(0).integers(0, m, n).astype(np.int32)
    seg, fin, valid = ctx.propagate_batch(0, nodes, controls, steps, parent_idx=idx)
    oseg, ofin, ovalid, _ = oracle.propagate_batch(0, nodes, controls, steps, parent_idx=idx, n_threads=8)
    frac, _ = _check(oracle, scene, seg, fin, valid, oseg, ofin, ovalid, lambda i: nodes[:, idx[i]], controls, steps)
    assert frac >= 0.999


def test_fixed_horizon_and_default_steps(pkg, orc, workloads, gpu_ctx_factory):
    scene = "Empty32x32_10robots"
    ctx = gpu_ctx_factory(scene)
    oracle = orc.Oracle(pkg.scenes.world(scene))
    parents, controls, _ = workloads.expansion_batch(scene, 4099, seed=4, interior=2.0)   # ragged size
    seg, fin, valid = ctx.propagate_batch(0, parents, controls, None)
    oseg, ofin, ovalid, _ = oracle.propagate_batch(0, parents, controls, None, n_threads=8)
    _check(oracle, scene, seg, fin, valid, oseg, ofin, ovalid, lambda i: parents[:, i], controls, None)
    # final state == last written segment state, or the parent when nothing was valid
    for i in range(0, 4099, 97):
        ref = seg[:, valid[i] - 1, i] if valid[i] > 0 else parents[:, i]
        assert (fin[:, i] == ref).all()


def test_edge_sizes(pkg, workloads, gpu_ctx_factory):
    scene = "Empty32x32_10robots"
    ctx = gpu_ctx_factory(scene)
    # empty batch
    seg, fin, valid = ctx.propagate_batch(0, np.zeros((5, 1), np.float32), np.zeros((2, 0), np.float32))
    assert valid.shape == (0,)
    # one sample, zero steps: nothing propagated, final == parent
    p = np.array([[3.0], [4.0], [0.5], [0.1], [1.0]], np.float32)
    seg, fin, valid = ctx.propagate_batch(0, p, np.array([[0.3], [0.2]], np.float32), np.array([0], np.int32))
    assert valid[0] == 0 and (fin[:, 0] == p[:, 0]).all()
    # max_steps smaller than requested steps clamps
    seg, fin, valid = ctx.propagate_batch(0, p, np.array([[0.3], [0.2]], np.float32), np.array([9], np.int32), max_steps=4)
    assert valid[0] == 4 and seg.shape == (5, 4, 1)


def test_invalid_parent_direction(pkg, orc, gpu_ctx_factory):
    # a sample whose first step leaves the workspace returns 0 valid steps and the parent state
    scene = "Empty32x32_10robots"
    ctx = gpu_ctx_factory(scene)
    p = np.array([[0.01], [5.0], [1.0], [0.0], [-np.pi]], np.float32)  # drives along -x out of bounds
    seg, fin, valid = ctx.propagate_batch(0, p, np.zeros((2, 1), np.float32), np.array([10], np.int32))
    o = orc.Oracle(pkg.scenes.world(scene))
    n, _, _ = o.propagate_while_valid(p[:, 0], [0, 0], 10)
    assert valid[0] == n == 0
    assert (fin[:, 0] == p[:, 0]).all()


def test_parent_headings_at_and_beyond_the_wrap(pkg, orc, gpu_ctx_factory):
    # the reference wraps the heading after every step (SO2 enforceBounds, StatePropagatorDatabase.h:67-74), whatever the
    # parent's heading was: parents at +-pi, just inside, and several turns away must give the oracle's segments
    scene = "Empty32x32_10robots"
    ctx = gpu_ctx_factory(scene)
    o = orc.Oracle(pkg.scenes.world(scene))
    pi32 = np.float32(np.pi)
    headings = np.array([-pi32, np.nextafter(pi32, np.float32(0)), np.nextafter(-pi32, np.float32(0)), 3.0, -3.1, 3.5, -4.0,
                         7.0, -9.5, 12.0, 40.0, -55.5], np.float32)
    rng = np.random.default_rng(7)
    n = len(headings) * 32
    parents = np.empty((5, n), np.float32)
    parents[0], parents[1] = 16.0 + rng.uniform(-3, 3, n), 16.0 + rng.uniform(-3, 3, n)
    parents[2], parents[3] = rng.uniform(-0.9, 0.9, n), rng.uniform(-0.9, 0.9, n)
    parents[4] = np.repeat(headings, 32)
    controls = rng.uniform(-1, 1, (2, n)).astype(np.float32)
    steps = rng.integers(1, 11, n).astype(np.int32)
    seg, fin, valid = ctx.propagate_batch(0, parents, controls, steps)
    for i in range(n):
        nv, oseg, ofin = o.propagate_while_valid(parents[:, i], controls[:, i], int(steps[i]))
        assert nv == valid[i], f"sample {i} (heading {parents[4, i]}): {valid[i]} vs {nv} valid steps"
        for j in range(nv):
            ok, d = state_close(seg[:, j, i:i + 1], np.asarray(oseg[j], np.float64)[:, None])
            assert ok.all(), f"sample {i} (heading {parents[4, i]}) step {j}: {d.ravel()}"
            assert -np.pi <= seg[4, j, i] < np.pi
    # packed layout, same arithmetic
    packed, offsets, pfin, pvalid, total = ctx.propagate_packed(0, parents, controls, steps)
    assert (pvalid == valid).all() and (pfin.view(np.uint32) == fin.view(np.uint32)).all()


def test_errors(pkg, gpu_ctx_factory):
    ctx = gpu_ctx_factory("Empty32x32_10robots")
    p = np.zeros((5, 3), np.float32)
    with pytest.raises(pkg.KcbsError) as e:   # n_parents neither 1 nor n without parent_idx
        ctx.propagate_batch(0, p, np.zeros((2, 7), np.float32))
    assert e.value.status == 1
    with pytest.raises(pkg.KcbsError) as e:   # unknown robot index mirrors unordered_map::at
        ctx.propagate_batch(99, p, np.zeros((2, 3), np.float32))
    assert e.value.status == 5


def test_long_horizon_20_steps(pkg, orc, workloads, gpu_ctx_factory):
    # setMinMaxControlDuration beyond the demos' (1, 10): up to KCBS_MAX_STEPS = 32 steps per segment
    scene = "Empty32x32_10robots"
    ctx = gpu_ctx_factory(scene)
    oracle = orc.Oracle(pkg.scenes.world(scene))
    n = 6000
    parents, controls, _ = workloads.expansion_batch(scene, n, seed=12, interior=3.0)
    controls *= np.float32(0.5)
    steps = np.random.default_rng(5).integers(1, 21, n).astype(np.int32)
    seg, fin, valid = ctx.propagate_batch(0, parents, controls, steps, max_steps=20)
    oseg, ofin, ovalid, _ = oracle.propagate_batch(0, parents, controls, steps, max_steps=20, n_threads=8)
    assert (valid == ovalid).mean() >= 0.998
    for j in range(20):
        idx = np.nonzero(np.minimum(valid, ovalid) > j)[0]
        if len(idx):
            ok, d = state_close(seg[:, j, idx], oseg[:, j, idx])
            assert ok.all(), f"step {j}: {d.max():.3e}"
    assert (valid > 10).any()
