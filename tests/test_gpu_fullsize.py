"""GPU checks at BASELINE.json's full sizes, where the CPU oracle would take minutes: size-independent properties of
the three kernels (permutation invariance, idempotence, prefix / composition rules, exact linearity of v and phi,
planted conflicts, shard min-reduction), plus the oracle on a subsample of the very same batch."""
import numpy as np
import pytest

from helpers import state_close

pytestmark = pytest.mark.gpu

N_EXP = 65536          # controls per expansion (configs[1])
N_VALID = 1 << 24
P, R, T = 64, 30, 10000  # config 5: 435 robot pairs x 10K time steps per plan (64 plans = 230 MB)


# ---- K1: Empty32x32, 10 robots, 64K controls per expansion ------------------------------------------------------------
@pytest.fixture(scope="module")
def expansion(workloads, gpu_ctx_factory):
    scene = "Empty32x32_10robots"
    ctx = gpu_ctx_factory(scene)
    parents, controls, steps = workloads.expansion_batch(scene, N_EXP, seed=11, shared_parent=True, interior=1.5)
    seg, fin, valid = ctx.propagate_batch(3, parents, controls, steps)
    return scene, ctx, parents, controls, steps, seg, fin, valid


def test_expansion_rows_and_final_state(expansion):
    _, _, parents, controls, steps, seg, fin, valid = expansion
    assert (valid >= 0).all() and (valid <= steps).all()
    j = np.arange(10)[:, None]
    live = j < valid[None, :]
    assert (seg[:, ~live] == 0).all()                       # rows past the valid count are zero
    # v and phi are exactly linear in time: the kernel's own expression, bit for bit
    t1 = (np.arange(1, 11, dtype=np.float32) * np.float32(0.1))[:, None]
    v = (controls[0][None, :].astype(np.float64) * t1.astype(np.float64) + np.float64(parents[2, 0])).astype(np.float32)
    p = (controls[1][None, :].astype(np.float64) * t1.astype(np.float64) + np.float64(parents[3, 0])).astype(np.float32)
    assert (seg[2][live] == v[live]).all() and (seg[3][live] == p[live]).all()
    # the final state is the last valid row (the parent when nothing is valid)
    last = np.clip(valid - 1, 0, 9)
    got = seg[:, last, np.arange(N_EXP)]
    moved = valid > 0
    assert (fin[:, moved] == got[:, moved]).all()
    assert (fin[:, ~moved] == parents[:, [0]]).all()
    assert ((seg[4][live] >= -np.float32(np.pi)) & (seg[4][live] < np.float32(np.pi))).all()
    assert 0.3 * steps.sum() < valid.sum() <= steps.sum()


def test_expansion_is_deterministic_and_permutation_invariant(expansion):
    _, ctx, parents, controls, steps, seg, fin, valid = expansion
    seg2, fin2, valid2 = ctx.propagate_batch(3, parents, controls, steps)
    assert (seg2 == seg).all() and (fin2 == fin).all() and (valid2 == valid).all()
    # the kernel sorts samples by duration inside every CTA: a permuted batch must give the permuted result exactly
    perm = np.random.default_rng(5).permutation(N_EXP)
    segp, finp, validp = ctx.propagate_batch(3, parents, np.ascontiguousarray(controls[:, perm]), steps[perm])
    assert (validp == valid[perm]).all() and (finp == fin[:, perm]).all() and (segp == seg[:, :, perm]).all()


def test_expansion_prefix_and_broadcast_parent(expansion):
    _, ctx, parents, controls, steps, seg, fin, valid = expansion
    # asking for fewer steps yields a prefix of the longer segment
    short = np.minimum(steps, 4).astype(np.int32)
    seg4, fin4, valid4 = ctx.propagate_batch(3, parents, controls, short)
    assert (valid4 == np.minimum(valid, 4)).all()
    live4 = np.arange(10)[:, None] < valid4[None, :]
    assert (seg4[:, live4] == seg[:, live4]).all()
    # one shared parent == the same parent replicated per sample == gathered through parent_idx
    rep = np.ascontiguousarray(np.repeat(parents, N_EXP, axis=1))
    segr, finr, validr = ctx.propagate_batch(3, rep, controls, steps)
    assert (segr == seg).all() and (validr == valid).all()
    segi, fini, validi = ctx.propagate_batch(3, parents, controls, steps, parent_idx=np.zeros(N_EXP, np.int32))
    assert (segi == seg).all() and (fini == fin).all()


def test_expansion_variants_agree(expansion):
    # the final-state-only call (direct kernel variant, what propagateWhileValid(state, control, steps, result) returns),
    # a batch whose size is not a multiple of 4 (scalar write-out) and a 12-step buffer (direct variant with segments)
    # all reproduce the staged 16 B path bit for bit
    _, ctx, parents, controls, steps, seg, fin, valid = expansion
    _, fin_only, valid_only = ctx.propagate_batch(3, parents, controls, steps, want_seg=False)
    assert (fin_only == fin).all() and (valid_only == valid).all()
    m = N_EXP - 3
    seg_m, fin_m, valid_m = ctx.propagate_batch(3, parents, np.ascontiguousarray(controls[:, :m]), steps[:m])
    assert (seg_m == seg[:, :, :m]).all() and (fin_m == fin[:, :m]).all() and (valid_m == valid[:m]).all()
    seg12, fin12, valid12 = ctx.propagate_batch(3, parents, controls, steps, max_steps=12)
    assert (valid12 == valid).all() and (fin12 == fin).all()
    assert (seg12[:, :10] == seg).all() and (seg12[:, 10:] == 0).all()


def test_expansion_composes(expansion):
    # 10 steps == 5 steps, then 5 more from the state reached (same control): the flow property of the integrator
    scene, ctx, parents, controls, steps, _, _, _ = expansion
    ten = np.full(N_EXP, 10, np.int32)
    five = np.full(N_EXP, 5, np.int32)
    seg10, fin10, valid10 = ctx.propagate_batch(3, parents, controls, ten)
    _, fin5, valid5 = ctx.propagate_batch(3, parents, controls, five)
    _, fin55, valid55 = ctx.propagate_batch(3, np.ascontiguousarray(fin5), controls, five)
    both = (valid10 == 10) & (valid5 == 5) & (valid55 == 5)
    assert both.mean() > 0.2
    ok, d = state_close(fin55[:, both], fin10[:, both].astype(np.float64))
    assert ok.all(), f"max |diff| {d.max():.3e}"


def test_expansion_subsample_against_oracle(expansion, pkg, orc):
    scene, _, parents, controls, steps, seg, fin, valid = expansion
    oracle = orc.Oracle(pkg.scenes.world(scene))
    idx = np.random.default_rng(9).choice(N_EXP, 4096, replace=False)
    oseg, ofin, ovalid, _ = oracle.propagate_batch(3, parents, np.ascontiguousarray(controls[:, idx]), steps[idx], n_threads=8)
    agree = valid[idx] == ovalid
    assert agree.mean() > 0.999
    ok, d = state_close(fin[:, idx][:, agree], ofin[:, agree])
    assert ok.all(), f"max |diff| {d.max():.3e}"


# ---- K2: Congested10x10, 2^24 states ------------------------------------------------------------------------------------
def test_validity_full_size_properties(pkg, orc, workloads, gpu_ctx_factory):
    scene = "Congested10x10_7robots"
    ctx = gpu_ctx_factory(scene)
    states = workloads.validity_states(scene, N_VALID, seed=4)
    mask = ctx.validity_batch(0, states)
    bits = ctx.unpack_mask(mask, N_VALID)
    assert (ctx.validity_batch(0, states) == mask).all()                      # idempotent
    # any slab of the batch gives the same verdicts (odd offsets and lengths: unaligned planes, ragged tails)
    for lo, hi in ((0, 1 << 20), (12345, 12345 + 999_999), (N_VALID - 77_777, N_VALID)):
        part = ctx.unpack_mask(ctx.validity_batch(0, np.ascontiguousarray(states[:, lo:hi])), hi - lo)
        assert (part == bits[lo:hi]).all()
    # permutation invariance on a strided subsample
    idx = np.random.default_rng(2).permutation(N_VALID)[: 1 << 21]
    sub = ctx.unpack_mask(ctx.validity_batch(0, np.ascontiguousarray(states[:, idx])), len(idx))
    assert (sub == bits[idx]).all()
    # a state outside any bound is invalid whatever the rest says
    lo_b, hi_b = pkg.scenes.world(scene).bounds()
    outside = (states[0] < lo_b[0]) | (states[0] > hi_b[0]) | (states[1] < lo_b[1]) | (states[1] > hi_b[1]) | \
              (np.abs(states[2]) > 1.0) | (np.abs(states[3]) > np.float32(np.pi / 3))
    assert not bits[outside].any()
    # the oracle on a subsample of the same batch: bit exact outside the 1e-6 band
    pick = idx[:200_000]
    ref, margin = orc.Oracle(pkg.scenes.world(scene)).validity_batch(0, np.ascontiguousarray(states[:, pick]), n_threads=8)
    clear = np.abs(margin) > 1e-6
    assert (bits[pick][clear] == ref[clear]).all() and clear.mean() > 0.999
    assert 0.1 < bits.mean() < 0.6


# ---- K3: Empty32x32, 30 robots, 435 pairs x 10K steps -----------------------------------------------------------------------
def test_conflict_scan_full_size_properties(pkg, orc, workloads, gpu_ctx_factory):
    scene = "Benchmark_Empty32x32_30robots"
    ctx = gpu_ctx_factory(scene)
    traj = workloads.cell_plans(scene, P, T, seed=3)            # conflict free by construction
    res = ctx.first_conflict(traj)
    assert all(tuple(r)[0] == -1 for r in res)
    # plant conflicts: late one in plan 1, a same-k tie in plan 2 (lexicographically smallest pair wins), an early
    # and a late one in plan 3 (the earlier k wins whatever the pair), a run of 37 steps in plan 4
    workloads.plant_conflict(traj, 1, 7, 22, 9999)
    workloads.plant_conflict(traj, 2, 11, 29, 5000)
    workloads.plant_conflict(traj, 2, 3, 17, 5000)
    workloads.plant_conflict(traj, 3, 28, 29, 128)
    workloads.plant_conflict(traj, 3, 0, 1, 8191)
    workloads.plant_conflict(traj, 4, 5, 6, 4090, length=37)
    res = [tuple(r) for r in ctx.first_conflict(traj)]
    assert res[1] == (7, 22, 9999, 9999)
    assert res[2] == (3, 17, 5000, 5000)
    assert res[3] == (28, 29, 128, 128)
    assert res[4] == (5, 6, 4090, 4126)
    assert all(r[0] == -1 for i, r in enumerate(res) if i not in (1, 2, 3, 4))
    # truncating a plan just before its first conflict makes it conflict free; just after keeps the conflict
    cut = np.ascontiguousarray(traj[3:4, :, :, :128])
    assert tuple(ctx.first_conflict(cut)[0])[0] == -1
    cut = np.ascontiguousarray(traj[3:4, :, :, :132])
    assert tuple(ctx.first_conflict(cut)[0]) == (28, 29, 128, 128)
    # relabelling the robots relabels the conflict
    perm = np.random.default_rng(1).permutation(R)
    inv = np.argsort(perm)
    rel = [tuple(r) for r in ctx.first_conflict(np.ascontiguousarray(traj[1:5][:, perm]))]
    for got, want in zip(rel, res[1:5]):
        a, b = sorted((int(inv[want[0]]), int(inv[want[1]])))
        assert got[2:] == want[2:] and (got[0], got[1]) == (a, b) or want is res[2]
    # the oracle agrees on the planted plans (full 435 x 10K scan of 4 plans)
    ores, _ = orc.Oracle(pkg.scenes.world(scene)).first_conflict_batch(np.ascontiguousarray(traj[1:5]), n_threads=8)
    assert [tuple(r) for r in ores] == res[1:5]
