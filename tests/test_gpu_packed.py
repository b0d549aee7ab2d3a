"""GPU tests of the packed (CSR) segment output of the expansion kernel (kcbs_propagate_packed[_dev]) and of the launch
shapes (kcbs_set_expansion_mode).  The dense output is gated against the oracle in test_gpu_propagate.py /
test_gpu_fullsize.py; here the packed rows must equal the dense rows BIT FOR BIT (same kernel arithmetic, other
layout) for every size, parent layout and launch shape.  A sample owns as many rows as it can have valid states at all
(its effective duration), starting at offsets[i]; the rows it did not reach are zero.  Blocks of samples are placed
first come, first served, so the offsets are not monotone across blocks: the runs must be disjoint and tight."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def dense_rows(seg, valid):
    """[5, S, n] dense segments -> [5, total] rows in (sample, step) order."""
    S, n = seg.shape[1], seg.shape[2]
    live = np.arange(S)[None, :] < valid[:, None]            # [n, S]
    return np.ascontiguousarray(seg.transpose(0, 2, 1)[:, live])


def packed_rows(packed, offsets, valid):
    """The valid rows of a packed result in (sample, step) order, and a mask of the rows that hold a state."""
    n = len(valid)
    assert (offsets[:n] >= 0).all()
    live = np.zeros(packed.shape[1], bool)
    idx = np.repeat(offsets[:n], valid) + (np.arange(valid.sum()) - np.repeat(np.cumsum(valid) - valid, valid))
    live[idx] = True
    return np.ascontiguousarray(packed[:, idx]), live


def check_packed(ctx, robot, parents, controls, steps, parent_idx=None, max_steps=10):
    seg, fin, valid = ctx.propagate_batch(robot, parents, controls, steps, parent_idx=parent_idx, max_steps=max_steps)
    packed, offsets, pfin, pvalid, total = ctx.propagate_packed(robot, parents, controls, steps, parent_idx=parent_idx,
                                                                max_steps=max_steps)
    n = controls.shape[1]
    assert total == offsets[n] and (pvalid == valid).all() and (pfin == fin).all()
    assert total == 0 or offsets[:n].min() == 0
    # sample i owns a run of rows from offsets[i] that holds its valid states: the runs of different samples are disjoint,
    # and all rows together are no more than the requested durations (+3 per block of samples: 16-byte alignment)
    req = np.full(n, max_steps) if steps is None else np.clip(steps, 0, max_steps)
    moved = np.flatnonzero(valid > 0)
    order = moved[np.argsort(offsets[moved], kind="stable")]
    assert (np.diff(np.append(offsets[order], total)) >= valid[order]).all()
    assert total <= req.sum() + 4 * (n // 256 + 2)
    want = dense_rows(seg, valid)
    got, live = packed_rows(packed, offsets, valid)
    assert (got.view(np.uint32) == want.view(np.uint32)).all(), "packed rows differ from the dense segments"
    assert (packed[:, ~live] == 0).all(), "rows that hold no state are not zero"
    return total


@pytest.mark.parametrize("scene", ["Empty32x32_10robots", "Congested10x10_7robots", "Corridor10x10_4robots"])
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_packed_equals_dense(pkg, workloads, scene, mode):
    with pkg.Context(pkg.scenes.world(scene), device=0) as ctx:
        ctx.set_expansion_mode(mode)
        for seed, n in ((1, 4096), (2, 20000), (3, 65536)):
            parents, controls, steps = workloads.expansion_batch(scene, n, seed=seed, interior=0.3)
            assert check_packed(ctx, 0, parents, controls, steps) > 0
        parents, controls, steps = workloads.expansion_batch(scene, 65536, seed=4, shared_parent=True, interior=1.5)
        check_packed(ctx, 1, parents, controls, steps)


@pytest.mark.parametrize("n", [1, 2, 7, 255, 256, 257, 511, 512, 513, 2047, 2049, 8191, 8193, 33000, 200000, 1 << 20])
def test_packed_sizes(pkg, workloads, n):
    # CTA boundaries of both launch shapes, pipeline chunk boundaries, grids of several waves
    scene = "Empty32x32_10robots"
    with pkg.Context(pkg.scenes.world(scene), device=0) as ctx:
        parents, controls, steps = workloads.expansion_batch(scene, n, seed=n, interior=0.2)
        check_packed(ctx, 0, parents, controls, steps)
        if n <= 33000:
            ctx.set_expansion_mode(2)
            check_packed(ctx, 0, parents, controls, steps)


def test_packed_parent_gather_and_short_horizon(pkg, workloads):
    scene = "Congested10x10_7robots"
    rng = np.random.default_rng(5)
    with pkg.Context(pkg.scenes.world(scene), device=0) as ctx:
        nodes, _, _ = workloads.expansion_batch(scene, 300, seed=9, interior=0.5)
        n = 30000
        _, controls, steps = workloads.expansion_batch(scene, n, seed=10, interior=0.5)
        idx = rng.integers(0, 300, n).astype(np.int32)
        check_packed(ctx, 2, nodes, controls, steps, parent_idx=idx)
        check_packed(ctx, 2, nodes, controls, np.minimum(steps, 4), parent_idx=idx, max_steps=4)
        check_packed(ctx, 2, nodes, controls, None, parent_idx=idx, max_steps=10)       # steps = NULL: max_steps everywhere
        zero = np.zeros(n, np.int32)
        packed, offsets, fin, valid, total = ctx.propagate_packed(2, nodes, controls, zero, parent_idx=idx)
        assert total == 0 and (offsets == 0).all() and (valid == 0).all() and (fin == nodes[:, idx]).all()


def test_packed_dev_repeated_launches_and_small_capacity(pkg, workloads):
    import torch
    scene = "Empty32x32_10robots"
    n = 100000
    with pkg.Context(pkg.scenes.world(scene), device=0) as ctx:
        parents, controls, steps = workloads.expansion_batch(scene, n, seed=21, interior=0.2)
        seg, fin, valid = ctx.propagate_batch(0, parents, controls, steps)
        want = dense_rows(seg, valid)
        total = int(valid.sum())
        dp, dc, ds = (torch.from_numpy(x).cuda() for x in (parents, controls, steps))
        st = torch.cuda.current_stream()
        ref_off = None
        for cap in (n * 10, None, -1001):
            cap = cap if cap and cap > 0 else int(ref_off[n]) + (cap or 0)      # roomy, exactly the rows, 1001 rows short
            packed = torch.full((5, cap), -7.0, device="cuda")
            offsets = torch.full((n + 1,), -1, dtype=torch.int32, device="cuda")
            dfin = torch.empty((5, n), device="cuda")
            dval = torch.empty(n, dtype=torch.int32, device="cuda")
            for rep in range(3):    # the row allocator must come back clean from every launch
                ctx.propagate_packed_dev(0, n, n, dp, dc, ds, None, 10, packed, cap, offsets, dfin, dval, st)
            torch.cuda.synchronize()
            off = offsets.cpu().numpy()
            if ref_off is None:
                ref_off = off
            assert off[n] == ref_off[n]     # the rows a launch needs never depend on the capacity
            got = packed.cpu().numpy()
            fits = off[:n] + valid <= cap       # samples whose rows all fit
            rows, live = packed_rows(got, off, np.where(fits, valid, 0))
            sel = np.repeat(fits, valid)
            assert (rows.view(np.uint32) == want[:, sel].view(np.uint32)).all()
            used = np.zeros(cap, bool)
            used[:min(cap, int(off[n]))] = True
            assert (got[:, ~used] == -7.0).all()      # nothing past the emitted rows, nothing past the capacity
            assert (dval.cpu().numpy() == valid).all() and (dfin.cpu().numpy() == fin).all()
        total = int(ref_off[n])
        # graph replay of packed launches (the row allocator is reset on the device)
        packed = torch.zeros((5, n * 10), device="cuda")
        offsets = torch.zeros(n + 1, dtype=torch.int32, device="cuda")
        side = torch.cuda.Stream()
        side.wait_stream(st)
        with torch.cuda.stream(side):
            ctx.propagate_packed_dev(0, n, n, dp, dc, ds, None, 10, packed, n * 10, offsets, None, None, side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                for _ in range(4):
                    ctx.propagate_packed_dev(0, n, n, dp, dc, ds, None, 10, packed, n * 10, offsets, None, None, side)
            packed.zero_()
            g.replay()
            g.replay()
            side.synchronize()
        rows, _ = packed_rows(packed.cpu().numpy(), offsets.cpu().numpy(), valid)
        assert (rows.view(np.uint32) == want.view(np.uint32)).all()
        assert int(offsets[n].item()) == total


def test_launch_shapes_are_bit_identical(pkg, workloads):
    # one packed pair per thread (latency shape) and two (throughput shape) integrate every sample with the same
    # arithmetic: dense segments, final states and counts agree bit for bit, also on the direct (final-state-only) path
    scene = "Congested10x10_7robots"
    out = []
    for mode in (1, 2, 0):
        with pkg.Context(pkg.scenes.world(scene), device=0) as ctx:
            ctx.set_expansion_mode(mode)
            parents, controls, steps = workloads.expansion_batch(scene, 50000, seed=33, interior=0.3)
            seg, fin, valid = ctx.propagate_batch(3, parents, controls, steps)
            _, fin2, valid2 = ctx.propagate_batch(3, parents, controls, steps, want_seg=False)
            seg20, fin20, valid20 = ctx.propagate_batch(3, parents, controls, steps * 2, max_steps=20)
            out.append((seg, fin, valid, fin2, valid2, seg20, fin20, valid20))
    for other in out[1:]:
        for a, b in zip(out[0], other):
            assert (a.view(np.uint32) == b.view(np.uint32)).all()
    assert (out[0][1] == out[0][3]).all() and (out[0][2] == out[0][4]).all()


def test_packed_argument_errors(pkg, workloads):
    from kcbs_b200.capi import KcbsError
    scene = "Empty32x32_10robots"
    with pkg.Context(pkg.scenes.world(scene), device=0) as ctx:
        parents, controls, steps = workloads.expansion_batch(scene, 64, seed=1)
        with pytest.raises(KcbsError):
            ctx.propagate_packed(0, parents, controls, steps, max_steps=11)
        with pytest.raises(KcbsError):
            ctx.set_expansion_mode(3)
        with pytest.raises(KcbsError):
            ctx.propagate_packed(99, parents, controls, steps)
