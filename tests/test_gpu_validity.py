"""GPU parity: kcbs_validity_batch (K2) against the oracle's isValid restatement.  Booleans must be
bit-exact outside the 1e-6 band around a decision boundary (north_star)."""
import numpy as np
import pytest

from helpers import BAND

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("scene,n", [("Congested10x10_7robots", 1 << 20), ("Corridor10x10_4robots", 1 << 18),
                                     ("Empty32x32_10robots", 1 << 18)])
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_states(pkg, orc, workloads, gpu_ctx_factory, scene, n, seed):
    ctx = gpu_ctx_factory(scene)
    oracle = orc.Oracle(pkg.scenes.world(scene))
    states = workloads.validity_states(scene, n, seed=seed)
    got = ctx.unpack_mask(ctx.validity_batch(0, states), n)
    ref, margin = oracle.validity_batch(0, states, n_threads=8)
    clear = np.abs(margin) > BAND
    assert (got[clear] == ref[clear]).all(), f"{(got[clear] != ref[clear]).sum()} booleans differ outside the band"
    assert clear.mean() > 0.999
    assert (got == ref).mean() > 0.9999
    assert 0.05 < ref.mean() < 0.95   # the batch exercises both outcomes


def test_boundary_values_exact(pkg, orc, gpu_ctx_factory):
    # float inputs sitting exactly on / next to every bound get the reference's double verdict
    scene = "Empty32x32_10robots"
    ctx = gpu_ctx_factory(scene)
    oracle = orc.Oracle(pkg.scenes.world(scene))
    lo, hi = pkg.scenes.world(scene).bounds()
    rows = []
    base = np.array([5.0, 5.0, 0.0, 0.0, 0.0], np.float32)
    for d in range(4):
        for b in (lo[d], hi[d]):
            f = np.float32(b)
            for v in (np.nextafter(f, np.float32(-np.inf)), f, np.nextafter(f, np.float32(np.inf))):
                q = base.copy()
                q[d] = v
                rows.append(q)
    for v in (-np.pi, np.pi):
        f = np.float32(v)
        for t in (np.nextafter(f, np.float32(-np.inf)), f, np.nextafter(f, np.float32(np.inf))):
            q = base.copy()
            q[4] = t
            rows.append(q)
    states = np.ascontiguousarray(np.array(rows, np.float32).T)
    got = ctx.unpack_mask(ctx.validity_batch(0, states), states.shape[1])
    ref, _ = oracle.validity_batch(0, states)
    assert (got == ref).all()


def test_footprint_cases(pkg, gpu_ctx_factory):
    # same hand cases as the oracle KAT (rotation by -theta, touching = collision)
    ctx = gpu_ctx_factory("Corridor10x10_4robots")
    q = np.array([[5.0, 2.3, 0, 0, 0], [5.0, 1.0, 0, 0, 0], [5.0, 1.2, 0, 0, np.pi / 4], [5.0, 1.05, 0, 0, np.pi / 4],
                  [5.0, 1.29, 0, 0, 0], [5.0, 1.31, 0, 0, 0]], np.float32).T
    got = ctx.unpack_mask(ctx.validity_batch(0, np.ascontiguousarray(q)), q.shape[1])
    assert got.tolist() == [False, True, False, True, True, False]


@pytest.mark.parametrize("n", [0, 1, 31, 32, 33, 127, 1025])
def test_ragged_sizes(pkg, orc, workloads, gpu_ctx_factory, n):
    scene = "Congested10x10_7robots"
    ctx = gpu_ctx_factory(scene)
    oracle = orc.Oracle(pkg.scenes.world(scene))
    states = workloads.validity_states(scene, max(n, 1), seed=11)[:, :n]
    mask = ctx.validity_batch(0, np.ascontiguousarray(states))
    assert mask.shape[0] == (n + 31) // 32
    if n == 0:
        return
    got = ctx.unpack_mask(mask, n)
    ref, margin = oracle.validity_batch(0, np.ascontiguousarray(states))
    clear = np.abs(margin) > BAND
    assert (got[clear] == ref[clear]).all()
    # bits past n in the last word are zero
    tail = np.unpackbits(mask.view(np.uint8), bitorder="little")[n:]
    assert not tail.any()


def test_per_robot_footprint(pkg, orc, workloads, gpu_ctx_factory):
    # robot_dims are looked up by index (robots_.at(name), StateValidityCheckerDatabase.h:47-48)
    w = pkg.World(x_max=10, y_max=10, robot_dims=[[1.0, 1.0], [0.4, 0.2]],
                  obstacles=pkg.scenes.SCENES["Congested10x10_7robots"]["obstacles"])
    ctx = gpu_ctx_factory(w)
    oracle = orc.Oracle(w)
    states = workloads.validity_states("Congested10x10_7robots", 1 << 16, seed=5)
    res = []
    for r in (0, 1):
        got = ctx.unpack_mask(ctx.validity_batch(r, states), states.shape[1])
        ref, margin = oracle.validity_batch(r, states, n_threads=8)
        clear = np.abs(margin) > BAND
        assert (got[clear] == ref[clear]).all()
        res.append(got)
    assert (res[0] != res[1]).any()


@pytest.mark.parametrize("n", [1 << 18, (1 << 18) + 4, (1 << 18) + 124, 300004, 1000000, (1 << 21) + 64])
def test_vector_and_scalar_load_paths_agree(pkg, workloads, n):
    # large 16-byte aligned batches take the 16 B loads (four states per lane); the same planes at a base that is not
    # 16-byte aligned take scalar loads.  Same verdict bits, also in the last partial pass, and nothing written past them.
    import torch

    scene = "Congested10x10_7robots"
    states = torch.from_numpy(workloads.validity_states(scene, n, seed=n % 97)).cuda()
    with pkg.Context(pkg.scenes.world(scene), device=0) as ctx:
        words = (n + 31) // 32
        shifted = torch.empty(5 * n + 1, dtype=torch.float32, device="cuda")[1:]
        assert shifted.data_ptr() % 16 != 0 and states.data_ptr() % 16 == 0
        shifted.copy_(states.reshape(-1))
        out_a = torch.full((words + 4,), 0x5a5a5a5a, dtype=torch.int32, device="cuda")
        out_b = out_a.clone()
        st = torch.cuda.current_stream()
        for robot in (0, 3):
            ctx.validity_batch_dev(robot, n, states, out_a, st)
            ctx.validity_batch_dev(robot, n, shifted, out_b, st)
            torch.cuda.synchronize()
            assert torch.equal(out_a, out_b)
            assert (out_a[words:] == 0x5a5a5a5a).all()
            bits = ctx.unpack_mask(out_a[:words].cpu().numpy().view(np.uint32), n)
            assert 0.05 < bits.mean() < 0.95
