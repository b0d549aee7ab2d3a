"""CPU check of the three-valued obstacle grid (kcbs_scene_grid, host code of the product library) against the
oracle's isValid restatement: a state in a "free" cell is never rejected because of an obstacle, a state in a "hit"
cell always is, whatever its heading; "maybe" cells name every obstacle that can decide."""
import numpy as np
import pytest

SCENES = ["Congested10x10_7robots", "Corridor10x10_4robots", "Benchmark_Corridor10x10_4robots"]


def _cells(world, dim, x, y):
    lo, hi = world.bounds()
    # the device lookup (kcbs_device.cuh, fine_cell) in float
    fx = x.astype(np.float32) * np.float32(dim / (hi[0] - lo[0])) + np.float32(-lo[0] * dim / (hi[0] - lo[0]))
    fy = y.astype(np.float32) * np.float32(dim / (hi[1] - lo[1])) + np.float32(-lo[1] * dim / (hi[1] - lo[1]))
    ix = np.clip(fx.astype(np.int64), 0, dim - 1)
    iy = np.clip(fy.astype(np.int64), 0, dim - 1)
    return iy, ix


@pytest.mark.parametrize("scene", SCENES)
def test_free_and_hit_cells_agree_with_the_oracle(pkg, orc, scene):
    from kcbs_b200 import capi
    world = pkg.scenes.world(scene)
    codes = capi.scene_grid(world)
    dim = codes.shape[0]
    assert codes.shape == (dim, dim) and dim >= 32
    lo, hi = world.bounds()
    g = np.random.Generator(np.random.Philox(key=[5, 7]))
    n = 400_000
    st = np.zeros((5, n), np.float32)
    st[0] = g.uniform(lo[0], hi[0], n)
    st[1] = g.uniform(lo[1], hi[1], n)
    st[4] = g.uniform(-np.pi, np.pi, n)
    st[4] = np.where(st[4] >= np.float32(np.pi), -np.float32(np.pi), st[4])
    # cell borders and obstacle edges are where a wrong margin would show: snap a third of the samples next to them
    k = n // 3
    cw = (hi[0] - lo[0]) / dim
    st[0, :k] = (np.round(st[0, :k] / cw) * cw + g.uniform(-2e-5, 2e-5, k)).astype(np.float32)
    st[1, k:2 * k] = (np.round(st[1, k:2 * k] / cw) * cw + g.uniform(-2e-5, 2e-5, k)).astype(np.float32)
    st[0] = np.clip(st[0], lo[0], hi[0])
    st[1] = np.clip(st[1], lo[1], hi[1])
    valid, _ = orc.Oracle(world).validity_batch(0, st, want_margin=False, n_threads=8)
    iy, ix = _cells(world, dim, st[0], st[1])
    code = codes[iy, ix]
    free, hit = code == 0, code == 0xFFFF
    assert valid[free].all(), "a state in a free cell was rejected"
    assert not valid[hit].any(), "a state in a hit cell was accepted"
    maybe = ~free & ~hit
    assert 0.02 < maybe.mean() < 0.5 and free.mean() > 0.2 and hit.mean() > 0.02
    assert valid[maybe].any() and not valid[maybe].all()   # the band really is undecided


def test_maybe_cells_list_every_deciding_obstacle(pkg, orc):
    # dropping the obstacles a cell does not list must not change the oracle's verdict for states in that cell
    from kcbs_b200 import capi
    scene = "Congested10x10_7robots"
    world = pkg.scenes.world(scene)
    codes = capi.scene_grid(world)
    dim = codes.shape[0]
    lo, hi = world.bounds()
    g = np.random.Generator(np.random.Philox(key=[6, 7]))
    n = 200_000
    st = np.zeros((5, n), np.float32)
    st[0] = g.uniform(lo[0], hi[0], n)
    st[1] = g.uniform(lo[1], hi[1], n)
    st[4] = g.uniform(-np.pi, np.pi, n)
    st[4] = np.where(st[4] >= np.float32(np.pi), -np.float32(np.pi), st[4])
    iy, ix = _cells(world, dim, st[0], st[1])
    code = codes[iy, ix].astype(np.int64)
    full, _ = orc.Oracle(world).validity_batch(0, st, want_margin=False, n_threads=8)
    obstacles = [list(o) for o in world.obstacles]
    import dataclasses
    listed = np.ones(n, bool)
    for o in range(len(obstacles)):
        only = dataclasses.replace(world, obstacles=[obstacles[o]])
        v_o, _ = orc.Oracle(only).validity_batch(0, st, want_margin=False, n_threads=8)
        names_o = (code == 0xFFFE) | (code == 0xFFFF) | ((code & 0xFF) == o + 1) | ((code >> 8) == o + 1)
        # obstacle o rejects the state => the cell must name it (or be hit / many)
        assert names_o[~v_o].all(), f"obstacle {o} decides a state whose cell does not list it"
        listed &= v_o | names_o
    assert listed.all()
    assert ((code != 0) & (code < 0xFFFE) & (code >> 8 != 0)).any()   # two-candidate cells occur


def test_empty_scene_grid_is_all_free(pkg):
    from kcbs_b200 import capi
    codes = capi.scene_grid(pkg.scenes.world("Empty32x32_10robots"))
    assert not codes.any()
