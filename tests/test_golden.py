"""Golden-vector tests.  tests/golden/*.npz were produced by tests/golden/make_golden.py, an
independent pure-numpy restatement of the cited reference semantics (the reference ships no vectors
of its own: PARITY UNPINNED).  CPU part: the C oracle must reproduce them.  GPU part: so must the
CUDA path through the C ABI."""
import os

import numpy as np
import pytest

from helpers import BAND, state_close

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PROP = ["Empty32x32_10robots", "Corridor10x10_4robots", "Congested10x10_7robots"]
VALID = ["Corridor10x10_4robots", "Congested10x10_7robots"]


def load(name):
    return dict(np.load(os.path.join(GOLD, name + ".npz")))


@pytest.mark.parametrize("scene", PROP)
def test_oracle_propagation_matches_golden(pkg, orc, scene):
    g = load("prop_" + scene)
    o = orc.Oracle(pkg.scenes.world(scene))
    seg, fin, valid, _ = o.propagate_batch(0, g["parents"], g["controls"], g["steps"])
    assert (valid == g["valid"]).all()
    assert np.abs(seg - g["seg"]).max() < 1e-11      # two double implementations of the same RK4
    assert np.abs(fin - g["fin"]).max() < 1e-11
    assert 0 < (valid < g["steps"]).sum() or scene.startswith("Empty")


@pytest.mark.parametrize("scene", VALID)
def test_oracle_validity_matches_golden(pkg, orc, scene):
    g = load("valid_" + scene)
    o = orc.Oracle(pkg.scenes.world(scene))
    valid, margin = o.validity_batch(0, g["states"])
    assert (valid == g["valid"]).all()
    inb = g["margin"] > -1e-3          # the oracle stops at the bounds test, numpy keeps the min over obstacles too
    assert np.abs(margin[g["valid"]] - g["margin"][g["valid"]]).max() < 1e-12
    assert inb.any()


def test_oracle_first_conflict_matches_golden(pkg, orc):
    g = load("conflict")
    o = orc.Oracle(pkg.scenes.world("Congested10x10_7robots"))
    res, _ = o.first_conflict_batch(g["traj"], g["lengths"])
    assert [tuple(r) for r in res] == [tuple(r) for r in g["result"]]
    assert (g["min_margin"] > BAND).all()
    assert (g["result"][:, 2] > 0).any()


@pytest.mark.gpu
@pytest.mark.parametrize("scene", PROP)
def test_gpu_propagation_matches_golden(pkg, gpu_ctx_factory, scene):
    g = load("prop_" + scene)
    ctx = gpu_ctx_factory(scene)
    seg, fin, valid = ctx.propagate_batch(0, g["parents"], g["controls"], g["steps"])
    agree = valid == g["valid"]
    assert agree.mean() >= 0.99
    for j in range(10):
        idx = np.nonzero(np.minimum(valid, g["valid"]) > j)[0]
        ok, d = state_close(seg[:, j, idx], g["seg"][:, j, idx])
        assert ok.all(), f"step {j}: {d.max():.3e}"
    ok, d = state_close(fin[:, agree], g["fin"][:, agree])
    assert ok.all()


@pytest.mark.gpu
@pytest.mark.parametrize("scene", VALID)
def test_gpu_validity_matches_golden(pkg, gpu_ctx_factory, scene):
    g = load("valid_" + scene)
    ctx = gpu_ctx_factory(scene)
    got = ctx.unpack_mask(ctx.validity_batch(0, g["states"]), g["states"].shape[1])
    clear = np.abs(g["margin"]) > BAND
    assert (got[clear] == g["valid"][clear]).all()


@pytest.mark.gpu
def test_gpu_first_conflict_matches_golden(pkg, gpu_ctx_factory):
    g = load("conflict")
    ctx = gpu_ctx_factory("Congested10x10_7robots")
    res = ctx.first_conflict(g["traj"], g["lengths"])
    assert [tuple(int(v) for v in r) for r in res] == [tuple(int(v) for v in r) for r in g["result"]]
