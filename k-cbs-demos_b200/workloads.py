"""Synthetic, seeded workloads of the shapes BASELINE.json names.
Counter-based RNG (numpy Philox) so every rank / test regenerates identical batches."""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

from . import scenes


def _rng(seed: int, stream: int = 0) -> np.random.Generator:
    return np.random.Generator(np.random.Philox(key=[seed, stream]))


def expansion_batch(scene: str, n: int, seed: int = 1, shared_parent: bool = False, fixed_steps: Optional[int] = None,
                    max_steps: int = 10, interior: float = 0.0) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """One tree expansion of n control samples (config 'batched propagation 64K controls/expansion').
    Returns parents [5, 1 or n], controls [2, n], steps [n].  Parent states are uniform over the
    state bounds (StateSpaceDatabase.h:49-59), controls over [-1, 1]^2 (ControlSpaceDatabase.h:48-51),
    steps over {1..max_steps} (demo_*.cpp:118) unless fixed."""
    s = scenes.SCENES[scene]
    g = _rng(seed, 1)
    m = 1 if shared_parent else n
    parents = np.empty((5, m), np.float32)
    parents[0] = g.uniform(interior, s["x_max"] - interior, m)
    parents[1] = g.uniform(interior, s["y_max"] - interior, m)
    parents[2] = g.uniform(-1.0, 1.0, m)
    parents[3] = g.uniform(-np.pi / 3, np.pi / 3, m)
    parents[4] = g.uniform(-np.pi, np.pi, m)
    parents[4] = np.where(parents[4] >= np.float32(np.pi), -np.float32(np.pi), parents[4])
    controls = g.uniform(-1.0, 1.0, (2, n)).astype(np.float32)
    if fixed_steps is None:
        steps = g.integers(1, max_steps + 1, n, dtype=np.int32)
    else:
        steps = np.full(n, fixed_steps, np.int32)
    return parents, controls, steps


def validity_states(scene: str, n: int, seed: int = 1) -> np.ndarray:
    """[5, n] states over a box slightly larger than the bounds so that every rejection reason occurs
    (config 'obstacle-dense validity checks')."""
    s = scenes.SCENES[scene]
    g = _rng(seed, 2)
    st = np.empty((5, n), np.float32)
    st[0] = g.uniform(-0.5, s["x_max"] + 0.5, n)
    st[1] = g.uniform(-0.5, s["y_max"] + 0.5, n)
    st[2] = g.uniform(-1.1, 1.1, n)
    st[3] = g.uniform(-1.15, 1.15, n)
    st[4] = g.uniform(-np.pi, np.pi, n)
    st[4] = np.where(st[4] >= np.float32(np.pi), -np.float32(np.pi), st[4])
    return st


def cell_plans(scene: str, P: int, T: int, seed: int = 1, margin: float = 0.435, xp=np) -> "np.ndarray":
    """[P, R, 3, T] conflict-free synthetic plans for full-scan throughput: robot r wanders smoothly
    inside its own cell of a square grid over the workspace, keeping `margin` from the cell edge.  The
    default margin exceeds the 0.7 x 0.5 robots' bounding radius (0.4301), which guarantees that no two
    footprints can touch; smaller margins give near misses (and occasional real conflicts).
    xp = numpy or torch (torch generates on the current CUDA device)."""
    s = scenes.SCENES[scene]
    R = len(s["robots"])
    side = int(np.ceil(np.sqrt(R)))
    cw, ch = s["x_max"] / side, s["y_max"] / side
    g = _rng(seed, 3)
    par = g.uniform(0.0, 1.0, (P, R, 6)).astype(np.float32)
    cx = ((np.arange(R) % side) + 0.5) * cw
    cy = ((np.arange(R) // side) + 0.5) * ch
    ax, ay = cw / 2 - margin, ch / 2 - margin
    if xp is np:
        k = np.arange(T, dtype=np.float32)[None, None, :]
        wx = (0.002 + 0.004 * par[..., 0])[..., None]
        wy = (0.002 + 0.004 * par[..., 1])[..., None]
        px = (2 * np.pi * par[..., 2])[..., None]
        py = (2 * np.pi * par[..., 3])[..., None]
        x = cx[None, :, None] + ax * np.sin(wx * k + px)
        y = cy[None, :, None] + ay * np.sin(wy * k + py)
        th = np.arctan2(ay * wy * np.cos(wy * k + py), ax * wx * np.cos(wx * k + px))
        th = np.where(th >= np.float32(np.pi), -np.float32(np.pi), th)
        return np.ascontiguousarray(np.stack([x, y, th], axis=2).astype(np.float32))
    torch = xp
    dev = "cuda"
    par_t = torch.from_numpy(par).to(dev)
    k = torch.arange(T, dtype=torch.float32, device=dev)[None, None, :]
    wx = (0.002 + 0.004 * par_t[..., 0])[..., None]
    wy = (0.002 + 0.004 * par_t[..., 1])[..., None]
    px = (2 * np.pi * par_t[..., 2])[..., None]
    py = (2 * np.pi * par_t[..., 3])[..., None]
    cxt = torch.from_numpy(cx.astype(np.float32)).to(dev)[None, :, None]
    cyt = torch.from_numpy(cy.astype(np.float32)).to(dev)[None, :, None]
    out = torch.empty((P, R, 3, T), dtype=torch.float32, device=dev)
    out[:, :, 0] = cxt + ax * torch.sin(wx * k + px)
    out[:, :, 1] = cyt + ay * torch.sin(wy * k + py)
    th = torch.atan2(ay * wy * torch.cos(wy * k + py), ax * wx * torch.cos(wx * k + px))
    out[:, :, 2] = torch.where(th >= float(np.float32(np.pi)), torch.full_like(th, -float(np.float32(np.pi))), th)
    return out


def plant_conflict(traj: np.ndarray, plan: int, r1: int, r2: int, k: int, length: int = 1, dx: float = 0.1) -> None:
    """Makes robots r1 and r2 of `plan` overlap at time indices [k, k + length): r2 is placed dx next to r1."""
    sl = slice(k, k + length)
    traj[plan, r2, 0, sl] = traj[plan, r1, 0, sl] + np.float32(dx)
    traj[plan, r2, 1, sl] = traj[plan, r1, 1, sl]
    traj[plan, r2, 2, sl] = traj[plan, r1, 2, sl]
