#!/usr/bin/env python
"""Generates tests/golden/*.npz: small seeded input/output vectors for the three parts of the hot
path, computed by an INDEPENDENT pure-numpy restatement (vectorised RK4, numpy SAT, Python scan
loop) -- not by the C oracle and not by the GPU.  The reference itself ships no golden vectors and
cannot run here (PARITY UNPINNED), so these fixtures pin the C oracle and the CUDA path to a second
implementation of the same cited semantics and guard both against regressions.

Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SCENES = json.load(open(os.path.join(ROOT, "k-cbs-demos_b200", "scenes.json")))
EPS = np.finfo(np.float64).eps


def rhs(q, u, L=0.5):
    # StatePropagatorDatabase.h:59-63
    return np.stack([q[2] * np.cos(q[4]), q[2] * np.sin(q[4]), u[0], u[1], q[2] / L * np.tan(q[3])])


def propagate(q, u, duration=0.1, h=0.01):
    n = int(round(duration / h))
    for _ in range(n):
        k1 = rhs(q, u)
        k2 = rhs(q + 0.5 * h * k1, u)
        k3 = rhs(q + 0.5 * h * k2, u)
        k4 = rhs(q + h * k3, u)
        q = q + h / 6 * k1 + h / 3 * k2 + h / 3 * k3 + h / 6 * k4
    th = np.fmod(q[4], 2 * np.pi)
    th = np.where(th < -np.pi, th + 2 * np.pi, np.where(th >= np.pi, th - 2 * np.pi, th))
    q = q.copy()
    q[4] = th
    return q


def sat_gap(ca, ta, ha, cb, tb, hb):
    """rectangles: centres c [2, n], headings t [n] (footprint rotated by -t), half extents h (hl, hw)."""
    def axes(t):
        c, s = np.cos(t), np.sin(t)
        return np.stack([c, -s]), np.stack([s, c])
    a1, a2 = axes(ta)
    b1, b2 = axes(tb)
    d = cb - ca
    best = np.full(d.shape[1], -np.inf)
    for n in (a1, a2, b1, b2):
        ra = ha[0] * np.abs((a1 * n).sum(0)) + ha[1] * np.abs((a2 * n).sum(0))
        rb = hb[0] * np.abs((b1 * n).sum(0)) + hb[1] * np.abs((b2 * n).sum(0))
        best = np.maximum(best, np.abs((d * n).sum(0)) - (ra + rb))
    return best


def is_valid(scene, q):
    s = SCENES[scene]
    lo = np.array([0, 0, -1, -np.pi / 3])[:, None]
    hi = np.array([s["x_max"], s["y_max"], 1, np.pi / 3])[:, None]
    ok = ((q[:4] - EPS <= hi) & (q[:4] + EPS >= lo)).all(0)
    ok &= (q[4] < np.pi + EPS) & (q[4] > -np.pi - EPS)
    margin = np.minimum((hi - q[:4]).min(0), (q[:4] - lo).min(0))
    margin = np.minimum(margin, np.minimum(np.pi - q[4], q[4] + np.pi))
    h = (s["robot_length"] / 2, s["robot_width"] / 2)
    for o in s["obstacles"]:
        cb = np.array([[o[0] + o[2] / 2], [o[1] + o[3] / 2]]) * np.ones((1, q.shape[1]))
        g = sat_gap(q[:2], q[4], h, cb, np.zeros(q.shape[1]), (o[2] / 2, o[3] / 2))
        ok &= g > 0
        margin = np.minimum(margin, g)
    return ok, margin


def propagate_while_valid(scene, q0, u, steps, max_steps=10):
    n = q0.shape[1]
    seg = np.zeros((5, max_steps, n))
    fin = q0.copy()
    valid = np.zeros(n, np.int32)
    alive = np.ones(n, bool)
    cur = q0.copy()
    for j in range(max_steps):
        nxt = propagate(cur, u)
        ok, _ = is_valid(scene, nxt)
        alive &= ok & (steps > j)
        seg[:, j, alive] = nxt[:, alive]
        fin[:, alive] = nxt[:, alive]
        valid[alive] = j + 1
        cur = np.where(alive[None], nxt, cur)
    return seg, fin, valid


def first_conflict(scene, traj, lengths=None):
    s = SCENES[scene]
    R, _, T = traj.shape
    h = (s["robot_length"] / 2, s["robot_width"] / 2)
    tr = traj.astype(np.float64).copy()
    if lengths is not None:
        for r in range(R):
            tr[r, :, lengths[r]:] = tr[r, :, lengths[r] - 1][:, None]
    min_abs = np.inf
    for k in range(T):
        for r1 in range(R):
            for r2 in range(r1 + 1, R):
                g = sat_gap(tr[r1, :2, k:k + 1], tr[r1, 2, k:k + 1], h, tr[r2, :2, k:k + 1], tr[r2, 2, k:k + 1], h)[0]
                min_abs = min(min_abs, abs(g))
                if g <= 0:
                    ke = k
                    while ke + 1 < T and sat_gap(tr[r1, :2, ke + 1:ke + 2], tr[r1, 2, ke + 1:ke + 2], h,
                                                 tr[r2, :2, ke + 1:ke + 2], tr[r2, 2, ke + 1:ke + 2], h)[0] <= 0:
                        ke += 1
                    return (r1, r2, k, ke), min_abs
    return (-1, -1, -1, -1), min_abs


def main():
    rng = np.random.Generator(np.random.Philox(key=[2026, 0]))
    out = {}
    # (1) expansion batches
    for scene in ("Empty32x32_10robots", "Corridor10x10_4robots", "Congested10x10_7robots"):
        s = SCENES[scene]
        n = 512
        q0 = np.stack([rng.uniform(0.3, s["x_max"] - 0.3, n), rng.uniform(0.3, s["y_max"] - 0.3, n), rng.uniform(-1, 1, n),
                       rng.uniform(-1, 1, n), rng.uniform(-np.pi, np.pi, n)]).astype(np.float32)
        u = rng.uniform(-1, 1, (2, n)).astype(np.float32)
        steps = rng.integers(1, 11, n).astype(np.int32)
        seg, fin, valid = propagate_while_valid(scene, q0.astype(np.float64), u.astype(np.float64), steps)
        out[f"prop_{scene}"] = dict(parents=q0, controls=u, steps=steps, seg=seg, fin=fin, valid=valid)
    # (2) validity
    for scene in ("Corridor10x10_4robots", "Congested10x10_7robots"):
        s = SCENES[scene]
        n = 4096
        q = np.stack([rng.uniform(-0.5, s["x_max"] + 0.5, n), rng.uniform(-0.5, s["y_max"] + 0.5, n), rng.uniform(-1.1, 1.1, n),
                      rng.uniform(-1.15, 1.15, n), rng.uniform(-np.pi, np.pi, n)]).astype(np.float32)
        ok, margin = is_valid(scene, q.astype(np.float64))
        out[f"valid_{scene}"] = dict(states=q, valid=ok, margin=margin)
    # (3) first conflict: smooth random walks of 7 robots in a 10x10 box, several plans, ragged lengths
    scene = "Congested10x10_7robots"
    R, T, P = 7, 160, 6
    trajs, res, lens, margins = [], [], [], []
    for p in range(P):
        lattice = np.array([[2, 2], [5, 2], [8, 2], [2, 5], [5, 5], [8, 5], [5, 8]], np.float64)
        pos = lattice + rng.uniform(-0.3, 0.3, (R, 2))
        vel = rng.uniform(-0.08, 0.08, (R, 2))
        tr = np.zeros((R, 3, T), np.float32)
        for k in range(T):
            vel += rng.normal(0, 0.01, (R, 2))
            pos = np.clip(pos + vel, 0.6, 9.4)
            tr[:, 0, k], tr[:, 1, k] = pos[:, 0], pos[:, 1]
            tr[:, 2, k] = np.arctan2(vel[:, 1], vel[:, 0])
        ln = rng.integers(T // 2, T + 1, R).astype(np.int32) if p % 2 else np.full(R, T, np.int32)
        r, m = first_conflict(scene, tr, ln)
        trajs.append(tr), res.append(r), lens.append(ln), margins.append(m)
    out["conflict"] = dict(traj=np.stack(trajs), lengths=np.stack(lens), result=np.array(res, np.int32),
                           min_margin=np.array(margins))
    for name, d in out.items():
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
        print(name, {k: v.shape for k, v in d.items()})
    print("conflicts:", out["conflict"]["result"].tolist(), out["conflict"]["min_margin"])


if __name__ == "__main__":
    main()
