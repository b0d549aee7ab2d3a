"""Shared helpers of the parity tests (oracle-side generators and comparison rules)."""
import numpy as np

ATOL, RTOL = 1e-5, 1e-4  # north_star: propagated states within 1e-4 relative / 1e-5 absolute
BAND = 1e-6              # north_star: booleans bit-exact except within 1e-6 of a decision boundary


def state_close(got, ref):
    """|got - ref| <= ATOL + RTOL * |ref| per component; heading (row 4) compared on the circle."""
    d = np.abs(got.astype(np.float64) - ref)
    d[4] = np.minimum(d[4], np.abs(2 * np.pi - d[4]))
    return d <= ATOL + RTOL * np.abs(ref), d


def rollout_plan(oracle, starts, T, seed, robot_offset=0):
    """Oracle-rolled plan: every robot starts at its demo start state and is
    driven with piece-wise constant random controls (new control every UniformInt{1..10} steps,
    re-drawn when the segment would leave the valid set).  Returns [R, 3, T] float32."""
    rng = np.random.default_rng(seed)
    R = starts.shape[1]
    traj = np.zeros((R, 3, T), np.float32)
    for r in range(R):
        q = starts[:, r].astype(np.float64)
        k = 0
        traj[r, :, 0] = q[[0, 1, 4]]
        k = 1
        fails = 0
        while k < T:
            u = rng.uniform(-1, 1, 2)
            steps = int(rng.integers(1, 11))
            n, seg, fin = oracle.propagate_while_valid(q, u, min(steps, T - k), robot=r + robot_offset)
            if n == 0:
                # heading into a bound faster than any control can brake: after a few re-draws the
                # robot is stopped in place (plans need not be dynamically feasible for the scan)
                fails += 1
                if fails > 8:
                    q = q.copy()
                    q[2] = 0.0
                    q[3] = 0.0
                continue
            fails = 0
            traj[r, 0, k:k + n] = seg[:, 0]
            traj[r, 1, k:k + n] = seg[:, 1]
            traj[r, 2, k:k + n] = seg[:, 4]
            q = fin
            k += n
    return traj
