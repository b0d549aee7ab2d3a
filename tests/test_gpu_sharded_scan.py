"""GPU parity of the sharded plan scan (kcbs_first_conflict_sharded_dev, SURVEY.md 8e / BASELINE config 5): every rank
scans a time slab, the per-plan keys are exchanged inside the kernel through mailboxes in peer memory, and every rank
must end up with exactly the records of the single-GPU scan (= the oracle's findConflicts restatement).

  * ranks as contexts of ONE process on cuda:0 (the protocol, epochs, empty slabs, ragged lengths) -- runs on one GPU;
  * ranks as PROCESSES under torchrun (CUDA IPC mailboxes; one GPU per rank when the box has them, else both ranks
    time-slice cuda:0) -- tests/mp_sharded_scan.py."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _t(res):
    return [tuple(int(v) for v in r) for r in res]


def _connect_local(ctxs, max_plans):
    handles = [c.peer_export(max_plans) for c in ctxs]
    for r, c in enumerate(ctxs):
        c.peer_connect(r, len(ctxs), handles)


@pytest.mark.parametrize("world", [2, 3, 4])
def test_contexts_of_one_process(pkg, orc, workloads, world):
    import torch
    scene = "Benchmark_Empty32x32_30robots"
    oracle = orc.Oracle(pkg.scenes.world(scene))
    P, R, T = 12, 30, 1000
    traj = workloads.cell_plans(scene, P, T, seed=3)
    workloads.plant_conflict(traj, 1, 4, 9, 500, length=7)        # middle slab
    workloads.plant_conflict(traj, 2, 20, 29, 640)                # same k, two pairs: (3, 17) wins
    workloads.plant_conflict(traj, 2, 3, 17, 640, length=3)
    workloads.plant_conflict(traj, 3, 28, 29, 0, length=2)        # first slab
    workloads.plant_conflict(traj, 4, 0, 1, T - 1)                # last slab, last index
    workloads.plant_conflict(traj, 5, 10, 11, 100, length=800)    # extension runs through every later slab
    workloads.plant_conflict(traj, 6, 5, 6, 900)                  # a later conflict of a smaller pair loses
    workloads.plant_conflict(traj, 6, 7, 8, 130)
    workloads.plant_conflict(traj, 7, 1, 2, 127)                  # two slabs report, the earlier one wins
    workloads.plant_conflict(traj, 7, 2, 3, 600)
    ref, _ = oracle.first_conflict_batch(traj, n_threads=8)
    ctxs = [pkg.Context(pkg.scenes.world(scene), device=0) for _ in range(world)]
    try:
        _connect_local(ctxs, 64)
        d = torch.from_numpy(traj).cuda()
        streams = [torch.cuda.Stream() for _ in range(world)]
        outs = [torch.full((P, 4), 7, dtype=torch.int32, device="cuda") for _ in range(world)]
        # ranks that share ONE process must not meet a device-wide synchronisation while a peer's scan is waiting for
        # them: the key workspace of every context is sized up front by a plain scan
        for r in range(world):
            ctxs[r].first_conflict_dev(P, R, T, d, None, outs[r], streams[r])
        torch.cuda.synchronize()
        for rep in range(5):    # consecutive calls: epochs advance, mailbox parities alternate
            order = range(world) if rep % 2 == 0 else reversed(range(world))
            for r in order:
                ctxs[r].first_conflict_sharded_dev(P, R, T, d, None, outs[r], streams[r])
            torch.cuda.synchronize()
            for r in range(world):
                ctxs[r].peer_status()
                assert _t(outs[r].cpu().numpy()) == _t(ref), f"rank {r} of {world}, call {rep}"
                outs[r].fill_(7)
        # ragged lengths and a horizon with fewer tiles than ranks (empty slabs still take part in the exchange)
        for T2 in (1, 100, 130, 257):
            t2 = np.ascontiguousarray(traj[:, :, :, :T2])
            rng = np.random.default_rng(T2)
            lengths = rng.integers(1, T2 + 1, size=(P, R)).astype(np.int32)
            want, _ = oracle.first_conflict_batch(t2, lengths=lengths, n_threads=8)
            d2, dl = torch.from_numpy(t2).cuda(), torch.from_numpy(lengths).cuda()
            for r in range(world):
                ctxs[r].first_conflict_sharded_dev(P, R, T2, d2, dl, outs[r], streams[r])
            torch.cuda.synchronize()
            for r in range(world):
                ctxs[r].peer_status()
                assert _t(outs[r].cpu().numpy()) == _t(want), f"T = {T2}, rank {r} of {world}"
        # the plain single-context call still works on a connected context (and leaves keys / tickets clean)
        ctxs[0].first_conflict_dev(P, R, T, d, None, outs[0], streams[0])
        torch.cuda.synchronize()
        assert _t(outs[0].cpu().numpy()) == _t(ref)
    finally:
        for c in ctxs:
            c.close()


def test_missing_peer_times_out_instead_of_hanging(pkg, workloads):
    # rank 1 never calls: rank 0's waits give up after KCBS_PEER_TIMEOUT_S, the records say so, the status call raises
    import torch
    from kcbs_b200.capi import KcbsError
    scene = "Benchmark_Empty32x32_30robots"
    ctxs = [pkg.Context(pkg.scenes.world(scene), device=0) for _ in range(2)]
    try:
        _connect_local(ctxs, 8)
        traj = torch.from_numpy(workloads.cell_plans(scene, 2, 256, seed=3)).cuda()
        out = torch.zeros((2, 4), dtype=torch.int32, device="cuda")
        ctxs[0].first_conflict_sharded_dev(2, 30, 256, traj, None, out, torch.cuda.current_stream())
        torch.cuda.synchronize()
        assert _t(out.cpu().numpy()) == [(-2, -2, -2, -2)] * 2
        with pytest.raises(KcbsError) as e:
            ctxs[0].peer_status()
        assert e.value.status == 6
        ctxs[0].peer_status()   # the flag is cleared by reporting it
    finally:
        for c in ctxs:
            c.close()


def test_time_slabs_partition_the_horizon(pkg):
    lib_ctx = pkg.Context(pkg.scenes.world("Empty32x32_10robots"), device=0)
    try:
        for T in (0, 1, 127, 128, 129, 1000, 10000, 12345):
            for world in (1, 2, 3, 8, 16):
                slabs = [lib_ctx.time_slab(T, r, world) for r in range(world)]
                assert slabs[0][0] == 0 and slabs[-1][1] == T
                for (a, b), (c, d) in zip(slabs, slabs[1:]):
                    assert b == c and a <= b and a % 128 == 0
                sizes = [b - a for a, b in slabs]
                assert max(sizes) - min(sizes) <= 128 + 127
    finally:
        lib_ctx.close()


@pytest.mark.parametrize("world", [2])
def test_ranks_as_processes_under_torchrun(world):
    import torch
    port = 29500 + (os.getpid() % 400)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "mp_sharded_scan.py")]
    env = dict(os.environ)
    env.setdefault("OMP_NUM_THREADS", "2")
    p = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, f"stdout:\n{p.stdout[-3000:]}\nstderr:\n{p.stderr[-3000:]}"
    assert p.stdout.count("sharded scan ok") == world, p.stdout[-2000:]
    assert torch.cuda.is_available()
