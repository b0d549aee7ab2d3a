"""world_size-2 CPU (gloo) tests of the host-side multi-GPU logic: time-slab partitioning of the conflict
scan with the unsigned-min key reduction, plan sharding + all-gather, and replan round-robin.  The per-rank
scan is played by the CPU oracle here (no GPU in this container).  On the GPU box the keys path is exercised by
tests/test_gpu_conflict.py (kcbs_conflict_keys_dev + kcbs_conflict_finish_dev on one GPU) and the production
multi-GPU scan -- keys exchanged inside the kernel over peer memory, kcbs_first_conflict_sharded_dev -- by
tests/test_gpu_sharded_scan.py (2 ranks under torchrun) and bench.py's strong-scaling section."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _slab_keys(oracle, sharding, traj, k_lo, k_hi):
    """Packed first-conflict key of every plan restricted to time indices [k_lo, k_hi) (CPU stand-in for
    kcbs_conflict_keys_dev)."""
    P = traj.shape[0]
    keys = torch.full((P,), sharding.NO_CONFLICT_KEY, dtype=torch.int64)
    if k_hi <= k_lo:
        return keys
    res, _ = oracle.first_conflict_batch(np.ascontiguousarray(traj[..., k_lo:k_hi]))
    for p in range(P):
        if res[p]["r1"] >= 0:
            key = sharding.pack_key(int(res[p]["k_start"]) + k_lo, int(res[p]["r1"]), int(res[p]["r2"]))
            keys[p] = key if key < (1 << 63) else key - (1 << 64)
    return keys


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    import __graft_entry__ as graft
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pkg = graft.load_package()
        orc = graft.load_oracle()
        from kcbs_b200 import sharding, workloads
        scene = "Empty32x32_10robots"
        oracle = orc.Oracle(pkg.scenes.world(scene))
        P, R, T = 6, 10, 700
        traj = workloads.cell_plans(scene, P, T, seed=21)
        workloads.plant_conflict(traj, 0, 2, 3, 50)          # rank 0's slab
        workloads.plant_conflict(traj, 1, 4, 9, 650, 20)     # rank 1's slab
        workloads.plant_conflict(traj, 2, 7, 8, 600)         # both slabs: earlier one must win
        workloads.plant_conflict(traj, 2, 0, 1, 100, 3)
        workloads.plant_conflict(traj, 3, 5, 6, 383, 4)      # run crossing the slab border (384)
        full, _ = oracle.first_conflict_batch(traj)

        # (1) one plan set scanned in time slabs + unsigned-min all-reduce of the keys
        k_lo, k_hi = sharding.time_slab(T, rank, world)
        assert k_lo % sharding.TILE == 0
        keys = _slab_keys(oracle, sharding, traj, k_lo, k_hi)
        sharding.reduce_min_keys(keys)
        got = [sharding.unpack_key(int(k)) for k in keys]
        want = [None if r["r1"] < 0 else (int(r["k_start"]), int(r["r1"]), int(r["r2"])) for r in full]
        assert got == want, (got, want)

        # (2) plans sharded across ranks, results all-gathered
        lo, hi = sharding.shard_range(P, rank, world)
        mine, _ = oracle.first_conflict_batch(np.ascontiguousarray(traj[lo:hi]))
        block = torch.from_numpy(np.stack([mine["r1"], mine["r2"], mine["k_start"], mine["k_end"]], axis=1).astype(np.int32))
        allres = sharding.gather_plans(block, world)
        ref = np.stack([full["r1"], full["r2"], full["k_start"], full["k_end"]], axis=1)
        assert (allres.numpy() == ref).all()

        # (2b) rows owned round-robin (replanned trajectories) merged on every rank
        owned = sharding.round_robin(R, rank, world)
        buf = torch.full((R, 3, T), 123.0)                      # stale values in rows this rank does not own
        for r in owned:
            buf[r] = torch.from_numpy(traj[0, r])
        sharding.merge_owned_rows(buf, owned)
        assert (buf.numpy() == traj[0]).all()

        # (3) replan jobs round-robin: every job exactly once
        jobs = torch.zeros(20, dtype=torch.int64)
        for j in sharding.round_robin(20, rank, world):
            jobs[j] += 1
        dist.all_reduce(jobs)
        assert (jobs == 1).all()

        # (4) the sharded-replans helper end to end with a stand-in for the GPU context: every rank plans its own
        # robots through rrt_plan_many, the merged plan and lengths are complete and identical on both ranks
        class StubCtx:
            calls = []

            def rrt_plan_many(self, robots, starts, goals, params=None, seeds=None, max_T=64):
                self.calls.append(list(robots))
                n = len(robots)
                tr = np.zeros((n, 5, max_T), np.float32)
                T = np.zeros(n, np.int32)
                stats = []
                for j, r in enumerate(robots):
                    T[j] = 5 + r
                    tr[j, :, :T[j]] = starts[:, j:j + 1] + np.arange(T[j], dtype=np.float32)[None, :] * (r + 1) + float(seeds[j] % 7)
                    stats.append(type("S", (), {"solved": 1})())
                return tr, T, stats

        stub = StubCtx()
        starts = np.arange(5 * R, dtype=np.float32).reshape(5, R)
        goals = np.zeros((R, 2), np.float32)
        plan, lengths, ok = sharding.plan_replans_sharded(stub, starts, goals, rank, world, 64, seed=3, device="cpu")
        assert ok and stub.calls == [sharding.round_robin(R, rank, world)]
        assert lengths.tolist() == [5 + r for r in range(R)]
        for r in range(R):
            want = starts[:, r:r + 1] + np.arange(5 + r, dtype=np.float32)[None, :] * (r + 1) + float((3 * 1000003 + r) % 7)
            assert (plan[r, :, :5 + r].numpy() == want).all() and (plan[r, :, 5 + r:] == 0).all()
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_gloo(orc):
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}


def test_partition_helpers(pkg):
    from kcbs_b200 import sharding
    for n, world in ((10, 3), (256, 8), (5, 8), (0, 2)):
        cover = []
        for r in range(world):
            lo, hi = sharding.shard_range(n, r, world)
            cover += list(range(lo, hi))
        assert cover == list(range(n))
    for T in (1, 127, 128, 1000, 10000):
        for world in (1, 2, 8):
            slabs = [sharding.time_slab(T, r, world) for r in range(world)]
            assert slabs[0][0] == 0 and slabs[-1][1] == T
            for (a, b), (c, d) in zip(slabs, slabs[1:]):
                assert b == c and a <= b
    assert sharding.unpack_key(sharding.pack_key(9999, 28, 29)) == (9999, 28, 29)
    assert sharding.unpack_key(-1) is None
    # unsigned order survives the int64 storage: a key with the top bit set is larger than a small key but
    # smaller than "no conflict"
    k = torch.tensor([-1, 5, (1 << 63) - (1 << 64) + 7], dtype=torch.int64)
    assert sharding.reduce_min_keys(k.clone()).tolist() == k.tolist()   # single process: identity
    bias = torch.iinfo(torch.int64).min
    assert int((k ^ bias).min() ^ bias) == 5
