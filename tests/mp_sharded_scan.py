"""Worker of tests/test_gpu_sharded_scan.py::test_ranks_as_processes_under_torchrun (launched by torch.distributed.run,
one process per rank): the sharded plan scan with CUDA-IPC mailboxes against the single-GPU scan and the oracle.
With fewer GPUs than ranks the ranks share cuda:0 (time-sliced) and the handles travel over gloo."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    n_dev = torch.cuda.device_count()
    dev = local % n_dev
    torch.cuda.set_device(dev)
    if n_dev >= world:
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
    else:
        dist.init_process_group("gloo")
    pkg = graft.load_package()
    orc = graft.load_oracle()
    from kcbs_b200 import sharding, workloads
    scene = "Benchmark_Empty32x32_30robots"
    oracle = orc.Oracle(pkg.scenes.world(scene))
    P, R, T = 16, 30, 2000
    traj = workloads.cell_plans(scene, P, T, seed=11)          # same seed on every rank = the all-gathered plans
    workloads.plant_conflict(traj, 1, 4, 9, 1500, length=7)
    workloads.plant_conflict(traj, 2, 3, 17, 40, length=3)
    workloads.plant_conflict(traj, 2, 20, 29, 1900)
    workloads.plant_conflict(traj, 5, 10, 11, 900, length=1000)
    workloads.plant_conflict(traj, 9, 0, 1, T - 1)
    ref, _ = oracle.first_conflict_batch(traj, n_threads=4)
    want = [tuple(int(v) for v in r) for r in ref]
    with pkg.Context(pkg.scenes.world(scene), device=dev) as ctx:
        sharding.connect_peers(ctx, rank, world, max_plans=64)
        d = torch.from_numpy(traj).cuda()
        out = torch.full((P, 4), 7, dtype=torch.int32, device="cuda")
        st = torch.cuda.current_stream()
        for rep in range(4):
            ctx.first_conflict_sharded_dev(P, R, T, d, None, out, st)
            torch.cuda.synchronize()
            ctx.peer_status()
            got = [tuple(int(v) for v in r) for r in out.cpu().numpy()]
            assert got == want, f"rank {rank} call {rep}: {got} != {want}"
            out.fill_(7)
        # back-to-back calls without host synchronisation in between (ranks stay in lock step through the mailboxes)
        for rep in range(20):
            ctx.first_conflict_sharded_dev(P, R, T, d, None, out, st)
        torch.cuda.synchronize()
        ctx.peer_status()
        assert [tuple(int(v) for v in r) for r in out.cpu().numpy()] == want
        single = torch.zeros((P, 4), dtype=torch.int32, device="cuda")
        ctx.first_conflict_dev(P, R, T, d, None, single, st)
        torch.cuda.synchronize()
        assert [tuple(int(v) for v in r) for r in single.cpu().numpy()] == want
        dist.barrier()
        ctx.peer_disconnect()
    print(f"sharded scan ok rank {rank}/{world} dev {dev} backend {dist.get_backend()}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
