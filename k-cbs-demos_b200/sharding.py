"""Host-side partitioning of the hot path across ranks (one process per GPU, torch.distributed).

Only two things on the path cross ranks:
  * per-robot replans / plan batches are independent -> contiguous or round-robin shards, no collective;
  * one plan scanned in time slabs -> one unsigned 64-bit min all-reduce of the packed (k, r1, r2) keys.
The same functions run under NCCL on the GPU box and under gloo in the CPU tests."""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist

NO_CONFLICT_KEY = -1  # 0xFFFF_FFFF_FFFF_FFFF viewed as int64
TILE = 128            # time indices per CTA tile of k_conflict_keys; slabs are aligned to it


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of n independent units (plans, states, samples)."""
    return n * rank // world, n * (rank + 1) // world


def round_robin(n_jobs: int, rank: int, world: int) -> List[int]:
    """(node, robot) replan jobs dealt round-robin over the ranks (config: 20 replans over 8 GPUs)."""
    return list(range(rank, n_jobs, world))


def time_slab(T: int, rank: int, world: int, align: int = TILE) -> Tuple[int, int]:
    """Time slab [k_lo, k_hi) of a T-step plan for this rank, aligned to the kernel's tile size so every
    rank keeps the bulk-copy path; empty slabs are possible when T < world * align."""
    tiles = (T + align - 1) // align
    lo, hi = shard_range(tiles, rank, world)
    return min(lo * align, T), min(hi * align, T)


def pack_key(k: int, r1: int, r2: int) -> int:
    return (k << 16) | (r1 << 8) | r2


def unpack_key(key: int):
    if key == NO_CONFLICT_KEY or key == 0xFFFFFFFFFFFFFFFF:
        return None
    key &= 0xFFFFFFFFFFFFFFFF
    return key >> 16, (key >> 8) & 0xFF, key & 0xFF


def connect_peers(ctx, rank: int, world: int, max_plans: int) -> None:
    """Set-up of the sharded plan scan (kcbs_first_conflict_sharded_dev): every rank exports its key mailbox, the
    128-byte handles are all-gathered (plain bytes; NCCL or gloo), every rank maps every peer's mailbox.  The
    all-gather doubles as the barrier the protocol needs between export (mailbox cleared) and the first scan."""
    handle = ctx.peer_export(max_plans)
    handles = [handle]
    if world > 1:
        handles = [None] * world
        dist.all_gather_object(handles, handle)
    ctx.peer_connect(rank, world, handles)
    if world > 1:
        dist.barrier()  # nobody scans before everybody has mapped everybody


def reduce_min_keys(keys: torch.Tensor) -> torch.Tensor:
    """In-place all-reduce of per-plan packed keys to the global first conflict.  Keys are unsigned 64-bit
    stored in int64 tensors (no-conflict = all ones = -1); flipping the sign bit turns the unsigned order
    into the signed order that ReduceOp.MIN implements."""
    assert keys.dtype == torch.int64
    bias = torch.iinfo(torch.int64).min
    keys.bitwise_xor_(bias)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(keys, op=dist.ReduceOp.MIN)
    keys.bitwise_xor_(bias)
    return keys


def gather_plans(results: torch.Tensor, world: int) -> torch.Tensor:
    """All-gathers equally sized per-rank result blocks ([P/world, 4] int32 conflicts or trajectories)."""
    if not (dist.is_initialized() and world > 1):
        return results
    out = [torch.empty_like(results) for _ in range(world)]
    dist.all_gather(out, results)
    return torch.cat(out, dim=0)


def merge_owned_rows(buf: torch.Tensor, owned: List[int]) -> torch.Tensor:
    """Every rank filled the rows it owns (all other rows zero); after the call every rank holds every row.
    A SUM all-reduce over disjoint owners is an all-gather that tolerates uneven row counts."""
    mask = torch.zeros(buf.shape[0], dtype=torch.bool, device=buf.device)
    if len(owned):
        mask[torch.as_tensor(owned, device=buf.device)] = True
    buf[~mask] = 0
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    return buf


def plan_replans_sharded(ctx, starts, goals, rank: int, world: int, max_T: int, seed: int = 1, device="cuda"):
    """Config 'per-robot replans sharded across GPUs': the R independent low-level replans of a K-CBS node are dealt
    round-robin over the ranks (kcbs_rrt_plan_many on each rank's own GPU: its jobs run concurrently), the planned trajectories are
    merged on every rank (NCCL over NVLink), and the merged plan is validated there with kcbs_first_conflict.
    The merge is ONE upload and ONE collective: every rank lays its rows, their lengths and its count of failed plans out in one
    host buffer (everything it does not own is zero), and a SUM all-reduce over disjoint owners gathers them (it tolerates
    uneven row counts; lengths and counts are small integers, exact in float32).
    Returns (plan [R, 5, max_T] tensor, lengths [R] tensor, all_solved flag)."""
    import numpy as np

    R = starts.shape[1]
    mine = round_robin(R, rank, world)
    n_plan = R * 5 * max_T
    host = np.zeros(n_plan + R + 1, dtype=np.float32)
    rows = host[:n_plan].reshape(R, 5, max_T)
    if len(mine):
        # this rank's replans run concurrently on the context's planner lanes (kcbs_rrt_plan_many)
        seeds = np.array([seed * 1000003 + r for r in mine], dtype=np.uint64)
        traj, T, stats = ctx.rrt_plan_many(mine, starts[:, mine], goals[mine], seeds=seeds, max_T=max_T)
        for j, r in enumerate(mine):
            if not stats[j].solved:
                host[n_plan + R] += 1.0
                continue
            rows[r, :, :T[j]] = traj[j, :, :T[j]]
            host[n_plan + r] = float(T[j])
    buf = torch.from_numpy(host).to(device)
    if dist.is_initialized() and world > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    plan = buf[:n_plan].view(R, 5, max_T)
    lengths = buf[n_plan:n_plan + R].to(torch.int32)
    return plan, lengths, float(buf[n_plan + R].item()) == 0.0
