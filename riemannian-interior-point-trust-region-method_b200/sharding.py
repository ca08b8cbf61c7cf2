"""Multi-GPU plumbing: (instance, initialpoint) pairs are independent, so a sweep shards over ranks with no
collective on the hot path; the only exchange is the final gather of the fixed-size per-pair result records
(SURVEY.md section 8e).  One process per GPU (torchrun), torch.distributed (NCCL on GPUs, gloo in CPU tests)."""
import os

import torch
import torch.distributed as dist


def env_rank_world():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def shard_range(total, rank, world):
    """Contiguous block of pair ids [lo, hi) for `rank`; sizes differ by at most one."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def weak_range(pairs_per_rank, rank):
    """Weak scaling (bench.py): every rank owns `pairs_per_rank` pairs, global ids rank*B .. rank*B+B-1."""
    return rank * pairs_per_rank, (rank + 1) * pairs_per_rank


def gather_records(local, world, out=None):
    """all_gather of equally sized [B, F] record tensors into [world * B, F] (rank-major), every rank gets all.
    With world == 1 returns `local`."""
    if world == 1:
        return local
    if out is None:
        out = local.new_empty((world * local.shape[0],) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, local.contiguous())
    return out


def gather_ragged_records(local, world):
    """Gather for shards of unequal size (shard_range): pads to the largest shard, returns the concatenation."""
    if world == 1:
        return local
    n = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s) for s in sizes]
    m = max(sizes)
    padded = local.new_zeros((m,) + tuple(local.shape[1:]))
    padded[:local.shape[0]] = local
    out = local.new_empty((world * m,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, padded)
    return torch.cat([out[r * m:r * m + sizes[r]] for r in range(world)])


def max_over_ranks(value, world, device):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)
