"""Multi-GPU plumbing: one process per GPU, worlds sharded contiguously, no data-path
collective for the env tick; the learner all-reduces (1) three advantage moments per update and (2) the flat
gradient buffer per optimizer step.  Everything here works on whatever device the tensors live on, so the
host-side logic is exercised by world_size-2 gloo tests on CPU (tests/test_parallel_gloo.py)."""
from __future__ import annotations

import torch


def shard_worlds(total_worlds: int, rank: int, world_size: int):
    """Contiguous split of `total_worlds` worlds: returns (world_offset, num_worlds) of this rank.
    The first (total % world_size) ranks take one extra world."""
    if not (0 <= rank < world_size):
        raise ValueError('rank out of range')
    base, rem = divmod(total_worlds, world_size)
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def allreduce_moments(moments: torch.Tensor, group=None):
    """moments = (sum, sum of squares, count) in float64; summed over ranks in place."""
    import torch.distributed as dist
    dist.all_reduce(moments, op=dist.ReduceOp.SUM, group=group)
    return moments


def normalize_from_moments(x: torch.Tensor, moments: torch.Tensor):
    """(x - mean) / std with numpy semantics (ddof = 0) from global moments (model/ppo.py:148)."""
    cnt = moments[2]
    mean = moments[0] / cnt
    var = moments[1] / cnt - mean * mean
    return ((x.double() - mean) / torch.sqrt(var)).to(x.dtype)


def average_gradients(flat_grad: torch.Tensor, group=None):
    """Sum the flat gradient over ranks (the 1/world_size is folded into the fused Adam step)."""
    import torch.distributed as dist
    dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    return flat_grad


def broadcast_parameters(flat: torch.Tensor, src=0, group=None):
    import torch.distributed as dist
    dist.broadcast(flat, src=src, group=group)
    return flat
