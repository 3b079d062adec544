"""Multi-GPU plumbing for the data-parallel policy update (one process per GPU, NCCL).

The reference has no distributed path (only single-process ``nn.DataParallel``,
tianshou/utils/net/common.py:473-515).  Here every rank holds a full replica of the 11k
parameters + Adam state and its own shard of the rollout; per optimiser step exactly ONE
all-reduce carries the flat gradient and the loss sums (n_params + 4 floats).  All functions work
on whatever backend the process group uses, so the host-side logic is testable with ``gloo`` on
CPU tensors (tests/test_parallel.py).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def world() -> tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(lo: int, hi: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous slice of minibatch positions [lo, hi) owned by ``rank`` (sizes differ by <= 1)."""
    n = hi - lo
    base, rem = divmod(n, world_size)
    start = lo + rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def allreduce_sum_(flat: torch.Tensor) -> torch.Tensor:
    """In-place SUM all-reduce of the flat gradient(+loss sums) buffer."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return flat


def allgather_moments(moments: torch.Tensor) -> torch.Tensor:
    """[world, 3] tensor of every rank's (count, mean, M2), identical on all ranks, rank order."""
    _, w = world()
    if w == 1:
        return moments.reshape(1, 3)
    out = torch.empty((w, 3), dtype=moments.dtype, device=moments.device)
    dist.all_gather_into_tensor(out, moments.reshape(1, 3).contiguous())
    return out


def broadcast_params_(flat: torch.Tensor, src: int = 0) -> torch.Tensor:
    """Make replicas bit-identical at start-up."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(flat, src=src)
    return flat
