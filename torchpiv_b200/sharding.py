"""Multi-GPU plumbing: one process per GPU, pairs sharded by index, NO data-path collective.

Image pairs are independent (the reference has no cross-pair state, PIVbackend.py:868-901), so a
sequence is split into contiguous blocks of pairs (``dataset.shard_range``), every rank runs the
ordinary single-GPU pipeline on its block, and the small result fields (tens of KB per pair) are
gathered on the host.  ``torch.distributed`` is only used for that gather, for barriers and for
the max-over-ranks of a timing -- NCCL on the GPU box, gloo in the CPU tests."""
from __future__ import annotations

import os
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from .dataset import shard_range

__all__ = ["dist_env", "shard_range", "max_over_ranks", "barrier", "gather_results"]


def dist_env() -> Tuple[int, int, int]:
    """(rank, world, local_rank) from the torchrun environment (1 process = 1 GPU)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    return rank, world, local


def _active() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def _reduce_device(device: Optional[torch.device]) -> torch.device:
    if device is not None:
        return torch.device(device)
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def barrier(device: Optional[torch.device] = None) -> None:
    """Process-group barrier followed by a device synchronize (both sides of a timed region)."""
    if _active():
        dist.barrier()
    if device is not None and torch.device(device).type == "cuda":
        torch.cuda.synchronize(device)


def max_over_ranks(value: float, device: Optional[torch.device] = None) -> float:
    """Maximum of a per-rank scalar (a device time in ms): the job is as slow as its slowest rank."""
    if not _active():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=_reduce_device(device))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_results(local: List[tuple], dst: int = 0) -> Optional[List[tuple]]:
    """Host-side gather of per-rank result lists ``[(pair_index, payload...), ...]`` to rank ``dst``,
    merged in pair order.  Returns None on the other ranks.  Python objects travel over the process
    group's CPU path (``gather_object``); nothing is reduced on the GPUs."""
    if not _active():
        return sorted(local, key=lambda r: r[0])
    rank, world = dist.get_rank(), dist.get_world_size()
    bucket = [None] * world if rank == dst else None
    dist.gather_object(local, bucket, dst=dst)
    if rank != dst:
        return None
    merged = [item for part in bucket for item in part]
    return sorted(merged, key=lambda r: r[0])
