"""Multi-GPU layout: streams are independent, so they shard across the GPUs of one box
with NO collective on the data path.  One process per GPU, each
owning an engine with its own weight copy and state arena; stream `s` lives on rank
`s % world_size` at local slot `s // world_size`.  The only cross-rank traffic is the
end-of-run reduction of timing / throughput numbers (`reduce_max`, `reduce_sum`).
"""
from __future__ import annotations

import os
from typing import List, Tuple

import numpy as np


def world() -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def owner_of(stream: int, world_size: int) -> Tuple[int, int]:
    """global stream id -> (rank, local slot)."""
    return stream % world_size, stream // world_size


def local_streams(n_streams: int, world_size: int, rank: int) -> np.ndarray:
    """Global ids of the streams this rank owns, in local-slot order."""
    return np.arange(rank, n_streams, world_size, dtype=np.int64)


def local_capacity(n_streams: int, world_size: int, rank: int) -> int:
    return (n_streams - rank + world_size - 1) // world_size if rank < n_streams else 0


def reduce_max(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def reduce_sum(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t[0])


def gather_events(local_events: List[Tuple[int, int, int]], world_size: int, rank: int):
    """All ranks' (local_slot, frame, kind) events as global (stream, frame, kind), stream-then-frame order."""
    import torch.distributed as dist
    mine = [(slot * world_size + rank, frame, kind) for (slot, frame, kind) in local_events]
    if not (dist.is_available() and dist.is_initialized()):
        return sorted(mine)
    bucket = [None] * world_size
    dist.all_gather_object(bucket, mine)
    return sorted(e for part in bucket for e in part)
