"""Scene sharding across ranks.

The path shards embarrassingly by scene: rank r owns scenes r, r+W, r+2W, ... (the same round-robin
the reference's eval drivers use, llava/eval/model_scanqa.py:244-245); depth never leaves its rank.
The only exchange is ONE all-gather of the fixed-size per-scene MC records (272 bytes each), written
by the greedy kernel straight into the send buffer.  PE rows are consumed on the rank that produced
them (data parallel) and are not gathered by default.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def world() -> tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_indices(n_items: int, rank: int | None = None, world_size: int | None = None) -> list[int]:
    """Indices of the scenes owned by `rank`: i % world_size == rank."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    return list(range(rank, n_items, world_size))


def padded_count(n_items: int, world_size: int) -> int:
    """Records per rank in the all-gather buffer (ranks with fewer scenes pad with empty records)."""
    return (n_items + world_size - 1) // world_size


def all_gather_records(local: torch.Tensor, n_items: int | None = None) -> torch.Tensor:
    """local: (n_local, R) int32 records of this rank's scenes (device tensor for NCCL, CPU for gloo).
    Returns the records of ALL scenes in global scene order, (n_items, R) when n_items is given
    (round-robin ownership is undone), otherwise the raw (world*n_local, R) rank-major buffer."""
    r, w = world()
    if w == 1:
        return local if n_items is None else local[:n_items]
    per = local.shape[0] if n_items is None else padded_count(n_items, w)
    send = local
    if local.shape[0] != per:
        send = torch.full((per, local.shape[1]), -1, dtype=local.dtype, device=local.device)
        send[: local.shape[0]] = local
    recv = torch.empty((w * per, local.shape[1]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(recv, send.contiguous())
    if n_items is None:
        return recv
    # rank-major (rank, slot) -> global index slot*w + rank
    out = recv.view(w, per, -1).transpose(0, 1).reshape(w * per, -1)
    return out[:n_items]
