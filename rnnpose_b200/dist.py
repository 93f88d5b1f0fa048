"""Multi-GPU plumbing of the refinement path: objects shard embarrassingly across ranks (no data-path
collective), and the per-object metrics are all-gathered once at the end.

Mirrors the reference's sharding (``DistributedSequatialSampler``: ``indices[rank::world]``,
reference utils/distributed_utils.py:154-169) and its metric gather (reference tools/train.py:724-741).
One process per GPU, ``torch.distributed`` with NCCL over NVLink/NVSwitch on GPUs; the same code runs
with the ``gloo`` backend on CPU tensors for the world_size-2 tests.
"""
from __future__ import annotations

import math
import os
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def env_world() -> Tuple[int, int, int]:
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def init_from_env(backend: str = "nccl") -> Tuple[int, int, int]:
    """Rendezvous from RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT (torchrun)."""
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, local_rank, world


def shard_indices(n_total: int, rank: int, world: int) -> List[int]:
    """Object indices of this rank: pad to a multiple of world by wrapping, then take rank::world
    (reference utils/distributed_utils.py:158-166)."""
    per = math.ceil(n_total / world)
    idx = list(range(n_total))
    idx += idx[: per * world - n_total]
    return idx[rank::world]


def all_gather_metrics(local: torch.Tensor, n_total: int | None = None) -> torch.Tensor:
    """local: [B_local, M] float32 (same B_local on every rank).  Returns [B_total, M] on every rank in
    OBJECT order (undoing the rank::world interleave) and drops the wrap-around padding."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        out = local
    else:
        world = dist.get_world_size()
        bufs = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(bufs, local.contiguous())
        out = torch.stack(bufs, dim=1).reshape(-1, local.shape[-1])      # [per, world, M] -> object order
    return out if n_total is None else out[:n_total]


def barrier():
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def max_over_ranks(value: float, device) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
