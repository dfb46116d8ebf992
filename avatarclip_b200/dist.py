"""View-sharded data parallelism: independent camera views shard one per rank; nets are replicated;
the only exchange is ONE all-reduce(sum) of the flat fp32 gradient per step (scaled by 1/world inside the fused Adam).
Semantics = single-GPU gradient accumulation over `world` views.

Host-side logic only (backend-agnostic: NCCL over NVLink on the GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.distributed as dist


def view_index(step: int, rank: int, world: int) -> int:
    """Index of the camera/background/light/jitter draw rank `rank` consumes at optimiser step `step`: one seeded
    stream, rank-strided, so that N ranks x 1 view == 1 rank x N views (gradient accumulation)."""
    return step * world + rank


def init_from_env(backend: str = "nccl", device: Optional[torch.device] = None):
    """torchrun-style rendezvous (RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT from the environment)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return None
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    kw = {}
    if backend == "nccl" and device is not None:
        kw["device_id"] = device
    dist.init_process_group(backend, **kw)
    return dist.group.WORLD


def allreduce_sum_(flat_grad: torch.Tensor, pg=None) -> torch.Tensor:
    """The step's single collective: in-place sum of the flat gradient vector over all view shards."""
    if pg is not None and dist.get_world_size(pg) > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=pg)
    return flat_grad


def grad_scale(pg=None) -> float:
    return 1.0 if pg is None else 1.0 / dist.get_world_size(pg)
