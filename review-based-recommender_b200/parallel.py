"""Data-parallel plumbing: one process per GPU, batch sharded across ranks, ONE all-reduce of the flat
parameter-gradient buffer per step.  The reference's only parallel mode is
nn.DataParallel (trainer/train_deepconn_pp.py:129-131), which re-broadcasts every parameter each step.

The forward/backward has no collective: every rank encodes its own shard.  After `loss.backward()` the
model's GradArena holds all parameter gradients in one contiguous fp32 buffer (ops.GradArena), so the
exchange is a single NCCL all-reduce over NVLink/NVSwitch followed by a scale by 1/world_size (each rank's
MSELoss averaged over its own shard → mean over the global batch).
"""
from __future__ import annotations

import os
from typing import Iterable, Optional

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None) -> tuple:
    """Initialise torch.distributed from torchrun's env (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local, world


def shard_range(n: int, rank: int, world: int) -> range:
    """Contiguous shard of n samples for this rank (first n % world ranks get one extra)."""
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def _arena_covers(model: torch.nn.Module) -> bool:
    arena = getattr(model, "last_arena", None)
    if arena is None or arena.flat is None:
        return False
    lo = arena.flat.data_ptr()
    hi = lo + arena.flat.numel() * 4
    for p in model.parameters():
        if p.requires_grad:
            if p.grad is None or not (lo <= p.grad.data_ptr() < hi):
                return False
    return True


def allreduce_gradients(model: torch.nn.Module, group=None, average: bool = True) -> int:
    """Sum (and average) parameter gradients across ranks.  Returns the number of collectives issued:
    1 when every `.grad` lives in the step's flat arena (the fast path), else one flattened bucket."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    world = dist.get_world_size(group)
    if _arena_covers(model):
        flat = model.last_arena.flat
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            flat.mul_(1.0 / world)
        return 1
    grads = [p.grad for p in model.parameters() if p.requires_grad and p.grad is not None]
    if not grads:
        return 0
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat.mul_(1.0 / world)
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()
    return 1


def broadcast_parameters(model: torch.nn.Module, src: int = 0, group=None) -> None:
    """Make every rank start from rank `src`'s parameters (replicated-parameter DP)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for p in model.parameters():
        dist.broadcast(p.data, src=src, group=group)
