"""Multi-GPU plumbing of the NeRF hot path: one process per GPU, torch.distributed.

Rendering shards rays / views across ranks with NO data-path collective; training is data parallel and its only
exchange step is one all-reduce (mean) of the flat fp32 gradient buffer of every block (2 x 595,848 floats = 4.77 MB)
over NCCL / NVLink.  The same functions run on the gloo backend, which is how the CPU tests cover the N>1 logic.
"""
from __future__ import annotations

import torch
import torch.distributed as td


def is_distributed() -> bool:
    return td.is_available() and td.is_initialized()


def world_size() -> int:
    return td.get_world_size() if is_distributed() else 1


def rank() -> int:
    return td.get_rank() if is_distributed() else 0


def shard_range(n_items: int, rank_: int | None = None, world: int | None = None) -> range:
    """Contiguous, balanced share of ``n_items`` (views of a test set, ray chunks of an image) of one rank:
    sizes differ by at most one, every item belongs to exactly one rank, empty when there are more ranks than items."""
    r = rank() if rank_ is None else rank_
    w = world_size() if world is None else world
    if not 0 <= r < w:
        raise ValueError(f'rank {r} outside world of {w}')
    lo = (n_items * r) // w
    hi = (n_items * (r + 1)) // w
    return range(lo, hi)


def allreduce_mean_(flat_grads: list[torch.Tensor]) -> None:
    """In place: every flat gradient buffer becomes the mean over ranks (sum all-reduce, then 1/world), which makes a
    step on N ranks x B rays equal to a single-GPU step on N*B rays (the loss is a mean over rays)."""
    w = world_size()
    if w == 1:
        return
    for g in flat_grads:
        td.all_reduce(g, op=td.ReduceOp.SUM)
        g.mul_(1.0 / w)


def allreduce_sum_(flat_grads: list[torch.Tensor]) -> None:
    """In place SUM over ranks; the 1/world of the mean is folded into the optimiser's gradient read (K7 grad_mult)."""
    if world_size() == 1:
        return
    for g in flat_grads:
        td.all_reduce(g, op=td.ReduceOp.SUM)


def broadcast_parameters_(flat_params: list[torch.Tensor], src: int = 0) -> None:
    """Makes every rank start from rank ``src``'s weights (ranks seeded differently for their ray batches)."""
    if world_size() == 1:
        return
    for p in flat_params:
        td.broadcast(p, src=src)
