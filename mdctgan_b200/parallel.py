"""Batch-sharded training / segment-sharded inference over the GPUs of one node (one process per GPU).

The reference is single-GPU (`DataParallel` is commented out, models/models.py:17-18); what is new here:

  * training: replicated G / D / Adam state, a different slice of the global batch per rank, and ONE
    `all_reduce(sum)` per step over the flat [grad_G | grad_D] bucket (optim.FlatBucket); the 1/world factor is folded
    into the Adam kernel (`FusedAdam.grad_scale`).  Both backward sweeps run before the exchange: loss_D only sees
    `sr.detach()` computed before the generator update, so this is the reference order (train.py:182-202) exactly.
  * inference / long-form generation: segments are independent (InstanceNorm is per sample, abs-norm constants are
    global), so ranks take contiguous runs of segments and nothing is exchanged on the data path.

Host-side logic only (no kernels): covered on CPU by world-size-2 gloo tests (tests/test_parallel_cpu.py).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def flat_layout(numels: Sequence[int], align: int = 4) -> Tuple[List[int], int]:
    """Element offsets of tensors packed back to back with every segment padded to `align` elements (16 bytes for
    fp32: float4 access in the Adam kernel).  Returns (offsets, total)."""
    offs, off = [], 0
    for n in numels:
        offs.append(off)
        off += (int(n) + align - 1) // align * align
    return offs, off


def shard_range(n_items: int, world: int, rank: int) -> range:
    """Contiguous, balanced run of items for `rank` (the first n_items % world ranks take one extra)."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


class GradExchange:
    """The single collective of a training step."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.grad_scale = 1.0 / self.world
        self.calls = 0

    def __call__(self, flat_grads: torch.Tensor) -> torch.Tensor:
        if flat_grads.dim() != 1 or not flat_grads.is_contiguous():
            raise ValueError("GradExchange: expects the flat contiguous gradient bucket")
        self.calls += 1
        if self.world > 1:
            dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM, group=self.group)
        return flat_grads


def broadcast_flat(flats: Sequence[torch.Tensor], src: int = 0, group=None) -> None:
    """Identical initial parameters / optimiser state on every rank."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        for t in flats:
            dist.broadcast(t, src=src, group=group)


def check_replicas_in_sync(flat_params: torch.Tensor, group=None, atol: float = 0.0) -> bool:
    """Debug aid: max |p - p_rank0| over the bucket is <= atol on every rank."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return True
    ref = flat_params.clone()
    dist.broadcast(ref, src=0, group=group)
    bad = torch.tensor([float((flat_params - ref).abs().max() > atol)], device=flat_params.device)
    dist.all_reduce(bad, op=dist.ReduceOp.MAX, group=group)
    return bad.item() == 0.0
