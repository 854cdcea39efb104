"""Data-parallel plumbing for the pre-training step: one process per GPU, bags sharded per rank.

The reference only knows single-process ``nn.DataParallel`` (train_MuRCL.py:145), which re-broadcasts the
parameters and scatters ~268 MB of features per forward call.  Here:

  * every rank packs / encodes / pools its own bags;
  * ONE all-gather per patch-step ships both views' projections (``2 x B_local x 128`` floats per rank) so each
    rank evaluates NT-Xent on the global ``2B x 2B`` matrix; the backward needs no collective - a rank keeps
    the gradient rows of its own samples;
  * ONE all-reduce(sum) per optimiser step carries all parameter gradients as a single flat bucket.
"""
from __future__ import annotations

from typing import Callable, Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced shard [lo, hi) of ``n_items`` for ``rank``."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class _GatherViews(torch.autograd.Function):
    """all_gather of a ``[2, B_local, d]`` block; backward = this rank's slice of the incoming gradient."""

    @staticmethod
    def forward(ctx, local: torch.Tensor, group):
        world = dist.get_world_size(group)
        ctx.rank = dist.get_rank(group)
        out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)      # concatenated along dim 0
        return out.view((world,) + tuple(local.shape))

    @staticmethod
    def backward(ctx, grad):
        return grad[ctx.rank].contiguous(), None


class DistributedNTXent(torch.nn.Module):
    """NT-Xent over the GLOBAL batch.  ``forward(z_i_local, z_j_local)`` returns the global loss (identical on
    every rank).  Summing the resulting parameter gradients over ranks (``allreduce_grads``) gives exactly the
    single-process gradient.  ``loss_fn(z_i, z_j, tau) -> (loss, cos)`` defaults to the fused CUDA kernel."""

    def __init__(self, local_batch_size: int, temperature: float, group=None, loss_fn: Optional[Callable] = None):
        super().__init__()
        self.local_batch_size = local_batch_size
        self.temperature = temperature
        self.group = group
        self.loss_fn = loss_fn
        self.last_cosine = None

    def forward(self, z_i: torch.Tensor, z_j: torch.Tensor) -> torch.Tensor:
        if z_i.shape[0] != self.local_batch_size or z_j.shape[0] != self.local_batch_size:
            raise RuntimeError(f"DistributedNTXent was built for local batch {self.local_batch_size}")
        fn = self.loss_fn
        if fn is None:
            from . import ops
            fn = ops.ntxent
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            both = _GatherViews.apply(torch.stack([z_i, z_j], 0), self.group)      # [W, 2, B_local, d]
            gi = both[:, 0].reshape(-1, z_i.shape[1])
            gj = both[:, 1].reshape(-1, z_i.shape[1])
            rank = dist.get_rank(self.group)
        else:
            gi, gj, rank = z_i, z_j, 0
        loss, cos = fn(gi, gj, float(self.temperature))
        b = self.local_batch_size
        self.last_cosine = cos[rank * b:(rank + 1) * b] if cos is not None else None
        return loss


def allreduce_grads(params: Iterable[torch.nn.Parameter], group=None) -> int:
    """Sum all parameter gradients across ranks with one flat all-reduce.  Returns the bucket size in bytes."""
    grads: List[torch.Tensor] = [p.grad for p in params if p.grad is not None]
    if not grads or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 0
    flat = torch._utils._flatten_dense_tensors(grads)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    for g, synced in zip(grads, torch._utils._unflatten_dense_tensors(flat, grads)):
        g.copy_(synced)
    return flat.numel() * flat.element_size()
