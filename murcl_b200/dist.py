"""Data-parallel plumbing for the pre-training step: one process per GPU, bags sharded per rank.

The reference only knows single-process ``nn.DataParallel`` (train_MuRCL.py:145), which re-broadcasts the
parameters and scatters ~268 MB of features per forward call.  Here:

  * every rank packs / encodes / pools its own bags;
  * ONE all-gather per patch-step ships both views' projections (``2 x B_local x 128`` floats per rank) so each
    rank evaluates NT-Xent on the global ``2B x 2B`` matrix; the backward needs no collective - a rank keeps
    the gradient rows of its own samples;
  * ONE all-reduce(sum) per optimiser step carries all parameter gradients as a single flat bucket.

Intra-bag sharding (BASELINE config 5: one 100k-patch bag over 2/4/8 GPUs): the rows of every bag are
split across the ranks; each rank pools its own rows and ONE all-gather of ``3 + L`` floats per bag merges the partial
results (online-softmax merge) - ``sharded_attention_pool``.  The backward pass needs no collective.
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


def exchange_row_stats(lse_own: torch.Tensor, inv_own: torch.Tensor, share: torch.Tensor, group=None):
    """One all-gather of every rank's ``[lse (2b) | inv_norm (2b) | loss share (1)]`` - the per-row statistics of ITS samples
    (view-i rows then view-j rows) - assembled into the global row order ``[view i: rank 0..W-1 | view j: rank 0..W-1]``:
    returns ``(lse [2B], inv_norm [2B], loss)``.  Without an initialised group it is the identity."""
    b2 = lse_own.numel()
    pack = torch.cat([lse_own.reshape(-1), inv_own.reshape(-1), share.reshape(-1)]).contiguous()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        world = dist.get_world_size(group)
        allp = torch.empty((world, 2 * b2 + 1), dtype=pack.dtype, device=pack.device)
        dist.all_gather_into_tensor(allp.view(-1), pack, group=group)
    else:
        world, allp = 1, pack.view(1, -1)
    b = b2 // 2
    lse = allp[:, :b2].reshape(world, 2, b).permute(1, 0, 2).reshape(-1).contiguous()
    inv = allp[:, b2:2 * b2].reshape(world, 2, b).permute(1, 0, 2).reshape(-1).contiguous()
    return lse, inv, allp[:, 2 * b2].sum()


class _TwoPhaseNTXent(torch.autograd.Function):
    """NT-Xent over the global batch with the log-sum-exp pass restricted to the rank's OWN rows (``murcl_ntxent_lse_slab``),
    one more small all-gather (``exchange_row_stats``: 4 b + 1 floats per rank) and the gradient slab from the complete
    statistics (``murcl_ntxent_grad_slab``).  Same loss and the same gradient rows as the every-rank-evaluates-all-rows
    form; per rank the O((2B)^2 d) log-sum-exp work shrinks by the number of ranks (measured on one B200: 149 -> ~40 us per
    patch-step at the global batch of 8 ranks), which is what pays for the second collective from 4 ranks on."""

    @staticmethod
    def forward(ctx, z_i, z_j, temperature, group, fns):
        from . import ops
        lse_fn, grad_fn = fns if fns is not None else (ops.ntxent_lse_slab, ops.ntxent_grad_slab)
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        b, d = z_i.shape
        local = torch.stack([z_i.detach().float(), z_j.detach().float()], 0).contiguous()          # [2, b, d]
        allz = torch.empty((world * 2, b, d), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(allz, local, group=group)
        Bg = world * b
        z_all = allz.view(world, 2, b, d).permute(1, 0, 2, 3).reshape(2 * Bg, d).contiguous()     # view i of all ranks, then view j
        slab = (rank * b, b)
        inv_n, lse, share, cos = lse_fn(z_all, Bg, float(temperature), slab)
        own = slice(rank * b, (rank + 1) * b)
        own_j = slice(Bg + rank * b, Bg + (rank + 1) * b)
        lse_f, inv_f, loss = exchange_row_stats(torch.cat([lse[own], lse[own_j]]), torch.cat([inv_n[own], inv_n[own_j]]), share, group)
        dz = grad_fn(z_all, Bg, float(temperature), slab, inv_f, lse_f)
        ctx.save_for_backward(dz[own], dz[own_j])
        cos_local = cos[own].contiguous()
        ctx.mark_non_differentiable(cos_local)
        return loss.reshape(()), cos_local

    @staticmethod
    def backward(ctx, g, _gcos):
        dzi, dzj = ctx.saved_tensors
        return dzi * g, dzj * g, None, None, None


def _local_lse_min_world() -> int:
    import os
    return int(os.environ.get("MURCL_NTX_LOCAL_LSE_MIN_WORLD", "4"))


class DistributedNTXent(torch.nn.Module):
    """NT-Xent over the GLOBAL batch.  ``forward(z_i_local, z_j_local)`` returns the global loss (identical on
    every rank).  Summing the resulting parameter gradients over ranks (``allreduce_grads``) gives exactly the
    single-process gradient.  ``loss_fn(z_i, z_j, tau) -> (loss, cos)`` defaults to the fused CUDA kernel."""

    def __init__(self, local_batch_size: int, temperature: float, group=None, loss_fn: Optional[Callable] = None,
                 phase_fns: Optional[Tuple[Callable, Callable]] = None):
        """``phase_fns = (lse_fn, grad_fn)`` substitutes the two kernels of the own-rows form (``ops.ntxent_lse_slab`` /
        ``ops.ntxent_grad_slab`` signatures) and forces that form at any world size - the CPU tests pass torch versions."""
        super().__init__()
        self.phase_fns = phase_fns
        self.local_batch_size = local_batch_size
        self.temperature = temperature
        self.group = group
        self.loss_fn = loss_fn
        self.last_cosine = None

    def forward(self, z_i: torch.Tensor, z_j: torch.Tensor) -> torch.Tensor:
        if z_i.shape[0] != self.local_batch_size or z_j.shape[0] != self.local_batch_size:
            raise RuntimeError(f"DistributedNTXent was built for local batch {self.local_batch_size}")
        fn = self.loss_fn
        b = self.local_batch_size
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1
        if multi and fn is None and (self.phase_fns is not None or (
                z_i.is_cuda and dist.get_world_size(self.group) >= _local_lse_min_world()
                and z_i.shape[1] % 4 == 0 and z_i.shape[1] <= 256)):
            # from 4 ranks on: every rank reduces only ITS rows of the global score matrix, the per-row statistics travel
            loss, cos = _TwoPhaseNTXent.apply(z_i, z_j, float(self.temperature), self.group, self.phase_fns)
            self.last_cosine = cos
            return loss
        if multi:
            both = _GatherViews.apply(torch.stack([z_i, z_j], 0), self.group)      # [W, 2, B_local, d]
            gi = both[:, 0].reshape(-1, z_i.shape[1])
            gj = both[:, 1].reshape(-1, z_i.shape[1])
            rank = dist.get_rank(self.group)
        else:
            gi, gj, rank = z_i, z_j, 0
        if fn is None:
            # the loss covers the global batch; only this rank's samples need gradient rows (_GatherViews.backward keeps
            # exactly those), so the gradient contraction runs on the rank's slab of rows
            from . import ops
            loss, cos = ops.ntxent(gi, gj, float(self.temperature), slab=(rank * b, b))
        else:
            loss, cos = fn(gi, gj, float(self.temperature))
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


# ------------------------------------------------------------------------------------------------
# intra-bag sharding: attention pooling of bags whose ROWS are split across the ranks
# ------------------------------------------------------------------------------------------------
def merge_pool_partials(m: torch.Tensor, l: torch.Tensor, M_loc: torch.Tensor, n: torch.Tensor, group=None):
    """Online-softmax merge of per-rank pooling partials of the same ``B`` bags.

    ``m``, ``l`` [B]: max and sum-exp of the rank's scores of each bag (``-inf`` / 0 where the rank holds no row of it);
    ``M_loc`` [B, L]: sum over the rank's rows of ``softmax_local(s) * h`` (normalised with the LOCAL ``l``);
    ``n`` [B]: the rank's row counts.  Returns ``(M, scale, n_total)``: the globally normalised pooled vectors, the factor
    ``l_r e^{m_r - m} / l`` that turns this rank's local softmax weights into global ones, and the global row counts.
    One all-gather of ``3 + L`` floats per bag; without an initialised process group it is the identity."""
    B, L = M_loc.shape
    packed = torch.cat([m.reshape(B, 1), l.reshape(B, 1), n.reshape(B, 1).to(M_loc.dtype), M_loc], 1).contiguous()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        allp = torch.empty((world * B, 3 + L), dtype=packed.dtype, device=packed.device)
        dist.all_gather_into_tensor(allp, packed, group=group)
        allp = allp.view(world, B, 3 + L)
    else:
        rank, allp = 0, packed.unsqueeze(0)
    m_r, l_r, n_r, M_r = allp[..., 0], allp[..., 1], allp[..., 2], allp[..., 3:]
    m_g = m_r.max(0).values                                                   # [B]
    safe = torch.where(torch.isfinite(m_g), m_g, torch.zeros_like(m_g))      # bags with no rows anywhere
    w_r = l_r * torch.exp(m_r - safe)                                         # [W, B]; exp(-inf) = 0 for empty shards
    l_g = w_r.sum(0)
    inv = torch.where(l_g > 0, 1.0 / l_g, torch.zeros_like(l_g))
    M = (w_r.unsqueeze(-1) * M_r).sum(0) * inv.unsqueeze(-1)
    return M, w_r[rank] * inv, n_r.sum(0)


class _KernelPoolFns:
    """The CUDA kernels behind the local part of the sharded pooling (murcl_seg_softmax, murcl_seg_wsum,
    murcl_pool_bwd_scores, murcl_pool_bwd_direct).  Tests on CPU substitute a torch implementation."""

    @staticmethod
    def local_pool(h, s, offsets, row_seg, B):
        from . import ops
        p, stats = ops.seg_softmax(s.contiguous(), offsets, B, 1, False)
        M = ops.seg_wsum(p, h, offsets, B, 1).reshape(B, -1)
        return p, stats[:, 0, 0], stats[:, 0, 1], M

    @staticmethod
    def backward(p, h, dM, M, offsets, row_seg, B):
        from . import ops
        L = h.shape[1]
        ds = ops.pool_bwd_scores(p, h, dM, M.reshape(B, 1, L), offsets, row_seg, B, 1, False)
        dh = torch.empty_like(h)
        ops.pool_bwd_direct(p, dM, row_seg, 1, L, dh, False)
        return dh, ds


class _ShardedAttentionPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, s, offsets, row_seg, inv_sqrt_n, group, fns):
        B = offsets.numel() - 1
        hd, sd = h.detach().contiguous(), s.detach().contiguous().float()
        p_loc, m, l, M_loc = fns.local_pool(hd, sd, offsets, row_seg, B)
        n_loc = (offsets[1:] - offsets[:-1]).to(torch.float32)
        M, scale, n_tot = merge_pool_partials(m, l, M_loc, n_loc, group)
        post = torch.rsqrt(n_tot.clamp_min(1.0)) if inv_sqrt_n else torch.ones_like(n_tot)   # abmil.py:41 on the GLOBAL count
        p = (p_loc * scale[row_seg.long()]).contiguous()                      # global softmax weights of the local rows
        ctx.save_for_backward(hd, p, M, post, offsets, row_seg)
        ctx.fns = fns
        ctx.mark_non_differentiable(p)
        return M * post.unsqueeze(1), p * post[row_seg.long()]

    @staticmethod
    def backward(ctx, dout, _dp):
        h, p, M, post, offsets, row_seg = ctx.saved_tensors
        B = offsets.numel() - 1
        dM = (dout.float() * post.unsqueeze(1)).contiguous()                  # gradient w.r.t. the un-scaled pooled vectors
        dh, ds = ctx.fns.backward(p, h, dM, M.contiguous(), offsets, row_seg, B)
        return dh, ds, None, None, None, None, None


def sharded_attention_pool(h: torch.Tensor, s: torch.Tensor, offsets: torch.Tensor, row_seg: torch.Tensor,
                           inv_sqrt_n: bool = False, group=None, fns=None):
    """Attention pooling ``M[b] = post_b * sum_n softmax_bag(s)[n] h[n]`` of ``B`` bags whose rows are split across the
    ranks of ``group``: ``h`` [n_local, L] and ``s`` [n_local] are this rank's rows and raw scores, ``offsets`` [B+1] /
    ``row_seg`` [n_local] their LOCAL CSR description (a rank may hold no row of a bag).  Returns ``(M [B, L], p [n_local])``
    - the same on every rank for ``M``, the global attention weights of the local rows for ``p``.  Differentiable in
    ``h`` and ``s``; every rank must feed the same ``dM`` (a replicated head does), then no backward collective is needed
    and the parameter gradients are summed by ``allreduce_grads`` as usual."""
    return _ShardedAttentionPool.apply(h, s, offsets, row_seg, bool(inv_sqrt_n), group, fns or _KernelPoolFns)
