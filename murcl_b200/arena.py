"""Flat parameter / gradient / shadow-weight arena for the training step.

One optimiser step of train_MuRCL.py:235-298 backpropagates through T x 2 bag passes that all use the same ~40
parameter tensors.  Left to autograd, every pass returns its own weight-gradient tensors and an ``add`` kernel per
parameter per pass folds them into ``.grad`` (~150 launches a step), every parameter is cast to the bf16 storage type
by its own launch, and the data-parallel gradient exchange first copies all gradients into a flat bucket and back.

``ParamArena`` lays the parameters of a set of modules out in ONE fp32 buffer (each parameter becomes a view of it):

  * ``grad``   - one fp32 buffer of the same layout; ``p.grad`` are views of it, and the weight-gradient / column-sum /
                 pooling-backward kernels ADD into those views directly (``ops.grad_target``), so no accumulation kernel
                 runs and ``zero_grad`` is one memset;
  * ``shadow`` - the bf16 copy the tcgen05 GEMMs read, refreshed by ONE cast launch after the optimiser step
                 (``refresh``), instead of one cast per parameter found stale (``ops.weight_as`` serves the views);
  * ``allreduce`` sums the flat gradient buffer across ranks in place (no flatten / unflatten copies);
  * ``flat_param`` exposes the whole arena as a single leaf so that the optimiser updates one tensor.

The parameters keep their names, shapes and ``state_dict`` behaviour (``load_state_dict`` copies into the views).
Moving a module to another device after building the arena detaches its parameters from it - build the arena last.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch

from . import ops

_ALIGN = 64          # elements: 256-byte fp32 / 128-byte bf16 alignment of every parameter (TMA needs 16 bytes)


class ParamArena:
    def __init__(self, params: Iterable[torch.nn.Parameter], shadow_dtype: Optional[torch.dtype] = torch.bfloat16):
        self.params: List[torch.nn.Parameter] = []
        seen = set()
        for p in params:
            if id(p) not in seen:
                seen.add(id(p))
                self.params.append(p)
        if not self.params:
            raise ValueError("ParamArena: no parameters")
        dev = self.params[0].device
        if dev.type != "cuda" or any(p.device != dev or p.dtype != torch.float32 for p in self.params):
            raise ops.MurclError("ParamArena: all parameters must be fp32 tensors on one CUDA device")
        self.offsets, total = [], 0
        for p in self.params:
            self.offsets.append(total)
            total += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.flat = torch.zeros(total, device=dev, dtype=torch.float32)
        self.grad = torch.zeros(total, device=dev, dtype=torch.float32)
        self.shadow = torch.zeros(total, device=dev, dtype=shadow_dtype) if shadow_dtype not in (None, torch.float32) else None
        with torch.no_grad():
            for p, off in zip(self.params, self.offsets):
                n = p.numel()
                view = self.flat[off:off + n].view(p.shape)
                view.copy_(p.data)
                p.data = view
                p.grad = self.grad[off:off + n].view(p.shape)
                p._murcl_accum = p.grad                      # ops.grad_target: kernels add into this view
                if self.shadow is not None:
                    p._murcl_shadow = self.shadow[off:off + n].view(p.shape)
        self.flat_param = torch.nn.Parameter(self.flat, requires_grad=True)
        self.flat_param.grad = self.grad
        self.refresh()

    # ---------------------------------------------------------------------------------------------------
    def zero_grad(self) -> None:
        """One memset; the ``.grad`` views stay attached (never ``set_to_none`` an arena's gradients)."""
        self.grad.zero_()
        for p, off in zip(self.params, self.offsets):
            if p.grad is None or p.grad.data_ptr() != self.grad.data_ptr() + 4 * off:
                p.grad = self.grad[off:off + p.numel()].view(p.shape)
                p._murcl_accum = p.grad

    def refresh(self) -> None:
        """Re-cast every weight to the shadow storage type: ONE launch, after the optimiser step (or after any other
        change of the fp32 values: ``load_state_dict``, manual edits)."""
        if self.shadow is not None:
            ops.cast_into(self.flat, self.shadow)
        # the optimiser updates the flat leaf: the parameter views' own version counters do not move, so every cache keyed
        # on them (bf16 casts outside the arena, the split-precision planes of the fp32 mode) is dropped explicitly
        ops.invalidate_weight_cache()

    def range_of(self, params) -> tuple:
        """Flat element range ``(lo, hi)`` covering ``params``, which must be consecutive members of the arena."""
        ids = [id(p) for p in self.params]
        idx = sorted(ids.index(id(p)) for p in params)
        if not idx or idx != list(range(idx[0], idx[-1] + 1)):
            raise ValueError("ParamArena.range_of: the parameters are not a consecutive run of the arena")
        hi = self.offsets[idx[-1] + 1] if idx[-1] + 1 < len(self.offsets) else self.flat.numel()
        return self.offsets[idx[0]], hi

    def allreduce(self, group=None, lo: int = 0, hi: Optional[int] = None, async_op: bool = False):
        """Sum the gradients ``[lo, hi)`` (default: all) across the ranks of ``group`` in place, straight out of the flat
        buffer.  ``async_op=True`` returns the collective's work handle (``.wait()`` before the optimiser reads the
        gradients): the head layers' gradients - 80 % of the bytes - are complete as soon as the recurrent-head tape has
        run, so their exchange can travel under the aggregators' backward passes."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return None
        hi = self.grad.numel() if hi is None else hi
        if hi <= lo:
            return None
        return dist.all_reduce(self.grad[lo:hi], op=dist.ReduceOp.SUM, group=group, async_op=async_op)

    def optimizer_params(self):
        """``[flat_param]``: hand this to the optimiser so that it updates one tensor (element-wise optimisers such as
        Adam / SGD with a uniform weight decay - the reference's, train_MuRCL.py:154-171 - are unchanged by the layout)."""
        return [self.flat_param]
