"""Optimiser step over a ``ParamArena`` as one kernel launch.

The reference trains with ``torch.optim.Adam`` (train_MuRCL.py:154-171: lr, weight decay, default betas / eps, no
amsgrad) and calls ``optimizer.step()`` once per batch (:296).  With the parameters in one flat buffer (arena.py) the
update is a single elementwise pass: ``ArenaAdam.step()`` launches ``murcl_adam_step`` - read p, g, m, v once, write p, m, v
and the bf16 shadow weights - instead of torch's ~19 multi-tensor launches followed by the arena's cast launch.  The step
counter lives on the device, so the launch replays inside a CUDA graph.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib, ops
from ._lib import check
from .arena import ParamArena


class ArenaAdam:
    """``torch.optim.Adam(arena.optimizer_params(), lr, betas, eps, weight_decay)`` in one launch.

    ``param_groups[0]["lr"]`` is read at every ``step()`` (learning-rate schedules assign to it, as with a torch
    optimiser); inside a captured CUDA graph a host value is frozen at capture time, so a schedule that must change
    between replays writes ``lr_tensor`` (a one-element fp32 device tensor) instead - pass ``lr_on_device=True``."""

    def __init__(self, arena: ParamArena, lr: float = 1e-3, betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, lr_on_device: bool = False):
        if not isinstance(arena, ParamArena):
            raise TypeError("ArenaAdam updates a murcl_b200.arena.ParamArena")
        if not (0.0 <= betas[0] < 1.0 and 0.0 <= betas[1] < 1.0) or eps < 0.0 or lr < 0.0 or weight_decay < 0.0:
            raise ValueError("ArenaAdam: invalid hyper-parameters")
        self.arena = arena
        self.param_groups = [{"params": arena.optimizer_params(), "lr": float(lr), "betas": (float(betas[0]), float(betas[1])),
                              "eps": float(eps), "weight_decay": float(weight_decay)}]
        dev = arena.flat.device
        self.exp_avg = torch.zeros_like(arena.flat)
        self.exp_avg_sq = torch.zeros_like(arena.flat)
        self._state = torch.zeros(2, device=dev, dtype=torch.int64)            # [steps taken, kernel scratch]
        self.lr_tensor: Optional[torch.Tensor] = torch.full((1,), float(lr), device=dev) if lr_on_device else None
        self.grad_scale = 1.0              # e.g. 1 / world_size after a summing all-reduce

    @property
    def steps(self) -> int:
        return int(self._state[0].item())

    def zero_grad(self, set_to_none: bool = False) -> None:
        self.arena.zero_grad()

    @torch.no_grad()
    def step(self) -> None:
        g = self.param_groups[0]
        a = self.arena
        sh = a.shadow if (a.shadow is not None and a.shadow.dtype == torch.bfloat16) else None
        check(_lib.load().murcl_adam_step(a.flat.data_ptr(), a.grad.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                                          None if sh is None else sh.data_ptr(), a.flat.numel(), g["lr"], g["betas"][0], g["betas"][1],
                                          g["eps"], g["weight_decay"], float(self.grad_scale),
                                          None if self.lr_tensor is None else self.lr_tensor.data_ptr(), self._state.data_ptr(),
                                          torch.cuda.current_stream().cuda_stream), "murcl_adam_step")
        # the shadow weights are fresh; every other cache keyed on the parameters (bf16 casts outside the arena, the
        # split-precision planes of the fp32 mode) is dropped, as ParamArena.refresh() does
        if a.shadow is not None and sh is None:
            ops.cast_into(a.flat, a.shadow)
        ops.invalidate_weight_cache()

    # torch.optim-style checkpointing of the moments (flat layout of the arena)
    def state_dict(self) -> dict:
        return {"exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(), "step": self.steps,
                "param_groups": [{k: v for k, v in self.param_groups[0].items() if k != "params"}]}

    def load_state_dict(self, sd: dict) -> None:
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self._state[0] = int(sd["step"])
        self._state[1] = 0
        for k, v in sd.get("param_groups", [{}])[0].items():
            self.param_groups[0][k] = v
