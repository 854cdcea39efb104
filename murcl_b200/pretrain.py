"""One MuRCL pre-training optimiser step on the B200 path (train_MuRCL.py:235-298).

This is the reference's call-site sequence - draw actions, ``get_feats``, ``mixup``, MIL aggregator,
``Full_layer``, ``NT_Xent``, rewards - written against the CSR store so that both views of a patch-step
are packed by ONE select + ONE gather/mixup launch.  The models, the projection head, the loss and the
PPO object are the drop-in modules, used through the same methods the reference trainer calls.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

from . import ops
from .csr import BagStore

Draw = Tuple[Sequence[torch.Tensor], Sequence[torch.Tensor], Sequence[torch.Tensor]]  # (actions, lams, perms) per view


def draw_patch_step(B: int, K: int, alpha: float, device, actions: Optional[List[torch.Tensor]] = None) -> Draw:
    """Random draws of one patch-step in the reference's RNG order: both action tensors
    (train_MuRCL.py:235,256-258), then per view ``rand`` + ``randperm`` (datasets.py:266-267)."""
    if actions is None:
        actions = [torch.rand((B, K), device=device) for _ in range(2)]
    lams, perms = [], []
    for _ in range(2):
        lams.append(alpha + torch.rand(size=(B, 1), device=device) * (1 - alpha))
        perms.append(torch.randperm(B, device=device))
    return actions, lams, perms


def pack_views(store: BagStore, draw: Draw, feat_size: int, out_dtype: torch.dtype, slot_bag=None) -> List[torch.Tensor]:
    """Both views of a patch-step through one packer pass: ``2B`` output slots over ``B`` bags."""
    actions, lams, perms = draw
    B = store.num_bags
    if slot_bag is None:
        slot_bag = torch.arange(B, dtype=torch.int32, device=store.device).repeat(2)
    act = torch.cat([a.to(torch.float32) for a in actions], 0)
    lam = torch.cat([l.reshape(-1) for l in lams], 0)
    perm = torch.cat([perms[0].reshape(-1), perms[1].reshape(-1) + B], 0)     # each view mixes within itself
    x = store.pack(act, feat_size, lam, perm, out_dtype, slot_bag)
    return [x[:B], x[B:]]


def pretrain_step(store: BagStore, model, fc, criterion, *, T: int = 6, feat_size: int = 1024, alpha: float = 0.9,
                  stage: int = 1, ppo=None, memories=None, draws: Optional[Sequence[Draw]] = None,
                  precision: Optional[str] = None, backward: bool = True):
    """Returns ``(loss, per-step losses)``.  ``stage`` follows train_MuRCL.py: 1 = random actions, 3 = PPO actor
    chooses the actions of patch-steps >= 1 (the actor is not updated); stage 2 (actor only) is not part of the
    MIL fwd+bwd hot path.  ``draws`` injects the random numbers (parity tests)."""
    if stage not in (1, 3):
        raise NotImplementedError("pretrain_step implements train stages 1 and 3")
    B, K, dev = store.num_bags, store.K, store.device
    dt = ops.storage_dtype(precision or ops.default_precision())
    slot_bag = torch.arange(B, dtype=torch.int32, device=dev).repeat(2)
    losses = []
    states = None
    sim_last = None
    for t in range(T):
        if draws is not None:
            draw = draws[t]
        else:
            actions = None
            if stage == 3 and t >= 1:
                actions = [ppo.select_action(s, m, restart_batch=(t == 1)) for s, m in zip(states, memories)]
            draw = draw_patch_step(B, K, alpha, dev, actions)
        x_views = pack_views(store, draw, feat_size, dt, slot_bag)
        outputs, states = model(x_views)
        outputs = [fc(o, restart=(t == 0)) for o in outputs]
        loss = criterion(outputs[0], outputs[1])
        losses.append(loss)
        sim = criterion.last_cosine.view(1, -1)          # by-product of the loss kernel (train_MuRCL.py:253,282)
        if t >= 1 and memories is not None:
            reward = sim_last - sim
            for m in memories:
                m.rewards.append(reward)
        sim_last = sim
    total = sum(losses) / T
    if backward:
        total.backward()
    if memories is not None:
        for m in memories:
            m.clear_memory()
    return total.detach(), [l.detach() for l in losses]
