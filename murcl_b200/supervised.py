"""Supervised RLMIL steps on the B200 path: the call-site sequences of
``train_CLAM`` (train_RLMIL.py:290-407) and ``train_DSMIL`` (train_RLMIL.py:475-604) written against the CSR store.

Per patch-step: select + gather (``get_feats``), the MIL aggregator, ``Full_layer`` (hidden state carried across the T
patch-steps), the classification loss, and the RL reward = change of the softmax confidence of the true class
(train_RLMIL.py:344,370-371).  Loss = mean over the T patch-steps (:377), backward.  The models, the head and the PPO
object are the drop-in modules; the cross-entropy over the ``[B, n_classes]`` logits is O(B) glue in torch.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.nn.functional as F

from . import ops
from .csr import BagStore


def _confidence(logits: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    return torch.softmax(logits.detach(), 1).gather(1, labels.view(-1, 1)).view(1, -1)


def _actions(t: int, stage: int, B: int, K: int, dev, ppo, memory, states, draws):
    if draws is not None:
        return draws[t]
    if t == 0 or stage == 1 or ppo is None:
        return torch.rand((B, K), device=dev)                       # train_RLMIL.py:325,348
    return ppo.select_action(states, memory, restart_batch=(t == 1))   # :350-353


def clam_step(store: BagStore, model, fc, labels: torch.Tensor, *, T: int = 6, feat_size: int = 1024, bag_weight: float = 0.7,
              stage: int = 1, ppo=None, memory=None, draws: Optional[Sequence[torch.Tensor]] = None,
              precision: Optional[str] = None, backward: bool = True):
    """``train_CLAM``: loss_t = bag_weight * CE(fc(M_t), y) + (1 - bag_weight) * instance_loss_t (train_RLMIL.py:336).
    Returns (mean loss, list of per-step logits)."""
    B, K, dev = store.num_bags, store.K, store.device
    dt = ops.storage_dtype(precision or ops.default_precision())
    labels = labels.to(dev).view(-1)
    losses, logits_all, states, conf_last = [], [], None, None
    for t in range(T):
        act = _actions(t, stage, B, K, dev, ppo, memory, states, draws)
        x = store.pack(act, feat_size, out_dtype=dt)
        pooled, states, res = model(x, label=labels, instance_eval=True)
        inst = res["instance_loss"] if isinstance(res, dict) else sum(r["instance_loss"] for r in res) / len(res)
        logits = fc(pooled, restart=(t == 0))
        losses.append(bag_weight * F.cross_entropy(logits, labels) + (1 - bag_weight) * inst)
        logits_all.append(logits.detach())
        conf = _confidence(logits, labels)
        if t >= 1 and memory is not None:
            memory.rewards.append(conf - conf_last)
        conf_last = conf
    loss = sum(losses) / T
    if backward:
        loss.backward()
    return loss.detach(), logits_all


def dsmil_step(store: BagStore, model, fc, labels: torch.Tensor, *, T: int = 6, feat_size: int = 1024, stage: int = 1,
               ppo=None, memory=None, draws: Optional[Sequence[torch.Tensor]] = None, precision: Optional[str] = None,
               backward: bool = True):
    """``train_DSMIL``: the bag embedding is the class-mean of the ``[C, D]`` bag tensor, the instance branch is the
    max-instance score; loss_t = 0.5 * CE(fc(bag), y) + 0.5 * CE(max-instance, y) (train_RLMIL.py:514-529)."""
    B, K, dev = store.num_bags, store.K, store.device
    dt = ops.storage_dtype(precision or ops.default_precision())
    labels = labels.to(dev).view(-1)
    losses, logits_all, states, conf_last = [], [], None, None
    for t in range(T):
        act = _actions(t, stage, B, K, dev, ppo, memory, states, draws)
        x = store.pack(act, feat_size, out_dtype=dt)
        bags = [x[b:b + 1] for b in range(B)] if B > 1 else x
        classes, bag, bag_det = model(bags)
        states = bag_det.mean(1)                                           # train_RLMIL.py:515
        if isinstance(classes, list):
            inst_max = torch.stack([c.max(0).values for c in classes], 0)  # :516 (per bag)
        else:
            inst_max = classes.max(0).values.view(1, -1)
        logits = fc(bag.mean(1), restart=(t == 0))                         # :517-518
        losses.append(0.5 * F.cross_entropy(logits, labels) + 0.5 * F.cross_entropy(inst_max, labels))
        logits_all.append(logits.detach())
        conf = _confidence(logits, labels)
        if t >= 1 and memory is not None:
            memory.rewards.append(conf - conf_last)
        conf_last = conf
    loss = sum(losses) / T
    if backward:
        loss.backward()
    return loss.detach(), logits_all
