"""One MuRCL pre-training optimiser step on the B200 path (train_MuRCL.py:235-298).

This is the reference's call-site sequence - draw actions, ``get_feats``, ``mixup``, MIL aggregator,
``Full_layer``, ``NT_Xent``, rewards - written against the CSR store so that both views of a patch-step
are packed by ONE select + ONE gather/mixup launch.  The models, the projection head, the loss and the
PPO object are the drop-in modules, used through the same methods the reference trainer calls.
"""
from __future__ import annotations

import contextlib
import os
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib, ops
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


def draw_step_batched(T: int, B: int, K: int, alpha: float, device, random_actions: bool):
    """Every random draw of one optimiser step in a handful of launches: ``(actions [T or 1, 2B, K], lam [T, 2B],
    perm [T, 2B] int32, order [T, 2B] int32)`` - rows ``[0, B)`` of a patch-step belong to view 0, ``[B, 2B)`` to view 1, each view's
    permutation stays inside its own half (``pack_views``).  Same distributions as ``draw_patch_step`` (uniform actions,
    ``lam ~ alpha + U(0,1)(1 - alpha)``, uniformly random permutations - the arg-sort of i.i.d. uniform keys), but the
    draws do not depend on anything the step computes, so they are issued once: per-step ``rand`` / ``randperm`` / ``cat``
    calls are ~20 dependent 2-3 us launches per patch-step on the step's critical path.  ``random_actions=False`` (stages
    2 / 3) draws only the first patch-step's actions; the actor chooses the rest."""
    lam = torch.rand((T, 2 * B), device=device).mul_(1.0 - alpha).add_(alpha)
    perm = torch.rand((T, 2, B), device=device).argsort(dim=2).to(torch.int32)
    perm[:, 1] += B
    act = torch.rand((T if random_actions else 1, 2 * B, K), device=device)
    perm = perm.view(T, 2 * B)
    from .csr import perm_cycle_order
    return act, lam, perm, perm_cycle_order(perm)       # + the gathers' slot walk order (csr.gather_rows_padded), all T at once


def pack_views(store: BagStore, draw: Draw, feat_size: int, out_dtype: torch.dtype, slot_bag=None) -> torch.Tensor:
    """Both views of a patch-step through one packer pass: ``2B`` output slots over ``B`` bags.  Returns the
    ``[2B, feat_size, D]`` tensor; rows ``[0,B)`` are view 0, ``[B,2B)`` view 1."""
    actions, lams, perms = draw
    if slot_bag is None:
        slot_bag = torch.arange(store.num_bags, dtype=torch.int32, device=store.device).repeat(2)
    B = slot_bag.numel() // 2
    act = torch.cat([a.to(torch.float32) for a in actions], 0)
    lam = torch.cat([l.reshape(-1) for l in lams], 0)
    perm = torch.cat([perms[0].reshape(-1), perms[1].reshape(-1) + B], 0)     # each view mixes within itself
    return store.pack(act, feat_size, lam, perm, out_dtype, slot_bag)


def _tape_decoder(model, tape):
    """The aggregator whose ``decoder`` layer the tape can take over (ABMIL behind the CL wrapper, same precision, bags not
    row-sharded), or None."""
    from .dropin.abmil import ABMIL
    from .dropin.cl import CL
    enc = getattr(model, "encoder", None) if isinstance(model, CL) else None
    if tape is None or not isinstance(enc, ABMIL) or enc.shard_rows or os.environ.get("MURCL_TAPE_DECODER", "1") == "0":
        return None
    if ops.storage_dtype(enc.precision or ops.default_precision()) != tape.dt or enc.decoder[0].out_features != tape.F:
        return None
    return enc


def _encode_stacked(model, x_all: torch.Tensor, n_views: int = 2, tape=None):
    """``encode_views`` that also hands back the un-split ``[n_views * B, F]`` encoder output (None when the model is not
    the CL wrapper): the recurrent-head tape takes it as one input.  With a tape that has taken over the aggregator's decoder
    layer (``HeadTape.attach_decoder``) the pooled vectors go through ``tape.decode``."""
    from .dropin.cl import CL
    B = x_all.shape[0] // n_views
    if tape is not None and tape.dec is not None:
        out = tape.decode(model.encoder.pooled(x_all))
        outs = [out[v * B:(v + 1) * B] for v in range(n_views)]
        return out, outs, outs
    if isinstance(model, CL):
        out = model.encoder(x_all)[0]
        outs = [out[v * B:(v + 1) * B] for v in range(n_views)]
        return out, outs, [o.detach() for o in outs]
    outs, states = model([x_all[v * B:(v + 1) * B] for v in range(n_views)])
    return None, outs, states


def encode_views(model, x_all: torch.Tensor, n_views: int = 2):
    """``model(x_views)`` of train_MuRCL.py:242,271 -> ``(outputs, detached states)``.  Bags are independent, so
    when the model is the CL wrapper all views go through its encoder in ONE batched call (half the launches,
    twice the rows per GEMM) and are split afterwards; any other model is called with the list of views."""
    from .dropin.cl import CL
    B = x_all.shape[0] // n_views
    if isinstance(model, CL):            # not duck-typed on `.encoder`: the drop-in ABMIL has an attribute of that name too
        out = model.encoder(x_all)[0]
        outs = [out[v * B:(v + 1) * B] for v in range(n_views)]
        return outs, [o.detach() for o in outs]
    return model([x_all[v * B:(v + 1) * B] for v in range(n_views)])


def pretrain_step(store: BagStore, model, fc, criterion, *, T: int = 6, feat_size: int = 1024, alpha: float = 0.9,
                  stage: int = 1, ppo=None, memories=None, draws: Optional[Sequence[Draw]] = None,
                  precision: Optional[str] = None, backward: bool = True, eps=None, keep_memory: bool = False,
                  slot_bag: Optional[torch.Tensor] = None, after_head_backward=None, rng: str = "reference",
                  overlap_heads: Optional[bool] = None):
    """Returns ``(loss, per-step losses)``.  ``stage`` follows train_MuRCL.py: 1 = random actions; 3 = the PPO actor
    chooses the actions of patch-steps >= 1 (the actor is not updated, :292-295); 2 = the same rollout under ``no_grad``
    with the MIL model frozen, then ``ppo.update(m)`` for each view's memory instead of the optimiser step (:244-247,
    :296-298).  ``draws`` injects the random numbers (parity tests): ``draws[t] = (actions | None, lams, perms)`` -
    ``None`` actions at ``t >= 1`` in stages 2/3 mean "ask the actor", with ``eps[t]`` (one ``[B, K]`` tensor per view) as
    its Gaussian draws.  ``keep_memory`` leaves the rollout in ``memories`` (the reference clears it, :301-302).
    ``slot_bag`` (int32 ``[2B]`` on the device: the store's bag index of every packed slot, view 0 then view 1) selects
    the step's B slides out of a larger resident store (``csr.ResidentSlides``); default: all bags of ``store``.
    ``after_head_backward`` (callable) runs once the projection head's parameter gradients are complete - i.e. right after
    the recurrent-head tape's batched backward and before the aggregators' - so that a data-parallel trainer can start
    exchanging them (``ParamArena.allreduce(..., async_op=True)``) under the rest of the backward pass.
    ``rng`` (used when ``draws`` is None): ``"reference"`` issues the random draws patch-step by patch-step in the
    reference's call order (train_MuRCL.py:235,256-258; datasets.py:266-267); ``"batched"`` draws everything the step
    needs up front (``draw_step_batched``: same distributions, ~110 fewer small launches per step).
    ``overlap_heads`` (default: on with the recurrent-head tape, ``MURCL_OVERLAP_HEADS=0`` turns it off): the projection
    head and the loss of patch-step t (``Full_layer`` on both views, NT-Xent, the reward) run on a side stream.  Nothing
    the NEXT patch-step needs depends on them - the actor reads the bag embeddings, not the loss - so their ~15 short,
    dependent launches overlap the actor's own chain and the window selection instead of preceding them; the streams
    join before the losses are summed."""
    if rng not in ("reference", "batched"):
        raise ValueError("rng must be 'reference' or 'batched'")
    if stage not in (1, 2, 3):
        raise ValueError("train_stage must be 1, 2 or 3")
    if stage != 1 and (ppo is None or memories is None):
        raise ValueError("stages 2 and 3 need the PPO object and one Memory per view")
    K, dev = store.K, store.device
    dt = ops.storage_dtype(precision or ops.default_precision())
    if slot_bag is None:
        B = store.num_bags
        slot_bag = torch.arange(B, dtype=torch.int32, device=dev).repeat(2)
    else:
        if slot_bag.dtype != torch.int32 or slot_bag.dim() != 1 or slot_bag.numel() % 2:
            raise ValueError("slot_bag must be an int32 vector of 2 * B store indices")
        B = slot_bag.numel() // 2
    losses = []
    states = None
    sim_last = None
    tape = None
    if (stage != 2 and backward and getattr(fc, "fc_rnn", False) and hasattr(fc, "forward_views")
            and os.environ.get("MURCL_DISABLE_HEADTAPE", "0") != "1"):
        # Full_layer over the T x 2 calls of this step on the recurrent-head tape: batched backward (headtape.py)
        from .headtape import HeadTape
        tape = HeadTape(fc, T, 2, B, dev, getattr(fc, "precision", None) or precision or ops.default_precision())
        dec_owner = _tape_decoder(model, tape)
        if dec_owner is not None:        # the aggregator's decoder layer joins the tape: one batched backward for all T calls
            tape.attach_decoder(dec_owner.decoder[0].weight, dec_owner.decoder[0].bias)
    if overlap_heads is None:
        overlap_heads = tape is not None and os.environ.get("MURCL_OVERLAP_HEADS", "1") != "0"
    side = head_stream(dev) if (overlap_heads and torch.device(dev).type == "cuda") else None
    pre = draw_step_batched(T, B, K, alpha, dev, stage == 1) if (draws is None and rng == "batched") else None
    grad_mode = torch.no_grad() if stage == 2 else torch.enable_grad()
    with grad_mode:
        for t in range(T):
            actions = lams = perms = None
            if draws is not None:
                actions, lams, perms = draws[t]
            if actions is None and stage != 1 and t >= 1:
                e = None if eps is None else eps[t]
                if hasattr(ppo, "select_action_views"):
                    actions = ppo.select_action_views(states, memories, restart_batch=(t == 1), eps=e)
                else:
                    actions = [ppo.select_action(s, m, restart_batch=(t == 1)) for s, m in zip(states, memories)]
            if pre is not None:
                act_all = pre[0][t if stage == 1 else 0] if actions is None else torch.cat(list(actions), 0)
                x_all = store.pack(act_all, feat_size, pre[1][t], pre[2][t], dt, slot_bag, order=pre[3][t])
            else:
                if lams is None:
                    draw = draw_patch_step(B, K, alpha, dev, actions)
                else:
                    if actions is None:
                        actions = [torch.rand((B, K), device=dev) for _ in range(2)]
                    draw = (actions, lams, perms)
                x_all = pack_views(store, draw, feat_size, dt, slot_bag)
            out_all, outputs, states = _encode_stacked(model, x_all, tape=tape)
            if side is not None:
                side.wait_stream(torch.cuda.current_stream())           # fork: the bag embeddings are complete
            with (torch.cuda.stream(side) if side is not None else contextlib.nullcontext()):
                if tape is not None:
                    outputs = tape.forward_views(outputs, restart=(t == 0), stacked=out_all)
                elif hasattr(fc, "forward_views"):
                    outputs = fc.forward_views(outputs, restart=(t == 0))
                else:
                    outputs = [fc(o, restart=(t == 0)) for o in outputs]
                loss = criterion(outputs[0], outputs[1])
                losses.append(loss)
                sim = criterion.last_cosine.view(1, -1)      # by-product of the loss kernel (train_MuRCL.py:253,282)
                if t >= 1 and memories is not None:
                    reward = sim_last - sim
                    for m in memories:
                        m.rewards.append(reward)
                sim_last = sim
    if side is not None:
        torch.cuda.current_stream().wait_stream(side)                   # join: every loss / reward is complete
    total = sum(losses) / T
    if stage == 2:
        for m in memories:
            ppo.update(m)
    elif backward and tape is not None:
        # loss -> projections (the loss kernels' own tiny graph), projections -> bag embeddings (the tape, one batched pass),
        # bag embeddings -> everything upstream (the aggregators' graphs, one per patch-step)
        dzs = torch.autograd.grad(total, tape.z_leaves)
        if side is not None:
            torch.cuda.current_stream().wait_stream(side)      # the loss nodes' backward ran where their forward did
        # the head's weight-gradient GEMMs go to the side stream unless a data-parallel hook wants those gradients right away
        wgrad_side = side if (after_head_backward is None and os.environ.get("MURCL_OVERLAP_HEAD_WGRAD", "1") != "0") else None
        d_outs = tape.backward(dzs, side=wgrad_side)
        if after_head_backward is not None:
            after_head_backward()
        torch.autograd.backward(tape.grad_roots, d_outs)
        if wgrad_side is not None:
            torch.cuda.current_stream().wait_stream(wgrad_side)
    elif backward:
        total.backward()
    if memories is not None and not keep_memory:
        for m in memories:
            m.clear_memory()
    # Full_layer.hidden carries its autograd graph across the T patch-steps (rlmil.py:216-219); the next optimiser step
    # restarts it (restart=True at t == 0), so once the backward has run the graph is dead weight: dropping it releases the
    # step's activations now and lets autograd rebuild its leaf nodes on whatever stream the next step runs on
    hid = getattr(fc, "hidden", None)
    if isinstance(hid, torch.Tensor) and hid.requires_grad:
        fc.hidden = hid.detach()
    return total.detach(), [l.detach() for l in losses]


_HEAD_STREAMS = {}


def head_stream(device) -> torch.cuda.Stream:
    """The per-device side stream of the projection-head / loss chain (high priority: its kernels are a few CTAs each and
    should not queue behind the pending CTAs of a persistent GEMM)."""
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    st = _HEAD_STREAMS.get(key)
    if st is None:
        st = torch.cuda.Stream(device=key, priority=-1)
        _HEAD_STREAMS[key] = st
    return st


class GraphedStep:
    """Captures ``step_fn`` (zero_grad + pretrain_step + optimiser) into one CUDA graph and replays it.

    One optimiser step issues ~700 short kernels; launched from Python they are CPU-bound (~40 us each).  All of
    them - the C-ABI kernels, torch's RNG draws, autograd's accumulations, Adam(capturable=True), NCCL
    collectives - are stream-ordered, so the whole step replays as a single graph launch.  The bag store and
    every tensor the step touches keep fixed addresses (torch's graph-private pool)."""

    def __init__(self, step_fn, warmup: int = 2, pool=None, capture_error_mode: str = "global"):
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                step_fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        # capture on the warm-up stream: autograd remembers the stream a leaf's gradient node was created on, and a
        # node that survived the warm-up would otherwise make the captured backward wait on an uncaptured stream
        with torch.cuda.graph(self.graph, pool=pool, stream=side, capture_error_mode=capture_error_mode):
            self.loss = step_fn()
        self.launches = _lib.launch_count() - n0          # libmurcl_b200 kernels per replay

    def pool(self):
        return self.graph.pool()

    def __call__(self) -> torch.Tensor:
        self.graph.replay()
        return self.loss
