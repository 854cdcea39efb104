"""CPU oracle for the MuRCL per-slide MIL hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import it.  Nothing under ``murcl_b200/`` imports it; the product path
fails loudly when the CUDA library is missing.

It is a functional (stateless, state-dict driven) restatement of the reference
algorithms, written from their definitions - not a copy of the modules:

  * integer selection arithmetic in numpy float32/int32 (bit-exact contract),
  * floating point in torch fp32 on CPU (fp64 on request) so that autograd supplies
    the reference gradients.

Parity pinning: every function below is checked by ``tests/test_oracle_golden.py``
against fixtures under ``tests/golden/`` that were produced by running the *real*
reference modules (``/root/reference``) in the build container with
``tests/golden/make_golden.py``.  The reference itself ships no tests or golden
vectors, so the fixtures are the pin.

All ``file:line`` citations are relative to the upstream repository root.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
StateDict = Dict[str, Tensor]


# --------------------------------------------------------------------------------------
# 1. Patch selection / gather  (utils/datasets.py:274-308)
# --------------------------------------------------------------------------------------
def select_windows(cluster_sizes: Sequence[int], num_patch: int, actions: np.ndarray,
                   feat_size: int) -> Tuple[np.ndarray, np.ndarray]:
    """Per-cluster half-open rank window [start, stop) that the reference's Python slice picks.

    utils/datasets.py:285-291 evaluates, with torch's type promotion (int64 tensor x python
    float -> float32 tensor):
        size = rint_half_even( f32(n) * f32(feat_size / num_patch) )        (:285-287)
        l    = floor( f32(a) * f32(n - size) )                              (:290)
        r    = l + size                                                     (:291)
    and then slices the cluster's patch list ``c[l:r]`` (:294) with Python semantics:
    a negative ``l`` counts from the end, ``r >= 0`` always, both clamp to ``len(c)``.
    """
    n = np.asarray(cluster_sizes, dtype=np.int64)
    ratio = np.float32(feat_size / num_patch)              # double division, then cast
    size = np.rint(n.astype(np.float32) * ratio).astype(np.int32)
    a = np.asarray(actions, dtype=np.float32)
    left = np.floor(a * (n - size).astype(np.float32)).astype(np.int32)
    right = left + size
    start = np.where(left >= 0, np.minimum(left, n), np.maximum(n + left, 0))
    stop = np.where(right >= 0, np.minimum(right, n), np.maximum(n + right, 0))
    stop = np.maximum(stop, start)
    return start.astype(np.int64), stop.astype(np.int64)


def select_indices(clusters: List[List[int]], num_patch: int, actions: np.ndarray,
                   feat_size: int) -> np.ndarray:
    """Ascending patch ids kept for one bag (utils/datasets.py:292-305): union of the
    per-cluster windows, sorted, truncated to the first ``feat_size``."""
    start, stop = select_windows([len(c) for c in clusters], num_patch, actions, feat_size)
    picked: List[int] = []
    for c, s, e in zip(clusters, start, stop):
        picked.extend(c[int(s):int(e)])
    picked.sort()
    return np.asarray(picked[:feat_size], dtype=np.int64)


def get_feats(feat_list: List[Tensor], clusters_list: List[List[List[int]]], actions: Tensor,
              feat_size: int = 1024) -> Tuple[Tensor, List[np.ndarray]]:
    """Dense ``[B, feat_size, D]`` selection with zero padding at the end
    (utils/datasets.py:299-307).  Also returns the kept patch ids per bag."""
    act = actions.detach().cpu().numpy().astype(np.float32)
    rows, kept = [], []
    for i, feat in enumerate(feat_list):
        f = feat.reshape(-1, feat.shape[-1])
        idx = select_indices(clusters_list[i], f.shape[0], act[i], feat_size)
        out = torch.zeros(feat_size, f.shape[1], dtype=f.dtype)
        out[: len(idx)] = f[torch.from_numpy(idx)]
        rows.append(out)
        kept.append(idx)
    return torch.stack(rows, 0), kept


def mixup_apply(x: Tensor, lam: Tensor, perm: Tensor) -> Tensor:
    """out_i = lam_i * x_i + (1 - lam_i) * x_perm(i): two rounded products, one rounded add
    (utils/datasets.py:268-270)."""
    lam = lam.reshape(-1, 1, 1).to(x.dtype)
    return lam * x + (1 - lam) * x[perm]


def mixup(x: Tensor, alpha: float, generator: Optional[torch.Generator] = None):
    """RNG order of utils/datasets.py:266-267: ``rand(B,1)`` first, ``randperm(B)`` second."""
    b = x.shape[0]
    lam = alpha + torch.rand(b, 1, generator=generator) * (1 - alpha)
    perm = torch.randperm(b, generator=generator)
    return mixup_apply(x, lam, perm), lam, perm


# --------------------------------------------------------------------------------------
# 2. MIL aggregators
# --------------------------------------------------------------------------------------
def softmax_pool(scores: Tensor, h: Tensor, post_scale: float = 1.0) -> Tuple[Tensor, Tensor]:
    """p = softmax_N(scores) * post_scale ; M = p^T h.  scores [N], h [N, L]."""
    p = torch.softmax(scores, dim=0) * post_scale
    return p @ h, p


def abmil_bag(x: Tensor, sd: StateDict, masks: Optional[Dict[str, Tensor]] = None,
              relu_masks: Optional[Sequence[Tensor]] = None) -> Tensor:
    """One bag through ABMIL (models/abmil.py:35-45): 3x(Linear+ReLU) encoder (:12-21),
    tanh attention (:23-27), softmax over N then / sqrt(N) (:40-41), A.H (:42), Linear+ReLU
    decoder (:29-32,44).  ``fc`` (:33) is never applied.  x [N, D_in] -> [1, L].
    ``relu_masks`` (tests only): three boolean [N, L] tensors that REPLACE the encoder's ReLU decisions
    (h = z * mask).  A pre-activation within rounding distance of zero may fall on either side of the ReLU in
    two correct fp32 evaluations; gradients then differ by that unit's whole contribution, which says nothing
    about arithmetic accuracy.  Adopting the device's decisions makes the gradient comparison exact."""
    h = x
    for j, i in enumerate((0, 3, 6)):
        z = F.linear(h, sd[f"encoder.{i}.weight"], sd[f"encoder.{i}.bias"])
        h = F.relu(z) if relu_masks is None else z * relu_masks[j].to(z.dtype)
        if masks is not None and f"enc{j}" in masks:      # train-mode nn.Dropout (:15,18): multiplicative keep/(1-p)
            h = h * masks[f"enc{j}"]
    u = torch.tanh(F.linear(h, sd["attention.0.weight"], sd["attention.0.bias"]))
    s = F.linear(u, sd["attention.2.weight"], sd["attention.2.bias"]).squeeze(-1)
    m, _ = softmax_pool(s, h, 1.0 / math.sqrt(h.shape[0]))
    return F.relu(F.linear(m.unsqueeze(0), sd["decoder.0.weight"], sd["decoder.0.bias"]))


def abmil_attention(x: Tensor, sd: StateDict) -> Tensor:
    """The attention weights ABMIL pools with (models/abmil.py:37-41): softmax over N of the tanh-attention scores,
    then / sqrt(N).  x [N, D_in] -> [N]."""
    h = x
    for i in (0, 3, 6):
        h = F.relu(F.linear(h, sd[f"encoder.{i}.weight"], sd[f"encoder.{i}.bias"]))
    u = torch.tanh(F.linear(h, sd["attention.0.weight"], sd["attention.0.bias"]))
    s = F.linear(u, sd["attention.2.weight"], sd["attention.2.bias"]).squeeze(-1)
    return torch.softmax(s, dim=0) / math.sqrt(h.shape[0])


def abmil_forward(bags: Sequence[Tensor], sd: StateDict) -> Tensor:
    """models/abmil.py:47-62: per-bag loop, concatenated -> [B, L]."""
    return torch.cat([abmil_bag(b.reshape(-1, b.shape[-1]), sd) for b in bags], 0)


def clam_attention_scores(h: Tensor, sd: StateDict, prefix: str, gate: bool,
                          masks: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """models/clam.py:18-60: gated ``W_c(tanh(W_a h) * sigmoid(W_b h))`` or plain
    ``W_2 tanh(W_1 h)`` raw scores [N] (dropout off)."""
    if gate:
        a = torch.tanh(F.linear(h, sd[f"{prefix}.attention_a.0.weight"], sd[f"{prefix}.attention_a.0.bias"]))
        b = torch.sigmoid(F.linear(h, sd[f"{prefix}.attention_b.0.weight"], sd[f"{prefix}.attention_b.0.bias"]))
        if masks is not None:                               # Dropout(0.25) after each branch (clam.py:46-48)
            a, b = a * masks["attn_a"], b * masks["attn_b"]
        return F.linear(a * b, sd[f"{prefix}.attention_c.weight"], sd[f"{prefix}.attention_c.bias"]).squeeze(-1)
    keys = sorted(k for k in sd if k.startswith(f"{prefix}.module.") and k.endswith(".weight"))
    u = torch.tanh(F.linear(h, sd[keys[0]], sd[keys[0].replace("weight", "bias")]))
    if masks is not None:
        u = u * masks["attn_a"]
    return F.linear(u, sd[keys[1]], sd[keys[1].replace("weight", "bias")]).squeeze(-1)


def clam_instance_loss(p: Tensor, h: Tensor, sd: StateDict, label: int, n_classes: int,
                       k_sample: int, subtyping: bool):
    """models/clam.py:103-132,146-168.  Top-k / bottom-k of the *post-softmax* attention,
    Linear(512->2) per class head, CE(mean) against [1]*k+[0]*k (in class) or [0]*k
    (out of class, subtyping only); summed, / n_classes when subtyping."""
    total = h.new_zeros(())
    preds: List[int] = []
    targets: List[int] = []
    for c in range(n_classes):
        w, b = sd[f"instance_classifiers.{c}.weight"], sd[f"instance_classifiers.{c}.bias"]
        top = torch.topk(p, k_sample).indices
        if c == label:
            bot = torch.topk(-p, k_sample).indices
            rows = torch.cat([h[top], h[bot]], 0)
            tgt = torch.cat([torch.ones(k_sample), torch.zeros(k_sample)]).long()
        elif subtyping:
            rows, tgt = h[top], torch.zeros(k_sample).long()
        else:
            continue
        logits = F.linear(rows, w, b)
        total = total + F.cross_entropy(logits, tgt)
        preds.extend(logits.argmax(1).tolist())
        targets.extend(tgt.tolist())
    if subtyping:
        total = total / n_classes
    return total, np.asarray(preds), np.asarray(targets)


def clam_sb_bag(x: Tensor, sd: StateDict, *, gate: bool = True, dropout_layers: bool = False,
                label: Optional[int] = None, instance_eval: bool = False, n_classes: int = 2,
                k_sample: int = 8, subtyping: bool = False, attention_only: bool = False,
                masks: Optional[Dict[str, Tensor]] = None, p_select: Optional[Tensor] = None):
    """One bag through CLAM_SB in eval mode (models/clam.py:134-181).  ``dropout_layers``
    only shifts the index of the attention sub-module in the Sequential (:69-77): it is
    ``attention_net.3`` when the model was built with dropout=True and ``.2`` otherwise.
    Returns (M [1,512], results dict) or raw scores [1,N] when ``attention_only`` (:141-142).  ``p_select`` (tests of
    the reduced-precision mode) replaces the attention weights the instance loss RANKS by - the top-k indices carry no
    gradient (:107-110) - so that both sides evaluate the loss on the same instances; ``results["attention"]`` holds
    this bag's own post-softmax weights."""
    att = "attention_net.3" if dropout_layers else "attention_net.2"
    h = F.relu(F.linear(x, sd["attention_net.0.weight"], sd["attention_net.0.bias"]))
    if masks is not None:                                   # train mode: Dropout(0.25) after the fc ReLU (clam.py:70-71)
        h = h * masks["enc"]
    s = clam_attention_scores(h, sd, att, gate, masks)
    if attention_only:
        return s.unsqueeze(0)
    m, p = softmax_pool(s, h)
    results = {"attention": p.detach()}
    if instance_eval:
        rank_by = p.detach() if p_select is None else p_select
        loss, preds, targets = clam_instance_loss(rank_by, h, sd, int(label), n_classes, k_sample, subtyping)
        results.update({"instance_loss": loss, "inst_preds": preds, "inst_labels": targets})
    return m.unsqueeze(0), results


def dsmil_bag(x: Tensor, sd: StateDict) -> Tuple[Tensor, Tensor]:
    """DSMIL (models/dsmil.py:11-16,64-81): instance classifier c = W x + b (:15);
    critical instance per class = arg-max over N (row 0 of the descending sort, :71-73);
    Q = W_q x + b_q (:67); A = softmax_N(Q q_max^T / sqrt(128)) (:76-77); B = A^T (W_v x + b_v)
    (:66,78).  Returns (classes [N,C], bag [1,C,D_in])."""
    c = F.linear(x, sd["i_classifier.fc.0.weight"], sd["i_classifier.fc.0.bias"])
    v = F.linear(x, sd["b_classifier.v.1.weight"], sd["b_classifier.v.1.bias"])
    q = F.linear(x, sd["b_classifier.q.weight"], sd["b_classifier.q.bias"])
    crit = torch.argmax(c, dim=0)
    q_max = F.linear(x[crit], sd["b_classifier.q.weight"], sd["b_classifier.q.bias"])
    a = torch.softmax((q @ q_max.t()) / math.sqrt(q.shape[1]), dim=0)
    return c, (a.t() @ v).unsqueeze(0)


# --------------------------------------------------------------------------------------
# 3. NT-Xent (utils/losses.py:5-41)
# --------------------------------------------------------------------------------------
def nt_xent(z_i: Tensor, z_j: Tensor, temperature: float) -> Tensor:
    """Closed form of utils/losses.py:24-41.  With z = [z_i; z_j], s = cos(z_a, z_b)/tau
    (:27-29, CosineSimilarity eps 1e-8), the CE over [positive | all b != a, b != pos(a)]
    logits (:30-39) equals  logsumexp_{b != a} s_ab - s_{a,pos(a)}; summed and / 2B (:39-40)."""
    b = z_i.shape[0]
    z = torch.cat([z_i, z_j], 0)
    zn = z / z.norm(dim=1, keepdim=True).clamp_min(1e-8)
    s = (zn @ zn.t()) / temperature
    pos = torch.cat([torch.arange(b, 2 * b), torch.arange(0, b)])
    s_pos = s[torch.arange(2 * b), pos]
    s_masked = s.masked_fill(torch.eye(2 * b, dtype=torch.bool), float("-inf"))
    return (torch.logsumexp(s_masked, dim=1) - s_pos).sum() / (2 * b)


def pair_cosine(z_i: Tensor, z_j: Tensor) -> Tensor:
    """Per-bag reward signal ``torch.cosine_similarity`` (train_MuRCL.py:253,282)."""
    return F.cosine_similarity(z_i, z_j)


# --------------------------------------------------------------------------------------
# 4. Heads: Full_layer and the PPO actor (models/rlmil.py)
# --------------------------------------------------------------------------------------
def gru_cell(x: Tensor, h: Tensor, w_ih: Tensor, w_hh: Tensor, b_ih: Tensor, b_hh: Tensor) -> Tensor:
    """One nn.GRU time step, gate order (r, z, n) as torch packs them."""
    gi = F.linear(x, w_ih, b_ih)
    gh = F.linear(h, w_hh, b_hh)
    i_r, i_z, i_n = gi.chunk(3, 1)
    h_r, h_z, h_n = gh.chunk(3, 1)
    r = torch.sigmoid(i_r + h_r)
    z = torch.sigmoid(i_z + h_z)
    n = torch.tanh(i_n + r * h_n)
    return (1 - z) * n + z * h


def full_layer_step(x: Tensor, h_prev: Optional[Tensor], sd: StateDict) -> Tuple[Tensor, Tensor]:
    """Full_layer with fc_rnn=True (models/rlmil.py:208-220): one GRU step from ``h_prev``
    (zeros on restart) then Linear(hidden -> class_num).  Returns (logits, h_new)."""
    if h_prev is None:
        h_prev = x.new_zeros(x.shape[0], sd["rnn.weight_hh_l0"].shape[1])
    h = gru_cell(x, h_prev, sd["rnn.weight_ih_l0"], sd["rnn.weight_hh_l0"], sd["rnn.bias_ih_l0"], sd["rnn.bias_hh_l0"])
    return F.linear(h, sd["fc.weight"], sd["fc.bias"]), h


def actor_act(state: Tensor, h_prev: Optional[Tensor], sd: StateDict, action_std: float, eps: Tensor):
    """ActorCritic.act with policy_conv=False, training=True (models/rlmil.py:66-97):
    MLP 512->2048->hidden with ReLU (:40-45,76), GRU step (:78), sigmoid head (:82), Gaussian
    sample mean + std*eps with scale_tril=diag(action_std) (:84-86; ``action_var`` holds the
    std, :56), clip to [0,1] through two ReLUs (:88-89), log-prob of the clipped action (:90).
    ``eps`` is the standard-normal draw.  Returns (action, logprob, h_new, mean)."""
    s = actor_encode_state(state, sd)
    if h_prev is None:
        h_prev = s.new_zeros(s.shape[0], sd["gru.weight_hh_l0"].shape[1])
    h = gru_cell(s, h_prev, sd["gru.weight_ih_l0"], sd["gru.weight_hh_l0"], sd["gru.bias_ih_l0"], sd["gru.bias_hh_l0"])
    mean = torch.sigmoid(F.linear(h, sd["actor.0.weight"], sd["actor.0.bias"]))
    action = torch.clamp(mean + action_std * eps, 0.0, 1.0)
    k = mean.shape[1]
    logprob = (-0.5 * (((action - mean) / action_std) ** 2).sum(1)
               - k * math.log(action_std) - 0.5 * k * math.log(2 * math.pi))
    return action, logprob, h, mean


def actor_encode_state(state: Tensor, sd: StateDict) -> Tensor:
    """The state encoder of ActorCritic (models/rlmil.py:29-45).  policy_conv=False: flatten -> Linear(2048) -> ReLU ->
    Linear(hidden) -> ReLU (:40-45, :71-72).  policy_conv=True (recognised by the ``state_encoder.3`` keys): 1x1 convolution
    without bias over the [N, F, r, r] state -> ReLU -> NCHW flatten -> Linear(hidden) -> ReLU (:30-37, :73-74)."""
    if "state_encoder.3.weight" in sd:
        w = sd["state_encoder.0.weight"]
        c = F.relu(torch.einsum("nfhw,cf->nchw", state, w.reshape(w.shape[0], w.shape[1])))
        return F.relu(F.linear(c.flatten(1), sd["state_encoder.3.weight"], sd["state_encoder.3.bias"]))
    s = F.relu(F.linear(state.flatten(1), sd["state_encoder.0.weight"], sd["state_encoder.0.bias"]))
    return F.relu(F.linear(s, sd["state_encoder.2.weight"], sd["state_encoder.2.bias"]))


def actor_evaluate(states: Tensor, actions: Tensor, sd: StateDict, action_std: float):
    """ActorCritic.evaluate with policy_conv=False (models/rlmil.py:99-127): the stored states ``[T, B, ...]`` go
    through the state MLP (:109), one ``nn.GRU`` pass over the T steps from a ZERO hidden state (:112), the sigmoid
    actor head (:115) and the critic (:124); the action log-probability and the entropy are those of
    ``MultivariateNormal(mean, scale_tril=diag(action_std))`` (:117-122).  Returns (logprob, value, entropy), each [T, B]."""
    T, B = states.shape[0], states.shape[1]
    s = actor_encode_state(states.reshape((T * B,) + tuple(states.shape[2:])), sd).reshape(T, B, -1)
    h = s.new_zeros(B, sd["gru.weight_hh_l0"].shape[1])
    feats = []
    for t in range(T):
        h = gru_cell(s[t], h, sd["gru.weight_ih_l0"], sd["gru.weight_hh_l0"], sd["gru.bias_ih_l0"], sd["gru.bias_hh_l0"])
        feats.append(h)
    feat = torch.cat(feats, 0)
    mean = torch.sigmoid(F.linear(feat, sd["actor.0.weight"], sd["actor.0.bias"]))
    value = F.linear(feat, sd["critic.0.weight"], sd["critic.0.bias"])
    k = mean.shape[1]
    a = actions.reshape(T * B, -1)
    logprob = (-0.5 * (((a - mean) / action_std) ** 2).sum(1)
               - k * math.log(action_std) - 0.5 * k * math.log(2 * math.pi))
    entropy = torch.full_like(logprob, 0.5 * k * (1.0 + math.log(2 * math.pi)) + k * math.log(action_std))
    return logprob.view(T, B), value.view(T, B), entropy.view(T, B)


def ppo_returns(rewards: Sequence[Tensor], gamma: float) -> Tensor:
    """models/rlmil.py:153-162: discounted return per patch-step (each reward is ``[1, B]``, train_MuRCL.py:283-288),
    concatenated to ``[T, B]`` and normalised with the mean / UNBIASED std over all T*B entries (+1e-5)."""
    out, running = [], 0
    for r in reversed(list(rewards)):
        running = r + gamma * running
        out.insert(0, running)
    ret = torch.cat(out, 0)
    return (ret - ret.mean()) / (ret.std() + 1e-5)


def ppo_loss(sd: StateDict, states: Tensor, actions: Tensor, old_logprobs: Tensor, returns: Tensor, *,
             action_std: float, eps_clip: float) -> Tensor:
    """One epoch's scalar objective (models/rlmil.py:169-181): mean over [T, B] of
    ``-min(ratio A, clip(ratio) A) + 0.5 * MSE(value, returns) - 0.01 * entropy`` with ``A = returns - value.detach()``;
    the MSE is already a mean over all entries (a scalar broadcast into the sum)."""
    logprob, value, entropy = actor_evaluate(states, actions, sd, action_std)
    ratio = torch.exp(logprob - old_logprobs.detach())
    adv = returns - value.detach()
    surr = torch.min(ratio * adv, torch.clamp(ratio, 1 - eps_clip, 1 + eps_clip) * adv)
    return (-surr + 0.5 * F.mse_loss(value, returns) - 0.01 * entropy).mean()


def ppo_update(sd: StateDict, states: Tensor, actions: Tensor, old_logprobs: Tensor, rewards: Sequence[Tensor], *,
               action_std: float, lr: float = 0.0003, betas=(0.9, 0.999), gamma: float = 0.7, K_epochs: int = 1,
               eps_clip: float = 0.2):
    """PPO.update (models/rlmil.py:152-184): K_epochs Adam steps (``torch.optim.Adam(lr, betas)``, :141) on ``ppo_loss``.
    Returns (new state dict, gradients of the FIRST epoch, per-epoch losses)."""
    params = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    opt = torch.optim.Adam(list(params.values()), lr=lr, betas=betas)
    returns = ppo_returns(rewards, gamma)
    first, losses = None, []
    for _ in range(K_epochs):
        loss = ppo_loss(params, states, actions, old_logprobs, returns, action_std=action_std, eps_clip=eps_clip)
        opt.zero_grad()
        loss.backward()
        if first is None:
            first = {k: (p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p)) for k, p in params.items()}
        losses.append(loss.detach())
        opt.step()
    return {k: p.detach() for k, p in params.items()}, first, losses


# --------------------------------------------------------------------------------------
# 5. The pre-training step (train_MuRCL.py:235-298), used as the CPU baseline workload
# --------------------------------------------------------------------------------------
def pretrain_step(feat_list, clusters_list, sd_model: StateDict, sd_fc: StateDict, *, arch: str = "ABMIL",
                  T: int = 6, feat_size: int = 1024, alpha: float = 0.9, temperature: float = 1.0,
                  generator: Optional[torch.Generator] = None, backward: bool = True,
                  clam_kwargs: Optional[dict] = None):
    """Stage-1 semantics of one optimiser step: for each of T patch-steps and 2 views draw random
    actions (:235,256-258), ``get_feats`` (:237,266), ``mixup`` (:239,268), aggregate every bag
    (:242,271), ``Full_layer`` with the hidden state carried across patch-steps (:243,272),
    NT-Xent (:249,277); loss = mean over T (:291), backward (:294).  Returns (loss, grads)."""
    b = len(feat_list)
    k = len(clusters_list[0])
    params = {n: p.detach().clone().requires_grad_(backward) for n, p in {**{"m." + a: v for a, v in sd_model.items()},
                                                                              **{"f." + a: v for a, v in sd_fc.items()}}.items()}
    sm = {n[2:]: p for n, p in params.items() if n.startswith("m.")}
    sf = {n[2:]: p for n, p in params.items() if n.startswith("f.")}
    hidden = None  # ONE state shared by both views, as Full_layer.hidden is (models/rlmil.py:216,219)
    losses = []
    for t in range(T):
        # RNG order of train_MuRCL.py:235-239 / :256-268: both action draws, then per view rand+randperm
        acts = [torch.rand(b, k, generator=generator) for _ in range(2)]
        views = [get_feats(feat_list, clusters_list, a, feat_size)[0] for a in acts]
        views = [mixup(x, alpha, generator)[0] for x in views]
        outs = []
        for v, x in enumerate(views):
            if arch == "ABMIL":
                pooled = abmil_forward(list(x), sm)
            else:
                kw = clam_kwargs or {}
                pooled = torch.cat([clam_sb_bag(xb, sm, **kw)[0] for xb in x], 0)
            # train_MuRCL.py:243,272 call the same Full_layer object for view 0 then view 1: with
            # restart=True both start from zeros; afterwards each call continues from the hidden state
            # the PREVIOUS CALL left behind (the other view's), not from its own view's history.
            z, hidden = full_layer_step(pooled, None if t == 0 else hidden, sf)
            outs.append(z)
        losses.append(nt_xent(outs[0], outs[1], temperature))
    loss = sum(losses) / T
    grads = None
    if backward:
        loss.backward()
        grads = {n: p.grad for n, p in params.items()}
    return loss.detach(), grads


def pretrain_step_stage3(feat_list, clusters_list, sd_model: StateDict, sd_fc: StateDict, sd_actor: StateDict, *,
                         first_actions: Sequence[Tensor], lams, perms, eps, action_std: float, T: int,
                         feat_size: int = 1024, temperature: float = 1.0, backward: bool = True):
    """Stage-3 semantics of train_MuRCL.py:235-298 for ABMIL with every random draw injected: ``first_actions`` (2 x
    [B, K]) are the uniform actions of patch-step 0 (:235); ``lams[t][v]`` / ``perms[t][v]`` the mixup draws (:239,268);
    ``eps[t][v]`` (t >= 1) the actor's standard-normal draws.  From patch-step 1 on each view's actions come from
    ``actor_act`` on the DETACHED bag embedding of the previous patch-step (:260-265; cl.py:15), the actor's GRU state
    restarting from zeros at patch-step 1 (``restart_batch``) and carried per view afterwards.  Rewards are
    ``similarity_last - similarity`` (:282-283).  Returns a dict: loss, losses, grads, actions, logprobs, rewards,
    states (the rollout a stage-2 ``ppo_update`` would consume)."""
    params = {n: p.detach().clone().requires_grad_(backward) for n, p in {**{"m." + a: v for a, v in sd_model.items()},
                                                                              **{"f." + a: v for a, v in sd_fc.items()}}.items()}
    sm = {n[2:]: p for n, p in params.items() if n.startswith("m.")}
    sf = {n[2:]: p for n, p in params.items() if n.startswith("f.")}
    hidden, actor_h = None, [None, None]
    states = None
    losses, rewards, actions_log = [], [], []
    roll = [dict(states=[], actions=[], logprobs=[]) for _ in range(2)]
    sim_last = None
    for t in range(T):
        if t == 0:
            acts = [a.float() for a in first_actions]
        else:
            acts = []
            for v in range(2):
                a, lp, h, _mean = actor_act(states[v], None if t == 1 else actor_h[v], sd_actor, action_std, eps[t][v])
                actor_h[v] = h
                roll[v]["states"].append(states[v]); roll[v]["actions"].append(a); roll[v]["logprobs"].append(lp)
                acts.append(a)
        actions_log.append(acts)
        outs, new_states = [], []
        for v in range(2):
            x, _ = get_feats(feat_list, clusters_list, acts[v], feat_size)
            x = mixup_apply(x, lams[t][v], perms[t][v])
            pooled = abmil_forward(list(x), sm)
            new_states.append(pooled.detach())
            z, hidden = full_layer_step(pooled, None if t == 0 else hidden, sf)
            outs.append(z)
        states = new_states
        losses.append(nt_xent(outs[0], outs[1], temperature))
        sim = pair_cosine(outs[0], outs[1]).view(1, -1)
        if t >= 1:
            rewards.append((sim_last - sim).detach())
        sim_last = sim
    loss = sum(losses) / T
    grads = None
    if backward:
        loss.backward()
        grads = {n: p.grad for n, p in params.items()}
    return dict(loss=loss.detach(), losses=[l.detach() for l in losses], grads=grads, actions=actions_log,
                logprobs=[torch.stack(r["logprobs"], 0) for r in roll], rewards=rewards,
                states=[torch.stack(r["states"], 0) for r in roll], roll_actions=[torch.stack(r["actions"], 0) for r in roll])
