"""Seeded synthetic inputs shaped like the reference's data.

Replaces the reference's file IO (``utils/datasets.py:12-165``: ``.npz`` features + JSON
cluster lists) for tests and benchmarks: there are no slides in this environment.  All
generators run on the CPU torch generator so a (seed, shape) pair names the same numbers
in the build container and on the GPU box.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import numpy as np
import torch

SEED = 985  # the reference's default seed (train_MuRCL.py:473)


def gen(seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    return g


def make_bag(n: int, d: int, k: int, g: torch.Generator) -> Tuple[torch.Tensor, List[List[int]], torch.Tensor]:
    """One slide: features ``[n, d]`` fp32, post-ReLU-like (non-negative), plus the inverted
    cluster lists in ascending patch order as ``features_clustering.py:19-25`` writes them and
    the per-patch cluster label they were built from."""
    feats = torch.clamp_min(0.5 * torch.randn(n, d, generator=g) + 0.3, 0.0)
    labels = torch.randint(0, k, (n,), generator=g)
    lab = labels.numpy()
    clusters = [np.nonzero(lab == j)[0].tolist() for j in range(k)]
    return feats, clusters, labels.to(torch.int32)


def make_bags(sizes, d: int, k: int, seed: int = SEED):
    g = gen(seed)
    feats, clusters, labels = [], [], []
    for n in sizes:
        f, c, l = make_bag(int(n), d, k, g)
        feats.append(f)
        clusters.append(c)
        labels.append(l)
    return feats, clusters, labels


def _linear(g, out_f, in_f, scale=1.0, bias_scale=1.0):
    bound = 1.0 / math.sqrt(in_f)
    w = (torch.rand(out_f, in_f, generator=g) * 2 - 1) * bound * scale
    b = (torch.rand(out_f, generator=g) * 2 - 1) * bound * bias_scale
    return w, b


def abmil_state(dim_in: int, L: int = 512, D: int = 128, dim_out: int = 2, seed: int = SEED,
                peak: float = 6.0) -> Dict[str, torch.Tensor]:
    """State dict with the reference's ABMIL parameter names (models/abmil.py:12-33).  ``peak``
    scales the attention layers so the softmax is not near-uniform (a near-uniform softmax
    hides online-softmax bugs)."""
    g = gen(seed)
    sd = {}
    for i, fin in ((0, dim_in), (3, L), (6, L)):
        sd[f"encoder.{i}.weight"], sd[f"encoder.{i}.bias"] = _linear(g, L, fin, 1.4)
    sd["attention.0.weight"], sd["attention.0.bias"] = _linear(g, D, L, peak)
    sd["attention.2.weight"], sd["attention.2.bias"] = _linear(g, 1, D, peak)
    sd["decoder.0.weight"], sd["decoder.0.bias"] = _linear(g, L, L, 1.4)
    sd["fc.weight"], sd["fc.bias"] = _linear(g, dim_out, L)
    return sd


def clam_state(in_dim: int, size_arg: str = "small", gate: bool = True, dropout: bool = False,
               n_classes: int = 2, seed: int = SEED, peak: float = 6.0) -> Dict[str, torch.Tensor]:
    """State dict with CLAM_SB's parameter names (models/clam.py:66-80)."""
    g = gen(seed)
    L, D = 512, (256 if size_arg == "small" else 384)
    att = "attention_net.3" if dropout else "attention_net.2"
    sd = {}
    sd["attention_net.0.weight"], sd["attention_net.0.bias"] = _linear(g, L, in_dim, 1.4)
    if gate:
        sd[f"{att}.attention_a.0.weight"], sd[f"{att}.attention_a.0.bias"] = _linear(g, D, L, peak)
        sd[f"{att}.attention_b.0.weight"], sd[f"{att}.attention_b.0.bias"] = _linear(g, D, L, peak)
        sd[f"{att}.attention_c.weight"], sd[f"{att}.attention_c.bias"] = _linear(g, 1, D, peak)
    else:
        last = 3 if dropout else 2
        sd[f"{att}.module.0.weight"], sd[f"{att}.module.0.bias"] = _linear(g, D, L, peak)
        sd[f"{att}.module.{last}.weight"], sd[f"{att}.module.{last}.bias"] = _linear(g, 1, D, peak)
    sd["classifiers.weight"], sd["classifiers.bias"] = _linear(g, n_classes, L)
    for c in range(n_classes):
        sd[f"instance_classifiers.{c}.weight"], sd[f"instance_classifiers.{c}.bias"] = _linear(g, 2, L, 4.0)
    return sd


def dsmil_state(dim_feat: int, num_classes: int = 2, seed: int = SEED, peak: float = 1.5) -> Dict[str, torch.Tensor]:
    """State dict with MILNet's parameter names (models/dsmil.py:6-62,103-119)."""
    g = gen(seed)
    sd = {}
    sd["i_classifier.fc.0.weight"], sd["i_classifier.fc.0.bias"] = _linear(g, num_classes, dim_feat, 2.0)
    sd["b_classifier.q.weight"], sd["b_classifier.q.bias"] = _linear(g, 128, dim_feat, peak)
    sd["b_classifier.v.1.weight"], sd["b_classifier.v.1.bias"] = _linear(g, dim_feat, dim_feat)
    w = (torch.rand(num_classes, num_classes, dim_feat, generator=g) * 2 - 1) / math.sqrt(dim_feat)
    sd["b_classifier.fcc.weight"] = w
    sd["b_classifier.fcc.bias"] = (torch.rand(num_classes, generator=g) * 2 - 1) / math.sqrt(dim_feat)
    return sd


def _gru(g, prefix, in_f, hid, sd):
    bound = 1.0 / math.sqrt(hid)
    sd[f"{prefix}.weight_ih_l0"] = (torch.rand(3 * hid, in_f, generator=g) * 2 - 1) * bound
    sd[f"{prefix}.weight_hh_l0"] = (torch.rand(3 * hid, hid, generator=g) * 2 - 1) * bound
    sd[f"{prefix}.bias_ih_l0"] = (torch.rand(3 * hid, generator=g) * 2 - 1) * bound
    sd[f"{prefix}.bias_hh_l0"] = (torch.rand(3 * hid, generator=g) * 2 - 1) * bound


def full_layer_state(feature_num: int, hidden: int = 1024, class_num: int = 128, seed: int = SEED):
    """State dict of Full_layer(fc_rnn=True) (models/rlmil.py:199-200)."""
    g = gen(seed)
    sd = {}
    _gru(g, "rnn", feature_num, hidden, sd)
    sd["fc.weight"], sd["fc.bias"] = _linear(g, class_num, hidden)
    return sd


def actor_state(state_dim: int, hidden: int = 512, action_size: int = 10, seed: int = SEED):
    """State dict of ActorCritic(policy_conv=False) (models/rlmil.py:39-53)."""
    g = gen(seed)
    sd = {}
    sd["state_encoder.0.weight"], sd["state_encoder.0.bias"] = _linear(g, 2048, state_dim)
    sd["state_encoder.2.weight"], sd["state_encoder.2.bias"] = _linear(g, hidden, 2048)
    _gru(g, "gru", hidden, hidden, sd)
    sd["actor.0.weight"], sd["actor.0.bias"] = _linear(g, action_size, hidden, 3.0)
    sd["critic.0.weight"], sd["critic.0.bias"] = _linear(g, 1, hidden)
    return sd


def camelyon_sizes(b: int, lo: int = 500, hi: int = 15500, seed: int = SEED) -> List[int]:
    """Ragged patch counts, uniform in [lo, hi] (mean ~8k, Camelyon16 at 20x)."""
    g = gen(seed + 17)
    return torch.randint(lo, hi + 1, (b,), generator=g).tolist()
