"""Supervised RLMIL steps (train_RLMIL.py train_CLAM / train_DSMIL call sites) against the oracle
composition on the same injected actions (fp32 mode, batch size 1 as in runs/scratch.sh)."""
import pytest
import torch
import torch.nn.functional as F

from murcl_b200 import synth
from oracle import murcl_oracle as O
from tests.helpers import assert_close, leaf_state

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _grads(module):
    return {n: p.grad.detach().cpu() for n, p in module.named_parameters() if p.grad is not None}


def test_clam_supervised_step():
    from murcl_b200 import supervised
    from murcl_b200.csr import BagStore
    from murcl_b200.dropin import clam, rlmil
    d, k, fs, T, hid = 48, 4, 64, 3, 40
    feats, clusters, _ = synth.make_bags([400], d, k, seed=101)
    sd_m = synth.clam_state(d, "small", True, False, 2, seed=102)
    sd_f = synth.full_layer_state(512, hid, 2, seed=103)
    label = torch.tensor([1])
    draws = [torch.rand(1, k, generator=synth.gen(104 + t)) for t in range(T)]
    # oracle
    pm, pf = leaf_state(sd_m), leaf_state(sd_f)
    h, losses = None, []
    for t in range(T):
        x, _ = O.get_feats(feats, clusters, draws[t], fs)
        M, res = O.clam_sb_bag(x[0], pm, gate=True, label=1, instance_eval=True, n_classes=2, subtyping=True)
        logits, h = O.full_layer_step(M, h, pf)
        losses.append(0.7 * F.cross_entropy(logits, label) + 0.3 * res["instance_loss"])
    want = sum(losses) / T
    want.backward()
    # device
    m = clam.CLAM_SB(gate=True, size_arg="small", n_classes=2, subtyping=True, in_dim=d, precision="fp32")
    m.load_state_dict(sd_m); m = m.to(DEV)
    fc = rlmil.Full_layer(512, hid, True, 2); fc.load_state_dict(sd_f); fc = fc.to(DEV)
    store = BagStore.from_cluster_lists(feats, clusters, DEV)
    got, logits = supervised.clam_step(store, m, fc, label, T=T, feat_size=fs, bag_weight=0.7,
                                       draws=[a.to(DEV) for a in draws], precision="fp32")
    assert_close(got, want, 1e-5, "loss")
    assert len(logits) == T and logits[0].shape == (1, 2)
    gm, gf = _grads(m), _grads(fc)
    for name, p in pm.items():
        if p.grad is not None and name in gm:
            assert_close(gm[name], p.grad, 2e-4, name, floor=1e-1 if name.endswith("attention_c.bias") else 1e-6)
    for name, p in pf.items():
        assert_close(gf[name], p.grad, 2e-4, name, floor=1e-6)


def test_dsmil_supervised_step_and_rewards():
    from murcl_b200 import supervised
    from murcl_b200.csr import BagStore
    from murcl_b200.dropin import dsmil, rlmil
    d, k, fs, T, hid = 64, 4, 96, 3, 32
    feats, clusters, _ = synth.make_bags([500], d, k, seed=111)
    sd_m = synth.dsmil_state(d, 2, seed=112)
    sd_f = synth.full_layer_state(d, hid, 2, seed=113)
    label = torch.tensor([0])
    draws = [torch.rand(1, k, generator=synth.gen(114 + t)) for t in range(T)]
    pm = leaf_state({a: v for a, v in sd_m.items() if "fcc" not in a})
    pf = leaf_state(sd_f)
    h, losses, confs = None, [], []
    for t in range(T):
        x, _ = O.get_feats(feats, clusters, draws[t], fs)
        classes, bag = O.dsmil_bag(x[0], pm)
        logits, h = O.full_layer_step(bag.mean(1), h, pf)
        inst_max = classes.max(0, keepdim=True).values
        losses.append(0.5 * F.cross_entropy(logits, label) + 0.5 * F.cross_entropy(inst_max, label))
        confs.append(float(torch.softmax(logits.detach(), 1)[0, 0]))
    want = sum(losses) / T
    want.backward()
    m = dsmil.build_dsmil(d, 2, precision="fp32"); m.load_state_dict(sd_m)
    fc = rlmil.Full_layer(d, hid, True, 2); fc.load_state_dict(sd_f); fc = fc.to(DEV)
    store = BagStore.from_cluster_lists(feats, clusters, DEV)
    mem = rlmil.Memory()
    got, _ = supervised.dsmil_step(store, m, fc, label, T=T, feat_size=fs, memory=mem,
                                   draws=[a.to(DEV) for a in draws], precision="fp32")
    assert_close(got, want, 1e-5, "loss")
    # rewards = change of the true-class confidence between consecutive patch-steps (train_RLMIL.py:370-371)
    assert len(mem.rewards) == T - 1
    for t in range(1, T):
        assert abs(float(mem.rewards[t - 1]) - (confs[t] - confs[t - 1])) < 1e-5
    gm, gf = _grads(m), _grads(fc)
    for name, p in pm.items():
        assert_close(gm[name], p.grad, 2e-4, name, floor=1e-6)
    for name, p in pf.items():
        assert_close(gf[name], p.grad, 2e-4, name, floor=1e-6)
