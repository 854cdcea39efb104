#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the REAL reference.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference ships no tests or known-answer vectors, so these files are
the pin for ``oracle/murcl_oracle.py``: each one stores the outputs (and gradients) of the
unmodified upstream modules on seeded synthetic inputs.  Inputs and weights are NOT stored when
they can be regenerated from ``murcl_b200.synth`` with the recorded seed; the fixtures stay small.

Hard-coded ``.cuda()`` calls in models/dsmil.py and models/rlmil.py are neutralised for the CPU
run by making ``Tensor.cuda`` / ``Module.cuda`` the identity before import.
"""
import os
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
REF = Path(os.environ.get("MURCL_REFERENCE", "/root/reference"))
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(REF))

torch.Tensor.cuda = lambda self, *a, **k: self          # noqa: E731
torch.nn.Module.cuda = lambda self, *a, **k: self       # noqa: E731

from models import abmil, clam, dsmil, rlmil            # noqa: E402  (reference)
from utils import datasets as ref_datasets              # noqa: E402  (reference)
from utils import losses as ref_losses                  # noqa: E402  (reference)

from murcl_b200 import synth                            # noqa: E402

torch.set_num_threads(1)


def npy(t):
    return t.detach().cpu().numpy()


def save(name, **arrays):
    np.savez_compressed(HERE / f"{name}.npz", **arrays)
    size = (HERE / f"{name}.npz").stat().st_size
    print(f"  {name}.npz  {size / 1024:.1f} KiB")


# ---------------------------------------------------------------------------------------------
def golden_selection():
    """Brute-force sweep of the selection arithmetic through the reference's own torch
    expressions (utils/datasets.py:285-291) and Python slicing (:294)."""
    rng = np.random.RandomState(985)
    cases = []
    for _ in range(6000):
        fs = int(rng.choice([16, 64, 512, 1024]))
        num_patch = int(rng.randint(1, 40000))
        n = int(rng.randint(0, min(num_patch, 6000) + 1))
        a = float(rng.choice([0.0, 1.0, rng.rand(), np.float32(rng.rand())]))
        cases.append((fs, num_patch, n, a))
    out = np.zeros((len(cases), 5), dtype=np.int64)
    acts = np.zeros(len(cases), dtype=np.float32)
    for i, (fs, num_patch, n, a) in enumerate(cases):
        ratio = fs / num_patch
        nt = torch.tensor([n])
        size = torch.round(nt * ratio).int()
        at = torch.tensor([a], dtype=torch.float32)
        left = torch.floor(at * (nt - size)).int()
        right = left + size
        picked = list(range(n))[left[0].item():right[0].item()]
        start = picked[0] if picked else 0
        out[i] = (fs, num_patch, n, start, len(picked))
        acts[i] = at[0].item()
    save("selection_sweep", cases=out, actions=acts)


def golden_get_feats():
    sizes = [300, 40, 64, 1000, 65, 7]
    k, d, fs = 6, 8, 64
    feats, clusters, _ = synth.make_bags(sizes, d, k, seed=11)
    clusters[2][3] = clusters[2][3] + clusters[2][4]      # make one cluster empty, keep ids unique
    clusters[2][4] = []
    clusters[2][3].sort()
    g = synth.gen(12)
    actions = torch.rand(len(sizes), k, generator=g)
    actions[0, 0], actions[0, 1], actions[1, 2], actions[3, 5] = 0.0, 1.0, 1.0, 0.0
    out = ref_datasets.get_feats([f.unsqueeze(0) for f in feats], clusters, actions, feat_size=fs)
    save("get_feats", sizes=np.asarray(sizes), k=k, d=d, fs=fs, seed_bags=11, actions=npy(actions), out=npy(out),
         merged_cluster=np.asarray([2, 3, 4]))


def golden_mixup():
    g = synth.gen(21)
    x = torch.randn(5, 12, 8, generator=g)
    torch.manual_seed(22)
    out, lam, perm = ref_datasets.mixup(x, 0.9)
    save("mixup", x=npy(x), seed=22, alpha=0.9, out=npy(out), lam=npy(lam), perm=npy(perm))


def sample(a, limit=4096):
    """Large gradients are stored as a deterministic strided sample (tests apply the same rule)."""
    a = np.asarray(a)
    if a.size <= limit:
        return a
    return a.reshape(-1)[:: -(-a.size // limit)]


def _grads(module, loss, extra=()):
    module.zero_grad()
    loss.backward()
    g = {f"grad.{n}": sample(npy(p.grad)) for n, p in module.named_parameters() if p.grad is not None}
    for n, t in extra:
        g[f"grad_input.{n}"] = npy(t.grad)
    return g


def golden_abmil():
    for tag, (dim_in, L, D, sizes) in {"small": (24, 32, 16, [50, 77, 1]), "full": (512, 512, 128, [300, 64])}.items():
        sd = synth.abmil_state(dim_in, L, D, 2, seed=31)
        m = abmil.ABMIL(dim_in, L=L, D=D, dim_out=2)
        m.load_state_dict(sd, strict=True)
        feats, _, _ = synth.make_bags(sizes, dim_in, 3, seed=32)
        bags = [f.clone().requires_grad_(tag == "small") for f in feats]
        out, det = m(bags)
        g = synth.gen(33)
        cot = torch.randn(out.shape, generator=g)
        arrays = dict(dims=np.asarray([dim_in, L, D]), sizes=np.asarray(sizes), out=npy(out), cot=npy(cot))
        if tag == "small":
            arrays.update(_grads(m, (out * cot).sum(), [(str(i), b) for i, b in enumerate(bags)]))
            # dense batch path [B,N,D] -> per-bag loop (abmil.py:57-60)
            xb = torch.stack([feats[0][:40], feats[1][:40]])
            arrays["out_dense"] = npy(m(xb)[0])
            arrays["out_single"] = npy(m(feats[0].unsqueeze(0))[0])
        save(f"abmil_{tag}", **arrays)


def golden_clam():
    in_dim, sizes = 24, [60, 33, 100]
    feats, _, _ = synth.make_bags(sizes, in_dim, 3, seed=42)
    for gate in (True, False):
        for dropout in (False, True):
            for subtyping in (False, True):
                tag = f"clam_g{int(gate)}_d{int(dropout)}_s{int(subtyping)}"
                n_classes = 3 if subtyping else 2
                sd = synth.clam_state(in_dim, "small", gate, dropout, n_classes, seed=41)
                m = clam.CLAM_SB(gate=gate, size_arg="small", dropout=dropout, k_sample=8, n_classes=n_classes,
                                 subtyping=subtyping, in_dim=in_dim)
                m.load_state_dict(sd, strict=True)
                m.eval()
                arrays = dict(sizes=np.asarray(sizes), in_dim=in_dim, n_classes=n_classes)
                labels = [1, 0, n_classes - 1]
                g = synth.gen(43)
                tot = 0.0
                for i, f in enumerate(feats):
                    x = f.clone().requires_grad_(True) if i == 0 else f
                    if i == 0:
                        x0 = x
                    out, det, res = m(x.unsqueeze(0), label=torch.tensor([labels[i]]), instance_eval=True)
                    cot = torch.randn(out.shape, generator=g)
                    arrays[f"out{i}"] = npy(out)
                    arrays[f"cot{i}"] = npy(cot)
                    arrays[f"inst_loss{i}"] = npy(res["instance_loss"])
                    arrays[f"inst_preds{i}"] = res["inst_preds"]
                    arrays[f"inst_labels{i}"] = res["inst_labels"]
                    arrays[f"raw_scores{i}"] = npy(m.bag_forward(f, attention_only=True))
                    tot = tot + (out * cot).sum() + 0.3 * res["instance_loss"]
                arrays["labels"] = np.asarray(labels)
                arrays.update(_grads(m, tot, [("0", x0)]))
                # batch (list) path without instance eval (clam.py:183-195)
                arrays["out_list"] = npy(m(list(feats))[0])
                save(tag, **arrays)
    # "big" size only changes D (clam.py:67)
    sd = synth.clam_state(in_dim, "big", True, False, 2, seed=44)
    m = clam.CLAM_SB(gate=True, size_arg="big", in_dim=in_dim)
    m.load_state_dict(sd, strict=True)
    m.eval()
    save("clam_big", out=npy(m(feats[0].unsqueeze(0))[0]), in_dim=in_dim, n=sizes[0])


def golden_dsmil():
    dim, c, sizes = 40, 2, [90, 17]
    sd = synth.dsmil_state(dim, c, seed=51)
    m = dsmil.build_dsmil(dim, c)
    m.load_state_dict(sd, strict=True)
    feats, _, _ = synth.make_bags(sizes, dim, 3, seed=52)
    g = synth.gen(53)
    arrays = dict(dim=dim, c=c, sizes=np.asarray(sizes))
    tot = 0.0
    xs = []
    for i, f in enumerate(feats):
        x = f.clone().requires_grad_(True)
        xs.append((str(i), x))
        classes, bag, bag_det = m(x.unsqueeze(0))
        cot_b = torch.randn(bag.shape, generator=g)
        cot_c = torch.randn(classes.shape, generator=g)
        arrays[f"classes{i}"], arrays[f"bag{i}"] = npy(classes), npy(bag)
        arrays[f"cot_b{i}"], arrays[f"cot_c{i}"] = npy(cot_b), npy(cot_c)
        tot = tot + (bag * cot_b).sum() + (classes * cot_c).sum()
    arrays.update(_grads(m, tot, xs))
    save("dsmil", **arrays)


def golden_ntxent():
    arrays = {}
    for i, (b, d, tau) in enumerate([(8, 16, 0.5), (5, 32, 1.0), (16, 128, 0.1)]):
        g = synth.gen(60 + i)
        zi = torch.randn(b, d, generator=g).requires_grad_(True)
        zj = (0.5 * zi.detach() + torch.randn(b, d, generator=g)).requires_grad_(True)
        crit = ref_losses.NT_Xent(b, tau)
        loss = crit(zi, zj)
        loss.backward()
        arrays.update({f"cfg{i}": np.asarray([b, d, tau]), f"zi{i}": npy(zi), f"zj{i}": npy(zj),
                       f"loss{i}": npy(loss), f"gzi{i}": npy(zi.grad), f"gzj{i}": npy(zj.grad),
                       f"cos{i}": npy(torch.cosine_similarity(zi, zj))})
    zi = torch.randn(6, 8, generator=synth.gen(69))
    arrays["loss_identical"] = npy(ref_losses.NT_Xent(6, 1.0)(zi, zi.clone()))
    save("ntxent", **arrays)


def golden_full_layer():
    fnum, hid, cls, b = 24, 40, 12, 5
    sd = synth.full_layer_state(fnum, hid, cls, seed=71)
    m = rlmil.Full_layer(fnum, hid, True, cls)
    m.load_state_dict(sd, strict=True)
    g = synth.gen(72)
    xs = [torch.randn(b, fnum, generator=g).requires_grad_(True) for _ in range(3)]
    cots = [torch.randn(b, cls, generator=g) for _ in range(3)]
    outs = [m(x, restart=(t == 0)) for t, x in enumerate(xs)]
    tot = sum((o * c).sum() for o, c in zip(outs, cots))
    arrays = dict(dims=np.asarray([fnum, hid, cls, b]))
    for t in range(3):
        arrays[f"x{t}"], arrays[f"cot{t}"], arrays[f"out{t}"] = npy(xs[t]), npy(cots[t]), npy(outs[t])
    arrays.update(_grads(m, tot, [(str(t), x) for t, x in enumerate(xs)]))
    save("full_layer", **arrays)


def golden_actor():
    import torch.distributions.multivariate_normal as mvn
    sdim, hid, k, b, std = 32, 24, 6, 7, 0.5
    sd = synth.actor_state(sdim, hid, k, seed=81)
    ppo = rlmil.PPO(sdim, sdim, hid, False, action_std=std, action_size=k)
    ppo.policy_old.load_state_dict(sd, strict=True)
    eps_log = []
    orig = mvn._standard_normal

    def recording(shape, dtype, device):
        e = orig(shape, dtype, device)
        eps_log.append(e.clone())
        return e

    mvn._standard_normal = recording
    mem = rlmil.Memory()
    g = synth.gen(82)
    arrays = dict(dims=np.asarray([sdim, hid, k, b]), std=std)
    torch.manual_seed(83)
    for t in range(3):
        state = torch.randn(b, sdim, generator=g)
        action = ppo.select_action(state, mem, restart_batch=(t == 0))
        arrays[f"state{t}"], arrays[f"action{t}"] = npy(state), npy(action)
        arrays[f"eps{t}"], arrays[f"logprob{t}"] = npy(eps_log[-1]), npy(mem.logprobs[-1])
        arrays[f"hidden{t}"] = npy(mem.hidden[-1][0])
    mvn._standard_normal = orig
    save("actor", **arrays)


def golden_actor_conv():
    """ActorCritic(policy_conv=True) (models/rlmil.py:30-37,71-74,104-107): act over two steps and evaluate with its
    gradients, states are [B, feature_dim, r, r] feature maps."""
    import torch.distributions.multivariate_normal as mvn
    fdim, r, hid, k, b, std = 8, 2, 24, 5, 6, 0.5
    sdim = fdim * r * r
    ppo = rlmil.PPO(fdim, sdim, hid, True, action_std=std, action_size=k)
    g = synth.gen(121)
    sd = {n: 0.3 * torch.randn(p.shape, generator=g) for n, p in ppo.policy.state_dict().items()}
    ppo.policy.load_state_dict(sd)
    ppo.policy_old.load_state_dict(sd)
    eps_log = []
    orig = mvn._standard_normal

    def recording(shape, dtype, device):
        e = orig(shape, dtype, device)
        eps_log.append(e.clone())
        return e

    mvn._standard_normal = recording
    mem = rlmil.Memory()
    arrays = dict(dims=np.asarray([fdim, r, hid, k, b]), std=std)
    arrays.update({f"sd.{n}": npy(v) for n, v in sd.items()})
    torch.manual_seed(122)
    for t in range(2):
        state = torch.randn(b, fdim, r, r, generator=g)
        action = ppo.select_action(state, mem, restart_batch=(t == 0))
        arrays[f"state{t}"], arrays[f"action{t}"] = npy(state), npy(action)
        arrays[f"eps{t}"], arrays[f"logprob{t}"] = npy(eps_log[-1]), npy(mem.logprobs[-1])
    mvn._standard_normal = orig
    lp, val, ent = ppo.policy.evaluate(torch.stack(mem.states, 0), torch.stack(mem.actions, 0))
    arrays["eval_logprob"], arrays["eval_value"] = npy(lp), npy(val)
    cot_l, cot_v = torch.randn(lp.shape, generator=g), torch.randn(val.shape, generator=g)
    arrays["cot_l"], arrays["cot_v"] = npy(cot_l), npy(cot_v)
    ppo.policy.zero_grad()
    ((lp * cot_l).sum() + (val * cot_v).sum()).backward()
    for n, p in ppo.policy.named_parameters():
        arrays[f"grad.{n}"] = npy(p.grad)
    save("actor_conv", **arrays)


def golden_ppo_update():
    """PPO.evaluate / PPO.update (models/rlmil.py:99-127,152-184) on a 4-step rollout produced by the reference's own
    ``select_action`` with recorded Gaussian draws and fixed rewards: the evaluate outputs, the first epoch's
    gradients and the weights after K_epochs=3 Adam steps."""
    import torch.distributions.multivariate_normal as mvn
    sdim, hid, k, b, std, T = 32, 24, 6, 7, 0.5, 4
    sd = synth.actor_state(sdim, hid, k, seed=101)
    ppo = rlmil.PPO(sdim, sdim, hid, False, action_std=std, lr=3e-4, gamma=0.1, K_epochs=3, action_size=k)
    ppo.policy.load_state_dict(sd, strict=True)
    ppo.policy_old.load_state_dict(sd, strict=True)
    eps_log = []
    orig = mvn._standard_normal

    def recording(shape, dtype, device):
        e = orig(shape, dtype, device)
        eps_log.append(e.clone())
        return e

    mvn._standard_normal = recording
    mem = rlmil.Memory()
    g = synth.gen(102)
    arrays = dict(dims=np.asarray([sdim, hid, k, b, T]), std=std, lr=3e-4, gamma=0.1, K_epochs=3, eps_clip=0.2, seed_actor=101)
    torch.manual_seed(103)
    for t in range(T):
        state = torch.randn(b, sdim, generator=g)
        ppo.select_action(state, mem, restart_batch=(t == 0))
        arrays[f"state{t}"], arrays[f"eps{t}"] = npy(state), npy(eps_log[-1])
        arrays[f"action{t}"], arrays[f"logprob{t}"] = npy(mem.actions[-1]), npy(mem.logprobs[-1])
        reward = 0.3 * torch.randn(1, b, generator=g)           # shape of train_MuRCL.py:283 (sim_last - sim, [1, B])
        mem.rewards.append(reward)
        arrays[f"reward{t}"] = npy(reward)
    mvn._standard_normal = orig
    states, actions = torch.stack(mem.states, 0), torch.stack(mem.actions, 0)
    lp, val, ent = ppo.policy.evaluate(states, actions)
    arrays["eval_logprob"], arrays["eval_value"], arrays["eval_entropy"] = npy(lp), npy(val), npy(ent)
    # first epoch's gradient = what update() computes before its first Adam step (same expressions, rlmil.py:153-181)
    disc, run = [], 0
    for r in reversed(mem.rewards):
        run = r + ppo.gamma * run
        disc.insert(0, run)
    ret = torch.cat(disc, 0)
    ret = (ret - ret.mean()) / (ret.std() + 1e-5)
    arrays["returns"] = npy(ret)
    ratios = torch.exp(lp - torch.stack(mem.logprobs, 0).detach())
    adv = ret - val.detach()
    loss = -torch.min(ratios * adv, torch.clamp(ratios, 0.8, 1.2) * adv) + 0.5 * ppo.MseLoss(val, ret) - 0.01 * ent
    ppo.policy.zero_grad()
    loss.mean().backward()
    arrays["loss0"] = npy(loss.mean())
    for n, p in ppo.policy.named_parameters():
        arrays[f"grad.{n}"] = sample(npy(p.grad))
    ppo.policy.zero_grad()
    ppo.update(mem)
    for n, p in ppo.policy.named_parameters():
        arrays[f"new.{n}"] = sample(npy(p))
        arrays[f"delta.{n}"] = sample(npy(p.detach() - sd[n]))
    assert all(torch.equal(a, b_) for a, b_ in zip(ppo.policy.state_dict().values(), ppo.policy_old.state_dict().values()))
    save("ppo_update", **arrays)


def golden_pretrain_step():
    """A miniature stage-1 optimiser step assembled exactly as train_MuRCL.py:235-294 does."""
    from models import cl
    b, k, d, fs, T, alpha, tau = 4, 3, 16, 32, 2, 0.9, 1.0
    L, D, hid, proj = 32, 16, 24, 8
    sizes = [90, 20, 45, 33]
    feats, clusters, _ = synth.make_bags(sizes, d, k, seed=91)
    sd_m = synth.abmil_state(d, L, D, proj, seed=92)
    sd_f = synth.full_layer_state(L, hid, proj, seed=93)
    enc = abmil.ABMIL(d, L=L, D=D, dim_out=proj)
    enc.load_state_dict(sd_m)
    model = cl.CL(enc, projection_dim=proj, n_features=L)
    fc = rlmil.Full_layer(L, hid, True, proj)
    fc.load_state_dict(sd_f)
    crit = ref_losses.NT_Xent(b, tau)
    feat_list = [f.unsqueeze(0) for f in feats]
    torch.manual_seed(94)
    losses = []
    for t in range(T):
        acts = [torch.rand((b, k)) for _ in range(2)]
        xv = [ref_datasets.get_feats(feat_list, clusters, action_sequence=a, feat_size=fs) for a in acts]
        xv = [ref_datasets.mixup(x, alpha)[0] for x in xv]
        outs, states = model(xv)
        outs = [fc(o, restart=(t == 0)) for o in outs]
        losses.append(crit(outs[0], outs[1]))
    loss = sum(losses) / T
    loss.backward()
    arrays = dict(cfg=np.asarray([b, k, d, fs, T, L, D, hid, proj]), sizes=np.asarray(sizes), alpha=alpha, tau=tau,
                  loss=npy(loss))
    arrays.update({f"grad.m.{n}": sample(npy(p.grad)) for n, p in enc.named_parameters() if p.grad is not None})
    arrays.update({f"grad.f.{n}": sample(npy(p.grad)) for n, p in fc.named_parameters() if p.grad is not None})
    save("pretrain_step", **arrays)


def golden_stage3_step():
    """A miniature stage-3 optimiser step (train_MuRCL.py:235-298 with ``train_stage == 3``: the PPO actor chooses the
    windows of patch-steps >= 1 from the detached bag embeddings; it is not updated), followed on the SAME rollout by
    what stage 2 does instead of the optimiser step (``ppo.update(m)`` for both memories, :296-298).  All random draws
    (first actions, mixup lambda / permutation, the actor's Gaussian noise) are recorded."""
    import torch.distributions.multivariate_normal as mvn
    from models import cl
    b, k, d, fs, T, alpha, tau, std = 6, 4, 16, 32, 3, 0.9, 1.0, 0.5
    L, D, hid, proj, phid = 32, 16, 24, 8, 20
    sizes = [90, 20, 45, 33, 150, 64]
    feats, clusters, _ = synth.make_bags(sizes, d, k, seed=111)
    sd_m = synth.abmil_state(d, L, D, proj, seed=112)
    sd_f = synth.full_layer_state(L, hid, proj, seed=113)
    sd_a = synth.actor_state(L, phid, k, seed=114)
    enc = abmil.ABMIL(d, L=L, D=D, dim_out=proj)
    enc.load_state_dict(sd_m)
    model = cl.CL(enc, projection_dim=proj, n_features=L)
    fc = rlmil.Full_layer(L, hid, True, proj)
    fc.load_state_dict(sd_f)
    ppo = rlmil.PPO(d, L, phid, False, action_std=std, lr=1e-3, gamma=0.1, K_epochs=2, action_size=k)
    ppo.policy.load_state_dict(sd_a)
    ppo.policy_old.load_state_dict(sd_a)
    crit = ref_losses.NT_Xent(b, tau)
    feat_list = [f.unsqueeze(0) for f in feats]
    memory_list = [rlmil.Memory(), rlmil.Memory()]
    eps_log = []
    orig = mvn._standard_normal

    def recording(shape, dtype, device):
        e = orig(shape, dtype, device)
        eps_log.append(e.clone())
        return e

    mvn._standard_normal = recording
    arrays = dict(cfg=np.asarray([b, k, d, fs, T, L, D, hid, proj, phid]), sizes=np.asarray(sizes), alpha=alpha, tau=tau,
                  std=std, ppo_lr=1e-3, ppo_gamma=0.1, ppo_K_epochs=2)
    torch.manual_seed(115)
    loss_list, states = [], None
    similarity_last = None
    for t in range(T):
        if t == 0:
            acts = [torch.rand((b, k)) for _ in range(2)]
        else:
            acts = [ppo.select_action(s, m, restart_batch=(t == 1)) for s, m in zip(states, memory_list)]
            arrays[f"eps{t}_0"], arrays[f"eps{t}_1"] = npy(eps_log[-2]), npy(eps_log[-1])
        xv = [ref_datasets.get_feats(feat_list, clusters, action_sequence=a, feat_size=fs) for a in acts]
        mixed = [ref_datasets.mixup(x, alpha) for x in xv]
        for v in range(2):
            arrays[f"act{t}_{v}"], arrays[f"lam{t}_{v}"], arrays[f"perm{t}_{v}"] = npy(acts[v]), npy(mixed[v][1]), npy(mixed[v][2])
        outs, states = model([m_[0] for m_ in mixed])
        outs = [fc(o, restart=(t == 0)) for o in outs]
        loss = crit(outs[0], outs[1])
        loss_list.append(loss)
        arrays[f"loss{t}"] = npy(loss)
        sim = torch.cosine_similarity(outs[0], outs[1]).view(1, -1)
        if t >= 1:
            reward = similarity_last - sim
            arrays[f"reward{t}"] = npy(reward)
            for m_ in memory_list:
                m_.rewards.append(reward)
        similarity_last = sim
    mvn._standard_normal = orig
    for v, m_ in enumerate(memory_list):
        arrays[f"logprobs_{v}"] = npy(torch.stack(m_.logprobs, 0))
    loss = sum(loss_list) / T
    loss.backward()
    arrays["loss"] = npy(loss)
    arrays.update({f"grad.m.{n}": sample(npy(p.grad)) for n, p in enc.named_parameters() if p.grad is not None})
    arrays.update({f"grad.f.{n}": sample(npy(p.grad)) for n, p in fc.named_parameters() if p.grad is not None})
    # stage 2 on the same rollout: PPO update from both memories in turn (train_MuRCL.py:296-298)
    for m_ in memory_list:
        m_.rewards = [r.detach() for r in m_.rewards]
        ppo.update(m_)
    for n, p in ppo.policy.named_parameters():
        arrays[f"ppo_delta.{n}"] = sample(npy(p.detach() - sd_a[n]))
    save("stage3_step", **arrays)


if __name__ == "__main__":
    assert REF.exists(), f"reference not found at {REF}"
    print("writing fixtures to", HERE)
    makers = [golden_selection, golden_get_feats, golden_mixup, golden_abmil, golden_clam, golden_dsmil, golden_ntxent,
              golden_full_layer, golden_actor, golden_actor_conv, golden_ppo_update, golden_pretrain_step, golden_stage3_step]
    only = set(sys.argv[1:])            # e.g. `make_golden.py golden_ppo_update` regenerates one fixture
    for fn in makers:
        if not only or fn.__name__ in only:
            fn()
