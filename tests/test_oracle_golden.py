"""Pins oracle/murcl_oracle.py to the fixtures generated from the real reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import torch

from murcl_b200 import synth
from oracle import murcl_oracle as O
from tests.helpers import assert_close, leaf_state, sample

TOL = 2e-5  # fp32 restatement vs fp32 reference: summation order only


def test_selection_sweep(golden):
    g = golden("selection_sweep")
    cases, acts = g["cases"], g["actions"]
    for (fs, num_patch, n, start_ref, count_ref), a in zip(cases, acts):
        s, e = O.select_windows([n], int(num_patch), np.asarray([a], dtype=np.float32), int(fs))
        assert e[0] - s[0] == count_ref, (fs, num_patch, n, a)
        if count_ref:
            assert s[0] == start_ref, (fs, num_patch, n, a)


def _get_feats_inputs(g):
    feats, clusters, _ = synth.make_bags(g["sizes"].tolist(), int(g["d"]), int(g["k"]), seed=int(g["seed_bags"]))
    b, src, dst = g["merged_cluster"].tolist()
    clusters[b][src] = sorted(clusters[b][src] + clusters[b][dst])
    clusters[b][dst] = []
    return feats, clusters


def test_get_feats(golden):
    g = golden("get_feats")
    feats, clusters = _get_feats_inputs(g)
    out, kept = O.get_feats(feats, clusters, torch.from_numpy(g["actions"]), int(g["fs"]))
    assert np.array_equal(out.numpy(), g["out"])          # bit-exact
    assert any(len(k) < int(g["fs"]) for k in kept) and any(len(k) == int(g["fs"]) for k in kept)


def test_mixup(golden):
    g = golden("mixup")
    out, lam, perm = O.mixup(torch.from_numpy(g["x"]), float(g["alpha"]), synth.gen(int(g["seed"])))
    assert np.array_equal(perm.numpy(), g["perm"])
    assert np.array_equal(lam.numpy(), g["lam"])
    assert np.array_equal(out.numpy(), g["out"])          # bit-exact: mul, mul, add


def _check_grads(g, params, prefix="grad."):
    n = 0
    for k, v in g.items():
        if k.startswith(prefix):
            p = params[k[len(prefix):]]
            got = p.grad if p.grad is not None else torch.zeros_like(p)
            assert_close(sample(got.numpy()), v, 5e-5, k, floor=1e-4)
            n += 1
    assert n > 0


def test_abmil_small(golden):
    g = golden("abmil_small")
    dim_in, L, D = g["dims"].tolist()
    sd = leaf_state(synth.abmil_state(dim_in, L, D, 2, seed=31))
    feats, _, _ = synth.make_bags(g["sizes"].tolist(), dim_in, 3, seed=32)
    bags = [f.clone().requires_grad_(True) for f in feats]
    out = O.abmil_forward(bags, sd)
    assert_close(out, g["out"], TOL, "out")
    (out * torch.from_numpy(g["cot"])).sum().backward()
    _check_grads(g, sd)
    for i, b in enumerate(bags):
        assert_close(b.grad, g[f"grad_input.{i}"], 5e-5, f"dx{i}")
    # attention.2.bias gradient is analytically zero (softmax shift invariance)
    assert float(sd["attention.2.bias"].grad.abs().max()) < 1e-6
    with torch.no_grad():
        dense = O.abmil_forward([feats[0][:40], feats[1][:40]], sd)
        assert_close(dense, g["out_dense"], TOL, "dense")
        assert_close(O.abmil_forward([feats[0]], sd), g["out_single"], TOL, "single")


def test_abmil_full(golden):
    g = golden("abmil_full")
    dim_in, L, D = g["dims"].tolist()
    sd = synth.abmil_state(dim_in, L, D, 2, seed=31)
    feats, _, _ = synth.make_bags(g["sizes"].tolist(), dim_in, 3, seed=32)
    assert_close(O.abmil_forward(feats, sd), g["out"], TOL, "out")


def test_clam(golden):
    for gate in (True, False):
        for dropout in (False, True):
            for subtyping in (False, True):
                g = golden(f"clam_g{int(gate)}_d{int(dropout)}_s{int(subtyping)}")
                in_dim, n_classes = int(g["in_dim"]), int(g["n_classes"])
                sd = leaf_state(synth.clam_state(in_dim, "small", gate, dropout, n_classes, seed=41))
                feats, _, _ = synth.make_bags(g["sizes"].tolist(), in_dim, 3, seed=42)
                labels = g["labels"].tolist()
                tot = 0.0
                x0 = feats[0].clone().requires_grad_(True)
                kw = dict(gate=gate, dropout_layers=dropout, n_classes=n_classes, subtyping=subtyping, k_sample=8)
                for i, f in enumerate(feats):
                    x = x0 if i == 0 else f
                    m, res = O.clam_sb_bag(x, sd, label=labels[i], instance_eval=True, **kw)
                    assert_close(m, g[f"out{i}"], TOL, f"M{i}")
                    assert_close(res["instance_loss"], g[f"inst_loss{i}"], TOL, f"inst{i}")
                    assert np.array_equal(res["inst_preds"], g[f"inst_preds{i}"])
                    assert np.array_equal(res["inst_labels"], g[f"inst_labels{i}"])
                    with torch.no_grad():
                        raw = O.clam_sb_bag(f, sd, attention_only=True, **kw)
                    assert_close(raw, g[f"raw_scores{i}"], TOL, f"raw{i}")
                    tot = tot + (m * torch.from_numpy(g[f"cot{i}"])).sum() + 0.3 * res["instance_loss"]
                tot.backward()
                _check_grads(g, sd)
                assert_close(x0.grad, g["grad_input.0"], 5e-5, "dx0")
                with torch.no_grad():
                    lst = torch.cat([O.clam_sb_bag(f, sd, **kw)[0] for f in feats], 0)
                assert_close(lst, g["out_list"], TOL, "list")
    g = golden("clam_big")
    sd = synth.clam_state(int(g["in_dim"]), "big", True, False, 2, seed=44)
    feats, _, _ = synth.make_bags([60, 33, 100], int(g["in_dim"]), 3, seed=42)
    assert_close(O.clam_sb_bag(feats[0], sd)[0], g["out"], TOL, "big")


def test_clam_instance_labels_known_answer():
    """n_classes=2, subtyping, label=1 -> targets [0]*8 + [1]*8 + [0]*8."""
    sd = synth.clam_state(24, "small", True, False, 2, seed=41)
    feats, _, _ = synth.make_bags([60], 24, 3, seed=42)
    _, res = O.clam_sb_bag(feats[0], sd, label=1, instance_eval=True, n_classes=2, subtyping=True)
    assert res["inst_labels"].tolist() == [0] * 8 + [1] * 8 + [0] * 8


def test_dsmil(golden):
    g = golden("dsmil")
    dim, c = int(g["dim"]), int(g["c"])
    sd = leaf_state(synth.dsmil_state(dim, c, seed=51))
    feats, _, _ = synth.make_bags(g["sizes"].tolist(), dim, 3, seed=52)
    tot = 0.0
    xs = []
    for i, f in enumerate(feats):
        x = f.clone().requires_grad_(True)
        xs.append(x)
        classes, bag = O.dsmil_bag(x, sd)
        assert_close(classes, g[f"classes{i}"], TOL, "classes")
        assert_close(bag, g[f"bag{i}"], TOL, "bag")
        tot = tot + (bag * torch.from_numpy(g[f"cot_b{i}"])).sum() + (classes * torch.from_numpy(g[f"cot_c{i}"])).sum()
    tot.backward()
    _check_grads(g, {k: v for k, v in sd.items() if "fcc" not in k})
    for i, x in enumerate(xs):
        assert_close(x.grad, g[f"grad_input.{i}"], 5e-5, f"dx{i}")


def test_ntxent(golden):
    g = golden("ntxent")
    for i in range(3):
        b, d, tau = g[f"cfg{i}"].tolist()
        zi = torch.from_numpy(g[f"zi{i}"]).requires_grad_(True)
        zj = torch.from_numpy(g[f"zj{i}"]).requires_grad_(True)
        loss = O.nt_xent(zi, zj, tau)
        assert_close(loss, g[f"loss{i}"], TOL, "loss")
        loss.backward()
        assert_close(zi.grad, g[f"gzi{i}"], 5e-5, "gzi")
        assert_close(zj.grad, g[f"gzj{i}"], 5e-5, "gzj")
        assert_close(O.pair_cosine(zi, zj), g[f"cos{i}"], TOL, "cos")
    z = torch.randn(6, 8, generator=synth.gen(69))
    assert_close(O.nt_xent(z, z.clone(), 1.0), g["loss_identical"], TOL, "identical views")
    # known answer: all 2B embeddings equal -> every logit equal -> ln(2B - 1)  
    same = z[:1].repeat(6, 1)
    assert abs(float(O.nt_xent(same, same.clone(), 1.0)) - np.log(11.0)) < 1e-5


def test_full_layer(golden):
    g = golden("full_layer")
    fnum, hid, cls, b = g["dims"].tolist()
    sd = leaf_state(synth.full_layer_state(fnum, hid, cls, seed=71))
    h, tot, xs = None, 0.0, []
    for t in range(3):
        x = torch.from_numpy(g[f"x{t}"]).requires_grad_(True)
        xs.append(x)
        out, h = O.full_layer_step(x, h, sd)
        assert_close(out, g[f"out{t}"], TOL, f"out{t}")
        tot = tot + (out * torch.from_numpy(g[f"cot{t}"])).sum()
    tot.backward()
    _check_grads(g, sd)
    for t, x in enumerate(xs):
        assert_close(x.grad, g[f"grad_input.{t}"], 5e-5, f"dx{t}")


def test_actor(golden):
    g = golden("actor")
    sdim, hid, k, b = g["dims"].tolist()
    sd = synth.actor_state(sdim, hid, k, seed=81)
    h = None
    for t in range(3):
        action, logprob, h, _ = O.actor_act(torch.from_numpy(g[f"state{t}"]), h, sd, float(g["std"]),
                                            torch.from_numpy(g[f"eps{t}"]))
        assert_close(action, g[f"action{t}"], TOL, "action")
        assert_close(logprob, g[f"logprob{t}"], TOL, "logprob")
        assert_close(h, g[f"hidden{t}"], TOL, "hidden")
        assert float(action.min()) >= 0.0 and float(action.max()) <= 1.0


def test_pretrain_step(golden):
    g = golden("pretrain_step")
    b, k, d, fs, T, L, D, hid, proj = g["cfg"].tolist()
    feats, clusters, _ = synth.make_bags(g["sizes"].tolist(), d, k, seed=91)
    sd_m = synth.abmil_state(d, L, D, proj, seed=92)
    sd_f = synth.full_layer_state(L, hid, proj, seed=93)
    loss, grads = O.pretrain_step(feats, clusters, sd_m, sd_f, arch="ABMIL", T=T, feat_size=fs,
                                  alpha=float(g["alpha"]), temperature=float(g["tau"]), generator=synth.gen(94))
    assert_close(loss, g["loss"], TOL, "loss")
    n = 0
    for key, v in g.items():
        if key.startswith("grad."):
            assert_close(sample(grads[key[5:]].numpy()), v, 2e-4, key, floor=1e-5)
            n += 1
    assert n >= 10


def _ppo_rollout(g):
    sdim, hid, k, b, T = g["dims"].tolist()
    states = torch.stack([torch.from_numpy(g[f"state{t}"]) for t in range(T)], 0)
    actions = torch.stack([torch.from_numpy(g[f"action{t}"]) for t in range(T)], 0)
    logprobs = torch.stack([torch.from_numpy(g[f"logprob{t}"]) for t in range(T)], 0)
    rewards = [torch.from_numpy(g[f"reward{t}"]) for t in range(T)]
    return (sdim, hid, k, b, T), states, actions, logprobs, rewards


def test_ppo_evaluate_and_update(golden):
    """PPO.evaluate / PPO.update (models/rlmil.py:99-127,152-184) on the reference's own rollout."""
    g = golden("ppo_update")
    (sdim, hid, k, b, T), states, actions, logprobs, rewards = _ppo_rollout(g)
    std = float(g["std"])
    sd = synth.actor_state(sdim, hid, k, seed=int(g["seed_actor"]))
    # the rollout itself: act() restated, chained over the T steps
    h = None
    for t in range(T):
        a, lp, h, _ = O.actor_act(states[t], h, sd, std, torch.from_numpy(g[f"eps{t}"]))
        assert_close(a, g[f"action{t}"], TOL, f"action{t}")
        assert_close(lp, g[f"logprob{t}"], TOL, f"logprob{t}")
    lp, val, ent = O.actor_evaluate(states, actions, sd, std)
    assert_close(lp, g["eval_logprob"], TOL, "evaluate.logprob")
    assert_close(val, g["eval_value"], TOL, "evaluate.value")
    assert_close(ent, g["eval_entropy"], TOL, "evaluate.entropy")
    assert_close(O.ppo_returns(rewards, float(g["gamma"])), g["returns"], TOL, "returns")
    new, first, losses = O.ppo_update(sd, states, actions, logprobs, rewards, action_std=std, lr=float(g["lr"]),
                                      gamma=float(g["gamma"]), K_epochs=int(g["K_epochs"]), eps_clip=float(g["eps_clip"]))
    assert_close(losses[0], g["loss0"], TOL, "loss0")
    n = 0
    for key, v in g.items():
        if key.startswith("grad."):
            assert_close(sample(first[key[5:]].numpy()), v, 5e-5, key, floor=1e-6)
        elif key.startswith("delta."):
            # Adam's first steps move every weight by ~lr: compare the DELTAS (1e-3 of the weights) to 1 %
            assert_close(sample((new[key[6:]] - sd[key[6:]]).numpy()), v, 1e-2, key, floor=1e-6)
            n += 1
    assert n == len(sd)


def _stage3_inputs(g):
    b, k, d, fs, T, L, D, hid, proj, phid = g["cfg"].tolist()
    feats, clusters, _ = synth.make_bags(g["sizes"].tolist(), d, k, seed=111)
    sd_m = synth.abmil_state(d, L, D, proj, seed=112)
    sd_f = synth.full_layer_state(L, hid, proj, seed=113)
    sd_a = synth.actor_state(L, phid, k, seed=114)
    first = [torch.from_numpy(g[f"act0_{v}"]) for v in range(2)]
    lams = [[torch.from_numpy(g[f"lam{t}_{v}"]) for v in range(2)] for t in range(T)]
    perms = [[torch.from_numpy(g[f"perm{t}_{v}"]) for v in range(2)] for t in range(T)]
    eps = [None] + [[torch.from_numpy(g[f"eps{t}_{v}"]) for v in range(2)] for t in range(1, T)]
    return (b, k, d, fs, T, L, D, hid, proj, phid), feats, clusters, sd_m, sd_f, sd_a, first, lams, perms, eps


def test_stage3_step(golden):
    """Actor-driven pre-training step (train_MuRCL.py:235-298, stage 3) + the stage-2 PPO update on its rollout."""
    g = golden("stage3_step")
    (b, k, d, fs, T, L, D, hid, proj, phid), feats, clusters, sd_m, sd_f, sd_a, first, lams, perms, eps = _stage3_inputs(g)
    r = O.pretrain_step_stage3(feats, clusters, sd_m, sd_f, sd_a, first_actions=first, lams=lams, perms=perms, eps=eps,
                               action_std=float(g["std"]), T=T, feat_size=fs, temperature=float(g["tau"]))
    for t in range(T):
        for v in range(2):
            assert_close(r["actions"][t][v], g[f"act{t}_{v}"], TOL, f"act{t}_{v}")
        assert_close(r["losses"][t], g[f"loss{t}"], TOL, f"loss{t}")
        if t >= 1:
            assert_close(r["rewards"][t - 1], g[f"reward{t}"], 1e-4, f"reward{t}", floor=1e-3)
    assert_close(r["loss"], g["loss"], TOL, "loss")
    for v in range(2):
        assert_close(r["logprobs"][v], g[f"logprobs_{v}"], TOL, f"logprobs_{v}")
    n = 0
    for key, val in g.items():
        if key.startswith("grad.m.") or key.startswith("grad.f."):
            assert_close(sample(r["grads"][key[5:]].numpy()), val, 5e-5, key, floor=1e-5)
            n += 1
    assert n >= 10
    # stage 2: both memories update the same policy in turn, Adam state carried across the two updates
    params = {kk: vv.detach().clone().requires_grad_(True) for kk, vv in sd_a.items()}
    opt = torch.optim.Adam(list(params.values()), lr=float(g["ppo_lr"]), betas=(0.9, 0.999))
    for v in range(2):
        ret = O.ppo_returns(r["rewards"], float(g["ppo_gamma"]))
        for _ in range(int(g["ppo_K_epochs"])):
            loss = O.ppo_loss(params, r["states"][v], r["roll_actions"][v], r["logprobs"][v], ret,
                              action_std=float(g["std"]), eps_clip=0.2)
            opt.zero_grad()
            loss.backward()
            opt.step()
    for key, val in g.items():
        if key.startswith("ppo_delta."):
            assert_close(sample((params[key[10:]].detach() - sd_a[key[10:]]).numpy()), val, 2e-2, key, floor=1e-5)


def test_actor_conv(golden):
    """ActorCritic(policy_conv=True): the 1x1-convolution state encoder (models/rlmil.py:30-37)."""
    g = golden("actor_conv")
    sd = leaf_state({n[3:]: torch.from_numpy(v) for n, v in g.items() if n.startswith("sd.")})
    std = float(g["std"])
    h, states, actions = None, [], []
    for t in range(2):
        st = torch.from_numpy(g[f"state{t}"])
        a, lp, h, _ = O.actor_act(st, h, sd, std, torch.from_numpy(g[f"eps{t}"]))
        assert_close(a, g[f"action{t}"], TOL, f"action{t}")
        assert_close(lp, g[f"logprob{t}"], TOL, f"logprob{t}")
        states.append(st)
        actions.append(a.detach())
    lp, val, _ = O.actor_evaluate(torch.stack(states, 0), torch.stack(actions, 0), sd, std)
    assert_close(lp, g["eval_logprob"], TOL, "evaluate.logprob")
    assert_close(val, g["eval_value"], TOL, "evaluate.value")
    ((lp * torch.from_numpy(g["cot_l"])).sum() + (val * torch.from_numpy(g["cot_v"])).sum()).backward()
    _check_grads(g, sd)
