"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the golden fixtures.

Tolerances are the north-star's: indices / gathers bit-exact; fp32 mode 1e-5 relative on attention
weights, bag vectors and losses (gradients 1e-4, they sum thousands of fp32 terms in a different order);
bf16 mode 2e-2 relative.  Run with ``pytest -m gpu`` on a B200.
"""
import math
import os

import numpy as np
import pytest
import torch

from murcl_b200 import synth
from oracle import murcl_oracle as O
from tests.helpers import assert_close, assert_close_elementwise, leaf_state, sample

pytestmark = pytest.mark.gpu

FP32_OUT, FP32_GRAD = 1e-5, 1e-4
BF16_OUT, BF16_GRAD = 2e-2, 6e-2
# attention weights, EVERY element against its own reference value (floor = the mean weight): fp32 mode sees the
# summation-order noise of |s| ~ 10 scores (1e-6 * |s| absolute on s = relative on p); bf16 mode the rounding of h / uv
FP32_ATTN, BF16_ATTN = 2e-5, 5e-2
DEV = "cuda"


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    assert torch.cuda.is_available(), "these tests need a GPU"
    from murcl_b200 import _lib
    _lib.load()
    torch.backends.cuda.matmul.allow_tf32 = False


def _load(module, sd):
    module.load_state_dict(sd, strict=True)
    return module.to(DEV)


def _zero_grad_key(k):
    """Final attention bias: its gradient is analytically zero (softmax shift invariance), only noise remains."""
    return k.endswith(("attention.2.bias", "attention_c.bias", "module.2.bias", "module.3.bias"))


def _grads(module):
    return {n: p.grad.detach().cpu() if p.grad is not None else torch.zeros_like(p).cpu() for n, p in module.named_parameters()}


# ------------------------------------------------------------------------------------------------
# (1) packer
# ------------------------------------------------------------------------------------------------
def test_get_feats_golden(golden):
    from murcl_b200.dropin import datasets
    g = golden("get_feats")
    feats, clusters, _ = synth.make_bags(g["sizes"].tolist(), int(g["d"]), int(g["k"]), seed=int(g["seed_bags"]))
    b, src, dst = g["merged_cluster"].tolist()
    clusters[b][src] = sorted(clusters[b][src] + clusters[b][dst])
    clusters[b][dst] = []
    feat_list = [f.unsqueeze(0).to(DEV) for f in feats]
    out = datasets.get_feats(feat_list, clusters, torch.from_numpy(g["actions"]).to(DEV), feat_size=int(g["fs"]))
    assert out.shape == g["out"].shape
    assert np.array_equal(out.cpu().numpy(), g["out"])


@pytest.mark.parametrize("fs,k,d", [(64, 5, 8), (1024, 10, 512), (100, 3, 20)])
def test_get_feats_random(fs, k, d):
    from murcl_b200.csr import BagStore
    sizes = [3 * fs, fs // 2, fs, fs + 3, 7, 2 * fs + 1, 5000]
    feats, clusters, labels = synth.make_bags(sizes, d, k, seed=5)
    g = synth.gen(6)
    actions = torch.rand(len(sizes), k, generator=g)
    actions[0, 0], actions[1, 1], actions[2, 0], actions[3, 2] = 0.0, 1.0, 1.0, 0.0
    want, kept = O.get_feats(feats, clusters, actions, fs)
    store = BagStore.from_cluster_lists(feats, clusters, DEV)
    sel_idx, sel_cnt = store.select(actions.to(DEV), fs)
    out = store.gather(sel_idx)
    assert torch.equal(out.cpu(), want)
    assert sel_cnt.cpu().tolist() == [len(x) for x in kept]
    offs = store.offsets_host
    for b, idx in enumerate(kept):
        got = sel_idx[b, : len(idx)].cpu().numpy().astype(np.int64) - offs[b]
        assert np.array_equal(got, idx)
        assert bool((sel_idx[b, len(idx):] == -1).all())
    # the on-disk ingest path (per-patch labels -> ranks on the device) builds the same store
    store2 = BagStore.from_labels(feats, labels, k, DEV)
    assert torch.equal(store2.patch_rank, store.patch_rank)
    assert torch.equal(store2.cluster_sizes, store.cluster_sizes)
    assert torch.equal(store2.pack(actions.to(DEV), fs).cpu(), want)
    # bf16 output is the fp32 result rounded to nearest even
    assert torch.equal(store.gather(sel_idx, out_dtype=torch.bfloat16).cpu(), want.to(torch.bfloat16))


def test_selection_sweep_on_device(golden):
    """The brute-force selection fixture, one single-cluster bag per case."""
    from murcl_b200.csr import BagStore
    g = golden("selection_sweep")
    cases, acts = g["cases"][:1500], g["actions"][:1500]
    for fs in sorted(set(cases[:, 0].tolist())):
        sel = np.nonzero(cases[:, 0] == fs)[0]
        feats, clusters = [], []
        for i in sel:
            n = cases[i][2]
            feats.append(torch.zeros(int(cases[i][1]), 1))
            clusters.append([list(range(int(n)))])
        store = BagStore.from_cluster_lists(feats, clusters, DEV)
        a = torch.from_numpy(acts[sel]).reshape(-1, 1).to(DEV)
        sel_idx, sel_cnt = store.select(a, int(fs))
        cnt = sel_cnt.cpu().numpy()
        first = sel_idx[:, 0].cpu().numpy()
        for j, i in enumerate(sel):
            want_cnt = min(int(cases[i][4]), int(fs))
            assert cnt[j] == want_cnt, cases[i]
            if want_cnt:
                assert first[j] - store.offsets_host[j] == cases[i][3], cases[i]


def test_mixup_bit_exact(golden):
    from murcl_b200.dropin import datasets
    x = torch.randn(6, 40, 24, generator=synth.gen(3)).to(DEV)
    torch.manual_seed(11)
    out, lam, perm = datasets.mixup(x, 0.9)
    want = O.mixup_apply(x.cpu(), lam.cpu(), perm.cpu())
    assert torch.equal(out.cpu(), want)
    assert float(lam.min()) >= 0.9 and float(lam.max()) < 1.0
    assert sorted(perm.cpu().tolist()) == list(range(6))
    # fused select + gather + mixup == get_feats then mixup
    from murcl_b200.csr import BagStore
    feats, clusters, _ = synth.make_bags([300, 90, 64, 200], 16, 4, seed=8)
    store = BagStore.from_cluster_lists(feats, clusters, DEV)
    actions = torch.rand(4, 4, generator=synth.gen(9))
    lam = 0.9 + 0.1 * torch.rand(4, 1, generator=synth.gen(10))
    perm = torch.randperm(4, generator=synth.gen(12))
    fused = store.pack(actions.to(DEV), 64, lam.to(DEV), perm.to(DEV))
    dense, _ = O.get_feats(feats, clusters, actions, 64)
    assert torch.equal(fused.cpu(), O.mixup_apply(dense, lam, perm))


def test_bf16_slide_store():
    """bf16 mode may keep the slide features in bf16 (half the store and the per-step H2D): selection is unchanged,
    the mix is still fp32 arithmetic on the stored values."""
    from murcl_b200.csr import BagStore, HostBags
    feats, clusters, labels = synth.make_bags([300, 90, 64, 200], 32, 4, seed=8)
    host = HostBags(feats, labels, 4, pin=True, dtype=torch.bfloat16)
    store = BagStore.empty_like_host(host, DEV)
    store.copy_from_host(host)
    ref_store = BagStore.from_cluster_lists(feats, clusters, DEV)
    assert torch.equal(store.patch_rank, ref_store.patch_rank) and torch.equal(store.cluster_sizes, ref_store.cluster_sizes)
    actions = torch.rand(4, 4, generator=synth.gen(9))
    lam = 0.9 + 0.1 * torch.rand(4, 1, generator=synth.gen(10))
    perm = torch.randperm(4, generator=synth.gen(12))
    got = store.pack(actions.to(DEV), 64, lam.to(DEV), perm.to(DEV), torch.bfloat16)
    dense, _ = O.get_feats([f.bfloat16().float() for f in feats], clusters, actions, 64)
    assert torch.equal(got.cpu(), O.mixup_apply(dense, lam, perm).to(torch.bfloat16))


# ------------------------------------------------------------------------------------------------
# dense layers
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K", [(300, 64, 48), (1000, 512, 512), (129, 130, 70), (5, 2, 1024), (2048, 128, 512)])
def test_linear_simt(M, N, K):
    from murcl_b200 import ops
    g = synth.gen(M + N + K)
    x = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / math.sqrt(K)
    b = torch.randn(N, generator=g)
    dy = torch.randn(M, N, generator=g)
    mask_src = torch.randn(M, K, generator=g)
    y_ref = torch.relu(x.double() @ w.double().t() + b.double())
    xd, wd, bd, dyd = x.to(DEV), w.to(DEV), b.to(DEV), dy.to(DEV)
    os.environ["MURCL_GEMM"] = "simt"
    try:
        y = ops.linear_fwd(xd, wd, bd, ops.ACT_RELU)
        assert_close(y, y_ref.float(), 2e-6, "fwd")
        cs = torch.zeros(K, device=DEV)
        dx = ops.linear_bwd_input(dyd, wd, mask_src.to(DEV), col_sum=cs)
        dx_ref = (dy.double() @ w.double()) * (mask_src > 0)
        assert_close(dx, dx_ref.float(), 2e-6, "bwd_input")
        assert_close(cs, dx_ref.sum(0).float(), 1e-5, "column sums")
        dw, db = ops.linear_bwd_weight(dyd, xd)
        assert_close(dw, (dy.double().t() @ x.double()).float(), 5e-6, "bwd_weight")
        assert_close(db, dy.double().sum(0).float(), 5e-6, "bias grad")
        # bf16 storage through the same kernels (fp32 accumulate)
        xb, wb = xd.bfloat16(), wd.bfloat16()
        yb = ops.linear_fwd(xb, wb, bd, ops.ACT_NONE, torch.float32)
        ref = xb.float().double().cpu() @ wb.float().double().cpu().t() + b.double()
        assert_close(yb, ref.float(), 2e-6, "bf16 storage fwd")
    finally:
        os.environ.pop("MURCL_GEMM", None)


# ------------------------------------------------------------------------------------------------
# (2) ABMIL / CLAM_SB
# ------------------------------------------------------------------------------------------------
def test_abmil_golden_small(golden):
    from murcl_b200.dropin import abmil
    g = golden("abmil_small")
    dim_in, L, D = g["dims"].tolist()
    m = _load(abmil.ABMIL(dim_in, L=L, D=D, dim_out=2, precision="fp32"), synth.abmil_state(dim_in, L, D, 2, seed=31))
    feats, _, _ = synth.make_bags(g["sizes"].tolist(), dim_in, 3, seed=32)
    bags = [f.to(DEV).requires_grad_(True) for f in feats]
    out, det = m(bags)
    assert not det.requires_grad
    assert_close(out, g["out"], FP32_OUT, "out")
    (out * torch.from_numpy(g["cot"]).to(DEV)).sum().backward()
    gr = _grads(m)
    for k, v in g.items():
        if k.startswith("grad."):
            assert_close(sample(gr[k[5:]].numpy()), v, FP32_GRAD, k, floor=1e-1 if _zero_grad_key(k) else 1e-4)
    for i, b in enumerate(bags):
        assert_close(b.grad, g[f"grad_input.{i}"], FP32_GRAD, f"dx{i}")
    with torch.no_grad():
        xb = torch.stack([feats[0][:40], feats[1][:40]]).to(DEV)
        assert_close(m(xb)[0], g["out_dense"], FP32_OUT, "dense batch")
        assert_close(m(feats[0].unsqueeze(0).to(DEV))[0], g["out_single"], FP32_OUT, "single")
    with pytest.raises(TypeError):
        m("not a tensor")


@pytest.mark.parametrize("precision,tol_out,tol_grad", [("fp32", FP32_OUT, FP32_GRAD), ("bf16", BF16_OUT, BF16_GRAD)])
def test_abmil_full_size(golden, precision, tol_out, tol_grad):
    """Camelyon16-shaped: D_in=512, L=512, D=128; ragged bags incl. cfg1's N~2000."""
    from murcl_b200.dropin import abmil
    sd = synth.abmil_state(512, 512, 128, 2, seed=31)
    m = _load(abmil.ABMIL(512, precision=precision), sd)
    if precision == "fp32":
        g = golden("abmil_full")
        feats, _, _ = synth.make_bags(g["sizes"].tolist(), 512, 3, seed=32)
        with torch.no_grad():
            assert_close(m([f.to(DEV) for f in feats])[0], g["out"], tol_out, "golden full")
    from murcl_b200 import ops
    sizes = [2000, 333, 1024]
    feats, _, _ = synth.make_bags(sizes, 512, 3, seed=77)
    sdl = leaf_state(sd)
    want = O.abmil_forward(feats, sdl)
    cot = torch.randn(want.shape, generator=synth.gen(78))
    ops._debug_save = {}
    try:
        out, _ = m([f.to(DEV) for f in feats])
        hs_dev = [h.float().cpu() for h in ops._debug_save["hs"]]
    finally:
        ops._debug_save = None
    assert_close(out, want, tol_out, "out")
    with torch.no_grad():
        p_want = torch.cat([O.abmil_attention(f, sd) for f in feats])
    assert_close(m.last_attention, p_want, tol_out, "attention weights (max-norm)")
    assert_close_elementwise(m.last_attention, p_want, FP32_ATTN if precision == "fp32" else BF16_ATTN, "attention weights")
    (out * cot.to(DEV)).sum().backward()
    gr = _grads(m)
    if precision == "fp32":
        # Gradients: fp64 evaluation of the reference that adopts the DEVICE's ReLU decisions.  Two correct fp32 evaluations
        # can put a pre-activation that is zero to rounding on different sides of the ReLU (measured here: 1 of 1.7 M units,
        # with the FFMA and with the tensor-core GEMMs alike); that unit's whole gradient contribution then differs (up to
        # 3e-4 of the layer's weight gradient for a high-attention row) although every arithmetic step is accurate to 1e-6.
        off = np.concatenate([[0], np.cumsum(sizes)])
        sd64 = leaf_state(sd, torch.float64)
        outs64 = []
        with torch.no_grad():
            ref_h = torch.cat(feats)
            flips = 0
            for j, i in enumerate((0, 3, 6)):
                ref_h = torch.relu(torch.nn.functional.linear(ref_h, sd[f"encoder.{i}.weight"], sd[f"encoder.{i}.bias"]))
                flips += int(((ref_h > 0) != (hs_dev[j + 1] > 0)).sum())
        assert flips <= 8, f"{flips} ReLU decisions differ from the fp32 reference (expected a handful of 1.7 M)"
        for b, f in enumerate(feats):
            masks = [hs_dev[j + 1][off[b]:off[b + 1]] > 0 for j in range(3)]
            outs64.append(O.abmil_bag(f.double(), sd64, relu_masks=masks))
        (torch.cat(outs64, 0) * cot.double()).sum().backward()
        for k, p in sd64.items():
            if k.startswith("fc."):
                continue
            assert_close(gr[k], p.grad.float(), 1e-5, k, floor=1e-1 if _zero_grad_key(k) else 1e-7)
    else:
        (want * cot).sum().backward()
        for k, p in sdl.items():
            if k.startswith("fc."):
                continue
            assert_close(gr[k], p.grad, tol_grad, k, floor=1e-1 if _zero_grad_key(k) else 1e-7)


def test_large_bag_stress_cfg5():
    """BASELINE config 5 shape: one bag of N=100k patches x 1024-d, bf16 fused pooling (CLAM_SB small).  Forward,
    attention weights and every parameter gradient are compared with the fp32 oracle at the bf16 tolerances; a ragged
    second bag rides along."""
    from murcl_b200.dropin import clam
    sd = synth.clam_state(1024, "small", True, False, 2, seed=71, peak=3.0)
    m = _load(clam.CLAM_SB(gate=True, size_arg="small", in_dim=1024, precision="bf16"), sd).eval()
    feats, _, _ = synth.make_bags([100000, 2500], 1024, 3, seed=72)
    sdl = leaf_state(sd)
    torch.set_num_threads(os.cpu_count() or 1)
    res = [O.clam_sb_bag(f, sdl, gate=True) for f in feats]
    want = torch.cat([r[0] for r in res], 0)
    cot = torch.randn(want.shape, generator=synth.gen(73))
    (want * cot).sum().backward()
    out, _ = m([f.to(DEV) for f in feats])
    assert_close(out, want, BF16_OUT, "100k-patch bag")
    p_want = torch.cat([r[1]["attention"] for r in res])
    assert_close(m.last_attention, p_want, BF16_OUT, "attention weights (max-norm)")
    assert_close_elementwise(m.last_attention, p_want, BF16_ATTN, "attention weights")
    (out * cot.to(DEV)).sum().backward()
    gr = _grads(m)
    n = 0
    for k, p in sdl.items():
        if p.grad is None or k.startswith("classifiers") or k.startswith("instance_classifiers"):
            continue
        assert_close(gr[k], p.grad, BF16_GRAD, k, floor=1e-1 if _zero_grad_key(k) else 1e-6)
        n += 1
    assert n >= 7


@pytest.mark.parametrize("gate", [True, False])
@pytest.mark.parametrize("dropout", [False, True])
@pytest.mark.parametrize("subtyping", [False, True])
def test_clam_golden(golden, gate, dropout, subtyping):
    from murcl_b200.dropin import clam
    g = golden(f"clam_g{int(gate)}_d{int(dropout)}_s{int(subtyping)}")
    in_dim, n_classes = int(g["in_dim"]), int(g["n_classes"])
    m = clam.CLAM_SB(gate=gate, size_arg="small", dropout=dropout, k_sample=8, n_classes=n_classes, subtyping=subtyping,
                     in_dim=in_dim, precision="fp32")
    m = _load(m, synth.clam_state(in_dim, "small", gate, dropout, n_classes, seed=41)).eval()
    feats, _, _ = synth.make_bags(g["sizes"].tolist(), in_dim, 3, seed=42)
    labels = g["labels"].tolist()
    tot = 0.0
    x0 = feats[0].to(DEV).requires_grad_(True)
    for i, f in enumerate(feats):
        x = x0 if i == 0 else f.to(DEV)
        out, det, res = m(x.unsqueeze(0), label=torch.tensor([labels[i]]), instance_eval=True)
        assert isinstance(res, dict)
        assert_close(out, g[f"out{i}"], FP32_OUT, f"M{i}")
        assert_close(res["instance_loss"], g[f"inst_loss{i}"], FP32_OUT, f"inst{i}")
        assert np.array_equal(res["inst_labels"], g[f"inst_labels{i}"])
        assert np.array_equal(res["inst_preds"], g[f"inst_preds{i}"])
        raw = m.bag_forward(f.to(DEV), attention_only=True)
        assert_close(raw, g[f"raw_scores{i}"], FP32_OUT, f"raw{i}")
        tot = tot + (out * torch.from_numpy(g[f"cot{i}"]).to(DEV)).sum() + 0.3 * res["instance_loss"]
    tot.backward()
    gr = _grads(m)
    for k, v in g.items():
        if k.startswith("grad."):
            assert_close(sample(gr[k[5:]].numpy()), v, FP32_GRAD, k, floor=1e-1 if _zero_grad_key(k) else 1e-4)
    assert_close(x0.grad, g["grad_input.0"], FP32_GRAD, "dx0")
    with torch.no_grad():
        outs, det = m([f.to(DEV) for f in feats])
        assert_close(outs, g["out_list"], FP32_OUT, "list batch")
        # batch + instance_eval returns a LIST of dicts (clam.py:183-195)
        o3 = m([f.to(DEV) for f in feats], label=torch.tensor(labels), instance_eval=True)
        assert isinstance(o3[2], list) and len(o3[2]) == len(feats)
        for i in range(len(feats)):
            assert_close(o3[2][i]["instance_loss"], g[f"inst_loss{i}"], FP32_OUT, f"batched inst{i}")


@pytest.mark.parametrize("gate", [True, False])
def test_clam_train_mode_dropout(gate):
    """MuRCL builds CLAM_SB with dropout=True and trains in model.train() (train_MuRCL.py:85,202).  The kernel draws
    its own keep masks (torch's Philox stream cannot be reproduced), so the masks are read back from the saved
    activations and handed to the oracle: forward and backward must then agree at the fp32 tolerance."""
    from murcl_b200 import ops
    from murcl_b200.dropin import clam
    in_dim, n = 48, 700
    sd = synth.clam_state(in_dim, "small", gate, True, 2, seed=61)
    m = _load(clam.CLAM_SB(gate=gate, size_arg="small", dropout=True, n_classes=2, in_dim=in_dim, precision="fp32"), sd)
    m.train()
    feats, _, _ = synth.make_bags([n], in_dim, 3, seed=62)
    x = feats[0].to(DEV).requires_grad_(True)
    ops._debug_save = {}
    try:
        out, _ = m(x.unsqueeze(0))
        saved = ops._debug_save
    finally:
        ops._debug_save = None
    q = 1.0 / 0.75
    h_d, uv_d = saved["hs"][1].float().cpu(), saved["uv"].float().cpu()
    D = 256
    masks = {"enc": (h_d != 0).float() * q, "attn_a": (uv_d[:, :D] != 0).float() * q}
    if gate:
        masks["attn_b"] = (uv_d[:, D:] != 0).float() * q
    keep = float((uv_d != 0).float().mean())
    assert abs(keep - 0.75) < 0.01, f"keep fraction {keep}"
    sdl = leaf_state(sd)
    xr = feats[0].clone().requires_grad_(True)
    want, _ = O.clam_sb_bag(xr, sdl, gate=gate, dropout_layers=True, masks=masks)
    assert_close(out, want, FP32_OUT, "train-mode M")
    cot = torch.randn(1, 512, generator=synth.gen(63))
    (want * cot).sum().backward()
    (out * cot.to(DEV)).sum().backward()
    gr = _grads(m)
    for k, p in sdl.items():
        if p.grad is None or k.startswith(("classifiers", "instance")):
            continue
        assert_close(gr[k], p.grad, FP32_GRAD, k, floor=1e-1 if _zero_grad_key(k) else 1e-6)
    assert_close(x.grad, xr.grad, FP32_GRAD, "dx")
    # a second forward draws different masks; eval() is deterministic
    out2, _ = m(x.detach().unsqueeze(0))
    assert not torch.equal(out2, out.detach())
    m.eval()
    with torch.no_grad():
        e1, _ = m(x.detach().unsqueeze(0))
        e2, _ = m(x.detach().unsqueeze(0))
    assert torch.equal(e1, e2)


def test_abmil_train_mode_dropout():
    from murcl_b200 import ops
    from murcl_b200.dropin import abmil
    sd = synth.abmil_state(40, 64, 32, 2, seed=64)
    m = _load(abmil.ABMIL(40, L=64, D=32, dropout=0.3, precision="fp32"), sd).train()
    feats, _, _ = synth.make_bags([500, 120], 40, 3, seed=65)
    ops._debug_save = {}
    try:
        out, _ = m([f.to(DEV) for f in feats])
        saved = ops._debug_save
    finally:
        ops._debug_save = None
    q = 1.0 / 0.7
    sdl = leaf_state(sd)
    wants, lo = [], 0
    for f in feats:
        hi = lo + f.shape[0]
        masks = {"enc0": (saved["hs"][1][lo:hi].float().cpu() != 0).float() * q,
                 "enc1": (saved["hs"][2][lo:hi].float().cpu() != 0).float() * q}
        wants.append(O.abmil_bag(f, sdl, masks))
        lo = hi
    want = torch.cat(wants, 0)
    assert_close(out, want, FP32_OUT, "train-mode out")
    cot = torch.randn(want.shape, generator=synth.gen(66))
    (want * cot).sum().backward()
    (out * cot.to(DEV)).sum().backward()
    gr = _grads(m)
    for k, p in sdl.items():
        if p.grad is None or k.startswith("fc."):
            continue
        assert_close(gr[k], p.grad, FP32_GRAD, k, floor=1e-1 if _zero_grad_key(k) else 1e-6)


def test_clam_big_and_errors(golden):
    from murcl_b200.dropin import clam
    g = golden("clam_big")
    m = _load(clam.CLAM_SB(gate=True, size_arg="big", in_dim=int(g["in_dim"]), precision="fp32"),
              synth.clam_state(int(g["in_dim"]), "big", True, False, 2, seed=44)).eval()
    feats, _, _ = synth.make_bags([60, 33, 100], int(g["in_dim"]), 3, seed=42)
    with torch.no_grad():
        assert_close(m(feats[0].unsqueeze(0).to(DEV))[0], g["out"], FP32_OUT, "big")
    with pytest.raises(RuntimeError):       # fewer than k_sample instances: torch.topk raises upstream
        m(feats[0][:5].unsqueeze(0).to(DEV), label=torch.tensor([1]), instance_eval=True)
    with pytest.raises(TypeError):
        m(3.0)


@pytest.mark.parametrize("precision,tol_out,tol_grad", [("fp32", FP32_OUT, FP32_GRAD), ("bf16", BF16_OUT, BF16_GRAD)])
def test_clam_ragged_cfg2(precision, tol_out, tol_grad):
    """BASELINE config 2 shape: CLAM_SB small + instance loss on ragged bags x 512-d.  Both modes check the bag vectors,
    every attention weight, the instance loss and all gradients.  The instance loss ranks instances by attention weight
    and the ranking carries no gradient (clam.py:107-110); in bf16 mode the bottom-k of near-zero weights legitimately
    differs from the fp32 ranking, so there the oracle ranks by the weights the device produced (same instances on both
    sides) - the loss values and gradients are then comparable at the bf16 tolerance."""
    from murcl_b200.dropin import clam
    sd = synth.clam_state(512, "small", True, False, 2, seed=45)
    m = _load(clam.CLAM_SB(gate=True, size_arg="small", k_sample=8, n_classes=2, subtyping=True, in_dim=512,
                           precision=precision), sd).eval()
    sizes = [2000, 3111, 5000]
    feats, _, _ = synth.make_bags(sizes, 512, 3, seed=46)
    labels = [1, 0, 1]
    cots = [torch.randn(1, 512, generator=synth.gen(50 + i)) for i in range(3)]
    out, det, res = m([f.to(DEV) for f in feats], label=torch.tensor(labels), instance_eval=True)
    p_dev = m.last_attention.detach().cpu()
    off = np.concatenate([[0], np.cumsum(sizes)])
    sdl = leaf_state(sd)
    tot_ref, wants = 0.0, []
    for i, f in enumerate(feats):
        rank_by = None if precision == "fp32" else p_dev[off[i]:off[i + 1]]
        mm, r = O.clam_sb_bag(f, sdl, gate=True, label=labels[i], instance_eval=True, n_classes=2, k_sample=8, subtyping=True,
                              p_select=rank_by)
        wants.append((mm, r))
        tot_ref = tot_ref + (mm * cots[i]).sum() + 0.3 * r["instance_loss"]
    tot_ref.backward()
    tot = 0.0
    for i in range(3):
        assert_close(out[i:i + 1], wants[i][0], tol_out, f"M{i}")
        assert_close(res[i]["instance_loss"], wants[i][1]["instance_loss"], tol_out, f"inst{i}")
        assert np.array_equal(res[i]["inst_labels"], wants[i][1]["inst_labels"])
        if precision == "fp32":
            assert np.array_equal(res[i]["inst_preds"], wants[i][1]["inst_preds"])
        p_i = p_dev[off[i]:off[i + 1]]
        assert_close(p_i, wants[i][1]["attention"], tol_out, f"attention{i} (max-norm)")
        assert_close_elementwise(p_i, wants[i][1]["attention"], FP32_ATTN if precision == "fp32" else BF16_ATTN, f"attention{i}")
        tot = tot + (out[i:i + 1] * cots[i].to(DEV)).sum() + 0.3 * res[i]["instance_loss"]
    tot.backward()
    gr = _grads(m)
    n = 0
    for k, p in sdl.items():
        if k.startswith("classifiers") or p.grad is None:
            continue
        assert_close(gr[k], p.grad, tol_grad, k, floor=1e-1 if _zero_grad_key(k) else 1e-5)
        n += 1
    assert n >= 9


# ------------------------------------------------------------------------------------------------
# (3) DSMIL
# ------------------------------------------------------------------------------------------------
def test_dsmil_golden(golden):
    from murcl_b200.dropin import dsmil
    g = golden("dsmil")
    dim, c = int(g["dim"]), int(g["c"])
    m = dsmil.build_dsmil(dim, c, precision="fp32")
    m.load_state_dict(synth.dsmil_state(dim, c, seed=51), strict=True)
    feats, _, _ = synth.make_bags(g["sizes"].tolist(), dim, 3, seed=52)
    tot, xs = 0.0, []
    for i, f in enumerate(feats):
        x = f.to(DEV).requires_grad_(True)
        xs.append(x)
        classes, bag, bag_det = m(x.unsqueeze(0))
        assert_close(classes, g[f"classes{i}"], FP32_OUT, "classes")
        assert_close(bag, g[f"bag{i}"], FP32_OUT, "bag")
        tot = tot + (bag * torch.from_numpy(g[f"cot_b{i}"]).to(DEV)).sum() + (classes * torch.from_numpy(g[f"cot_c{i}"]).to(DEV)).sum()
    tot.backward()
    gr = _grads(m)
    for k, v in g.items():
        if k.startswith("grad."):
            assert_close(sample(gr[k[5:]].numpy()), v, FP32_GRAD, k, floor=1e-4)
    for i, x in enumerate(xs):
        assert_close(x.grad, g[f"grad_input.{i}"], FP32_GRAD, f"dx{i}")
    # batch form: lists of per-bag outputs, bags concatenated (dsmil.py:26-36,83-91)
    with torch.no_grad():
        classes, bag, _ = m([f.unsqueeze(0).to(DEV) for f in feats])
        assert isinstance(classes, list) and bag.shape == (2, c, dim)
        assert_close(bag[1:2], g["bag1"], FP32_OUT, "batched bag")
    with pytest.raises(TypeError):
        m(1.0)


@pytest.mark.parametrize("precision,tol,tol_grad", [("fp32", FP32_OUT, FP32_GRAD), ("bf16", BF16_OUT, BF16_GRAD)])
def test_dsmil_tcga_shape(precision, tol, tol_grad):
    """BASELINE config 4 shape: N=10k x 1024-d, values and gradients.  V is applied after pooling (a reassociation):
    the fp32 comparison is made against the fp64 evaluation of the reference formula, which both the reference's fp32
    result and ours must match to 1e-5."""
    from murcl_b200.dropin import dsmil
    sd = synth.dsmil_state(1024, 2, seed=55)
    m = dsmil.build_dsmil(1024, 2, precision=precision)
    m.load_state_dict(sd, strict=True)
    feats, _, _ = synth.make_bags([10000], 1024, 3, seed=56)
    torch.set_num_threads(os.cpu_count() or 1)
    sdl = leaf_state(sd, torch.float64)
    want_c, want_b = O.dsmil_bag(feats[0].double(), sdl)
    g = synth.gen(57)
    cot_b, cot_c = torch.randn(want_b.shape, generator=g), torch.randn(want_c.shape, generator=g)
    ((want_b * cot_b.double()).sum() + (want_c * cot_c.double()).sum()).backward()
    with torch.no_grad():                                   # the reference's own fp32 evaluation sits within the same bound
        ref32_c, ref32_b = O.dsmil_bag(feats[0], sd)
        assert_close(ref32_b, want_b.float(), FP32_OUT, "fp32 reference vs fp64")
    classes, bag, _ = m(feats[0].unsqueeze(0).to(DEV))
    assert_close(classes, want_c.float(), tol, "classes")
    assert_close(bag, want_b.float(), tol, "bag")
    ((bag * cot_b.to(DEV)).sum() + (classes * cot_c.to(DEV)).sum()).backward()
    gr = _grads(m)
    n = 0
    for k, p in sdl.items():
        if p.grad is None:
            continue
        assert_close(gr[k], p.grad.float(), tol_grad, k, floor=1e-6)
        n += 1
    assert n >= 6


# ------------------------------------------------------------------------------------------------
# (4) NT-Xent
# ------------------------------------------------------------------------------------------------
def test_ntxent_golden(golden):
    from murcl_b200.dropin import losses
    g = golden("ntxent")
    for i in range(3):
        b, d, tau = g[f"cfg{i}"].tolist()
        zi = torch.from_numpy(g[f"zi{i}"]).to(DEV).requires_grad_(True)
        zj = torch.from_numpy(g[f"zj{i}"]).to(DEV).requires_grad_(True)
        crit = losses.NT_Xent(int(b), tau)
        loss = crit(zi, zj)
        assert loss.dim() == 0
        assert_close(loss, g[f"loss{i}"], FP32_OUT, "loss")
        loss.backward()
        assert_close(zi.grad, g[f"gzi{i}"], FP32_GRAD, "gzi")
        assert_close(zj.grad, g[f"gzj{i}"], FP32_GRAD, "gzj")
        assert_close(crit.last_cosine, g[f"cos{i}"], FP32_OUT, "cos")
    same = torch.randn(1, 8, generator=synth.gen(69)).repeat(6, 1).to(DEV)
    assert abs(float(losses.NT_Xent(6, 1.0)(same, same.clone())) - math.log(11.0)) < 1e-5
    with pytest.raises(RuntimeError):
        losses.NT_Xent(4, 1.0)(same, same)


@pytest.mark.parametrize("B,d,tau", [(128, 128, 1.0), (128, 128, 0.07), (1024, 128, 0.5)])
def test_ntxent_pretrain_shape(B, d, tau):
    from murcl_b200 import ops
    g = synth.gen(B + d)
    zi = torch.randn(B, d, generator=g, dtype=torch.float64)
    zj = 0.7 * zi + torch.randn(B, d, generator=g, dtype=torch.float64)
    zi32, zj32 = zi.float().requires_grad_(True), zj.float().requires_grad_(True)
    zi.requires_grad_(True); zj.requires_grad_(True)
    want = O.nt_xent(zi, zj, tau)          # fp64 evaluation of the same formula
    want.backward()
    a, b = zi32.detach().to(DEV).requires_grad_(True), zj32.detach().to(DEV).requires_grad_(True)
    loss, cos = ops.ntxent(a, b, tau)
    loss.backward()
    assert_close(loss, want.float(), FP32_OUT, "loss")
    assert_close(a.grad, zi.grad.float(), FP32_GRAD, "gzi")
    assert_close(b.grad, zj.grad.float(), FP32_GRAD, "gzj")


# ------------------------------------------------------------------------------------------------
# (5) heads
# ------------------------------------------------------------------------------------------------
def test_full_layer_golden(golden):
    from murcl_b200.dropin import rlmil
    g = golden("full_layer")
    fnum, hid, cls, b = g["dims"].tolist()
    m = _load(rlmil.Full_layer(fnum, hid, True, cls), synth.full_layer_state(fnum, hid, cls, seed=71))
    tot, xs = 0.0, []
    for t in range(3):
        x = torch.from_numpy(g[f"x{t}"]).to(DEV).requires_grad_(True)
        xs.append(x)
        out = m(x, restart=(t == 0))
        assert_close(out, g[f"out{t}"], FP32_OUT, f"out{t}")
        tot = tot + (out * torch.from_numpy(g[f"cot{t}"]).to(DEV)).sum()
    tot.backward()
    gr = _grads(m)
    for k, v in g.items():
        if k.startswith("grad."):
            assert_close(sample(gr[k[5:]].numpy()), v, FP32_GRAD, k, floor=1e-4)
    for t, x in enumerate(xs):
        assert_close(x.grad, g[f"grad_input.{t}"], FP32_GRAD, f"dx{t}")


def test_actor_golden(golden):
    from murcl_b200.dropin import rlmil
    g = golden("actor")
    sdim, hid, k, b = g["dims"].tolist()
    ppo = rlmil.PPO(sdim, sdim, hid, False, action_std=float(g["std"]), action_size=k)
    ppo.policy_old.load_state_dict(synth.actor_state(sdim, hid, k, seed=81), strict=True)
    mem = rlmil.Memory()
    for t in range(3):
        state = torch.from_numpy(g[f"state{t}"]).to(DEV)
        action = ppo.policy_old.act(state, mem, restart_batch=(t == 0), training=True, eps=torch.from_numpy(g[f"eps{t}"]).to(DEV))
        assert_close(action, g[f"action{t}"], FP32_OUT, "action")
        assert_close(mem.logprobs[-1], g[f"logprob{t}"], FP32_OUT, "logprob")
        assert_close(mem.hidden[-1][0], g[f"hidden{t}"], FP32_OUT, "hidden")
    assert len(mem.states) == 3 and len(mem.actions) == 3
    # PPO.update runs end to end on the kernels (rewards as train_MuRCL.py:283-288 stores them)
    for _ in range(2):
        mem.rewards.append(torch.randn(1, b, device=DEV))
    mem.states, mem.actions, mem.logprobs = mem.states[:2], mem.actions[:2], mem.logprobs[:2]
    before = ppo.policy.actor[0].weight.detach().clone()
    ppo.update(mem)
    assert not torch.equal(before, ppo.policy.actor[0].weight.detach())
    assert torch.equal(ppo.policy.actor[0].weight, ppo.policy_old.actor[0].weight)


def test_ppo_update_golden(golden):
    """PPO.evaluate / PPO.update (models/rlmil.py:99-127,152-184) against the reference's own run: rollout, evaluate
    outputs, the first epoch's gradients, and the weights after K_epochs Adam steps."""
    from murcl_b200.dropin import rlmil
    g = golden("ppo_update")
    sdim, hid, k, b, T = g["dims"].tolist()
    std = float(g["std"])
    sd = synth.actor_state(sdim, hid, k, seed=int(g["seed_actor"]))
    ppo = rlmil.PPO(sdim, sdim, hid, False, action_std=std, lr=float(g["lr"]), gamma=float(g["gamma"]),
                    K_epochs=int(g["K_epochs"]), eps_clip=float(g["eps_clip"]), action_size=k)
    ppo.policy.load_state_dict(sd, strict=True)
    ppo.policy_old.load_state_dict(sd, strict=True)
    mem = rlmil.Memory()
    for t in range(T):
        state = torch.from_numpy(g[f"state{t}"]).to(DEV)
        action = ppo.policy_old.act(state, mem, restart_batch=(t == 0), training=True, eps=torch.from_numpy(g[f"eps{t}"]).to(DEV))
        assert_close(action, g[f"action{t}"], FP32_OUT, f"action{t}")
        assert_close(mem.logprobs[-1], g[f"logprob{t}"], FP32_OUT, f"logprob{t}")
        mem.rewards.append(torch.from_numpy(g[f"reward{t}"]).to(DEV))
    states, actions = torch.stack(mem.states, 0), torch.stack(mem.actions, 0)
    lp, val, ent = ppo.policy.evaluate(states, actions)
    assert_close(lp, g["eval_logprob"], FP32_OUT, "evaluate.logprob")
    assert_close(val, g["eval_value"], FP32_OUT, "evaluate.value")
    assert_close(ent, g["eval_entropy"], FP32_OUT, "evaluate.entropy")
    # first epoch's objective and gradients (rlmil.py:169-181)
    ret = torch.from_numpy(g["returns"]).to(DEV)
    ratios = torch.exp(lp - torch.stack(mem.logprobs, 0).detach())
    adv = ret - val.detach()
    loss = -torch.min(ratios * adv, torch.clamp(ratios, 0.8, 1.2) * adv) + 0.5 * ppo.MseLoss(val, ret) - 0.01 * ent
    ppo.policy.zero_grad()
    loss.mean().backward()
    assert_close(loss.mean(), g["loss0"], FP32_OUT, "loss0")
    gr = _grads(ppo.policy)
    for key, v in g.items():
        if key.startswith("grad."):
            assert_close(sample(gr[key[5:]].numpy()), v, FP32_GRAD, key, floor=1e-6)
    ppo.policy.zero_grad()
    ppo.update(mem)
    n = 0
    for name, p in ppo.policy.named_parameters():
        # Adam's first steps move every weight by ~lr whatever the gradient's size: compare the DELTAS to 1 %
        assert_close(sample((p.detach().cpu() - sd[name]).numpy()), g[f"delta.{name}"], 1e-2, f"delta.{name}", floor=1e-6)
        n += 1
    assert n == len(sd)
    for a, b_ in zip(ppo.policy.state_dict().values(), ppo.policy_old.state_dict().values()):
        assert torch.equal(a, b_)


def _stage3_fixture(g):
    b, k, d, fs, T, L, D, hid, proj, phid = g["cfg"].tolist()
    feats, clusters, _ = synth.make_bags(g["sizes"].tolist(), d, k, seed=111)
    draws, eps = [], [None]
    for t in range(T):
        acts = [torch.from_numpy(g[f"act0_{v}"]).to(DEV) for v in range(2)] if t == 0 else None
        draws.append((acts, [torch.from_numpy(g[f"lam{t}_{v}"]).to(DEV) for v in range(2)],
                      [torch.from_numpy(g[f"perm{t}_{v}"]).to(DEV) for v in range(2)]))
        if t >= 1:
            eps.append([torch.from_numpy(g[f"eps{t}_{v}"]).to(DEV) for v in range(2)])
    return (b, k, d, fs, T, L, D, hid, proj, phid), feats, clusters, draws, eps


def _stage3_objects(cfg, g, lr):
    from murcl_b200.dropin import abmil, cl, losses, rlmil
    b, k, d, fs, T, L, D, hid, proj, phid = cfg
    enc = _load(abmil.ABMIL(d, L=L, D=D, dim_out=proj, precision="fp32"), synth.abmil_state(d, L, D, proj, seed=112))
    model = cl.CL(enc, projection_dim=proj, n_features=L)
    fc = _load(rlmil.Full_layer(L, hid, True, proj), synth.full_layer_state(L, hid, proj, seed=113))
    sd_a = synth.actor_state(L, phid, k, seed=114)
    ppo = rlmil.PPO(d, L, phid, False, action_std=float(g["std"]), lr=lr, gamma=float(g["ppo_gamma"]),
                    K_epochs=int(g["ppo_K_epochs"]), action_size=k)
    ppo.policy.load_state_dict(sd_a)
    ppo.policy_old.load_state_dict(sd_a)
    return enc, model, fc, ppo, sd_a, losses.NT_Xent(b, float(g["tau"]))


def test_stage3_step_golden(golden):
    """The benchmarked call sequence - actor-chosen windows (train_MuRCL.py:254-288, stage 3) through
    ``pretrain.pretrain_step`` / ``act_views`` / ``forward_views`` - against the reference's own run of that loop:
    chosen actions, log-probs, per-step losses, rewards, total loss and every gradient."""
    from murcl_b200 import pretrain
    from murcl_b200.csr import BagStore
    from murcl_b200.dropin import rlmil
    g = golden("stage3_step")
    cfg, feats, clusters, draws, eps = _stage3_fixture(g)
    b, k, d, fs, T, L, D, hid, proj, phid = cfg
    enc, model, fc, ppo, sd_a, crit = _stage3_objects(cfg, g, float(g["ppo_lr"]))
    store = BagStore.from_cluster_lists(feats, clusters, DEV)
    mems = [rlmil.Memory(), rlmil.Memory()]
    loss, per_step = pretrain.pretrain_step(store, model, fc, crit, T=T, feat_size=fs, alpha=float(g["alpha"]), stage=3,
                                            ppo=ppo, memories=mems, draws=draws, eps=eps, precision="fp32", keep_memory=True)
    for t in range(T):
        assert_close(per_step[t], g[f"loss{t}"], FP32_OUT, f"loss{t}")
    assert_close(loss, g["loss"], FP32_OUT, "loss")
    for v, m in enumerate(mems):
        assert len(m.actions) == T - 1 and len(m.rewards) == T - 1
        for t in range(1, T):
            assert_close(m.actions[t - 1], g[f"act{t}_{v}"], FP32_OUT, f"act{t}_{v}")
            assert_close(m.rewards[t - 1], g[f"reward{t}"], 1e-4, f"reward{t}", floor=1e-3)
        assert_close(torch.stack(m.logprobs, 0), g[f"logprobs_{v}"], FP32_OUT, f"logprobs_{v}")
    gm, gf = _grads(enc), _grads(fc)
    n = 0
    for key, v in g.items():
        if key.startswith("grad.m."):
            assert_close(sample(gm[key[7:]].numpy()), v, 3e-4, key, floor=1e-5)
            n += 1
        elif key.startswith("grad.f."):
            assert_close(sample(gf[key[7:]].numpy()), v, 3e-4, key, floor=1e-5)
            n += 1
    assert n >= 10
    # the actor was not updated in stage 3
    for name, p in ppo.policy.named_parameters():
        assert torch.equal(p.detach().cpu(), sd_a[name])


def test_stage2_step_golden(golden):
    """Stage 2 (train_MuRCL.py:244-247,296-298): the same rollout under no_grad, then ``ppo.update`` from both views'
    memories; the policy's weight deltas are compared with the reference's."""
    from murcl_b200 import pretrain
    from murcl_b200.csr import BagStore
    from murcl_b200.dropin import rlmil
    g = golden("stage3_step")
    cfg, feats, clusters, draws, eps = _stage3_fixture(g)
    b, k, d, fs, T, L, D, hid, proj, phid = cfg
    enc, model, fc, ppo, sd_a, crit = _stage3_objects(cfg, g, float(g["ppo_lr"]))
    store = BagStore.from_cluster_lists(feats, clusters, DEV)
    mems = [rlmil.Memory(), rlmil.Memory()]
    loss, _ = pretrain.pretrain_step(store, model, fc, crit, T=T, feat_size=fs, alpha=float(g["alpha"]), stage=2,
                                     ppo=ppo, memories=mems, draws=draws, eps=eps, precision="fp32")
    assert_close(loss, g["loss"], FP32_OUT, "loss")
    assert all(p.grad is None for p in enc.parameters()), "stage 2 must not touch the MIL model"
    assert len(mems[0].actions) == 0, "the reference clears the memories after the update"
    n = 0
    for name, p in ppo.policy.named_parameters():
        assert_close(sample((p.detach().cpu() - sd_a[name]).numpy()), g[f"ppo_delta.{name}"], 2e-2, f"ppo_delta.{name}", floor=1e-5)
        n += 1
    assert n == len(sd_a)


# ------------------------------------------------------------------------------------------------
# composition: one miniature pre-training step against the reference-generated fixture
# ------------------------------------------------------------------------------------------------
def _mini_step_draws(g_cfg, seed):
    b, k = g_cfg[0], g_cfg[1]
    gen = synth.gen(seed)
    draws = []
    for _ in range(g_cfg[4]):
        acts = [torch.rand((b, k), generator=gen) for _ in range(2)]
        lams, perms = [], []
        for _ in range(2):
            lams.append(0.9 + torch.rand(b, 1, generator=gen) * (1 - 0.9))
            perms.append(torch.randperm(b, generator=gen))
        draws.append((acts, lams, perms))
    return draws


def test_pretrain_step_golden(golden):
    from murcl_b200 import pretrain
    from murcl_b200.csr import BagStore
    from murcl_b200.dropin import abmil, cl, losses, rlmil
    g = golden("pretrain_step")
    cfg = g["cfg"].tolist()
    b, k, d, fs, T, L, D, hid, proj = cfg
    feats, clusters, _ = synth.make_bags(g["sizes"].tolist(), d, k, seed=91)
    enc = _load(abmil.ABMIL(d, L=L, D=D, dim_out=proj, precision="fp32"), synth.abmil_state(d, L, D, proj, seed=92))
    model = cl.CL(enc, projection_dim=proj, n_features=L)
    fc = _load(rlmil.Full_layer(L, hid, True, proj), synth.full_layer_state(L, hid, proj, seed=93))
    crit = losses.NT_Xent(b, float(g["tau"]))
    store = BagStore.from_cluster_lists(feats, clusters, DEV)
    draws = [([a.to(DEV) for a in acts], [l.to(DEV) for l in lams], [p.to(DEV) for p in perms])
             for acts, lams, perms in _mini_step_draws(cfg, 94)]
    loss, _ = pretrain.pretrain_step(store, model, fc, crit, T=T, feat_size=fs, alpha=float(g["alpha"]), draws=draws,
                                     precision="fp32")
    assert_close(loss, g["loss"], FP32_OUT, "loss")
    gm, gf = _grads(enc), _grads(fc)
    n = 0
    for key, v in g.items():
        if key.startswith("grad.m."):
            assert_close(sample(gm[key[7:]].numpy()), v, 3e-4, key, floor=1e-5)
            n += 1
        elif key.startswith("grad.f."):
            assert_close(sample(gf[key[7:]].numpy()), v, 3e-4, key, floor=1e-5)
            n += 1
    assert n >= 10


@pytest.mark.parametrize("world,b,d", [(4, 128, 128), (8, 16, 64), (3, 20, 256)])
def test_ntxent_own_rows_passes_match_the_full_kernel(world, b, d):
    """murcl_ntxent_lse_slab + (exchange of the per-row statistics) + murcl_ntxent_grad_slab - what every rank runs from 4
    ranks on - reproduce the all-rows kernel (utils/losses.py:24-41): loss = sum of the ranks' shares, the same gradient
    rows, the same cosines; and both agree with the fp64 oracle."""
    from murcl_b200 import ops
    Bg = world * b
    z = torch.randn(2 * Bg, d, generator=synth.gen(77)).to(DEV)
    loss_full, dz_full, cos_full = ops.ntxent_raw(z, Bg, 0.7, True)
    lse = torch.empty(2 * Bg, device=DEV); inv = torch.empty(2 * Bg, device=DEV)
    shares, cos = [], torch.empty(Bg, device=DEV)
    for r in range(world):
        inv_r, lse_r, share, cos_r = ops.ntxent_lse_slab(z, Bg, 0.7, (r * b, b))
        for sl in (slice(r * b, (r + 1) * b), slice(Bg + r * b, Bg + (r + 1) * b)):
            lse[sl], inv[sl] = lse_r[sl], inv_r[sl]
        cos[r * b:(r + 1) * b] = cos_r[r * b:(r + 1) * b]
        shares.append(share)
    loss = torch.stack(shares).sum()
    assert_close(loss, loss_full.reshape(()), 2e-6, "loss from the ranks' shares")
    assert_close(cos, cos_full, 1e-6, "cosines")
    dz = torch.zeros_like(z)
    for r in range(world):
        dz_r = ops.ntxent_grad_slab(z, Bg, 0.7, (r * b, b), inv, lse)
        rows = torch.cat([torch.arange(r * b, (r + 1) * b), torch.arange(Bg + r * b, Bg + (r + 1) * b)]).to(DEV)
        other = torch.ones(2 * Bg, dtype=torch.bool, device=DEV); other[rows] = False
        assert float(dz_r[other].abs().max()) == 0.0
        dz[rows] = dz_r[rows]
    assert_close(dz, dz_full, 2e-6, "gradient rows", floor=1e-9)
    z64 = z.double().cpu().requires_grad_(True)
    ref = O.nt_xent(z64[:Bg], z64[Bg:], 0.7)
    ref.backward()
    assert_close(loss, ref.float(), FP32_OUT, "loss vs oracle")
    assert_close(dz, z64.grad.float(), FP32_GRAD, "gradient vs oracle", floor=1e-9)


def test_gather_slot_order_never_changes_the_result():
    """murcl_perm_cycle_order lists the slots cycle by cycle (partner after partner); murcl_pack_gather_ordered walks them in
    that order.  The packed batch is bit-identical to the plain-order gather (datasets.py:263-308), for cycle orders, for
    arbitrary orders and for a malformed permutation."""
    from murcl_b200.csr import BagStore, perm_cycle_order
    g = synth.gen(31)
    sizes = [int(v) for v in torch.randint(40, 900, (24,), generator=g)]
    feats, clusters, _ = synth.make_bags(sizes, 64, 5, seed=32)
    store = BagStore.from_cluster_lists(feats, clusters, DEV)
    S = 48
    slot_bag = torch.arange(24, dtype=torch.int32, device=DEV).repeat(2)
    act = torch.rand(S, 5, generator=g).to(DEV)
    lam = (0.9 + 0.1 * torch.rand(S, generator=g)).to(DEV)
    perm = torch.cat([torch.randperm(24, generator=g), torch.randperm(24, generator=g) + 24]).to(DEV, torch.int32)
    sel, _ = store.select(act, 128, slot_bag)
    os.environ["MURCL_GATHER_ORDER"] = "0"
    try:
        plain = store.gather(sel, lam, perm, torch.float32)
    finally:
        os.environ.pop("MURCL_GATHER_ORDER")
    order = perm_cycle_order(perm)
    o = order.cpu().tolist()
    p = perm.cpu().tolist()
    assert sorted(o) == list(range(S))
    follows = sum(1 for a, b in zip(o[:-1], o[1:]) if p[a] == b)
    cycles = sum(1 for a, b in zip(o[:-1], o[1:]) if p[a] != b) + 1
    assert follows + cycles == S                       # every step either follows the permutation or opens a new cycle
    for od in (order, None, torch.randperm(S, generator=g).to(DEV, torch.int32)):
        assert torch.equal(store.gather(sel, lam, perm, torch.float32, order=od), plain)
    assert torch.equal(store.gather(sel, lam, perm, torch.bfloat16, order=order), plain.to(torch.bfloat16))
    bad = perm.clone(); bad[3] = bad[4]; bad[7] = 1000                      # not a bijection, one entry out of range
    ob = perm_cycle_order(bad)
    assert sorted(ob.cpu().tolist()) == list(range(S))
    batch = perm_cycle_order(torch.stack([perm, perm.flip(0).contiguous() * 0 + perm]))
    assert torch.equal(batch[0], order) and torch.equal(batch[1], order)


def test_batched_step_draws_are_the_reference_draws_in_one_go(golden, monkeypatch):
    """``rng="batched"``: every draw of the step issued up front (``draw_step_batched``).  The distributions are the
    reference's (datasets.py:266-267: lam in [alpha, 1), one permutation per view inside its own half) and the step
    computed from them is bit-identical to the same numbers injected patch-step by patch-step."""
    from murcl_b200 import pretrain
    from murcl_b200.csr import BagStore
    from murcl_b200.dropin import abmil, cl, losses, rlmil
    g = golden("pretrain_step")
    b, k, d, fs, T, L, D, hid, proj = g["cfg"].tolist()
    torch.manual_seed(5)
    act, lam, perm, order = pretrain.draw_step_batched(T, b, k, 0.9, DEV, True)
    assert order.shape == perm.shape and order.dtype == torch.int32
    assert act.shape == (T, 2 * b, k) and lam.shape == (T, 2 * b) and perm.shape == (T, 2 * b) and perm.dtype == torch.int32
    assert float(lam.min()) >= 0.9 and float(lam.max()) < 1.0 and float(act.min()) >= 0.0 and float(act.max()) < 1.0
    ident = torch.arange(b, device=DEV, dtype=torch.int32)
    for t in range(T):
        assert torch.equal(perm[t, :b].sort().values, ident) and torch.equal(perm[t, b:].sort().values, ident + b)
    assert pretrain.draw_step_batched(T, b, k, 0.9, DEV, False)[0].shape == (1, 2 * b, k)
    feats, clusters, _ = synth.make_bags(g["sizes"].tolist(), d, k, seed=91)
    store = BagStore.from_cluster_lists(feats, clusters, DEV)
    crit = losses.NT_Xent(b, float(g["tau"]))
    results = []
    for mode in ("batched", "injected"):
        enc = _load(abmil.ABMIL(d, L=L, D=D, dim_out=proj, precision="fp32"), synth.abmil_state(d, L, D, proj, seed=92))
        model = cl.CL(enc, projection_dim=proj, n_features=L)
        fc = _load(rlmil.Full_layer(L, hid, True, proj), synth.full_layer_state(L, hid, proj, seed=93))
        if mode == "batched":
            monkeypatch.setattr(pretrain, "draw_step_batched", lambda *a, **kw: (act, lam, perm, order))
            loss, _ = pretrain.pretrain_step(store, model, fc, crit, T=T, feat_size=fs, alpha=0.9, precision="fp32", rng="batched")
        else:
            draws = [([act[t, :b], act[t, b:]], [lam[t, :b].view(b, 1), lam[t, b:].view(b, 1)],
                      [perm[t, :b].long(), perm[t, b:].long() - b]) for t in range(T)]
            loss, _ = pretrain.pretrain_step(store, model, fc, crit, T=T, feat_size=fs, alpha=0.9, precision="fp32", draws=draws)
        results.append((loss.clone(), _grads(enc), _grads(fc)))
    assert torch.equal(results[0][0], results[1][0])
    for ga, gb in ((results[0][1], results[1][1]), (results[0][2], results[1][2])):
        for key in ga:
            assert_close(ga[key], gb[key], 1e-6, f"grad {key} batched vs injected draws", floor=1e-6)
    with pytest.raises(ValueError):
        pretrain.pretrain_step(store, model, fc, crit, T=T, feat_size=fs, rng="nope")


# ------------------------------------------------------------------------------------------------
# resident slide cache, slot_bag addressing, row-sharded bags
# ------------------------------------------------------------------------------------------------
def test_resident_slides_slot_bag_matches_per_batch_store():
    """csr.ResidentSlides: slides uploaded on first touch into one arena; a step addresses its slides through slot_bag.
    The packed batch must be bit-identical to packing a store built from just those slides (get_feats, datasets.py:274-308)."""
    from murcl_b200.csr import BagStore, HostBags, ResidentSlides
    sizes = [300, 40, 1000, 65, 7, 512]
    k, d, fs = 5, 16, 64
    feats, clusters, labels = synth.make_bags(sizes, d, k, seed=201)
    res = ResidentSlides(HostBags(feats, labels, k, pin=True), DEV)
    pick = [4, 1, 2]
    moved = res.ensure(pick)
    assert moved == sum(sizes[i] for i in pick) * (d * 4 + 8) + len(pick) * k * 4
    assert res.ensure(pick) == 0 and res.ensure([2, 5]) == sizes[5] * (d * 4 + 8) + k * 4
    g = synth.gen(202)
    act = torch.rand(2 * len(pick), k, generator=g).to(DEV)
    lam = (0.9 + 0.1 * torch.rand(2 * len(pick), generator=g)).to(DEV)
    perm = torch.cat([torch.randperm(3, generator=g), torch.randperm(3, generator=g) + 3]).to(DEV)
    slot_bag = torch.tensor(pick, dtype=torch.int32, device=DEV).repeat(2)
    got = res.store.pack(act, fs, lam, perm, torch.float32, slot_bag)
    small = BagStore.from_cluster_lists([feats[i] for i in pick], [clusters[i] for i in pick], DEV)
    want = small.pack(act, fs, lam, perm, torch.float32, torch.arange(3, dtype=torch.int32, device=DEV).repeat(2))
    assert torch.equal(got, want)
    for v in range(2):        # and against the oracle's list slicing + mixup of each view
        x, _ = O.get_feats([feats[i] for i in pick], [clusters[i] for i in pick], act[3 * v:3 * v + 3].cpu(), fs)
        x = O.mixup_apply(x, lam[3 * v:3 * v + 3].cpu().reshape(-1, 1), (perm[3 * v:3 * v + 3] - 3 * v).cpu())
        assert torch.equal(got[3 * v:3 * v + 3].cpu(), x)


@pytest.mark.parametrize("kind", ["clam", "abmil"])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_row_sharded_mode_single_rank_is_identity(kind, precision):
    """shard_bags() without a process group: the merge of pooling partials is the identity, so values and gradients must
    equal the un-sharded model's (the 2-rank merge is covered on CPU by tests/test_dist_gloo.py and on GPUs by
    tools/check_sharded_model.py)."""
    from murcl_b200.dropin import abmil, clam

    def build():
        if kind == "clam":
            return _load(clam.CLAM_SB(gate=True, size_arg="small", in_dim=64, precision=precision),
                         synth.clam_state(64, "small", True, False, 2, seed=301, peak=3.0)).eval()
        return _load(abmil.ABMIL(64, precision=precision), synth.abmil_state(64, 512, 128, 2, seed=302, peak=2.0)).eval()

    feats, _, _ = synth.make_bags([700, 33, 1500], 64, 3, seed=303)
    cot = torch.randn(3, 512, generator=synth.gen(304)).to(DEV)
    a, b = build(), build().shard_bags(True)
    outs = []
    for m in (a, b):
        out, _ = m([f.to(DEV) for f in feats])
        (out * cot).sum().backward()
        outs.append(out)
    tol = 1e-6 if precision == "fp32" else 1e-3
    assert_close(outs[1], outs[0], tol, "sharded vs whole")
    assert_close(b.last_attention, a.last_attention, tol, "attention")
    ga, gb = _grads(a), _grads(b)
    for n in ga:
        if not n.startswith(("fc.", "classifiers", "instance_classifiers")):
            assert_close(gb[n], ga[n], 50 * tol, n, floor=1e-1 if _zero_grad_key(n) else 1e-6)   # atomics: summation order


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_pretrain_step_golden_with_param_arena(golden, precision):
    """The same reference-generated miniature step with the parameters laid out in a ParamArena: gradients are summed
    inside the kernels into the flat buffer (no autograd accumulation), bf16 weights come from the shadow copy; a second
    step after ``zero_grad`` must reproduce the first (nothing leaks between steps), and ``refresh`` must track an update."""
    from murcl_b200 import pretrain
    from murcl_b200.arena import ParamArena
    from murcl_b200.csr import BagStore
    from murcl_b200.dropin import abmil, cl, losses, rlmil
    g = golden("pretrain_step")
    cfg = g["cfg"].tolist()
    b, k, d, fs, T, L, D, hid, proj = cfg
    feats, clusters, _ = synth.make_bags(g["sizes"].tolist(), d, k, seed=91)
    enc = _load(abmil.ABMIL(d, L=L, D=D, dim_out=proj, precision=precision), synth.abmil_state(d, L, D, proj, seed=92))
    model = cl.CL(enc, projection_dim=proj, n_features=L)
    fc = _load(rlmil.Full_layer(L, hid, True, proj), synth.full_layer_state(L, hid, proj, seed=93))
    fc.precision = precision
    arena = ParamArena(list(enc.parameters()) + list(fc.parameters()),
                       shadow_dtype=torch.bfloat16 if precision == "bf16" else None)
    assert all(p.grad is not None and p._murcl_accum is p.grad for p in enc.parameters())
    crit = losses.NT_Xent(b, float(g["tau"]))
    store = BagStore.from_cluster_lists(feats, clusters, DEV)
    draws = [([a.to(DEV) for a in acts], [l.to(DEV) for l in lams], [p.to(DEV) for p in perms])
             for acts, lams, perms in _mini_step_draws(cfg, 94)]
    tol_l, tol_g = (FP32_OUT, 3e-4) if precision == "fp32" else (BF16_OUT, 1e-1)
    snaps = []
    for rep in range(2):
        arena.zero_grad()
        loss, _ = pretrain.pretrain_step(store, model, fc, crit, T=T, feat_size=fs, alpha=float(g["alpha"]), draws=draws,
                                         precision=precision)
        assert_close(loss, g["loss"], tol_l, "loss")
        gm, gf = _grads(enc), _grads(fc)
        n = 0
        for key, v in g.items():
            if key.startswith("grad.m.") and not key.endswith("fc.weight") and not key.endswith("fc.bias"):
                assert_close(sample(gm[key[7:]].numpy()), v, tol_g, key, floor=1e-5 if precision == "fp32" else 1e-3)
                n += 1
            elif key.startswith("grad.f."):
                assert_close(sample(gf[key[7:]].numpy()), v, tol_g, key, floor=1e-5 if precision == "fp32" else 1e-3)
                n += 1
        assert n >= 10
        snaps.append(arena.grad.clone())
    assert_close(snaps[1], snaps[0], 1e-5 if precision == "fp32" else 1e-3, "second step after zero_grad", floor=1e-6)
    # the optimiser sees one flat leaf; the bf16 shadow follows after refresh()
    opt = torch.optim.SGD(arena.optimizer_params(), lr=0.1)
    before = enc.encoder[0].weight.detach().clone()
    opt.step()
    arena.refresh()
    assert not torch.equal(before, enc.encoder[0].weight.detach())
    if precision == "bf16":
        assert torch.equal(enc.encoder[0].weight._murcl_shadow, enc.encoder[0].weight.detach().bfloat16())


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_step_side_stream_and_graph_replay_match_the_plain_step(golden, precision):
    """The scheduling of the step must not change its numbers: (a) projection head / loss chain and the tape's weight-gradient
    GEMMs on the side stream vs everything on one stream; (b) the whole optimiser step (zero_grad, forward, backward, fused
    Adam) captured by ``GraphedStep`` - fork / join edges included - and replayed vs launched eagerly.  Differences allowed:
    the summation order of the atomically accumulated gradients."""
    from murcl_b200 import pretrain
    from murcl_b200.arena import ParamArena
    from murcl_b200.csr import BagStore
    from murcl_b200.dropin import abmil, cl, losses, rlmil
    from murcl_b200.optim import ArenaAdam
    g = golden("pretrain_step")
    b, k, d, fs, T, L, D, hid, proj = g["cfg"].tolist()
    feats, clusters, _ = synth.make_bags(g["sizes"].tolist(), d, k, seed=91)
    store = BagStore.from_cluster_lists(feats, clusters, DEV)
    crit = losses.NT_Xent(b, float(g["tau"]))
    draws = [([a.to(DEV) for a in acts], [l.to(DEV) for l in lams], [p.to(DEV) for p in perms])
             for acts, lams, perms in _mini_step_draws(g["cfg"].tolist(), 94)]

    def build():
        enc = _load(abmil.ABMIL(d, L=L, D=D, dim_out=proj, precision=precision), synth.abmil_state(d, L, D, proj, seed=92))
        model = cl.CL(enc, projection_dim=proj, n_features=L)
        fc = _load(rlmil.Full_layer(L, hid, True, proj), synth.full_layer_state(L, hid, proj, seed=93))
        fc.precision = precision
        arena = ParamArena(list(enc.parameters()) + list(fc.parameters()), shadow_dtype=torch.bfloat16 if precision == "bf16" else None)
        return model, fc, arena, ArenaAdam(arena, lr=1e-3, weight_decay=1e-5)

    def step(model, fc, arena, opt, overlap):
        arena.zero_grad()
        loss, _ = pretrain.pretrain_step(store, model, fc, crit, T=T, feat_size=fs, alpha=float(g["alpha"]), draws=draws,
                                         precision=precision, overlap_heads=overlap)
        opt.step()
        return loss

    # strict part - ONE step from identical weights: only the summation order of the atomically accumulated gradients differs
    tol = 5e-6 if precision == "fp32" else 5e-3
    first = {}
    for name, overlap in (("one_stream", False), ("side_stream", True)):
        model, fc, arena, opt = build()
        arena.zero_grad()
        loss, _ = pretrain.pretrain_step(store, model, fc, crit, T=T, feat_size=fs, alpha=float(g["alpha"]), draws=draws,
                                         precision=precision, overlap_heads=overlap)
        first[name] = (loss.clone(), arena.grad.clone())
    assert_close(first["one_stream"][0], g["loss"], FP32_OUT if precision == "fp32" else BF16_OUT, "loss vs the reference fixture")
    assert_close(first["side_stream"][0], first["one_stream"][0], tol, "loss: side stream vs one stream")
    assert_close(first["side_stream"][1], first["one_stream"][1], 4 * tol, "gradients: side stream vs one stream", floor=1e-7)
    # loose part - three optimiser steps, eagerly and as a replayed graph: Adam (update = m / sqrt(v)) amplifies the summation
    # noise of near-cancelling gradient entries, so the trajectories are compared at the precision mode's own tolerance
    loose = FP32_GRAD if precision == "fp32" else BF16_OUT
    model, fc, arena, opt = build()
    eager = torch.stack([step(model, fc, arena, opt, True).clone() for _ in range(3)])
    model, fc, arena, opt = build()
    graphed = pretrain.GraphedStep(lambda: step(model, fc, arena, opt, True), warmup=1)      # the warm-up step is step 1
    replayed = torch.stack([graphed().clone() for _ in range(2)])
    torch.cuda.synchronize()
    assert opt.steps == 3                                          # device-side step counter: warm-up + two replays
    assert float(eager[2]) != float(eager[0])                      # the optimiser moved the weights: the later losses differ
    assert_close(replayed, eager[1:], loose, "losses of steps 2, 3: graph replay vs eager")


@pytest.mark.parametrize("B,d,b0,nb", [(128, 128, 0, 128), (128, 128, 32, 16), (1024, 128, 896, 128), (24, 32, 5, 7), (8, 256, 0, 3)])
def test_ntxent_gradient_slab(B, d, b0, nb):
    """murcl_ntxent_fwd_bwd_slab: the loss covers the whole batch, the gradient only the samples [b0, b0+nb) of both views
    (what a data-parallel rank keeps); compared with the fp64 evaluation of utils/losses.py:24-41."""
    from murcl_b200 import ops
    g = synth.gen(B + d + b0)
    zi = torch.randn(B, d, generator=g, dtype=torch.float64)
    zj = 0.7 * zi + torch.randn(B, d, generator=g, dtype=torch.float64)
    zi.requires_grad_(True); zj.requires_grad_(True)
    want = O.nt_xent(zi, zj, 0.5)
    want.backward()
    z = torch.cat([zi.detach(), zj.detach()], 0).float().to(DEV)
    loss, dz, cos = ops.ntxent_raw(z, B, 0.5, True, slab=(b0, nb))
    assert_close(loss, want.float().reshape(1), FP32_OUT, "loss")
    assert_close(cos, torch.cosine_similarity(zi.detach(), zj.detach()).float(), FP32_OUT, "cos")
    full = torch.cat([zi.grad, zj.grad], 0).float()
    rows = torch.cat([torch.arange(b0, b0 + nb), torch.arange(B + b0, B + b0 + nb)])
    assert_close(dz.cpu()[rows], full[rows], FP32_GRAD, "slab rows", floor=float(full.abs().max()) * 1e-3)
    mask = torch.ones(2 * B, dtype=torch.bool)
    mask[rows] = False
    assert float(dz.cpu()[mask].abs().max() if mask.any() else 0.0) == 0.0, "rows outside the slab must stay zero"


@pytest.mark.parametrize("k", [1, 8, 12])
def test_seg_topk_ends_matches_torch_topk(k):
    """murcl_seg_topk_ends (single pass for k <= 8, multi-pass beyond) against torch.topk per bag (clam.py:107-110), on
    ragged bags incl. one with exactly k rows, a 100k-row bag and heavy ties (identical zero-pad rows give identical
    weights; the kernel breaks ties by lowest index, and the VALUES at the chosen indices must equal topk's)."""
    from murcl_b200 import ops
    g = synth.gen(400 + k)
    sizes = [k, 37, 5000, 100000, 1024, 300]
    parts = [torch.softmax(3 * torch.randn(n, generator=g), 0) for n in sizes]
    parts[4][200:] = parts[4][200]                      # 824 tied weights (zero-pad rows)
    parts[5][:] = 1.0 / 300                             # everything tied
    p = torch.cat(parts).to(DEV)
    off = torch.tensor([0] + list(torch.tensor(sizes).cumsum(0)), dtype=torch.int64, device=DEV)
    top, bot = ops.seg_topk_ends(p, off, len(sizes), k)
    lo = 0
    for b, n in enumerate(sizes):
        pb = parts[b]
        tv, ti = torch.topk(pb, k)
        bv, bi = torch.topk(-pb, k)
        got_t, got_b = top[b].cpu().long() - lo, bot[b].cpu().long() - lo
        assert got_t.min() >= 0 and got_t.max() < n and len(set(got_t.tolist())) == k
        assert got_b.min() >= 0 and got_b.max() < n and len(set(got_b.tolist())) == k
        assert torch.equal(pb[got_t], tv), f"bag {b}: top-{k} values"
        assert torch.equal(-pb[got_b], bv), f"bag {b}: bottom-{k} values"
        if b < 4:                                       # no ties: the indices themselves are determined
            assert torch.equal(got_t, ti) and torch.equal(got_b, bi)
        lo += n


def test_actor_conv_golden(golden):
    """ActorCritic(policy_conv=True) (rlmil.py:30-37): the 1x1-convolution state encoder on the GEMM kernels, against the
    reference's own act / evaluate run incl. all parameter gradients."""
    from murcl_b200.dropin import rlmil
    g = golden("actor_conv")
    fdim, r, hid, k, b = g["dims"].tolist()
    ppo = rlmil.PPO(fdim, fdim * r * r, hid, True, action_std=float(g["std"]), action_size=k)
    sd = {n[3:]: torch.from_numpy(v) for n, v in g.items() if n.startswith("sd.")}
    ppo.policy.load_state_dict(sd, strict=True)
    ppo.policy_old.load_state_dict(sd, strict=True)
    mem = rlmil.Memory()
    for t in range(2):
        state = torch.from_numpy(g[f"state{t}"]).to(DEV)
        action = ppo.policy_old.act(state, mem, restart_batch=(t == 0), training=True, eps=torch.from_numpy(g[f"eps{t}"]).to(DEV))
        assert_close(action, g[f"action{t}"], FP32_OUT, f"action{t}")
        assert_close(mem.logprobs[-1], g[f"logprob{t}"], FP32_OUT, f"logprob{t}")
    lp, val, _ = ppo.policy.evaluate(torch.stack(mem.states, 0), torch.stack(mem.actions, 0))
    assert_close(lp, g["eval_logprob"], FP32_OUT, "evaluate.logprob")
    assert_close(val, g["eval_value"], FP32_OUT, "evaluate.value")
    ppo.policy.zero_grad()
    ((lp * torch.from_numpy(g["cot_l"]).to(DEV)).sum() + (val * torch.from_numpy(g["cot_v"]).to(DEV)).sum()).backward()
    gr = _grads(ppo.policy)
    for key, v in g.items():
        if key.startswith("grad."):
            assert_close(gr[key[5:]].numpy(), v, FP32_GRAD, key, floor=1e-6)
