"""Fused attention pooling (murcl_attnpool_fwd: tcgen05 projection + gating + scores + online softmax + weighted sum in
one pass over H) against (a) the separate projection / score / softmax / weighted-sum kernels it replaces and (b) an fp64
evaluation of the same bf16 inputs (abmil.py:36-45, clam.py:37-60,170)."""
import math
import os

import pytest
import torch

from murcl_b200 import synth
from tests.helpers import assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda"

RAGGED = [1, 5, 127, 128, 129, 300, 1000, 64, 2, 2049]
CASES = [
    # L, D, gated, inv_sqrt_n, sizes
    (512, 128, False, True, RAGGED),              # ABMIL
    (512, 256, True, False, RAGGED),              # CLAM_SB small, gated: 512 projection columns = 4 column passes
    (512, 256, False, False, [700, 3, 1500]),     # CLAM_SB small, not gated
    (256, 64, True, False, [130, 126, 1]),        # narrow rows: 4 k-blocks, one column pass
    (512, 128, False, True, [1024] * 24),         # tile-aligned bags (the pre-training layout), more tiles than SMs
]


def _inputs(L, D, gated, sizes, seed):
    g = synth.gen(seed)
    n = sum(sizes)
    nc = D * (2 if gated else 1)
    h = torch.clamp_min(0.5 * torch.randn(n, L, generator=g) + 0.2, 0).bfloat16()
    wab = (torch.randn(nc, L, generator=g) * (2.0 / math.sqrt(L))).bfloat16()
    bab = 0.3 * torch.randn(nc, generator=g)
    wc = torch.randn(D, generator=g) * (3.0 / math.sqrt(D))
    bc = torch.randn(1, generator=g)
    offsets = torch.tensor([0] + list(torch.tensor(sizes).cumsum(0)), dtype=torch.int64)
    return h, wab, bab, wc, bc, offsets


def _reference(h, wab, bab, wc, bc, offsets, D, gated, inv_sqrt_n):
    z = h.double() @ wab.double().t() + bab.double()
    g = torch.tanh(z[:, :D]) * torch.sigmoid(z[:, D:]) if gated else torch.tanh(z)
    s = g @ wc.double() + bc.double()
    p = torch.empty_like(s)
    M = []
    for b in range(len(offsets) - 1):
        lo, hi = int(offsets[b]), int(offsets[b + 1])
        pb = torch.softmax(s[lo:hi], 0)
        if inv_sqrt_n:
            pb = pb / math.sqrt(hi - lo)
        p[lo:hi] = pb
        M.append(pb @ h[lo:hi].double())
    return s, p, torch.stack(M)


@pytest.mark.parametrize("L,D,gated,inv_sqrt_n,sizes", CASES)
def test_attnpool_matches_separate_kernels_and_fp64(L, D, gated, inv_sqrt_n, sizes):
    from murcl_b200 import ops, _lib
    assert _lib.load().murcl_attnpool_supported(L, D, int(gated), _lib.BF16) == 1
    h, wab, bab, wc, bc, offsets = _inputs(L, D, gated, sizes, 11 * L + D + len(sizes))
    B = len(sizes)
    hd, wd, bd, wcd, bcd, od = (t.to(DEV) for t in (h, wab, bab, wc, bc, offsets))
    row_seg = ops.row_segments(od, hd.shape[0])
    uv, s, p, M, stats = ops.attnpool_fwd(hd, wd, bd, wcd, bcd, od, row_seg, B, D, gated, inv_sqrt_n)
    # (a) the kernels it replaces: same MMA order, same MUFU activations -> the saved activations agree to the bit
    os.environ["MURCL_GEMM"] = "tcgen05"
    try:
        uv2 = ops.linear_fwd(hd, wd, bd, ops.ACT_TANH_SIGMOID if gated else ops.ACT_TANH)
    finally:
        os.environ.pop("MURCL_GEMM", None)
    mism = (uv != uv2).float().mean().item()
    assert mism < 1e-3, f"saved activations differ from the GEMM path in {mism:.2e} of the entries"
    assert_close(uv.float(), uv2.float(), 1e-3, "uv")
    s2 = ops.attn_score_fwd(uv, wcd, bcd, D, gated)
    p2, st2 = ops.seg_softmax(s2, od, B, 1, inv_sqrt_n)
    M2 = ops.seg_wsum(p2, hd, od, B, 1)
    assert_close(s, s2, 2e-6, "scores vs separate kernel")
    assert_close(p, p2, 2e-5, "weights vs separate kernels")
    assert_close(M, M2, 2e-5, "pooled vs separate kernels")
    assert_close(stats[:, 0, 0], st2[:, 0, 0], 1e-6, "row max")
    # (b) fp64 evaluation of the same bf16 inputs (MUFU tanh + bf16 rounding of the activations: bf16-mode budget)
    s_ref, p_ref, M_ref = _reference(h, wab, bab, wc, bc, offsets, D, gated, inv_sqrt_n)
    assert_close(s.cpu(), s_ref.float(), 2e-2, "scores vs fp64")
    assert_close(M.cpu().reshape(B, L), M_ref.float(), 2e-2, "pooled vs fp64")
    for b in range(B):
        lo, hi = int(offsets[b]), int(offsets[b + 1])
        tot = p[lo:hi].double().sum().item() * (math.sqrt(hi - lo) if inv_sqrt_n else 1.0)
        assert abs(tot - 1.0) < 1e-5, f"bag {b}: weights sum to {tot}"


def test_attnpool_without_saved_activations_and_empty_bag():
    """Inference form (uv = NULL) and a bag with no rows between two others."""
    from murcl_b200 import ops
    L, D = 512, 128
    sizes = [200, 0, 333]
    h, wab, bab, wc, bc, offsets = _inputs(L, D, False, sizes, 5)
    hd, wd, bd, wcd, bcd, od = (t.to(DEV) for t in (h, wab, bab, wc, bc, offsets))
    row_seg = ops.row_segments(od, hd.shape[0])
    uv, s, p, M, stats = ops.attnpool_fwd(hd, wd, bd, wcd, bcd, od, row_seg, 3, D, False, True, save_uv=False)
    assert uv is None
    _, s1, p1, M1, _ = ops.attnpool_fwd(hd, wd, bd, wcd, bcd, od, row_seg, 3, D, False, True, save_uv=True)
    assert torch.equal(s, s1) and torch.equal(p, p1) and torch.equal(M, M1)
    assert float(M[1].abs().max()) == 0.0 and float(stats[1, 0, 1]) == 0.0
    s_ref, p_ref, M_ref = _reference(h, wab, bab, wc, bc, torch.tensor([0, 200, 200, 533]), D, False, True)
    assert_close(M.cpu()[[0, 2]].reshape(2, L), M_ref.float()[[0, 2]], 2e-2, "pooled vs fp64")


def test_attnpool_rejects_unsupported_shapes():
    from murcl_b200 import ops
    assert not ops.attnpool_supported(1024, 128, False, torch.bfloat16)      # rows wider than the resident tile
    assert not ops.attnpool_supported(512, 384, True, torch.bfloat16)        # 768 projection columns > TMEM
    assert not ops.attnpool_supported(512, 128, False, torch.float32)        # fp32 mode keeps the exact SIMT path
    assert ops.attnpool_supported(512, 256, False, torch.bfloat16)
    assert not ops.attnpool_supported(512, 256, True, torch.bfloat16)        # supported by the kernel, not preferred (512 columns)


def test_sharded_attention_pool_kernels_single_rank():
    """dist.sharded_attention_pool on the CUDA kernels (no process group: the merge is the identity; the 2-rank merge
    itself is covered on CPU by tests/test_dist_gloo.py and on 2 GPUs by tools/check_sharded_pool.py)."""
    from murcl_b200 import dist as mdist, ops
    from oracle import murcl_oracle as O
    g = synth.gen(21)
    sizes = [300, 1, 0, 77]
    L = 512
    H = [torch.randn(n, L, generator=g) for n in sizes]
    S = [2.0 * torch.randn(n, generator=g) for n in sizes]
    G = torch.randn(len(sizes), L, generator=g)
    offsets = torch.tensor([0, 300, 301, 301, 378], dtype=torch.int64, device=DEV)
    row_seg = ops.row_segments(offsets, 378)
    for inv_sqrt_n in (False, True):
        h = torch.cat(H).to(DEV).requires_grad_(True)
        s = torch.cat(S).to(DEV).requires_grad_(True)
        M, p = mdist.sharded_attention_pool(h, s, offsets, row_seg, inv_sqrt_n)
        (M * G.to(DEV)).sum().backward()
        Hf = [x.clone().requires_grad_(True) for x in H]
        Sf = [x.clone().requires_grad_(True) for x in S]
        ref = torch.stack([O.softmax_pool(Sf[b], Hf[b], sizes[b] ** -0.5 if inv_sqrt_n else 1.0)[0] if sizes[b] else torch.zeros(L)
                           for b in range(len(sizes))])
        (ref * G).sum().backward()
        assert_close(M.cpu(), ref.detach(), 1e-5, "pooled")
        assert_close(h.grad.cpu(), torch.cat([x.grad for x, n in zip(Hf, sizes) if n]), 1e-5, "dh")
        # ds = p (dM.h - dM.M): a difference of two O(sqrt(L)) dot products in fp32 (measured 1.8e-5)
        assert_close(s.grad.cpu(), torch.cat([x.grad for x, n in zip(Sf, sizes) if n]), 1e-4, "ds")


BWD_CASES = [
    # L, D, gated, inv_sqrt_n, sizes, dtype, drop_scale
    (512, 128, False, True, RAGGED, torch.bfloat16, 1.0),          # ABMIL, the pre-training shape
    (512, 128, False, True, [1024] * 24, torch.bfloat16, 1.0),     # tile-aligned bags, more chunks than one wave
    (512, 256, True, False, RAGGED, torch.bfloat16, 1.0),          # CLAM_SB small, gated
    (512, 384, True, False, [700, 3, 1500], torch.bfloat16, 1.0),  # CLAM_SB big
    (512, 256, True, False, [300, 64], torch.bfloat16, 1.0 / 0.75),  # train-mode dropout after the activations
    (512, 128, False, True, RAGGED, torch.float32, 1.0),           # exact mode
    (512, 256, True, False, [130, 126, 1], torch.float32, 1.0),
    (32, 16, False, True, [50, 77, 1], torch.float32, 1.0),        # the golden "small" shape: few active lanes
    (1024, 128, False, False, [333, 2], torch.bfloat16, 1.0),      # widest supported rows
]


@pytest.mark.parametrize("L,D,gated,inv_sqrt_n,sizes,dtype,q", BWD_CASES)
def test_attnpool_bwd_matches_separate_kernels_and_fp64(L, D, gated, inv_sqrt_n, sizes, dtype, q):
    """murcl_attnpool_bwd (one pass over H: ds, d(pre-activation) over uv, dwc, dbc, bias column sums) against
    (a) murcl_pool_bwd_scores + murcl_attn_score_bwd, which it replaces, and (b) an fp64 evaluation of SURVEY 7.3's
    formulas on the same stored operands."""
    from murcl_b200 import ops
    assert ops.attnpool_bwd_supported(L, D, gated, dtype)
    g = synth.gen(7 * L + D + len(sizes))
    n, B = sum(sizes), len(sizes)
    nc = D * (2 if gated else 1)
    h = torch.clamp_min(0.5 * torch.randn(n, L, generator=g) + 0.2, 0).to(dtype)
    u = torch.tanh(torch.randn(n, D, generator=g))
    act = torch.cat([u, torch.sigmoid(torch.randn(n, D, generator=g))], 1) if gated else u
    if q != 1.0:                                                 # inverted dropout: kept entries scaled by q, dropped = 0
        act = act * q * (torch.rand(n, nc, generator=g) > 0.25)
    act = act.to(dtype)
    wc = torch.randn(D, generator=g) * (3.0 / math.sqrt(D))
    s = 2.0 * torch.randn(n, generator=g)
    dM = torch.randn(B, L, generator=g)
    offsets = torch.tensor([0] + list(torch.tensor(sizes).cumsum(0)), dtype=torch.int64)
    p = torch.empty(n)
    M = torch.zeros(B, L)
    for b in range(B):
        lo, hi = int(offsets[b]), int(offsets[b + 1])
        p[lo:hi] = torch.softmax(s[lo:hi], 0) / (math.sqrt(hi - lo) if inv_sqrt_n else 1.0)
        M[b] = p[lo:hi] @ h[lo:hi].float()
    hd, pd, Md, dMd, wcd, od = (t.to(DEV).contiguous() for t in (h, p, M, dM, wc, offsets))
    row_seg = ops.row_segments(od, n)
    uv1, uv2 = act.to(DEV).contiguous(), act.to(DEV).contiguous()
    dwc, dbc, dpre, ds = ops.attnpool_bwd_(hd, uv1, pd, Md, dMd, wcd, od, row_seg, B, D, gated, inv_sqrt_n, q, want_ds=True)
    # (a) the two kernels it replaces
    ds2 = ops.pool_bwd_scores(pd, hd, dMd, Md.reshape(B, 1, L), od, row_seg, B, 1, inv_sqrt_n)
    dwc2, dbc2, dpre2 = ops.attn_score_bwd_(uv2, wcd, ds2, D, gated, q)
    assert_close(ds, ds2, 1e-5, "ds vs pool_bwd_scores", floor=1e-6)
    assert_close(uv1.float(), uv2.float(), 1e-2 if dtype == torch.bfloat16 else 1e-5, "d(pre-activation) vs attn_score_bwd", floor=1e-6)
    assert_close(dwc, dwc2, 1e-4, "dwc vs attn_score_bwd", floor=1e-5)
    assert_close(dpre, dpre2, 1e-3 if dtype == torch.bfloat16 else 1e-4, "bias column sums vs attn_score_bwd", floor=1e-5)
    # (b) fp64
    hh, aa = h.double(), act.double()
    seg = torch.repeat_interleave(torch.arange(B), torch.tensor(sizes))
    alpha = torch.tensor([1.0 / math.sqrt(max(x, 1)) if inv_sqrt_n else 1.0 for x in sizes], dtype=torch.float64)
    t = (dM.double()[seg] * hh).sum(1)
    K = (dM.double() * M.double()).sum(1) / alpha
    ds_ref = p.double() * (t - K[seg])
    ua = aa[:, :D] / q
    keep_u = (aa[:, :D] != 0) if q != 1.0 else torch.ones_like(ua, dtype=torch.bool)
    if gated:
        va = aa[:, D:] / q
        keep_v = (aa[:, D:] != 0) if q != 1.0 else torch.ones_like(va, dtype=torch.bool)
        du = ds_ref[:, None] * wc.double() * aa[:, D:] * q * (1 - ua * ua) * keep_u
        dv = ds_ref[:, None] * wc.double() * aa[:, :D] * q * va * (1 - va) * keep_v
        d_ref = torch.cat([du, dv], 1)
        dwc_ref = (ds_ref[:, None] * aa[:, :D] * aa[:, D:]).sum(0)
    else:
        d_ref = ds_ref[:, None] * wc.double() * q * (1 - ua * ua) * keep_u
        dwc_ref = (ds_ref[:, None] * aa).sum(0)
    tol = 1e-2 if dtype == torch.bfloat16 else 2e-5
    assert_close(ds.cpu(), ds_ref.float(), 2e-5, "ds vs fp64", floor=1e-6)
    assert_close(uv1.float().cpu(), d_ref.float(), tol, "d(pre-activation) vs fp64", floor=1e-6)
    assert_close(dwc.cpu(), dwc_ref.float(), 1e-4, "dwc vs fp64", floor=1e-5)
    assert_close(dpre.cpu(), d_ref.sum(0).float(), 2e-3 if dtype == torch.bfloat16 else 1e-4, "bias column sums vs fp64", floor=1e-4)
    assert_close(dbc.cpu(), ds_ref.sum().reshape(1).float(), 1.0, "dbc", floor=1e-3)   # analytically zero: noise only
