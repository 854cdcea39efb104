"""tcgen05 GEMM family (bf16 operands, fp32 TMEM accumulators) against an fp64 evaluation of the same bf16 inputs.
The comparison isolates the kernel: inputs are already bf16, so the only error is fp32 accumulation order and the
final rounding of the output to its storage type."""
import math
import os

import numpy as np
import pytest
import torch

from murcl_b200 import synth
from tests.helpers import assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(autouse=True)
def _force_tc():
    os.environ["MURCL_GEMM"] = "tcgen05"
    yield
    os.environ.pop("MURCL_GEMM", None)


def _mk(M, N, K, seed):
    g = synth.gen(seed)
    x = torch.randn(M, K, generator=g).bfloat16()
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).bfloat16()
    b = torch.randn(N, generator=g)
    dy = torch.randn(M, N, generator=g).bfloat16()
    return x, w, b, dy


SHAPES = [(4096, 512, 512), (1000, 128, 512), (131, 384, 1024), (2048, 256, 64), (777, 512, 200), (128, 128, 64),
          (5000, 256, 512), (8200, 512, 128), (40000, 512, 512),      # these three exercise the CTA-pair kernel + row tails
          # B-stationary variants (many tiles per CTA): 1-CTA 128-wide tiles, a ragged reduction, 3 column tiles (grid 147)
          (40000, 128, 512), (40000, 512, 200), (40000, 384, 512)]


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_tc_forward(M, N, K):
    from murcl_b200 import ops
    x, w, b, _ = _mk(M, N, K, M + N + K)
    ref = x.double() @ w.double().t() + b.double()
    xd, wd, bd = x.to(DEV), w.to(DEV), b.to(DEV)
    y32 = ops.linear_fwd(xd, wd, bd, ops.ACT_NONE, torch.float32)
    assert_close(y32, ref.float(), 3e-6, "fp32 out")
    y16 = ops.linear_fwd(xd, wd, bd, ops.ACT_RELU)
    assert y16.dtype == torch.bfloat16
    assert_close(y16.float(), torch.relu(ref).float(), 4e-3, "bf16 relu out")
    if N % 64 == 0:
        # 1-bit ReLU mask written by the forward epilogue, consumed by the input-gradient epilogue
        y2, bits = ops.linear_fwd(xd, wd, bd, ops.ACT_RELU, relu_bits=True)
        assert torch.equal(y2, y16)
        n_idx = torch.arange(N, device=DEV)
        got = ((bits[n_idx // 64] >> (n_idx % 64).unsqueeze(1)) & 1).t().bool()
        assert torch.equal(got, y16 > 0)
    if N % 2 == 0:
        yg = ops.linear_fwd(xd, wd, bd, ops.ACT_TANH_SIGMOID, torch.float32)
        want = torch.cat([torch.tanh(ref[:, : N // 2]), torch.sigmoid(ref[:, N // 2:])], 1)
        assert_close(yg, want.float(), 1e-5, "gated activation")


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_tc_input_grad(M, N, K):
    """dx[M,K] = dy[M,N] w[N,K] (+ row term, ReLU mask): B operand is MN-major."""
    from murcl_b200 import ops
    if K < 128:
        pytest.skip("output width below the tcgen05 tile")
    x, w, _, dy = _mk(M, N, K, 7 * M + N + K)
    g = synth.gen(M)
    relu_src = torch.randn(M, K, generator=g).bfloat16()
    n_seg = 3
    seg = (torch.arange(M) * n_seg // M).to(torch.int32)
    rs = torch.rand(M, generator=g)
    rv = torch.randn(n_seg, K, generator=g)
    ref = dy.double() @ w.double()
    dx = ops.linear_bwd_input(dy.to(DEV), w.to(DEV))
    assert_close(dx.float(), ref.float(), 4e-3, "plain")
    ref2 = (ref + rs.double()[:, None] * rv.double()[seg.long()]) * (relu_src.double() > 0)
    dx2 = ops.linear_bwd_input(dy.to(DEV), w.to(DEV), relu_src.to(DEV), rs.to(DEV), rv.to(DEV), seg.to(DEV))
    assert_close(dx2.float(), ref2.float(), 4e-3, "row term + mask")
    assert bool(((dx2 == 0) | (relu_src.to(DEV) > 0)).all())
    if K % 64 == 0:
        m_idx = torch.arange(M).unsqueeze(1)
        mask = relu_src > 0
        words = torch.zeros(K // 64, M, dtype=torch.int64)
        for kk in range(K):
            words[kk // 64] |= mask[:, kk].to(torch.int64) << (kk % 64)
        dx4 = ops.linear_bwd_input(dy.to(DEV), w.to(DEV), None, rs.to(DEV), rv.to(DEV), seg.to(DEV), relu_bits=words.to(DEV))
        assert torch.equal(dx4, dx2), "bit-mask path must equal the relu_src path"
    # fused column sums of the stored (bf16-rounded) result = bias gradient of the next layer
    cs = torch.zeros(K, device=DEV)
    dx3 = ops.linear_bwd_input(dy.to(DEV), w.to(DEV), relu_src.to(DEV), rs.to(DEV), rv.to(DEV), seg.to(DEV), col_sum=cs)
    assert torch.equal(dx3, dx2)
    assert_close(cs, dx3.double().sum(0).float(), 2e-5, "fused column sums")


@pytest.mark.parametrize("M,N,K", [(128, 3072, 1024), (256, 3072, 512), (131, 384, 1024), (64, 64, 128), (1000, 200, 136)])
def test_tc_input_grad_accumulate(M, N, K):
    """dx (fp32) += dy[M,N] w[N,K] with the reduction split over the GPU (vector atomics): the W_hh step of the GRU head's
    backward recurrence.  Against fp64 on the bf16 operands; twice in a row accumulates twice."""
    from murcl_b200 import ops
    _x, w, _, dy = _mk(M, N, K, 3 * M + N + K)
    base = torch.randn(M, K, generator=synth.gen(M + 1))
    out = base.to(DEV).contiguous()
    assert ops.linear_bwd_input_accum_(dy.to(DEV), w.to(DEV), out)
    ref = dy.double() @ w.double()
    assert_close(out, (base.double() + ref).float(), 5e-6, "accumulated once")
    assert ops.linear_bwd_input_accum_(dy.to(DEV), w.to(DEV), out)
    assert_close(out, (base.double() + 2 * ref).float(), 5e-6, "accumulated twice")
    # fp32 operands are not taken: the caller falls back to linear_bwd_input
    assert not ops.linear_bwd_input_accum_(dy.float().to(DEV), w.float().to(DEV), out)


@pytest.mark.parametrize("M,N,K", SHAPES + [(131072, 512, 512), (20000, 128, 512)])
def test_tc_weight_grad(M, N, K):
    """dw[N,K] = dy^T x over M rows: both operands MN-major, split-K across the SMs."""
    from murcl_b200 import ops
    if K < 128:
        pytest.skip("output width below the tcgen05 tile")
    x, _, _, dy = _mk(M, N, K, 3 * M + N + K)
    xd, dyd = x.to(DEV), dy.to(DEV)
    dw, db = ops.linear_bwd_weight(dyd, xd)
    ref = (dyd.double().t() @ xd.double()).float()
    assert_close(dw, ref, 2e-5, "dw")
    assert_close(db, dyd.double().sum(0).float(), 2e-5, "db")


def test_tc_matches_simt_on_pretrain_shapes():
    """Same bf16 operands through both back ends (SIMT accumulates in fp32 too): only rounding order differs."""
    from murcl_b200 import ops
    x, w, b, dy = _mk(8192, 512, 512, 99)
    xd, wd, bd, dyd = x.to(DEV), w.to(DEV), b.to(DEV), dy.to(DEV)
    y_tc = ops.linear_fwd(xd, wd, bd, ops.ACT_RELU, torch.float32)
    os.environ["MURCL_GEMM"] = "simt"
    y_simt = ops.linear_fwd(xd, wd, bd, ops.ACT_RELU, torch.float32)
    assert_close(y_tc, y_simt, 2e-6, "tc vs simt")


# ------------------------------------------------------------------------------------------------
# exact-fp32 dense layers on the bf16 tensor cores (split precision)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("planes,tol", [(2, 8e-6), (3, 1e-6)])
@pytest.mark.parametrize("M,N,K", [(4096, 512, 512), (3000, 128, 512), (1111, 512, 128), (2048, 256, 64)])
def test_split_precision_gemms_match_fp64(monkeypatch, planes, tol, M, N, K):
    """murcl_linear_{fwd,bwd_input,bwd_weight}_split against fp64 on the SAME fp32 operands.  Tolerance is relative to the
    largest output magnitude: two planes (three products) keep everything above 3 * 2^-18 |x||y|, three planes 2^-24."""
    from murcl_b200 import ops
    monkeypatch.setenv("MURCL_FP32_GEMM", f"split{planes}")
    g = synth.gen(M + N + K + planes)
    x = torch.clamp_min(0.5 * torch.randn(M, K, generator=g) + 0.2, 0)          # post-ReLU-like activations
    w = torch.randn(N, K, generator=g) / math.sqrt(K)
    b = 0.1 * torch.randn(N, generator=g)
    dy = torch.randn(M, N, generator=g)
    xd, wd, bd, dyd = (t.to(DEV) for t in (x, w, b, dy))
    assert ops._split_ok(M, N, K, xd, wd) == planes
    # forward + ReLU + bit mask
    y, bits = ops.linear_fwd(xd, wd, bd, ops.ACT_RELU, relu_bits=(N % 64 == 0))
    want = torch.relu(x.double() @ w.double().t() + b.double())
    assert_close(y.cpu(), want.float(), tol, "forward")
    # input gradient through the ReLU mask of a layer of width K (mask from a forward of that width)
    if K % 64 == 0 and K >= 128:
        pre = torch.randn(M, K, generator=g)
        # the layout murcl_linear_fwd writes: [K / 64][M] 64-bit words, bit (c % 64) of word [c // 64][m] = (pre[m, c] > 0)
        on = (pre > 0).numpy().astype(np.uint64).reshape(M, K // 64, 64)
        words = (on << np.arange(64, dtype=np.uint64)).sum(-1, dtype=np.uint64).T.copy()
        kb = torch.from_numpy(words.view(np.int64)).to(DEV)
        dx = ops.linear_bwd_input(dyd, wd, relu_bits=kb)
        want_dx = (dy.double() @ w.double()) * (pre > 0)
        assert_close(dx.cpu(), want_dx.float(), tol, "input gradient (masked)")
    dx = ops.linear_bwd_input(dyd, wd) if K >= 128 else None
    if dx is not None:
        assert_close(dx.cpu(), (dy.double() @ w.double()).float(), tol, "input gradient")
    # weight gradient (+ bias), plain and accumulated into an existing buffer
    if K >= 128:
        dw, db = ops.linear_bwd_weight(dyd, xd, True)
        want_dw = dy.double().t() @ x.double()
        assert_close(dw.cpu(), want_dw.float(), tol, "weight gradient")
        assert_close(db.cpu(), dy.double().sum(0).float(), 1e-5, "bias gradient")
        acc_w, acc_b = torch.ones(N, K, device=DEV), torch.ones(N, device=DEV)
        ops.linear_bwd_weight(dyd, xd, True, dw_into=acc_w, db_into=acc_b)
        assert_close(acc_w.cpu(), (want_dw + 1).float(), tol, "accumulated weight gradient")
        assert_close(acc_b.cpu(), (dy.double().sum(0) + 1).float(), 1e-5, "accumulated bias gradient")
