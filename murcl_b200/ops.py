"""Tensor-level wrappers over the C ABI (include/murcl_b200.h).

PyTorch supplies device memory, streams and autograd bookkeeping; every computation below is a
call into libmurcl_b200.so on the caller's current CUDA stream.  There is no CPU path: a CPU
tensor raises ``MurclError``.
"""
from __future__ import annotations

import os
from typing import Optional

import torch

from . import _lib
from ._lib import (ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_TANH, ACT_TANH_SIGMOID, BF16, F32, GEMM_AUTO, GEMM_SIMT,
                   GEMM_TCGEN05, MurclError, check)

_DT = {torch.float32: F32, torch.bfloat16: BF16}


def default_precision() -> str:
    """'fp32' (exact, FFMA) or 'bf16' (tcgen05, fp32 accumulate).  Env var MURCL_PRECISION."""
    p = os.environ.get("MURCL_PRECISION", "fp32").lower()
    if p not in ("fp32", "bf16"):
        raise MurclError(f"MURCL_PRECISION must be fp32 or bf16, got {p!r}")
    return p


def storage_dtype(precision: str) -> torch.dtype:
    return torch.bfloat16 if precision == "bf16" else torch.float32


def _backend() -> int:
    return {"auto": GEMM_AUTO, "simt": GEMM_SIMT, "tcgen05": GEMM_TCGEN05}[os.environ.get("MURCL_GEMM", "auto").lower()]


def _chk(t: torch.Tensor, name: str, dtype=None) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise MurclError(f"{name}: expected a CUDA tensor (libmurcl_b200 has no CPU path)")
    if dtype is not None and t.dtype != dtype:
        raise MurclError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise MurclError(f"{name}: expected a contiguous tensor")
    if t.device.index != torch.cuda.current_device():
        # launches go to the CURRENT device's stream: a tensor of another GPU would be dereferenced there
        raise MurclError(f"{name}: tensor lives on {t.device} but the current CUDA device is {torch.cuda.current_device()} "
                         f"(wrap the call in `with torch.cuda.device(t.device)`)")
    return t


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _s():
    return torch.cuda.current_stream().cuda_stream


def _dt(t: torch.Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise MurclError(f"unsupported storage dtype {t.dtype}") from None


# ------------------------------------------------------------------------------------------------
# optional per-launch timing of the dense layers (bench.py's roofline probe)
# ------------------------------------------------------------------------------------------------
_profile = None
_debug_save = None          # tests set this to a dict to receive the (post-dropout) activations of the last aggregate


def set_profile(log):
    """``log`` (a list) receives ``(family, flops, start_event, end_event)`` per dense-layer launch; None = off."""
    global _profile
    _profile = log


class _Timed:
    def __init__(self, family, flops):
        self.family, self.flops = family, flops

    def __enter__(self):
        if _profile is not None:
            self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if _profile is not None:
            self.e1.record()
            _profile.append((self.family, self.flops, self.e0, self.e1))
        return False


# ------------------------------------------------------------------------------------------------
# raw (non-differentiable) wrappers
# ------------------------------------------------------------------------------------------------
def set_row_order(descending: bool) -> bool:
    """Scheduling hint for the next dense-layer / fused-pooling launches of this thread (murcl_set_row_order): walk the row
    tiles from the last rows to the first.  Returns the previous setting.  Never changes results."""
    return bool(_lib.load().murcl_set_row_order(1 if descending else 0))


def _serpentine() -> bool:
    return os.environ.get("MURCL_SERPENTINE", "1") != "0"


def cast(src: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    _chk(src, "cast.src")
    if src.dtype == dtype:
        return src
    dst = torch.empty_like(src, dtype=dtype)
    check(_lib.load().murcl_cast(_p(src), _dt(src), _p(dst), _DT[dtype], src.numel(), _s()), "murcl_cast")
    return dst


def cast_into(src: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
    """dst[:] = src converted to dst's storage type (one launch; same element count)."""
    _chk(src, "cast_into.src"); _chk(dst, "cast_into.dst")
    if src.numel() != dst.numel():
        raise MurclError("cast_into: element counts differ")
    check(_lib.load().murcl_cast(_p(src), _dt(src), _p(dst), _dt(dst), src.numel(), _s()), "murcl_cast")
    return dst


_weight_epoch = 0


def invalidate_weight_cache() -> None:
    """Drop every cached storage-dtype weight copy.  Needed only after a mutation that autograd cannot see:
    ``param.data.mul_(...)``-style edits do not bump ``param._version``.  Optimiser steps, ``load_state_dict``,
    ``nn.init.*_`` and ``module.to(device)`` are detected without this call."""
    global _weight_epoch
    _weight_epoch += 1


def weight_as(w: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """Storage-dtype copy of a parameter.  The copy is cached ON the parameter object and keyed by
    ``(data_ptr, device, _version, dtype, shape, epoch)``: weights are cast once per optimiser update, not once per
    forward; a device move / ``param.data = ...`` (new storage) or any versioned in-place update misses the cache.
    Un-versioned ``.data`` edits need ``invalidate_weight_cache()``."""
    if w.dtype == dtype:
        d = w.detach()
        d = d if d.is_contiguous() else d.contiguous()
        try:
            d._murcl_src = w             # lets the split-precision path cache the weight's bf16 planes on the parameter
        except AttributeError:
            pass
        return d
    sh = getattr(w, "_murcl_shadow", None)          # ParamArena: one multi-tensor cast per optimiser step keeps it fresh
    if sh is not None and sh.dtype == dtype and sh.device == w.device:
        return sh
    key = (w.data_ptr(), w.device, w._version, dtype, tuple(w.shape), _weight_epoch)
    hit = getattr(w, "_murcl_cast", None)
    if hit is not None and hit[0] == key:
        return hit[1]
    out = cast(w.detach().contiguous(), dtype)
    try:
        w._murcl_cast = (key, out)
    except AttributeError:
        pass
    return out


# ---- exact-fp32 GEMMs on the bf16 tensor cores (split precision; include/murcl_b200.h) -------------------------------
def fp32_gemm_planes() -> int:
    """MURCL_FP32_GEMM = split3 (default: 3 bf16 planes, 6 products, every term down to 2^-24: holds the 1e-5 budget on
    attention weights with a 10x margin) | split2 (2 planes, 3 products: ~6e-6 of the output scale per layer, measured
    1.1e-5 on attention weights - NOT within the parity budget, a faster approximate mode) | simt (FFMA only)."""
    m = os.environ.get("MURCL_FP32_GEMM", "split3").lower()
    if m not in ("simt", "split2", "split3"):
        raise MurclError(f"MURCL_FP32_GEMM must be simt, split2 or split3, got {m!r}")
    return {"simt": 0, "split2": 2, "split3": 3}[m]


def _split_ok(M, N, K, *tensors) -> int:
    planes = fp32_gemm_planes()
    if planes == 0 or _backend() == GEMM_SIMT or any(t.dtype != torch.float32 for t in tensors):
        return 0
    return planes if _lib.load().murcl_linear_split_supported(int(M), int(N), int(K)) else 0


def split_planes(t: torch.Tensor, planes: int):
    """fp32 [rows, cols] -> (bf16 [planes, plane_rows, cols], plane_rows) with plane_rows = rows rounded up to 64 (zero rows)."""
    _chk(t, "split_planes.t", torch.float32)
    rows, cols = t.shape
    pr = (rows + 63) // 64 * 64
    buf = torch.empty((planes, pr, cols), device=t.device, dtype=torch.bfloat16)
    check(_lib.load().murcl_split_planes(_p(t), rows, cols, planes, pr, _p(buf), _s()), "murcl_split_planes")
    return buf, pr


_last_planes = [None]        # (tensor object, version, planes, (buf, pr)): the gradient a layer's dgrad and wgrad both split


def _act_planes(t: torch.Tensor, planes: int, remember: bool = False):
    hit = _last_planes[0]
    if hit is not None and hit[0] is t and hit[1] == t._version and hit[2] == planes:
        return hit[3]
    out = split_planes(t, planes)
    if remember:
        _last_planes[0] = (t, t._version, planes, out)      # holds `t` alive, so its address cannot be recycled under the key
    return out


def _weight_planes(w: torch.Tensor, planes: int):
    """Planes of a weight, cached on the PARAMETER it was detached from (``weight_as`` tags the detached view)."""
    src = getattr(w, "_murcl_src", None)
    if src is None:
        return split_planes(w, planes)
    key = (src.data_ptr(), src.device, src._version, planes, tuple(src.shape), _weight_epoch)
    hit = getattr(src, "_murcl_planes", None)
    if hit is not None and hit[0] == key:
        return hit[1]
    out = split_planes(w, planes)
    try:
        src._murcl_planes = (key, out)
    except AttributeError:
        pass
    return out


def linear_fwd(x, w, bias, act=ACT_NONE, out_dtype=None, relu_bits=False, out=None):
    """act(x w^T + bias).  ``relu_bits=True`` (ReLU layers, N % 64 == 0) also returns the 1-bit-per-output mask
    ``[N/64, M]`` int64 that ``linear_bwd_input`` takes instead of re-reading the activation.  ``out`` (contiguous
    ``[M, N]``) receives the result instead of a fresh tensor."""
    _chk(x, "linear_fwd.x"); _chk(w, "linear_fwd.w")
    if x.dtype != w.dtype:
        raise MurclError(f"linear_fwd: x is {x.dtype} but w is {w.dtype}")
    M, K = x.shape
    N = w.shape[0]
    if w.shape[1] != K:
        raise MurclError(f"linear_fwd: shape mismatch x{tuple(x.shape)} w{tuple(w.shape)}")
    if out is not None:
        _chk(out, "linear_fwd.out")
        if tuple(out.shape) != (M, N):
            raise MurclError(f"linear_fwd: out is {tuple(out.shape)}, expected {(M, N)}")
        out_dtype, y = out.dtype, out
    else:
        out_dtype = out_dtype or x.dtype
        y = torch.empty((M, N), device=x.device, dtype=out_dtype)
    if bias is not None:
        _chk(bias, "linear_fwd.bias", torch.float32)
    bits = torch.empty((N // 64, M), device=x.device, dtype=torch.int64) if relu_bits else None
    planes = _split_ok(M, N, K, x, w, y)
    if planes:
        (xp, xpr), (wp, wpr) = _act_planes(x, planes), _weight_planes(w, planes)
        with _Timed("linear_fwd", 2.0 * M * N * K):
            check(_lib.load().murcl_linear_fwd_split(_p(xp), _p(wp), _p(bias), _p(y), M, N, K, act, planes, xpr, wpr, _p(bits), _s()),
                  "murcl_linear_fwd_split")
        return (y, bits) if relu_bits else y
    with _Timed("linear_fwd" if M >= 4096 else "head_fwd", 2.0 * M * N * K):
        check(_lib.load().murcl_linear_fwd(_p(x), _p(w), _p(bias), _p(y), M, N, K, act, _dt(x), _DT[out_dtype], _backend(),
                                           _p(bits), _s()), "murcl_linear_fwd")
    return (y, bits) if relu_bits else y


def linear_bwd_input(dy, w, relu_src=None, row_scale=None, row_vec=None, row_seg=None, col_sum=None, out_scale=1.0,
                     relu_bits=None, out=None):
    _chk(dy, "linear_bwd_input.dy"); _chk(w, "linear_bwd_input.w")
    if dy.dtype != w.dtype:
        raise MurclError(f"linear_bwd_input: dy is {dy.dtype} but w is {w.dtype}")
    M, N = dy.shape
    K = w.shape[1]
    if w.shape[0] != N:
        raise MurclError(f"linear_bwd_input: shape mismatch dy{tuple(dy.shape)} w{tuple(w.shape)}")
    if out is not None:
        _chk(out, "linear_bwd_input.out", dy.dtype)
        if tuple(out.shape) != (M, K):
            raise MurclError(f"linear_bwd_input: out is {tuple(out.shape)}, expected {(M, K)}")
    dx = out if out is not None else torch.empty((M, K), device=dy.device, dtype=dy.dtype)
    if relu_src is not None:
        _chk(relu_src, "linear_bwd_input.relu_src", dy.dtype)
    planes = _split_ok(M, K, N, dy, w) if (relu_src is None or relu_bits is not None) else 0
    if planes:
        (dp, dpr), (wp, wpr) = _act_planes(dy, planes, remember=True), _weight_planes(w, planes)
        with _Timed("linear_bwd_input", 2.0 * M * N * K):
            check(_lib.load().murcl_linear_bwd_input_split(_p(dp), _p(wp), _p(dx), M, N, K, _p(row_scale), _p(row_vec), _p(row_seg),
                                                           _p(col_sum), float(out_scale), _p(relu_bits), planes, dpr, wpr, _s()),
                  "murcl_linear_bwd_input_split")
        return dx
    with _Timed("linear_bwd_input" if M >= 4096 else "head_bwd_input", 2.0 * M * N * K):
        check(_lib.load().murcl_linear_bwd_input(_p(dy), _p(w), _p(dx), M, N, K, _p(relu_src), _p(row_scale), _p(row_vec),
                                                 _p(row_seg), _p(col_sum), float(out_scale), _p(relu_bits), _dt(dy), _backend(),
                                                 _s()),
              "murcl_linear_bwd_input")
    return dx


def linear_bwd_input_accum_(dy, w, dx_f32) -> bool:
    """``dx_f32 += dy @ w`` (fp32 buffer, split-K with atomics: murcl_linear_bwd_input_accum).  Returns False - and does
    nothing - when the shape / storage type is not taken (fp32 operands, tiny layers): the caller then uses
    ``linear_bwd_input`` and adds."""
    if dy.dtype != torch.bfloat16 or w.dtype != torch.bfloat16 or _backend() == GEMM_SIMT:
        return False
    M, N = dy.shape
    K = w.shape[1]
    lib = _lib.load()
    if not lib.murcl_linear_bwd_input_accum_supported(int(M), int(N), int(K), _DT[dy.dtype]):
        return False
    _chk(dy, "linear_bwd_input_accum.dy"); _chk(w, "linear_bwd_input_accum.w"); _chk(dx_f32, "linear_bwd_input_accum.dx", torch.float32)
    if w.shape[0] != N or tuple(dx_f32.shape) != (M, K):
        raise MurclError(f"linear_bwd_input_accum: shape mismatch dy{tuple(dy.shape)} w{tuple(w.shape)} dx{tuple(dx_f32.shape)}")
    with _Timed("head_bwd_input", 2.0 * M * N * K):
        check(lib.murcl_linear_bwd_input_accum(_p(dy), _p(w), _p(dx_f32), M, N, K, _DT[dy.dtype], _s()), "murcl_linear_bwd_input_accum")
    return True


def linear_bwd_weight(dy, x, want_bias=True, dw_into=None, db_into=None):
    """(dw [N, K], db [N] | None) fp32.  ``dw_into`` / ``db_into`` (fp32, contiguous, e.g. the ``.grad`` views of a
    ``ParamArena``) make the kernels ADD their result to those buffers (no separate accumulation launch); the returned
    entry is then the buffer itself."""
    _chk(dy, "linear_bwd_weight.dy"); _chk(x, "linear_bwd_weight.x", dy.dtype)
    M, N = dy.shape
    K = x.shape[1]
    lib = _lib.load()
    acc = dw_into is not None
    if acc:
        _chk(dw_into, "linear_bwd_weight.dw_into", torch.float32)
        if dw_into.numel() != N * K:
            raise MurclError(f"linear_bwd_weight: dw_into has {dw_into.numel()} elements, expected {N * K}")
        if want_bias and db_into is None:
            raise MurclError("linear_bwd_weight: accumulating dw needs db_into as well when a bias gradient is wanted")
    dw = dw_into if acc else torch.empty((N, K), device=dy.device, dtype=torch.float32)
    db = None
    if want_bias:
        db = db_into if acc else torch.empty((N,), device=dy.device, dtype=torch.float32)
    planes = _split_ok(M, N, K, dy, x) if (K >= 128 and K % 64 == 0) else 0
    if planes:
        (dp, pr), (xp, pr2) = _act_planes(dy, planes, remember=True), _act_planes(x, planes)
        ws = torch.empty((max(int(lib.murcl_linear_bwd_weight_split_workspace(M, N, K)), 1),), device=dy.device, dtype=torch.float32)
        with _Timed("linear_bwd_weight", 2.0 * M * N * K):
            check(lib.murcl_linear_bwd_weight_split(_p(dp), _p(xp), _p(dw), M, N, K, planes, pr, _p(ws), int(acc), _s()),
                  "murcl_linear_bwd_weight_split")
        if want_bias:
            part = torch.empty((N,), device=dy.device, dtype=torch.float32)
            check(lib.murcl_colsum(_p(dy), M, N, _dt(dy), _p(part), _s()), "murcl_colsum")
            if acc:
                db.add_(part)
            else:
                db = part
        return dw, db
    nws = int(lib.murcl_linear_bwd_weight_workspace(M, N, K))
    ws = torch.empty((max(nws, 1),), device=dy.device, dtype=torch.float32)
    with _Timed("linear_bwd_weight" if M >= 4096 else "head_bwd_weight", 2.0 * M * N * K):
        check(lib.murcl_linear_bwd_weight(_p(dy), _p(x), _p(dw), _p(db), M, N, K, _dt(dy), _backend(), _p(ws), int(acc), _s()),
              "murcl_linear_bwd_weight")
    return dw, db


def grad_target(param):
    """The persistent fp32 gradient view the kernels accumulate into when ``param`` belongs to a ``ParamArena``
    (murcl_b200/arena.py), else None (the gradient is returned to autograd as usual)."""
    t = getattr(param, "_murcl_accum", None) if param is not None else None
    return t if (t is not None and t.is_cuda) else None


def relu_bwd(dy, y):
    _chk(dy, "relu_bwd.dy"); _chk(y, "relu_bwd.y", dy.dtype)
    dz = torch.empty_like(dy)
    check(_lib.load().murcl_relu_bwd(_p(dy), _p(y), _p(dz), dy.numel(), _dt(dy), _s()), "murcl_relu_bwd")
    return dz


def row_segments(offsets: torch.Tensor, n_rows: int) -> torch.Tensor:
    _chk(offsets, "row_segments.offsets", torch.int64)
    seg = torch.empty((n_rows,), device=offsets.device, dtype=torch.int32)
    check(_lib.load().murcl_row_segments(_p(offsets), offsets.numel() - 1, _p(seg), _s()), "murcl_row_segments")
    return seg


def attn_score_fwd(uv, wc, bc, D, gated):
    n = uv.shape[0]
    s = torch.empty((n,), device=uv.device, dtype=torch.float32)
    check(_lib.load().murcl_attn_score_fwd(_p(uv), _p(wc), _p(bc), _p(s), n, D, int(gated), _dt(uv), _s()),
          "murcl_attn_score_fwd")
    return s


def seg_softmax(s, offsets, B, C, inv_sqrt_n):
    _chk(s, "seg_softmax.s", torch.float32)
    p = torch.empty_like(s)
    stats = torch.empty((B, C, 2), device=s.device, dtype=torch.float32)
    check(_lib.load().murcl_seg_softmax(_p(s), _p(offsets), B, C, int(inv_sqrt_n), _p(p), _p(stats), _s()),
          "murcl_seg_softmax")
    return p, stats


def seg_wsum(p, h, offsets, B, C):
    _chk(p, "seg_wsum.p", torch.float32); _chk(h, "seg_wsum.h")
    n_rows, L = h.shape
    lib = _lib.load()
    out = torch.empty((B, C, L), device=h.device, dtype=torch.float32)
    ws = torch.empty((max(int(lib.murcl_seg_wsum_workspace(n_rows, B, C, L)), 1),), device=h.device, dtype=torch.float32)
    check(lib.murcl_seg_wsum(_p(p), _p(h), _p(offsets), n_rows, B, C, L, _dt(h), _p(out), _p(ws), _s()), "murcl_seg_wsum")
    return out


def attnpool_supported(L, D, gated, dtype) -> bool:
    """The fused attention-pooling kernel covers bf16 rows with L <= 512 and D*(1+gated) <= 512 (a multiple of 128);
    ``MURCL_DISABLE_ATTNPOOL=1`` forces the separate projection / score / softmax / weighted-sum kernels."""
    if os.environ.get("MURCL_DISABLE_ATTNPOOL", "0") == "1" or dtype != torch.bfloat16:
        return False
    # with more than 256 projection columns the kernel re-streams the row tile and the weights once per 128 columns
    # (L2-bandwidth bound) and has a single TMEM accumulator stage: the CTA-pair GEMM + separate kernels are as fast
    if D * (2 if gated else 1) > 256 and os.environ.get("MURCL_FORCE_ATTNPOOL", "0") != "1":
        return False
    return bool(_lib.load().murcl_attnpool_supported(int(L), int(D), int(gated), _lib.BF16))


def attnpool_fwd(h, wab, bab, wc, bc, offsets, row_seg, B, D, gated, inv_sqrt_n, save_uv=True):
    """Fused uv / scores / softmax / pooled bag vectors in one pass over ``h`` (murcl_attnpool_fwd).
    Returns (uv or None, s [n], p [n], M [B, 1, L], stats [B, 1, 2])."""
    _chk(h, "attnpool_fwd.h", torch.bfloat16); _chk(wab, "attnpool_fwd.wab", torch.bfloat16)
    n_rows, L = h.shape
    nc = D * (2 if gated else 1)
    if tuple(wab.shape) != (nc, L):
        raise MurclError(f"attnpool_fwd: projection weight {tuple(wab.shape)} != ({nc}, {L})")
    lib = _lib.load()
    dev = h.device
    uv = torch.empty((n_rows, nc), device=dev, dtype=torch.bfloat16) if save_uv else None
    s = torch.empty((n_rows,), device=dev, dtype=torch.float32)
    p = torch.empty((n_rows,), device=dev, dtype=torch.float32)
    M = torch.empty((B, 1, L), device=dev, dtype=torch.float32)
    stats = torch.empty((B, 1, 2), device=dev, dtype=torch.float32)
    ws = torch.empty((max(int(lib.murcl_attnpool_workspace(n_rows, B, L)), 1),), device=dev, dtype=torch.float32)
    # profile record: "flops" slot carries the algorithmic HBM bytes of this HBM-bound kernel (h once, uv, s, p)
    with _Timed("attnpool_fwd", float(n_rows) * (L * 2 + (nc * 2 if save_uv else 0) + 8)):
        check(lib.murcl_attnpool_fwd(_p(h), _p(wab), _p(bab), _p(wc), _p(bc), _p(offsets), _p(row_seg), n_rows, B, L, D,
                                     int(gated), int(inv_sqrt_n), _lib.BF16, _p(uv) if uv is not None else None, _p(s), _p(p),
                                     _p(M), _p(stats), _p(ws), _s()), "murcl_attnpool_fwd")
    return uv, s, p, M, stats


def pool_bwd_scores(p, h, dM, M, offsets, row_seg, B, C, inv_sqrt_n):
    n_rows, L = h.shape
    _chk(dM, "pool_bwd_scores.dM", torch.float32); _chk(M, "pool_bwd_scores.M", torch.float32)
    ds = torch.empty((n_rows, C) if C > 1 else (n_rows,), device=h.device, dtype=torch.float32)
    kbuf = torch.empty((B * C,), device=h.device, dtype=torch.float32)
    check(_lib.load().murcl_pool_bwd_scores(_p(p), _p(h), _p(dM), _p(M), _p(offsets), _p(row_seg), n_rows, B, C, L,
                                            int(inv_sqrt_n), _dt(h), _p(ds), _p(kbuf), _s()), "murcl_pool_bwd_scores")
    return ds


def pool_bwd_direct(p, dM, row_seg, C, L, out, accumulate):
    n_rows = out.shape[0]
    check(_lib.load().murcl_pool_bwd_direct(_p(p), _p(dM), _p(row_seg), n_rows, C, L, _dt(out), _p(out), int(accumulate),
                                            _s()), "murcl_pool_bwd_direct")
    return out


def dropout_(x: torch.Tensor, p: float, seed: torch.Tensor) -> torch.Tensor:
    """In-place inverted dropout; ``seed`` is a 1-element int64 CUDA tensor (device-side, so graph replays differ)."""
    _chk(x, "dropout.x"); _chk(seed, "dropout.seed", torch.int64)
    check(_lib.load().murcl_dropout(_p(x), x.numel(), float(p), _p(seed), _dt(x), _s()), "murcl_dropout")
    return x


def attn_score_bwd_(uv, wc, ds, D, gated, drop_scale=1.0):
    """In place: uv becomes the gradient w.r.t. the pre-activations.  Returns (dwc [D], dbc [1], column sums of the
    new uv [D | 2D] = bias gradient of the attention projection)."""
    buf = torch.zeros((D + 1 + uv.shape[1],), device=uv.device, dtype=torch.float32)      # one memset for all three
    dwc, dbc, dpre = buf[:D], buf[D:D + 1], buf[D + 1:]
    if dpre.data_ptr() % 16:
        dpre = torch.zeros((uv.shape[1],), device=uv.device, dtype=torch.float32)
    check(_lib.load().murcl_attn_score_bwd(_p(uv), _p(wc), _p(ds), _p(dwc), _p(dbc), _p(dpre), uv.shape[0], D, int(gated),
                                           float(drop_scale), _dt(uv), _s()), "murcl_attn_score_bwd")
    return dwc, dbc, dpre


def attnpool_bwd_supported(L, D, gated, dtype) -> bool:
    if os.environ.get("MURCL_DISABLE_ATTNPOOL_BWD", "0") == "1" or dtype not in _DT:
        return False
    return bool(_lib.load().murcl_attnpool_bwd_supported(int(L), int(D), int(gated), _DT[dtype]))


def attnpool_bwd_(h, uv, p, M, dM, wc, offsets, row_seg, B, D, gated, inv_sqrt_n, drop_scale=1.0, want_ds=False,
                  into=None):
    """Fused pooling backward (murcl_attnpool_bwd): one pass over ``h``; ``uv`` becomes the gradient w.r.t. the
    pre-activations IN PLACE.  Returns (dwc [D], dbc [1], column sums of the new uv, ds or None).  ``into`` =
    (dwc, dbc, dpre) fp32 buffers the kernel's atomics ADD to (persistent gradient views); default: fresh zeros."""
    _chk(h, "attnpool_bwd.h"); _chk(uv, "attnpool_bwd.uv", h.dtype)
    _chk(p, "attnpool_bwd.p", torch.float32); _chk(M, "attnpool_bwd.M", torch.float32); _chk(dM, "attnpool_bwd.dM", torch.float32)
    n_rows, L = h.shape
    if into is not None:
        dwc, dbc, dpre = into
    else:
        buf = torch.zeros((D + 4 + uv.shape[1],), device=uv.device, dtype=torch.float32)      # one memset for all three
        dwc, dbc, dpre = buf[:D], buf[D:D + 1], buf[D + 4:]
    ds = torch.empty((n_rows,), device=h.device, dtype=torch.float32) if want_ds else None
    # profile record: "flops" slot carries the algorithmic HBM bytes (h once, uv read + rewritten, p)
    es = h.element_size()
    with _Timed("attnpool_bwd", float(n_rows) * (L * es + 2 * uv.shape[1] * es + 4)):
        check(_lib.load().murcl_attnpool_bwd(_p(h), _p(uv), _p(p), _p(M), _p(dM), _p(wc), _p(offsets), _p(row_seg), n_rows, B, L,
                                             D, int(gated), int(inv_sqrt_n), float(drop_scale), _dt(h), _p(ds), _p(dwc), _p(dbc),
                                             _p(dpre), _s()), "murcl_attnpool_bwd")
    return dwc, dbc, dpre, ds


def seg_topk_ends(p, offsets, B, k):
    top = torch.empty((B, k), device=p.device, dtype=torch.int32)
    bot = torch.empty((B, k), device=p.device, dtype=torch.int32)
    check(_lib.load().murcl_seg_topk_ends(_p(p), _p(offsets), B, k, _p(top), _p(bot), _s()), "murcl_seg_topk_ends")
    return top, bot


def seg_argmax(c, offsets, B, C):
    idx = torch.empty((B, C), device=c.device, dtype=torch.int32)
    check(_lib.load().murcl_seg_argmax(_p(c), _p(offsets), B, C, _p(idx), _s()), "murcl_seg_argmax")
    return idx


def gather_rows(h, idx):
    _chk(idx, "gather_rows.idx", torch.int32)
    out = torch.empty((idx.numel(), h.shape[1]), device=h.device, dtype=torch.float32)
    check(_lib.load().murcl_gather_rows(_p(h), _p(idx), idx.numel(), h.shape[1], _dt(h), _p(out), _s()), "murcl_gather_rows")
    return out


def scatter_add_rows_(dh, idx, rows):
    _chk(rows, "scatter_add_rows.rows", torch.float32)
    check(_lib.load().murcl_scatter_add_rows(_p(dh), _p(idx), idx.numel(), dh.shape[1], _dt(dh), _p(rows), _s()),
          "murcl_scatter_add_rows")
    return dh


# ------------------------------------------------------------------------------------------------
# autograd: small dense heads (fp32 storage; decoder, GRU gates, actor MLP, classifier heads)
# ------------------------------------------------------------------------------------------------
class _Linear(torch.autograd.Function):
    """y = act(x w^T + b).  Operands are stored in ``dtype`` for the GEMMs (fp32 or bf16); y is fp32."""

    @staticmethod
    def forward(ctx, x, w, b, act, dtype):
        xs = cast(x.detach().contiguous(), dtype)
        ws = weight_as(w, dtype)
        bs = None if b is None else b.detach().contiguous().float()
        y = linear_fwd(xs, ws, bs, act, torch.float32)
        ctx.act = act
        ctx.has_bias = b is not None
        ctx.dtype = dtype
        ctx.x_dtype = x.dtype
        ctx.gw, ctx.gb = grad_target(w), grad_target(b)
        ctx.save_for_backward(xs, ws, y)
        return y

    @staticmethod
    def backward(ctx, dy):
        xs, ws, y = ctx.saved_tensors
        dy = dy.contiguous().float()
        if ctx.act == ACT_RELU:
            dz = relu_bwd(dy, y)
        elif ctx.act == ACT_TANH:
            dz = dy * (1 - y * y)
        elif ctx.act == ACT_SIGMOID:
            dz = dy * y * (1 - y)
        else:
            dz = dy
        dz = cast(dz.contiguous(), ctx.dtype)
        dx = linear_bwd_input(dz, ws).to(ctx.x_dtype) if ctx.needs_input_grad[0] else None
        dw = db = None
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            if ctx.gw is not None and (not ctx.has_bias or ctx.gb is not None):
                linear_bwd_weight(dz, xs, want_bias=ctx.has_bias, dw_into=ctx.gw, db_into=ctx.gb)     # summed in place
            else:
                dw, db = linear_bwd_weight(dz, xs, want_bias=ctx.has_bias)
        return dx, dw, db, None, None


def linear(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor] = None, act: int = ACT_NONE,
           dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """act(x @ w.T + b) on libmurcl_b200 kernels, differentiable; x [M,K] CUDA, result fp32."""
    if not x.is_cuda:
        raise MurclError("linear: expected a CUDA tensor (libmurcl_b200 has no CPU path)")
    return _Linear.apply(x, w, b, act, dtype)


class _ScaleGrad(torch.autograd.Function):
    """Identity in the forward pass; the gradient is multiplied by ``factor`` on the way back."""

    @staticmethod
    def forward(ctx, t, factor):
        ctx.factor = factor
        return t.view_as(t)

    @staticmethod
    def backward(ctx, g):
        return g * ctx.factor, None


def scale_grad(t: torch.Tensor, factor: float) -> torch.Tensor:
    return _ScaleGrad.apply(t, float(factor))


class _GRUCell(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gi, gh, h_prev):
        gi, gh, h_prev = gi.contiguous(), gh.contiguous(), h_prev.contiguous()
        B, H = h_prev.shape
        h_new = torch.empty_like(h_prev)
        gates = torch.empty_like(gi)
        check(_lib.load().murcl_gru_cell_fwd(_p(gi), _p(gh), _p(h_prev), _p(h_new), _p(gates), B, H, _s()),
              "murcl_gru_cell_fwd")
        ctx.save_for_backward(gates, gh, h_prev)
        return h_new

    @staticmethod
    def backward(ctx, dh_new):
        gates, gh, h_prev = ctx.saved_tensors
        B, H = h_prev.shape
        dh_new = dh_new.contiguous()
        dgi, dgh, dh_prev = torch.empty_like(gates), torch.empty_like(gates), torch.empty_like(h_prev)
        check(_lib.load().murcl_gru_cell_bwd(_p(dh_new), _p(gates), _p(gh), _p(h_prev), _p(dgi), _p(dgh), _p(dh_prev), B, H,
                                             _s()), "murcl_gru_cell_bwd")
        return dgi, dgh, dh_prev


def gru_cell(gi, gh, h_prev):
    """Fused GRU cell on precomputed gate pre-activations (differentiable)."""
    return _GRUCell.apply(gi, gh, h_prev)


def gru_step(x, h_prev, w_ih, w_hh, b_ih, b_hh, dtype: torch.dtype = torch.float32):
    """One nn.GRU time step (gate order r,z,n) built from two dense layers and the fused cell kernel."""
    gi = linear(x, w_ih, b_ih, ACT_NONE, dtype)
    gh = linear(h_prev, w_hh, b_hh, ACT_NONE, dtype)
    return _GRUCell.apply(gi, gh, h_prev)


def actor_head(logits, eps, std):
    logits, eps = _chk(logits.contiguous(), "actor_head.logits", torch.float32), _chk(eps.contiguous(), "actor_head.eps", torch.float32)
    B, K = logits.shape
    action, mean = torch.empty_like(logits), torch.empty_like(logits)
    logprob = torch.empty((B,), device=logits.device, dtype=torch.float32)
    check(_lib.load().murcl_actor_head(_p(logits), _p(eps), float(std), _p(action), _p(logprob), _p(mean), B, K, _s()),
          "murcl_actor_head")
    return action, logprob, mean


# ------------------------------------------------------------------------------------------------
# autograd: NT-Xent
# ------------------------------------------------------------------------------------------------
def ntxent_raw(z: torch.Tensor, B: int, temperature: float, want_grad=True, slab=None):
    """(loss [1], dz [2B, d] | None, cos [B]).  ``slab = (b0, nb)`` restricts the gradient to the samples
    [b0, b0 + nb) of both views (murcl_ntxent_fwd_bwd_slab); the other rows of dz are zero."""
    _chk(z, "ntxent.z", torch.float32)
    R, d = z.shape
    loss = torch.empty((1,), device=z.device, dtype=torch.float32)
    cos = torch.empty((B,), device=z.device, dtype=torch.float32)
    ws = torch.empty((int(_lib.load().murcl_ntxent_workspace(B, d)),), device=z.device, dtype=torch.float32)
    if slab is None or tuple(slab) == (0, B):
        dz = torch.empty_like(z) if want_grad else None
        check(_lib.load().murcl_ntxent_fwd_bwd(_p(z), B, d, float(temperature), _p(loss), _p(dz), _p(cos), _p(ws), _s()),
              "murcl_ntxent_fwd_bwd")
    else:
        b0, nb = int(slab[0]), int(slab[1])
        dz = torch.zeros_like(z) if want_grad else None
        check(_lib.load().murcl_ntxent_fwd_bwd_slab(_p(z), B, d, float(temperature), b0, nb, _p(loss), _p(dz), _p(cos), _p(ws),
                                                    _s()), "murcl_ntxent_fwd_bwd_slab")
    return loss, dz, cos


def ntxent_lse_slab(z: torch.Tensor, B: int, temperature: float, slab):
    """Log-sum-exp pass over the rows of the samples ``slab = (b0, nb)`` of both views only (murcl_ntxent_lse_slab):
    ``(inv_norm [2B], lse [2B], loss_share [1], cos [B])`` - only the slab's entries of the three vectors are defined."""
    _chk(z, "ntxent.z", torch.float32)
    R, d = z.shape
    b0, nb = int(slab[0]), int(slab[1])
    inv_norm = torch.empty((R,), device=z.device, dtype=torch.float32)
    lse = torch.empty((R,), device=z.device, dtype=torch.float32)
    share = torch.empty((1,), device=z.device, dtype=torch.float32)
    cos = torch.empty((B,), device=z.device, dtype=torch.float32)
    ws = torch.empty((int(_lib.load().murcl_ntxent_slab_workspace(B, d, nb)),), device=z.device, dtype=torch.float32)
    check(_lib.load().murcl_ntxent_lse_slab(_p(z), B, d, float(temperature), b0, nb, _p(inv_norm), _p(lse), _p(share), _p(cos),
                                            _p(ws), _s()), "murcl_ntxent_lse_slab")
    return inv_norm, lse, share, cos


def ntxent_grad_slab(z: torch.Tensor, B: int, temperature: float, slab, inv_norm: torch.Tensor, lse: torch.Tensor) -> torch.Tensor:
    """Gradient rows of the samples ``slab`` of both views from the complete ``inv_norm`` / ``lse`` (murcl_ntxent_grad_slab);
    the other rows of the returned ``[2B, d]`` tensor are zero."""
    _chk(z, "ntxent.z", torch.float32); _chk(inv_norm, "ntxent.inv_norm", torch.float32); _chk(lse, "ntxent.lse", torch.float32)
    R, d = z.shape
    b0, nb = int(slab[0]), int(slab[1])
    if inv_norm.numel() != R or lse.numel() != R:
        raise MurclError("ntxent_grad_slab: inv_norm / lse must have one entry per row of z")
    dz = torch.zeros_like(z)
    ws = torch.empty((int(_lib.load().murcl_ntxent_slab_workspace(B, d, nb)),), device=z.device, dtype=torch.float32)
    check(_lib.load().murcl_ntxent_grad_slab(_p(z), B, d, float(temperature), b0, nb, _p(inv_norm), _p(lse), _p(dz), _p(ws), _s()),
          "murcl_ntxent_grad_slab")
    return dz


class _NTXent(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z_i, z_j, temperature, slab):
        B = z_i.shape[0]
        z = torch.cat([z_i.detach(), z_j.detach()], 0).float().contiguous()
        loss, dz, cos = ntxent_raw(z, B, temperature, True, slab)
        ctx.save_for_backward(dz)
        ctx.B = B
        ctx.mark_non_differentiable(cos)
        return loss.reshape(()), cos

    @staticmethod
    def backward(ctx, g, _gcos):
        (dz,) = ctx.saved_tensors
        dz = dz * g
        return dz[: ctx.B], dz[ctx.B:], None, None


def ntxent(z_i, z_j, temperature, slab=None):
    """(loss, cos_pair): fused NT-Xent forward+gradient; cos_pair[b] = cos(z_i[b], z_j[b]) (the reward signal).
    ``slab = (b0, nb)``: only the samples [b0, b0 + nb) receive a gradient (what a data-parallel rank needs for the rows it
    contributed to an all-gathered batch; the other rows' gradients are returned as zeros)."""
    if not z_i.is_cuda:
        raise MurclError("ntxent: expected CUDA tensors (libmurcl_b200 has no CPU path)")
    return _NTXent.apply(z_i, z_j, temperature, slab)


# ------------------------------------------------------------------------------------------------
# autograd: the MIL aggregator (encoder -> attention -> segmented softmax pooling)
# ------------------------------------------------------------------------------------------------
class _MILAggregate(torch.autograd.Function):
    """x [n_rows, D_in] (CSR rows of all bags) -> M [B, L].

    args: x, offsets, row_seg, meta, wab, bab, wc, bc, inst_w, inst_b, *enc (w0, b0, w1, b1, ...)
    meta: dict(B, gated, inv_sqrt_n, dtype, inst=None | dict(groups...))
    Also returns p [n_rows] (attention weights incl. post-scale), raw scores s, inst loss per group.
    """

    @staticmethod
    def forward(ctx, x, offsets, row_seg, meta, wab, bab, wc, bc, inst_w, inst_b, *enc):
        dt = meta["dtype"]
        B, gated = meta["B"], meta["gated"]
        D = wc.numel()
        hs = [cast(x.detach().contiguous(), dt)]
        hbits = [None]                          # per activation: 1-bit ReLU mask (or None -> the backward reads h itself)
        enc_w = []
        drop = meta.get("drop")                 # train-mode dropout: {"enc": [p after layer 1..n], "attn": p}
        seeds = None
        if drop is not None:
            seeds = torch.randint(0, 2 ** 62, (len(enc) // 2 + 1,), device=x.device, dtype=torch.int64)
        # serpentine row order: an activation (268 MB at the pre-training shape) is larger than L2, so a consumer that starts
        # at row 0 finds the rows its producer wrote FIRST - evicted by then.  Alternate layers walk the rows in opposite
        # directions: every layer starts with the rows the previous one wrote last, which are still in L2.
        serp = _serpentine()
        for i in range(0, len(enc), 2):
            w = weight_as(enc[i], dt)
            enc_w.append(w)
            dropped = drop is not None and drop["enc"][i // 2] > 0
            prev_order = set_row_order(serp and (i // 2) % 2 == 1)
            try:
                if w.shape[0] % 64 == 0 and not dropped:
                    h, hb = linear_fwd(hs[-1], w, enc[i + 1].detach().contiguous(), ACT_RELU, relu_bits=True)
                else:
                    h, hb = linear_fwd(hs[-1], w, enc[i + 1].detach().contiguous(), ACT_RELU), None
            finally:
                set_row_order(prev_order)
            if dropped:
                dropout_(h, drop["enc"][i // 2], seeds[i // 2:i // 2 + 1])   # mask = zeros of h (inactive or dropped)
            hs.append(h)
            hbits.append(hb)
        H = hs[-1]
        wab_s = weight_as(wab, dt)
        wc_f = wc.detach().reshape(-1).contiguous().float()
        bc_f = bc.detach().reshape(-1).contiguous().float()
        attn_drop = drop is not None and drop["attn"] > 0
        if not attn_drop and attnpool_supported(H.shape[1], D, gated, H.dtype):
            # one pass over H: projection (tcgen05) + gating + score + online softmax + weighted sum
            prev_order = set_row_order(serp and (len(enc) // 2) % 2 == 1)
            try:
                uv, s, p, M, _ = attnpool_fwd(H, wab_s, bab.detach().contiguous().float(), wc_f, bc_f, offsets, row_seg, B, D,
                                              gated, meta["inv_sqrt_n"] and not meta.get("shard"))
            finally:
                set_row_order(prev_order)
            M = M.reshape(B, -1)
        else:
            uv = linear_fwd(H, wab_s, bab.detach().contiguous(), ACT_TANH_SIGMOID if gated else ACT_TANH)
            if attn_drop:
                dropout_(uv, drop["attn"], seeds[-1:])
            s = attn_score_fwd(uv, wc_f, bc_f, D, gated)
            p, _ = seg_softmax(s, offsets, B, 1, meta["inv_sqrt_n"] and not meta.get("shard"))
            M = seg_wsum(p, H, offsets, B, 1).reshape(B, -1)
        shard = bool(meta.get("shard"))
        inv_sqrt_bwd = meta["inv_sqrt_n"]
        if shard:
            # intra-bag sharding (BASELINE config 5): the rows of every bag are split across the ranks of
            # meta["shard_group"]; p / M above are normalised over the LOCAL rows only.  One all-gather of (m, l, n, M_loc)
            # per bag merges the partial softmax sums; afterwards p holds the GLOBAL weights of the local rows and M the
            # global pooled vectors (identical on every rank).  The backward needs no collective.
            M, p, inv_sqrt_bwd = _merge_shards(meta, offsets, row_seg, p, M, H, s)
        if _debug_save is not None:
            _debug_save.update(hs=[h.clone() for h in hs], uv=uv.clone())
        inst = meta.get("inst")
        if inst is not None and shard:
            raise MurclError("the instance-clustering loss ranks the instances of a whole bag: not available with row sharding")
        inst_loss = s.new_zeros((0,))
        preds = s.new_zeros((0,), dtype=torch.int32)
        saved_inst = ()
        if inst is not None:
            k = inst["k"]
            top, bot = seg_topk_ends(p, offsets, B, k)
            # per-group row index lists; layout decided on the host (tiny), data stays on device
            both = torch.cat([top, bot], 1)                                   # [B, 2k]
            idx_parts = [(both if in_cls else top)[b] for (b, _c, in_cls) in inst["groups"]]
            idx = torch.cat(idx_parts).contiguous()
            rows = gather_rows(H, idx)
            G = len(inst["groups"])
            inst_loss = torch.empty((G,), device=x.device, dtype=torch.float32)
            preds = torch.empty((idx.numel(),), device=x.device, dtype=torch.int32)
            dlogits = torch.empty((idx.numel(), 2), device=x.device, dtype=torch.float32)
            iw = inst_w.detach().contiguous().float()
            ib = inst_b.detach().contiguous().float()
            check(_lib.load().murcl_clam_inst_ce_fwd(_p(rows), _p(inst["targets"]), _p(inst["group_off"]), _p(inst["group_cls"]),
                                                     G, _p(iw), _p(ib), rows.shape[1], _p(inst_loss), _p(preds), _p(dlogits),
                                                     _s()), "murcl_clam_inst_ce_fwd")
            saved_inst = (idx, rows, dlogits, iw)
        ctx.meta = meta
        ctx.n_enc = len(enc) // 2
        ctx.D = D
        ctx.consumed = False
        ctx.x_dtype = x.dtype
        ctx.hbits = hbits
        ctx.inv_sqrt_bwd = inv_sqrt_bwd
        # persistent gradient views (ParamArena): the backward kernels add into them and autograd gets None
        ctx.g_attn = tuple(grad_target(t) for t in (wab, bab, wc, bc))
        ctx.g_enc = tuple(grad_target(t) for t in enc)
        M_out = M
        if shard:
            M, M_out = M          # (merged vectors without the post scale: what the backward's K_b uses, pooled output)
        ctx.save_for_backward(offsets, row_seg, wab_s, wc_f, uv, p, M, *hs, *enc_w, *saved_inst)
        ctx.mark_non_differentiable(p, s, preds)
        return M_out, p, s, inst_loss, preds

    @staticmethod
    def backward(ctx, dM, _dp, _ds, dinst, _dpreds):
        if ctx.consumed:
            raise MurclError("MIL aggregate: backward called twice (saved activations are consumed in place)")
        ctx.consumed = True
        meta = ctx.meta
        B, gated, D, n_enc = meta["B"], meta["gated"], ctx.D, ctx.n_enc
        sv = ctx.saved_tensors
        offsets, row_seg, wab_s, wc_f, uv, p, M = sv[:7]
        hs = sv[7:7 + n_enc + 1]
        enc_w = sv[7 + n_enc + 1:7 + 2 * n_enc + 1]
        rest = sv[7 + 2 * n_enc + 1:]
        H = hs[-1]
        L = H.shape[1]
        dM = dM.contiguous().float()
        drop = meta.get("drop")
        q_attn = 1.0 / (1.0 - drop["attn"]) if drop is not None and drop["attn"] > 0 else 1.0
        q_enc = [1.0 / (1.0 - pe) if drop is not None and pe > 0 else 1.0 for pe in (drop["enc"] if drop else [0.0] * n_enc)]
        g_wab, g_bab, g_wc, g_bc = ctx.g_attn
        g_enc = ctx.g_enc
        inst = meta.get("inst")
        attn_in_place = all(t is not None for t in (g_wab, g_bab, g_wc, g_bc))
        # every encoder parameter has a persistent gradient view and no instance-loss scatter touches dZ afterwards
        enc_in_place = n_enc > 0 and inst is None and all(t is not None for t in g_enc)
        if attnpool_bwd_supported(L, D, gated, H.dtype):
            # one pass over H: t_n = dM.h_n, ds, d(pre-activation) over uv, dwc / bias column sums
            into = (g_wc.reshape(-1), g_bc.reshape(-1), g_bab.reshape(-1)) if attn_in_place else None
            dwc, dbc, dbab, _ = attnpool_bwd_(H, uv, p, M.reshape(B, L).contiguous(), dM, wc_f, offsets, row_seg, B, D, gated,
                                              ctx.inv_sqrt_bwd, q_attn, into=into)
        else:
            ds = pool_bwd_scores(p, H, dM, M.reshape(B, 1, L), offsets, row_seg, B, 1, ctx.inv_sqrt_bwd)
            dwc, dbc, dbab = attn_score_bwd_(uv, wc_f, ds, D, gated, q_attn)    # uv now holds d(pre-activation)
            if attn_in_place:
                g_wc.reshape(-1).add_(dwc); g_bc.reshape(-1).add_(dbc); g_bab.add_(dbab)
        if attn_in_place:
            linear_bwd_weight(uv, H, want_bias=False, dw_into=g_wab)
            dwab = dbab = dwc = dbc = None
        else:
            dwab, _ = linear_bwd_weight(uv, H, want_bias=False)
        hbits = ctx.hbits
        relu_src = H if (n_enc > 0 and hbits[n_enc] is None) else None
        # bias gradients ride along as fused column sums of each dZ (the instance-loss scatter below changes dZ
        # after the fact, so that case takes the separate column-sum pass)
        fuse_db = n_enc > 0 and inst is None
        if enc_in_place:
            db_next = [g_enc[2 * l + 1] for l in range(n_enc)]            # atomics add straight into the bias gradients
        else:
            db_next = torch.zeros((n_enc, L), device=uv.device, dtype=torch.float32) if fuse_db else None
        dz = linear_bwd_input(uv, wab_s, relu_src, p, dM, row_seg, col_sum=db_next[n_enc - 1] if fuse_db else None,
                              out_scale=q_enc[n_enc - 1] if n_enc > 0 else 1.0,
                              relu_bits=hbits[n_enc] if n_enc > 0 else None)       # + p_n dM[b] direct term, ReLU mask
        d_inst_w = d_inst_b = None
        if inst is not None:
            idx, rows, dlogits, iw = rest
            G = len(inst["groups"])
            drows = torch.empty_like(rows)
            d_inst_w = torch.zeros_like(iw)
            d_inst_b = torch.zeros((iw.shape[0], 2), device=iw.device, dtype=torch.float32)
            gl = dinst.contiguous().float()
            check(_lib.load().murcl_clam_inst_ce_bwd(_p(rows), _p(dlogits), _p(gl), _p(inst["group_off"]), _p(inst["group_cls"]),
                                                     G, _p(iw), L, _p(drows), _p(d_inst_w), _p(d_inst_b), _s()),
                  "murcl_clam_inst_ce_bwd")
            if n_enc > 0:
                drows = drows * (rows > 0) * q_enc[n_enc - 1]           # same ReLU (+dropout) mask as the fused epilogue
            scatter_add_rows_(dz, idx, drows.contiguous())
        # The input-gradient CHAIN first, in serpentine row order (each launch starts with the rows the previous one wrote
        # last: still in L2, the 268 MB gradient as a whole is not), then the weight gradients, the freshest gradient first.
        # MURCL_SERPENTINE=0: one direction and the layer-by-layer order (weight gradient, then input gradient).
        serp = _serpentine()
        dzs = [None] * (n_enc + 1)
        dzs[n_enc] = dz
        dx_raw = None

        def wgrad(l):
            if enc_in_place:
                linear_bwd_weight(dzs[l], hs[l - 1], want_bias=False, dw_into=g_enc[2 * (l - 1)])
                return None, None
            dw, db = linear_bwd_weight(dzs[l], hs[l - 1], want_bias=not fuse_db)
            return dw, (db_next[l - 1] if fuse_db else db)

        def dgrad(l):
            prev_order = set_row_order(serp and (n_enc - l) % 2 == 0)
            try:
                if l > 1:
                    return linear_bwd_input(dzs[l], enc_w[l - 1], hs[l - 1] if hbits[l - 1] is None else None,
                                            col_sum=db_next[l - 2] if fuse_db else None, out_scale=q_enc[l - 2],
                                            relu_bits=hbits[l - 1])
                return linear_bwd_input(dzs[1], enc_w[0]) if ctx.needs_input_grad[0] else None
            finally:
                set_row_order(prev_order)

        gw = {}
        if serp:
            for l in range(n_enc, 0, -1):
                nxt = dgrad(l)
                if l > 1:
                    dzs[l - 1] = nxt
                else:
                    dx_raw = nxt
            for l in range(1, n_enc + 1):
                gw[l] = wgrad(l)
        else:
            for l in range(n_enc, 0, -1):
                gw[l] = wgrad(l)
                nxt = dgrad(l)
                if l > 1:
                    dzs[l - 1] = nxt
                else:
                    dx_raw = nxt
        grads_enc = []
        for l in range(1, n_enc + 1):
            grads_enc += list(gw[l])
        dx = None
        if ctx.needs_input_grad[0]:
            dx = (dx_raw if n_enc > 0 else dz).to(ctx.x_dtype)
        return (dx, None, None, None, dwab, dbab, None if dwc is None else dwc.reshape(1, -1), dbc, d_inst_w, d_inst_b,
                *grads_enc)


def _merge_shards(meta, offsets, row_seg, p_loc, M_loc, H, s):
    """Local pooling partials -> global attention weights and pooled vectors (``dist.merge_pool_partials``).  Returns
    ``((M_merged, M_out), p_global, False)``: ``M_out = post * M_merged`` is the module output, ``M_merged`` what the
    backward's ``K_b = dM . M / post`` needs (with the post scale folded into ``p`` the kernel runs with inv_sqrt_n off)."""
    from . import dist as mdist
    B = meta["B"]
    # (m, l) of the local softmax: recomputed from s and p would cost a pass; both pooling paths return the local
    # normalisation, so m_b = max s, l_b = sum exp(s - m_b) are taken from one segmented softmax over the raw scores
    _, stats = seg_softmax(s, offsets, B, 1, False)
    m, l = stats[:, 0, 0].contiguous(), stats[:, 0, 1].contiguous()
    n_loc = (offsets[1:] - offsets[:-1]).to(torch.float32)
    M_merged, scale, n_tot = mdist.merge_pool_partials(m, l, M_loc.reshape(B, -1).float(), n_loc, meta.get("shard_group"))
    post = torch.rsqrt(n_tot.clamp_min(1.0)) if meta["inv_sqrt_n"] else torch.ones_like(n_tot)
    p = (p_loc * (scale * post)[row_seg.long()]).contiguous()
    return (M_merged.contiguous(), (M_merged * post.unsqueeze(1)).contiguous()), p, False


@torch.no_grad()
def mil_attention_scores(x, meta, wab, bab, wc, bc, enc_params):
    """Raw (pre-softmax) attention scores [n_rows] of every instance: CLAM's ``attention_only`` path
    (clam.py:141-142) used by the heat-map script.  Not differentiable."""
    if not x.is_cuda:
        raise MurclError("mil_attention_scores: expected CUDA tensors (libmurcl_b200 has no CPU path)")
    dt = meta["dtype"]
    h = cast(x.detach().contiguous(), dt)
    for i in range(0, len(enc_params), 2):
        h = linear_fwd(h, weight_as(enc_params[i], dt), enc_params[i + 1].detach().contiguous(), ACT_RELU)
    uv = linear_fwd(h, weight_as(wab, dt), bab.detach().contiguous(), ACT_TANH_SIGMOID if meta["gated"] else ACT_TANH)
    return attn_score_fwd(uv, wc.detach().reshape(-1).contiguous().float(), bc.detach().reshape(-1).contiguous().float(),
                          wc.numel(), meta["gated"])


def mil_aggregate(x, offsets, row_seg, meta, wab, bab, wc, bc, inst_w, inst_b, enc_params):
    if not x.is_cuda:
        raise MurclError("mil_aggregate: expected CUDA tensors (libmurcl_b200 has no CPU path)")
    return _MILAggregate.apply(x, offsets, row_seg, meta, wab, bab, wc, bc, inst_w, inst_b, *enc_params)


# ------------------------------------------------------------------------------------------------
# autograd: DSMIL critical-instance attention
# ------------------------------------------------------------------------------------------------
class _DSMILAggregate(torch.autograd.Function):
    """BClassifier (dsmil.py:64-81): x [n_rows, D], instance scores c [n_rows, C] (used for the arg-max
    only) -> bag [B, C, D] fp32.  V is applied AFTER pooling: A^T (X Wv^T + bv) = (A^T X) Wv^T + bv because
    every softmax column sums to 1, which removes 89% of the reference's FLOPs."""

    @staticmethod
    def forward(ctx, x, classes, offsets, row_seg, meta, wq, bq, wv, bv):
        dt, B = meta["dtype"], meta["B"]
        xs = cast(x.detach().contiguous(), dt)
        classes = classes.detach().contiguous().float()
        C = classes.shape[1]
        wq_s = weight_as(wq, dt)
        crit = seg_argmax(classes, offsets, B, C)
        q = linear_fwd(xs, wq_s, bq.detach().contiguous().float(), ACT_NONE, torch.float32)
        n_rows, Dq = q.shape
        a = torch.empty((n_rows, C), device=x.device, dtype=torch.float32)
        check(_lib.load().murcl_dsmil_scores_fwd(_p(q), _p(crit), _p(row_seg), n_rows, C, Dq, _p(a), _s()),
              "murcl_dsmil_scores_fwd")
        p, _ = seg_softmax(a, offsets, B, C, False)
        mx = seg_wsum(p, xs, offsets, B, C)                               # [B, C, D]
        wv_f = wv.detach().contiguous().float()
        bag = linear_fwd(mx.reshape(B * C, -1), wv_f, bv.detach().contiguous().float(), ACT_NONE).reshape(B, C, -1)
        ctx.meta = meta
        ctx.x_dtype = x.dtype
        ctx.save_for_backward(offsets, row_seg, xs, wq_s, wv_f, q, crit, p, mx)
        return bag

    @staticmethod
    def backward(ctx, dbag):
        meta = ctx.meta
        dt, B = meta["dtype"], meta["B"]
        offsets, row_seg, xs, wq_s, wv_f, q, crit, p, mx = ctx.saved_tensors
        n_rows, D = xs.shape
        C, Dq = p.shape[1], q.shape[1]
        dbag2 = dbag.contiguous().float().reshape(B * C, D)
        dmx = linear_bwd_input(dbag2, wv_f).reshape(B, C, D)
        dwv, dbv = linear_bwd_weight(dbag2, mx.reshape(B * C, D))
        da = pool_bwd_scores(p, xs, dmx, mx, offsets, row_seg, B, C, False).reshape(n_rows, C)
        dq = torch.empty_like(q)
        check(_lib.load().murcl_dsmil_scores_bwd(_p(q), _p(da), _p(crit), _p(row_seg), _p(offsets), n_rows, B, C, Dq, _p(dq),
                                                 _s()), "murcl_dsmil_scores_bwd")
        dq_s = cast(dq, dt)
        dwq, dbq = linear_bwd_weight(dq_s, xs)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = linear_bwd_input(dq_s, wq_s)
            pool_bwd_direct(p, dmx.contiguous(), row_seg, C, D, dx, True)
            dx = dx.to(ctx.x_dtype)
        return dx, None, None, None, None, dwq, dbq, dwv, dbv


def dsmil_aggregate(x, classes, offsets, row_seg, meta, wq, bq, wv, bv):
    if not x.is_cuda:
        raise MurclError("dsmil_aggregate: expected CUDA tensors (libmurcl_b200 has no CPU path)")
    return _DSMILAggregate.apply(x, classes, offsets, row_seg, meta, wq, bq, wv, bv)
