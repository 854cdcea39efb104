#!/usr/bin/env python
"""CUDA-event timing of the tcgen05 dense-layer entry points on given shapes (bf16): fwd / input gradient / weight gradient.
cuBLAS (torch.matmul, bf16, no epilogue) on the same shapes is printed beside them as a yardstick.
usage: bench_gemm_shapes.py M,N,K [M,N,K ...]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from murcl_b200 import ops  # noqa: E402

DEV = "cuda"


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


for spec in sys.argv[1:]:
    M, N, K = (int(v) for v in spec.split(","))
    g = torch.Generator().manual_seed(1)
    x = torch.randn(M, K, generator=g).bfloat16().to(DEV)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).bfloat16().to(DEV)
    b = torch.randn(N, generator=g).to(DEV)
    dy = torch.randn(M, N, generator=g).bfloat16().to(DEV)
    t_f = timeit(lambda: ops.linear_fwd(x, w, b, ops.ACT_NONE))
    t_f32 = timeit(lambda: ops.linear_fwd(x, w, b, ops.ACT_NONE, torch.float32))
    t_d = timeit(lambda: ops.linear_bwd_input(dy, w))
    dw_acc = torch.zeros(N, K, device=DEV)
    t_w = timeit(lambda: ops.linear_bwd_weight(dy, x, False, dw_into=dw_acc))
    wt = w.t().contiguous()
    c_f = timeit(lambda: torch.matmul(x, wt))                 # [M,K] @ [K,N]
    c_d = timeit(lambda: torch.matmul(dy, w))                  # [M,N] @ [N,K]
    c_w = timeit(lambda: torch.matmul(dy.t(), x))              # [N,M] @ [M,K]
    fl = 2.0 * M * N * K
    print(f"M={M} N={N} K={K}: fwd {t_f:.1f} us ({fl / t_f / 1e6:.0f} TF)  fwd->fp32 {t_f32:.1f} us  dgrad {t_d:.1f} us ({fl / t_d / 1e6:.0f} TF)  "
          f"wgrad {t_w:.1f} us ({fl / t_w / 1e6:.0f} TF) | cuBLAS fwd {c_f:.1f} dgrad {c_d:.1f} wgrad {c_w:.1f} us")
