#!/usr/bin/env python
"""Cost split of the input-gradient epilogue's fused terms on the attention projection's shape (dy [M,128] x w [128,512]) and
the encoder's (dy [M,512] x w [512,512]): plain, + ReLU bit mask, + pooling row term, + bias column sums, all."""
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


M, FS = 262144, 1024
g = torch.Generator().manual_seed(1)
for N, K in ((128, 512), (512, 512)):
    dy = torch.randn(M, N, generator=g).bfloat16().to(DEV)
    w = (torch.randn(N, K, generator=g) / N ** 0.5).bfloat16().to(DEV)
    bits = torch.randint(-2 ** 62, 2 ** 62, (K // 64, M), generator=g, dtype=torch.int64).to(DEV)
    p = torch.rand(M, generator=g).to(DEV)
    dM = torch.randn(M // FS, K, generator=g).to(DEV)
    offsets = torch.arange(0, M + 1, FS, dtype=torch.int64, device=DEV)
    row_seg = ops.row_segments(offsets, M)
    cs = torch.zeros(K, device=DEV)
    out = torch.empty(M, K, device=DEV, dtype=torch.bfloat16)
    cases = {
        "plain": dict(),
        "+mask": dict(relu_bits=bits),
        "+rowvec": dict(row_scale=p, row_vec=dM, row_seg=row_seg),
        "+colsum": dict(col_sum=cs),
        "mask+rowvec": dict(relu_bits=bits, row_scale=p, row_vec=dM, row_seg=row_seg),
        "mask+colsum": dict(relu_bits=bits, col_sum=cs),
        "all": dict(relu_bits=bits, row_scale=p, row_vec=dM, row_seg=row_seg, col_sum=cs),
    }
    line = [f"dy[{M},{N}] x w[{N},{K}]:"]
    for name, kw in cases.items():
        t = timeit(lambda: ops.linear_bwd_input(dy, w, out=out, **kw))
        line.append(f"{name} {t:.1f}")
    print("  ".join(line), "us")
