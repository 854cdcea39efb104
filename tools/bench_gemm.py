#!/usr/bin/env python
"""Per-launch timing of the dense-layer kernels on the pre-training shapes (CUDA events, L2 flushed between reps)."""
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from murcl_b200 import ops  # noqa: E402


def timeit(fn, reps=10):
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e-3


def main():
    dev = "cuda"
    dt = torch.bfloat16 if (len(sys.argv) < 2 or sys.argv[1] == "bf16") else torch.float32
    shapes = [(131072, 512, 512), (131072, 128, 512), (131072, 512, 1024), (256, 3072, 1024), (256, 512, 512), (128, 3072, 1024)]
    for M, N, K in shapes:
        x = torch.randn(M, K, device=dev).to(dt)
        w = (torch.randn(N, K, device=dev) / K ** 0.5).to(dt)
        b = torch.randn(N, device=dev)
        dy = torch.randn(M, N, device=dev).to(dt)
        seg = torch.zeros(M, dtype=torch.int32, device=dev)
        rs = torch.rand(M, device=dev)
        rv = torch.randn(1, K, device=dev)
        fl = 2.0 * M * N * K
        rows = []
        rows.append(("fwd relu", timeit(lambda: ops.linear_fwd(x, w, b, ops.ACT_RELU))))
        rows.append(("fwd tanh", timeit(lambda: ops.linear_fwd(x, w, b, ops.ACT_TANH))))
        rows.append(("fwd f32out", timeit(lambda: ops.linear_fwd(x, w, b, ops.ACT_NONE, torch.float32))))
        rows.append(("dgrad", timeit(lambda: ops.linear_bwd_input(dy, w))))
        rows.append(("dgrad+mask+row", timeit(lambda: ops.linear_bwd_input(dy, w, x, rs, rv, seg))))
        rows.append(("wgrad", timeit(lambda: ops.linear_bwd_weight(dy, x))))
        print(f"M={M} N={N} K={K} {dt}")
        for name, t in rows:
            print(f"   {name:16s} {t*1e6:9.1f} us  {fl/t/1e12:8.1f} TFLOP/s")


if __name__ == "__main__":
    main()
