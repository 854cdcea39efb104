#!/usr/bin/env python
"""BASELINE config 5 across GPUs: ONE bag of N x 1024 rows split over the ranks (intra-bag sharding).  Every rank pools
its rows with the CUDA kernels, one all-gather of (3 + L) floats merges the partial results, and rank 0 checks the pooled
vector and its own gradient rows against the un-sharded computation of the whole bag on its GPU.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_sharded_pool.py
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from murcl_b200 import dist as mdist, ops  # noqa: E402


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    if world > 1:
        dist.init_process_group("nccl")
    dev = "cuda"
    N, L = int(os.environ.get("ROWS", 100000)), 1024
    g = torch.Generator().manual_seed(5)
    h_all = torch.randn(N, L, generator=g)
    s_all = 3.0 * torch.randn(N, generator=g)
    G = torch.randn(1, L, generator=g).to(dev)
    lo, hi = mdist.shard_range(N, rank, world)
    h = h_all[lo:hi].to(dev).requires_grad_(True)
    s = s_all[lo:hi].to(dev).requires_grad_(True)
    off = torch.tensor([0, hi - lo], dtype=torch.int64, device=dev)
    seg = ops.row_segments(off, hi - lo)
    M, p = mdist.sharded_attention_pool(h, s, off, seg, inv_sqrt_n=True)
    (M * G).sum().backward()
    torch.cuda.synchronize()
    # timing of the sharded forward (events, max over ranks)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        mdist.sharded_attention_pool(h.detach(), s.detach(), off, seg, inv_sqrt_n=True)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 10], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = True
    if rank == 0:
        hf = h_all.to(dev).requires_grad_(True)
        sf = s_all.to(dev).requires_grad_(True)
        pf = torch.softmax(sf.double(), 0) / (N ** 0.5)
        ref = (pf.unsqueeze(0) @ hf.double()).float()
        (ref * G).sum().backward()
        err_m = float((M - ref).abs().max() / ref.abs().max())
        err_h = float((h.grad - hf.grad[lo:hi]).abs().max() / hf.grad.abs().max())
        err_s = float((s.grad - sf.grad[lo:hi]).abs().max() / sf.grad.abs().max())
        ok = err_m < 1e-5 and err_h < 1e-5 and err_s < 1e-4
        print(f"sharded pooling of one {N} x {L} bag over {world} GPU(s): {float(t):.3f} ms per forward "
              f"({N / float(t) / 1e3:.1f} M rows/s); rel err M {err_m:.1e}, dh {err_h:.1e}, ds {err_s:.1e}: {'OK' if ok else 'MISMATCH'}")
    if world > 1:
        torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0 if ok else 1)


if __name__ == "__main__":
    main()
