#!/usr/bin/env python
"""torchrun check of the two data-parallel NT-Xent forms (murcl_b200/dist.py) on real GPUs: every rank evaluates the
log-sum-exp of all global rows (default below 4 ranks) vs every rank reduces only its own rows + one all-gather of the per-row
statistics (default from 4 ranks on).  Same loss, same local gradient rows; prints the per-call times of both.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/check_dist_ntxent.py"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from murcl_b200 import dist as mdist  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    b, d = 128, 128
    g = torch.Generator().manual_seed(100 + rank)
    zi = torch.randn(b, d, generator=g).cuda().requires_grad_(True)
    zj = torch.randn(b, d, generator=g).cuda().requires_grad_(True)
    out = {}
    for name, min_world in (("all_rows", 99), ("own_rows", 2)):
        os.environ["MURCL_NTX_LOCAL_LSE_MIN_WORLD"] = str(min_world)
        crit = mdist.DistributedNTXent(b, 1.0)
        for p in (zi, zj):
            p.grad = None
        loss = crit(zi, zj)
        loss.backward()
        out[name] = (loss.detach().clone(), zi.grad.clone(), zj.grad.clone(), crit.last_cosine.clone())
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            crit(zi, zj)
        e1.record()
        torch.cuda.synchronize()
        out[name + "_us"] = e0.elapsed_time(e1) * 1e3 / 20
    a, o = out["all_rows"], out["own_rows"]
    err = max(float((a[i] - o[i]).abs().max()) / max(float(a[i].abs().max()), 1e-12) for i in range(4))
    print(f"[rank {rank}/{world}] loss {float(a[0]):.7f} vs {float(o[0]):.7f}; max rel diff (loss, dzi, dzj, cos) {err:.2e}; "
          f"eager forward per call: all-rows {out['all_rows_us']:.0f} us, own-rows {out['own_rows_us']:.0f} us", flush=True)
    assert err < 5e-6, err
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
