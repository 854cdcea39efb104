"""Minimal torchrun probe: does NCCL come up on this box, and does the bench's collective pattern run?"""
import os, sys, time
import torch, torch.distributed as dist
t0 = time.time()
def log(msg):
    sys.stderr.write(f"[probe r{os.environ.get('RANK')} +{time.time()-t0:5.1f}s] {msg}\n"); sys.stderr.flush()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
log("init_process_group ...")
dist.init_process_group("nccl", device_id=dev)
log("init done; all_reduce ...")
x = torch.ones(4, device=dev) * (rank + 1)
dist.all_reduce(x); torch.cuda.synchronize()
log(f"all_reduce ok {x.tolist()}")
y = torch.empty(world * 2, 3, device=dev)
dist.all_gather_into_tensor(y, torch.full((2, 3), float(rank), device=dev)); torch.cuda.synchronize()
log(f"all_gather ok {y[:, 0].tolist()}")
dist.barrier(); torch.cuda.synchronize()
log("barrier ok")
dist.destroy_process_group()
log("done")
