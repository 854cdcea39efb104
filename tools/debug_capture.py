#!/usr/bin/env python
"""Find the call that invalidates a CUDA-graph capture of the pre-training step: checks the capture status after every
phase of one captured step."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from murcl_b200 import pretrain, ops  # noqa: E402
from murcl_b200.csr import ResidentSlides  # noqa: E402

sys.argv = ["bench.py", "--dataset-slides", "256"]
a = bench.parse()
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
host = bench.make_host_batch(a, seed=1000, n_slides=a.dataset_slides, pin=False)
slides = ResidentSlides(host, dev)
slides.ensure(range(a.dataset_slides))
job = bench.Job(a, 0, 1, dev)
slot_bag = torch.arange(a.bags, dtype=torch.int32, device=dev).repeat(2)


LOG = []


def status(tag):
    try:
        s = torch.cuda.is_current_stream_capturing()
        if s:
            LOG.append(tag)
    except Exception as e:  # noqa: BLE001
        print(f"  [{tag}] STATUS ERROR: {type(e).__name__}: {str(e).splitlines()[0]}", flush=True)
        raise


def wrap(obj, name, tag):
    fn = getattr(obj, name)

    def inner(*args, **kw):
        out = fn(*args, **kw)
        status(tag)
        return out

    setattr(obj, name, inner)


wrap(pretrain, "pack_views", "pack_views")
wrap(pretrain, "encode_views", "encode_views")
wrap(job.fc, "forward_views", "fc")
wrap(job.arena, "zero_grad", "arena.zero_grad")
wrap(job.arena, "refresh", "arena.refresh")
wrap(job.opt, "step", "opt.step")
wrap(job.ppo, "select_action_views", "actor")
orig_bwd = torch.Tensor.backward


def bwd(self, *args, **kw):
    status("before backward")
    r = orig_bwd(self, *args, **kw)
    status("after backward")
    return r


torch.Tensor.backward = bwd
for name in ("linear_bwd_weight", "linear_bwd_input", "attnpool_bwd_", "attnpool_fwd", "linear_fwd", "cast", "relu_bwd",
             "ntxent_raw", "actor_head", "cast_into", "row_segments"):
    wrap(ops, name, name)

side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(2):
        job.step(slides.store, slot_bag)
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
print("warm-up done; capturing", flush=True)
g = torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(g, stream=side):
        job.step(slides.store, slot_bag)
    print("capture OK")
except Exception as e:  # noqa: BLE001
    import traceback
    print("capture failed:", type(e).__name__, str(e).splitlines()[0])
    print("last ops before the failure:", LOG[-12:], "of", len(LOG))
    print("".join(traceback.format_exc().splitlines(keepends=True)[-30:]))
os._exit(0)
