#!/usr/bin/env python
"""Where a single-bag (launch-bound) supervised step spends its time: wall-clock per `ops.*` call with a device
synchronisation after each (so kernel time and host overhead are both counted), for cfg1 (ABMIL) and cfg4 (DSMIL)."""
import collections
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from murcl_b200 import ops, synth  # noqa: E402
from murcl_b200.dropin import abmil, dsmil  # noqa: E402

DEV = "cuda"
acc = collections.defaultdict(lambda: [0, 0.0])


def wrap(name):
    fn = getattr(ops, name)

    def inner(*a, **k):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn(*a, **k)
        torch.cuda.synchronize()
        acc[name][0] += 1
        acc[name][1] += time.perf_counter() - t0
        return out

    setattr(ops, name, inner)


for n in ("cast", "linear_fwd", "linear_bwd_input", "linear_bwd_weight", "attnpool_fwd", "attnpool_bwd_", "seg_softmax", "seg_wsum",
          "pool_bwd_scores", "pool_bwd_direct", "seg_argmax", "relu_bwd", "row_segments", "weight_as", "attn_score_fwd", "split_planes"):
    wrap(n)


def run(tag, step, reps=10):
    for _ in range(3):
        step()
    acc.clear()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        step()
    torch.cuda.synchronize()
    tot = (time.perf_counter() - t0) / reps
    print(f"== {tag}: {tot * 1e3:.2f} ms per step (with a sync around every op)")
    for k, (c, t) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
        print(f"   {k:20s} {c / reps:5.1f} calls  {t / reps * 1e3:7.3f} ms")


for prec in ("bf16", "fp32"):
    m = abmil.ABMIL(512, precision=prec); m.load_state_dict(synth.abmil_state(512, 512, 128, 2, seed=1)); m = m.to(DEV)
    x = synth.make_bags([2000], 512, 10, seed=2)[0][0].to(DEV)

    def step1():
        m.zero_grad(set_to_none=True)
        m(x.unsqueeze(0))[0].sum().backward()
    run(f"cfg1 ABMIL {prec}", step1)
    d = dsmil.build_dsmil(1024, 2, precision=prec); d.load_state_dict(synth.dsmil_state(1024, 2, seed=6))
    xd = synth.make_bags([10000], 1024, 10, seed=7)[0][0].unsqueeze(0).to(DEV)

    def step4():
        d.zero_grad(set_to_none=True)
        c, b, _ = d(xd)
        (c.sum() + b.sum()).backward()
    run(f"cfg4 DSMIL {prec}", step4)
