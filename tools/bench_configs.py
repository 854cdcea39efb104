#!/usr/bin/env python
"""Throughput of the other BASELINE.json configurations (parity-test shapes, not the headline bench line):

  cfg1  ABMIL fwd+bwd, one bag N=2000 x 512 (the reference's CPU-runnable case)
  cfg2  CLAM_SB(small) + top-k instance loss, 64 ragged bags N~U[2000,20000] x 512
  cfg4  DSMIL, one bag N=10000 x 1024
  cfg5  one bag N=100000 x 1024, CLAM_SB pooling, bf16

Each is timed with CUDA events (median of `reps`, inputs resident, L2 flushed between reps) in fp32 and bf16 mode and,
for the CPU column, through the oracle port on all host cores.  One JSON line per config.
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from murcl_b200 import synth  # noqa: E402
from murcl_b200.dropin import abmil, clam, dsmil  # noqa: E402
from oracle import murcl_oracle as O  # noqa: E402

DEV = "cuda"


def gpu_time(fn, reps=7):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    ts.sort()
    return ts[len(ts) // 2]


def graphed(step):
    """The same step replayed as ONE CUDA graph (pretrain.GraphedStep): the single-bag configurations are bound by the
    host's ~40 us per launch, not by the kernels; fixed shapes (RLMIL trains on fixed-size windows) make them capturable."""
    from murcl_b200 import pretrain
    try:
        g = pretrain.GraphedStep(step, warmup=2)
        return lambda: g()
    except Exception as e:                                   # noqa: BLE001
        sys.stderr.write(f"[bench_configs] graph capture failed: {type(e).__name__}: {e}\n")
        return None


def cpu_time(fn, reps=2):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return min(ts)


def leaf(sd):
    return {k: v.clone().requires_grad_(True) for k, v in sd.items()}


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    out = []

    # ---- cfg1 ---------------------------------------------------------------------------------
    sd = synth.abmil_state(512, 512, 128, 2, seed=1)
    feats, _, _ = synth.make_bags([2000], 512, 10, seed=2)
    row = {"config": "cfg1 ABMIL fwd+bwd, 1 bag x 2000 x 512", "unit": "bags/s"}
    for prec in ("fp32", "bf16"):
        m = abmil.ABMIL(512, precision=prec); m.load_state_dict(sd); m = m.to(DEV)
        x = feats[0].to(DEV)

        def step():
            m.zero_grad(set_to_none=False)
            m(x.unsqueeze(0))[0].sum().backward()
        row[prec] = round(1.0 / gpu_time(step), 1)
        gs = graphed(step)
        if gs is not None:
            row[prec + "_cuda_graph"] = round(1.0 / gpu_time(gs), 1)
    sdl = leaf(sd)
    row["cpu_oracle"] = round(1.0 / cpu_time(lambda: O.abmil_forward(feats, sdl).sum().backward()), 2)
    out.append(row)

    # ---- cfg2 ---------------------------------------------------------------------------------
    g = synth.gen(3)
    sizes = torch.randint(2000, 20001, (64,), generator=g).tolist()
    feats, _, _ = synth.make_bags(sizes, 512, 10, seed=4)
    labels = torch.randint(0, 2, (64,), generator=g)
    sd = synth.clam_state(512, "small", True, False, 2, seed=5)
    row = {"config": f"cfg2 CLAM_SB(small)+instance loss fwd+bwd, 64 ragged bags ({sum(sizes)} patches) x 512", "unit": "bags/s"}
    for prec in ("fp32", "bf16"):
        m = clam.CLAM_SB(gate=True, size_arg="small", k_sample=8, n_classes=2, subtyping=True, in_dim=512, precision=prec)
        m.load_state_dict(sd); m = m.to(DEV).eval()
        bags = [f.to(DEV) for f in feats]

        def step():
            m.zero_grad(set_to_none=True)
            o, _, res = m(bags, label=labels, instance_eval=True)
            (o.sum() + sum(r["instance_loss"] for r in res)).backward()
        row[prec] = round(64.0 / gpu_time(step, reps=5), 1)
    sdl = leaf(sd)

    def cpu_step():
        tot = 0.0
        for f, l in list(zip(feats, labels.tolist()))[:8]:
            mm, res = O.clam_sb_bag(f, sdl, gate=True, label=l, instance_eval=True, n_classes=2, subtyping=True)
            tot = tot + mm.sum() + res["instance_loss"]
        tot.backward()
    row["cpu_oracle"] = round(8.0 / cpu_time(cpu_step, reps=1), 2)
    row["cpu_sample"] = "first 8 of the 64 bags"
    out.append(row)

    # ---- cfg4 ---------------------------------------------------------------------------------
    sd = synth.dsmil_state(1024, 2, seed=6)
    feats, _, _ = synth.make_bags([10000], 1024, 10, seed=7)
    row = {"config": "cfg4 DSMIL fwd+bwd, 1 bag x 10000 x 1024", "unit": "bags/s"}
    for prec in ("fp32", "bf16"):
        m = dsmil.build_dsmil(1024, 2, precision=prec); m.load_state_dict(sd)
        x = feats[0].unsqueeze(0).to(DEV)

        def step():
            m.zero_grad(set_to_none=False)
            c, b, _ = m(x)
            (c.sum() + b.sum()).backward()
        row[prec] = round(1.0 / gpu_time(step), 1)
        gs = graphed(step)
        if gs is not None:
            row[prec + "_cuda_graph"] = round(1.0 / gpu_time(gs), 1)
    sdl = leaf(sd)

    def cpu_step4():
        c, b = O.dsmil_bag(feats[0], sdl)
        (c.sum() + b.sum()).backward()
    row["cpu_oracle"] = round(1.0 / cpu_time(cpu_step4), 2)
    out.append(row)

    # ---- cfg5 ---------------------------------------------------------------------------------
    sd = synth.clam_state(1024, "small", True, False, 2, seed=8, peak=3.0)
    feats, _, _ = synth.make_bags([100000], 1024, 10, seed=9)
    row = {"config": "cfg5 CLAM_SB pooling fwd+bwd, 1 bag x 100000 x 1024 (bf16)", "unit": "bags/s"}
    m = clam.CLAM_SB(gate=True, size_arg="small", in_dim=1024, precision="bf16"); m.load_state_dict(sd); m = m.to(DEV).eval()
    x = feats[0].to(DEV)

    def step5():
        m.zero_grad(set_to_none=True)
        m(x.unsqueeze(0))[0].sum().backward()
    t = gpu_time(step5, reps=5)
    row["bf16"] = round(1.0 / t, 2)
    row["bf16_patches_per_s"] = round(100000 / t)
    out.append(row)

    for r in out:
        print(json.dumps(r))


if __name__ == "__main__":
    main()
