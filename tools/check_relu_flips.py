#!/usr/bin/env python
"""Evidence for DESIGN.md section 2: fp32-mode gradients against an fp64 evaluation of the reference that adopts the DEVICE's
ReLU decisions, for the FFMA GEMMs (MURCL_FP32_GEMM=simt) and the split-precision tensor-core GEMMs (split3).  With the
device's masks every gradient agrees to ~2e-6; against the reference's own masks a single unit whose pre-activation is zero to
rounding (1 of 1.7 M here, in BOTH modes) moves a layer's weight gradient by up to 3e-4.  Run on a GPU box."""
import os, sys, math, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from murcl_b200 import synth, ops
from murcl_b200.dropin import abmil
import torch.nn.functional as F
DEV = "cuda"
sd = synth.abmil_state(512, 512, 128, 2, seed=31)
sizes = [2000, 333, 1024]
feats, _, _ = synth.make_bags(sizes, 512, 3, seed=77)
cot = torch.randn(3, 512, generator=synth.gen(78))
for mode in ("simt", "split3"):
    os.environ["MURCL_FP32_GEMM"] = mode
    m = abmil.ABMIL(512, precision="fp32"); m.load_state_dict(sd); m = m.to(DEV)
    ops._debug_save = {}
    out, _ = m([f.to(DEV) for f in feats])
    hs = [h.cpu() for h in ops._debug_save["hs"]]
    ops._debug_save = None
    (out * cot.to(DEV)).sum().backward()
    # fp64 reference that uses the DEVICE's ReLU masks (so that boundary flips do not count as errors)
    P = {k: v.detach().clone().double().requires_grad_(True) for k, v in sd.items()}
    x = torch.cat(feats).double()
    h = x
    for j, i in enumerate((0, 3, 6)):
        z = F.linear(h, P[f"encoder.{i}.weight"], P[f"encoder.{i}.bias"])
        h = z * (hs[j + 1] > 0).double()
    outs, lo = [], 0
    for n in sizes:
        hb = h[lo:lo + n]; lo += n
        u = torch.tanh(F.linear(hb, P["attention.0.weight"], P["attention.0.bias"]))
        s = F.linear(u, P["attention.2.weight"], P["attention.2.bias"]).squeeze(-1)
        p = torch.softmax(s, 0) / math.sqrt(n)
        outs.append(F.relu(F.linear((p @ hb).unsqueeze(0), P["decoder.0.weight"], P["decoder.0.bias"])))
    ref = torch.cat(outs, 0)
    (ref * cot.double()).sum().backward()
    errs = {}
    for n_, p_ in m.named_parameters():
        if n_.startswith("fc.") or p_.grad is None or n_ == "attention.2.bias": continue
        w = P[n_].grad.float()
        errs[n_] = float((p_.grad.cpu() - w).abs().max() / w.abs().max().clamp_min(1e-12))
    print(mode, {k: f"{v:.1e}" for k, v in errs.items()})
