import os, sys, math, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from murcl_b200 import synth, ops
from murcl_b200.dropin import abmil
from oracle import murcl_oracle as O
DEV = "cuda"
sd = synth.abmil_state(512, 512, 128, 2, seed=31)
feats, _, _ = synth.make_bags([2000, 333, 1024], 512, 3, seed=77)
sdl = {k: v.detach().clone().double().requires_grad_(True) for k, v in sd.items()}
want = O.abmil_forward([f.double() for f in feats], sdl)
cot = torch.randn(want.shape, generator=synth.gen(78))
(want * cot.double()).sum().backward()
for mode in ("simt", "split3", "split2"):
    os.environ["MURCL_FP32_GEMM"] = mode
    m = abmil.ABMIL(512, precision="fp32"); m.load_state_dict(sd); m = m.to(DEV)
    out, _ = m([f.to(DEV) for f in feats])
    (out * cot.to(DEV)).sum().backward()
    errs = {}
    for n, p in m.named_parameters():
        if n.startswith("fc.") or p.grad is None: continue
        w = sdl[n].grad.float()
        errs[n] = float((p.grad.cpu() - w).abs().max() / w.abs().max().clamp_min(1e-12))
    print(mode, "out", float((out.cpu() - want.float()).abs().max() / want.abs().max()), {k: f"{v:.1e}" for k, v in errs.items()})
# isolated weight gradient with a wide dynamic range across rows
g = synth.gen(5)
for M in (3357, 3392, 4096):
    x = torch.relu(torch.randn(M, 512, generator=g))
    dy = torch.randn(M, 512, generator=g) * torch.exp(4 * torch.randn(M, 1, generator=g))
    ref = (dy.double().t() @ x.double())
    for mode in ("simt", "split3"):
        os.environ["MURCL_FP32_GEMM"] = mode
        dw, _ = ops.linear_bwd_weight(dy.to(DEV), x.to(DEV), False)
        print(M, mode, "wgrad rel err", float((dw.cpu().double() - ref).abs().max() / ref.abs().max()))
