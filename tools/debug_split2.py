import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from murcl_b200 import synth, ops
from murcl_b200.dropin import abmil
import torch.nn.functional as F
DEV = "cuda"
sd = synth.abmil_state(512, 512, 128, 2, seed=31)
feats, _, _ = synth.make_bags([2000, 333, 1024], 512, 3, seed=77)
x = torch.cat(feats).double()
hs64 = [x]
for i in (0, 3, 6):
    hs64.append(F.relu(F.linear(hs64[-1], sd[f"encoder.{i}.weight"].double(), sd[f"encoder.{i}.bias"].double())))
for mode in ("simt", "split3"):
    os.environ["MURCL_FP32_GEMM"] = mode
    m = abmil.ABMIL(512, precision="fp32"); m.load_state_dict(sd); m = m.to(DEV)
    ops._debug_save = {}
    out, _ = m([f.to(DEV) for f in feats])
    hs = ops._debug_save["hs"]
    ops._debug_save = None
    for l in (1, 2, 3):
        got = hs[l].cpu()
        flips = ((got > 0) != (hs64[l] > 0))
        print(mode, "layer", l, "mask flips", int(flips.sum()), "max |h| at flips", float(hs64[l][flips].abs().max()) if flips.any() else 0.0,
              "fwd rel err", float((got.double() - hs64[l]).abs().max() / hs64[l].abs().max()))
