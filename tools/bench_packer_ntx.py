#!/usr/bin/env python
"""(1) pack_gather with the slots walked in plain order vs along the cycles of the mixup permutation (DRAM reads of the
source rows: twice vs once), at the bench's shapes; (2) NT-Xent loss + gradient slab at the global batch sizes of 1-8 ranks."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from murcl_b200 import ops, synth  # noqa: E402
from murcl_b200.csr import BagStore, HostBags, perm_cycle_order  # noqa: E402

DEV = "cuda"


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


B, K, FS, D = 128, 10, 1024, 512
g = synth.gen(7)
sizes = synth.camelyon_sizes(2 * B, 500, 15500, seed=1000)
feats, _cl, labels = synth.make_bags(sizes, D, K, seed=1000)
host = HostBags(feats, labels, K, pin=False, dtype=torch.bfloat16)
store = BagStore.empty_like_host(host, DEV)
store.copy_from_host(host)
slot_bag = torch.randperm(2 * B, generator=g)[:B].to(torch.int32).repeat(2).to(DEV)
act = torch.rand(2 * B, K, generator=g).to(DEV)
lam = (0.9 + 0.1 * torch.rand(2 * B, generator=g)).to(DEV)
perm = torch.cat([torch.randperm(B, generator=g), torch.randperm(B, generator=g) + B]).to(DEV, torch.int32)
sel, _ = store.select(act, FS, slot_bag)
order = perm_cycle_order(perm)
ident = torch.arange(2 * B, dtype=torch.int32, device=DEV)
t_plain = timeit(lambda: store.gather(sel, lam, perm, torch.bfloat16, order=ident))
t_cycle = timeit(lambda: store.gather(sel, lam, perm, torch.bfloat16, order=order))
t_order = timeit(lambda: perm_cycle_order(perm))
print(f"pack_gather 256 slots x 1024 x 512 bf16: plain order {t_plain:.1f} us, cycle order {t_cycle:.1f} us (cycle-order kernel {t_order:.1f} us)")

for world in (1, 2, 4, 8):
    Bg = 128 * world
    z = torch.randn(2 * Bg, 128, generator=g).to(DEV)
    t_full = timeit(lambda: ops.ntxent_raw(z, Bg, 1.0, True))
    t_slab = timeit(lambda: ops.ntxent_raw(z, Bg, 1.0, True, slab=(0, 128)))
    print(f"NT-Xent global batch {Bg} (2B = {2 * Bg} rows, d = 128): full gradient {t_full:.1f} us, one rank's slab {t_slab:.1f} us")
