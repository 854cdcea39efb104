"""World-size-2 gloo tests (CPU) of the data-parallel plumbing: sharded NT-Xent over all-gathered embeddings
plus a summed gradient all-reduce reproduces the single-process loss and gradient."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from murcl_b200 import dist as mdist
from oracle import murcl_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_loss(zi, zj, tau):
    return O.nt_xent(zi, zj, tau), O.pair_cosine(zi, zj).detach()


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    B, d_in, d = 6, 10, 8
    w = torch.nn.Linear(d_in, d)                      # same init on every rank (same seed)
    xi, xj = torch.randn(B, d_in), torch.randn(B, d_in)
    lo, hi = mdist.shard_range(B, rank, world)
    crit = mdist.DistributedNTXent(hi - lo, 0.5, loss_fn=_oracle_loss)
    loss = crit(w(xi[lo:hi]), w(xj[lo:hi]))
    loss.backward()
    nbytes = mdist.allreduce_grads(w.parameters())
    # single-process reference
    w2 = torch.nn.Linear(d_in, d)
    w2.load_state_dict(w.state_dict())
    ref = O.nt_xent(w2(xi), w2(xj), 0.5)
    ref.backward()
    ok = (abs(float(loss) - float(ref)) < 1e-6
          and torch.allclose(w.weight.grad, w2.weight.grad, atol=1e-6)
          and torch.allclose(w.bias.grad, w2.bias.grad, atol=1e-6)
          and nbytes == 4 * (d_in * d + d)
          and torch.allclose(crit.last_cosine, O.pair_cosine(w2(xi), w2(xj))[lo:hi], atol=1e-6))
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_sharded_ntxent_matches_single_process():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert dict(ret) == {0: True, 1: True}


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 128, 131):
        for world in (1, 2, 3, 8):
            spans = [mdist.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
