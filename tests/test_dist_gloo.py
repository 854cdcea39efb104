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


# ---- torch (CPU) versions of the two kernels of the own-rows NT-Xent form (murcl_ntxent_lse_slab / _grad_slab) ----------
def _rows_of(Bg, slab):
    b0, nb = slab
    return torch.cat([torch.arange(b0, b0 + nb), torch.arange(Bg + b0, Bg + b0 + nb)])


def _cpu_lse_slab(z, Bg, tau, slab):
    R = 2 * Bg
    inv = 1.0 / z.norm(dim=1).clamp_min(1e-8)
    zn = z * inv[:, None]
    rows = _rows_of(Bg, slab)
    s = zn[rows] @ zn.t() / tau                                   # only the slab's rows of the score matrix
    s[torch.arange(rows.numel()), rows] = -float("inf")
    lse_rows = torch.logsumexp(s, 1)
    pos = torch.where(rows < Bg, rows + Bg, rows - Bg)
    s_pos = (zn[rows] * zn[pos]).sum(1) / tau
    lse = torch.full((R,), float("nan")); inv_out = torch.full((R,), float("nan"))
    lse[rows] = lse_rows; inv_out[rows] = inv[rows]
    cos = torch.full((Bg,), float("nan"))
    cos[slab[0]:slab[0] + slab[1]] = s_pos[:slab[1]] * tau
    return inv_out, lse, ((lse_rows - s_pos).sum() / R).reshape(1), cos


def _cpu_grad_slab(z, Bg, tau, slab, inv, lse):
    R = 2 * Bg
    assert not torch.isnan(inv).any() and not torch.isnan(lse).any()      # the exchange filled every row
    zn = z * inv[:, None]
    rows = _rows_of(Bg, slab)
    s = zn[rows] @ zn.t() / tau
    coef = torch.exp(s - lse[rows][:, None]) + torch.exp(s - lse[None, :])
    pos = torch.where(rows < Bg, rows + Bg, rows - Bg)
    coef[torch.arange(rows.numel()), pos] -= 2.0
    coef[torch.arange(rows.numel()), rows] = 0.0
    g = coef @ zn / (tau * R)                                       # d loss / d zn of the slab's rows
    dz = torch.zeros_like(z)
    dz[rows] = (g - (g * zn[rows]).sum(1, keepdim=True) * zn[rows]) * inv[rows][:, None]
    return dz


def _worker_two_phase(rank, world, port, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    B, d_in, d = 6, 10, 8
    w = torch.nn.Linear(d_in, d)
    xi, xj = torch.randn(B, d_in), torch.randn(B, d_in)
    lo, hi = mdist.shard_range(B, rank, world)
    # layout of the statistics exchange: rank r contributes rows [r b, (r+1) b) of view i and of view j
    b = hi - lo
    lse_f, inv_f, tot = mdist.exchange_row_stats(torch.arange(2 * b, dtype=torch.float32) + 100 * rank,
                                                 torch.arange(2 * b, dtype=torch.float32) + 1000 * (rank + 1), torch.tensor([rank + 1.0]))
    want = torch.cat([torch.arange(b) + 100.0 * r for r in range(world)] + [torch.arange(b, 2 * b) + 100.0 * r for r in range(world)])
    ok = torch.equal(lse_f, want) and torch.equal(inv_f - 1000, want + (torch.arange(2 * B) % B // b) * 900.0) and float(tot) == 3.0
    crit = mdist.DistributedNTXent(b, 0.5, phase_fns=(_cpu_lse_slab, _cpu_grad_slab))
    loss = crit(w(xi[lo:hi]), w(xj[lo:hi]))
    loss.backward()
    mdist.allreduce_grads(w.parameters())
    w2 = torch.nn.Linear(d_in, d)
    w2.load_state_dict(w.state_dict())
    ref = O.nt_xent(w2(xi), w2(xj), 0.5)
    ref.backward()
    ok = (ok and abs(float(loss) - float(ref)) < 1e-6
          and torch.allclose(w.weight.grad, w2.weight.grad, atol=1e-6)
          and torch.allclose(w.bias.grad, w2.bias.grad, atol=1e-6)
          and torch.allclose(crit.last_cosine, O.pair_cosine(w2(xi), w2(xj))[lo:hi], atol=1e-6))
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_own_rows_ntxent_with_statistics_exchange_matches_single_process():
    """The >= 4-rank form (every rank reduces only its rows of the global score matrix, one all-gather of the per-row
    statistics, gradient slab from the complete statistics) against the single-process loss and gradient, world size 2."""
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker_two_phase, args=(world, port, ret), nprocs=world, join=True)
        assert dict(ret) == {0: True, 1: True}


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 128, 131):
        for world in (1, 2, 3, 8):
            spans = [mdist.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


# ---- intra-bag sharding: rows of every bag split across ranks ------------------------------------------------------
class _TorchPoolFns:
    """Torch stand-in for the local pooling kernels (the CUDA kernels do not run in this container)."""

    @staticmethod
    def local_pool(h, s, offsets, row_seg, B):
        p = torch.zeros_like(s)
        m = torch.full((B,), float("-inf"))
        l = torch.zeros(B)
        M = torch.zeros(B, h.shape[1])
        for b in range(B):
            lo, hi = int(offsets[b]), int(offsets[b + 1])
            if hi > lo:
                m[b] = s[lo:hi].max()
                e = torch.exp(s[lo:hi] - m[b])
                l[b] = e.sum()
                p[lo:hi] = e / l[b]
                M[b] = p[lo:hi] @ h[lo:hi]
        return p, m, l, M

    @staticmethod
    def backward(p, h, dM, M, offsets, row_seg, B):
        seg = row_seg.long()
        k = (dM * M).sum(1)
        ds = p * ((dM[seg] * h).sum(1) - k[seg])
        return p.unsqueeze(1) * dM[seg], ds


def _pool_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(3)
    sizes = [37, 1, 12, 0, 5]                                # rows per bag (one empty bag, one single-row bag)
    L = 16
    H = [torch.randn(n, L, generator=g) for n in sizes]
    S = [3.0 * torch.randn(n, generator=g) for n in sizes]
    G = torch.randn(len(sizes), L, generator=g)              # upstream gradient, the same on every rank
    ok = True
    for inv_sqrt_n in (False, True):
        # this rank's slice of every bag (rank 1 gets nothing of the single-row bag)
        parts = [mdist.shard_range(n, rank, world) for n in sizes]
        h = torch.cat([H[b][lo:hi] for b, (lo, hi) in enumerate(parts)]).requires_grad_(True)
        s = torch.cat([S[b][lo:hi] for b, (lo, hi) in enumerate(parts)]).requires_grad_(True)
        counts = torch.tensor([hi - lo for lo, hi in parts])
        offsets = torch.cat([torch.zeros(1, dtype=torch.int64), counts.cumsum(0)])
        row_seg = torch.repeat_interleave(torch.arange(len(sizes), dtype=torch.int32), counts)
        M, p = mdist.sharded_attention_pool(h, s, offsets, row_seg, inv_sqrt_n, fns=_TorchPoolFns)
        (M * G).sum().backward()
        # single-process reference on whole bags
        Hf = [x.clone().requires_grad_(True) for x in H]
        Sf = [x.clone().requires_grad_(True) for x in S]
        ref = torch.stack([O.softmax_pool(Sf[b], Hf[b], sizes[b] ** -0.5 if inv_sqrt_n else 1.0)[0] if sizes[b] else torch.zeros(L)
                           for b in range(len(sizes))])
        (ref * G).sum().backward()
        dh_ref = torch.cat([Hf[b].grad[lo:hi] for b, (lo, hi) in enumerate(parts) if sizes[b]])
        ds_ref = torch.cat([Sf[b].grad[lo:hi] for b, (lo, hi) in enumerate(parts) if sizes[b]])
        ok = ok and torch.allclose(M, ref, atol=1e-6) and torch.allclose(h.grad, dh_ref, atol=1e-6) \
            and torch.allclose(s.grad, ds_ref, atol=1e-6)
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_sharded_attention_pool_matches_whole_bags():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_pool_worker, args=(world, port, ret), nprocs=world, join=True)
        assert dict(ret) == {0: True, 1: True}
