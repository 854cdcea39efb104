"""Turns the input forms the reference's ``forward`` methods accept into CSR rows + offsets."""
from __future__ import annotations

from typing import List, Sequence, Union

import torch

from . import ops
from ._lib import MurclError

_dense_cache = {}


class Rows:
    """All instances of a batch of bags as one ``[n_rows, D]`` tensor plus the CSR bookkeeping."""

    __slots__ = ("rows", "offsets", "offsets_host", "row_seg", "B")

    def __init__(self, rows, offsets, offsets_host, row_seg):
        self.rows, self.offsets, self.offsets_host, self.row_seg = rows, offsets, offsets_host, row_seg
        self.B = len(offsets_host) - 1

    @property
    def sizes(self) -> List[int]:
        return [b - a for a, b in zip(self.offsets_host[:-1], self.offsets_host[1:])]


def _bookkeeping(sizes: Sequence[int], device):
    key = (tuple(sizes), str(device)) if len(sizes) <= 1024 else None
    if key is not None and key in _dense_cache:
        return _dense_cache[key]
    offs = [0]
    for n in sizes:
        offs.append(offs[-1] + int(n))
    offsets = torch.tensor(offs, dtype=torch.int64, device=device)
    row_seg = ops.row_segments(offsets, offs[-1])
    out = (offsets, offs, row_seg)
    if key is not None and len(_dense_cache) < 256:
        _dense_cache[key] = out
    return out


def to_rows(x: Union[torch.Tensor, List[torch.Tensor]]) -> Rows:
    """list of ``[N_i, D]`` / ``[1, N_i, D]`` tensors, or a dense ``[B, N, D]`` tensor (no copy)."""
    if isinstance(x, (list, tuple)):
        bags = [b.reshape(-1, b.shape[-1]) for b in x]
        if not bags:
            raise MurclError("empty bag list")
        rows = bags[0] if len(bags) == 1 else torch.cat(bags, 0)
        sizes = [int(b.shape[0]) for b in bags]
    elif isinstance(x, torch.Tensor):
        if x.dim() == 2:
            x = x.unsqueeze(0)
        if x.dim() != 3:
            raise MurclError(f"expected [B, N, D], got {tuple(x.shape)}")
        rows = x.reshape(-1, x.shape[-1])
        sizes = [int(x.shape[1])] * int(x.shape[0])
    else:
        raise TypeError
    if not rows.is_cuda:
        raise MurclError("bags must live on a CUDA device (libmurcl_b200 has no CPU path)")
    offsets, offs, row_seg = _bookkeeping(sizes, rows.device)
    return Rows(rows, offsets, offs, row_seg)
