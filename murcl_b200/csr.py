"""Device-resident ragged-bag store (CSR) and the packer calls on it.

Replaces the per-step Python list handling of ``get_feats`` (utils/datasets.py:274-308): all slides
of a batch (or of the whole dataset - Camelyon16 is ~4.4 GB) live in ONE
feature buffer ``feats[n_rows, D]`` with ``offsets[B+1]``; every patch carries its cluster label and
its rank inside that cluster, which is all the selection rule needs.
"""
from __future__ import annotations

import os

from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import BF16, F32, MurclError, check


def _s():
    return torch.cuda.current_stream().cuda_stream


class BagStore:
    """CSR store of B slides on one CUDA device."""

    def __init__(self, feats: torch.Tensor, offsets_host: Sequence[int], patch_cluster: torch.Tensor,
                 patch_rank: torch.Tensor, cluster_sizes: torch.Tensor, num_clusters: int):
        self.feats = feats                      # [n_rows, D] fp32 cuda
        self.offsets_host = [int(o) for o in offsets_host]
        self.offsets = torch.tensor(self.offsets_host, dtype=torch.int64, device=feats.device)
        self.patch_cluster = patch_cluster      # [n_rows] int32
        self.patch_rank = patch_rank            # [n_rows] int32
        self.cluster_sizes = cluster_sizes      # [B, K] int32
        self.K = int(num_clusters)

    # -- construction ---------------------------------------------------------------------------
    @staticmethod
    def _stack_feats(feat_list, device, dtype=torch.float32) -> tuple:
        sizes = [int(f.shape[-2]) for f in feat_list]
        d = int(feat_list[0].shape[-1])
        offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        feats = torch.empty((int(offsets[-1]), d), dtype=dtype, device=device)
        for f, lo, hi in zip(feat_list, offsets[:-1], offsets[1:]):
            # host (ideally pinned) or device source: one copy straight into the CSR slot
            feats[lo:hi].copy_(f.reshape(-1, d), non_blocking=True)
        return feats, offsets

    @classmethod
    def from_cluster_lists(cls, feat_list: List[torch.Tensor], clusters_list: List[List[List[int]]], device=None):
        """From the reference's in-memory format: per slide a ``[N,D]`` (or ``[1,N,D]``) tensor and the
        JSON inverted lists ``clusters[j] = [patch ids]`` (datasets.py:145-165).  The rank of a patch
        is its POSITION in its list, so non-ascending lists behave as in the reference's slicing."""
        device = torch.device(device) if device is not None else (
            feat_list[0].device if feat_list[0].is_cuda else torch.device("cuda"))
        if device.type != "cuda":
            raise MurclError("BagStore needs a CUDA device (libmurcl_b200 has no CPU path)")
        feats, offsets = cls._stack_feats(feat_list, device)
        K = len(clusters_list[0])
        n_rows = int(offsets[-1])
        cluster = np.full(n_rows, -1, dtype=np.int32)
        rank = np.full(n_rows, -1, dtype=np.int32)
        sizes = np.zeros((len(feat_list), K), dtype=np.int32)
        for b, clusters in enumerate(clusters_list):
            if len(clusters) != K:
                raise MurclError("all slides must use the same number of clusters")
            lo = offsets[b]
            for j, c in enumerate(clusters):
                if len(c):
                    ids = np.asarray(c, dtype=np.int64) + lo
                    cluster[ids] = j
                    rank[ids] = np.arange(len(c), dtype=np.int32)
                sizes[b, j] = len(c)
        return cls(feats, offsets, torch.from_numpy(cluster).to(device), torch.from_numpy(rank).to(device),
                   torch.from_numpy(sizes).to(device), K)

    @classmethod
    def from_labels(cls, feat_list: List[torch.Tensor], labels_list: List[torch.Tensor], num_clusters: int, device=None):
        """From the on-disk format: per slide ``img_features [N,D]`` and ``features_cluster_indices [N]``
        (features_clustering.py:10-16).  Ranks and cluster sizes are computed on the device."""
        device = torch.device(device) if device is not None else torch.device("cuda")
        feats, offsets = cls._stack_feats(feat_list, device)
        n_rows, B = int(offsets[-1]), len(feat_list)
        labels = torch.empty((n_rows,), dtype=torch.int32, device=device)
        for l, lo, hi in zip(labels_list, offsets[:-1], offsets[1:]):
            labels[lo:hi].copy_(torch.as_tensor(l).reshape(-1).to(torch.int32), non_blocking=True)
        rank = torch.empty_like(labels)
        sizes = torch.empty((B, num_clusters), dtype=torch.int32, device=device)
        off_dev = torch.tensor(offsets, dtype=torch.int64, device=device)
        check(_lib.load().murcl_csr_rank_patches(labels.data_ptr(), off_dev.data_ptr(), B, num_clusters, rank.data_ptr(),
                                                 sizes.data_ptr(), _s()), "murcl_csr_rank_patches")
        return cls(feats, offsets, labels, rank, sizes, num_clusters)

    @classmethod
    def empty_like_host(cls, host: "HostBags", device=None):
        """Device buffers sized for ``host`` (contents undefined until ``copy_from_host``)."""
        device = torch.device(device) if device is not None else torch.device("cuda")
        n_rows, d = host.feats.shape
        return cls(torch.empty((n_rows, d), dtype=host.feats.dtype, device=device), host.offsets,
                   torch.empty((n_rows,), dtype=torch.int32, device=device),
                   torch.empty((n_rows,), dtype=torch.int32, device=device),
                   torch.empty(tuple(host.cluster_sizes.shape), dtype=torch.int32, device=device), host.K)

    def copy_from_host(self, host: "HostBags") -> int:
        """Asynchronous H2D of a whole batch from pinned memory on the current stream.  Returns bytes moved."""
        if tuple(host.feats.shape) != tuple(self.feats.shape) or list(host.offsets) != self.offsets_host:
            raise MurclError("copy_from_host: host batch does not match the device buffers")
        self.feats.copy_(host.feats, non_blocking=True)
        self.patch_cluster.copy_(host.patch_cluster, non_blocking=True)
        self.patch_rank.copy_(host.patch_rank, non_blocking=True)
        self.cluster_sizes.copy_(host.cluster_sizes, non_blocking=True)
        return host.nbytes

    # -- properties -------------------------------------------------------------------------------
    @property
    def num_bags(self) -> int:
        return len(self.offsets_host) - 1

    @property
    def dim(self) -> int:
        return int(self.feats.shape[1])

    @property
    def device(self):
        return self.feats.device

    # -- packer -----------------------------------------------------------------------------------
    def select(self, actions: torch.Tensor, feat_size: int, slot_bag: Optional[torch.Tensor] = None):
        """Selection only: ``sel_idx [S, FS]`` int32 global CSR rows (-1 = pad) and ``sel_cnt [S]``."""
        if not actions.is_cuda:
            raise MurclError("BagStore.select: actions must be a CUDA tensor")
        actions = actions.detach().to(torch.float32).contiguous()
        S, K = actions.shape
        if K != self.K:
            raise MurclError(f"actions have {K} columns but the store has {self.K} clusters")
        if slot_bag is None and S != self.num_bags:
            raise MurclError(f"{S} action rows for {self.num_bags} bags (pass slot_bag to map slots to bags)")
        sel_idx = torch.empty((S, feat_size), dtype=torch.int32, device=self.device)
        sel_cnt = torch.empty((S,), dtype=torch.int32, device=self.device)
        check(_lib.load().murcl_pack_select(self.patch_cluster.data_ptr(), self.patch_rank.data_ptr(), self.offsets.data_ptr(),
                                            self.cluster_sizes.data_ptr(), None if slot_bag is None else slot_bag.data_ptr(),
                                            actions.data_ptr(), S, K, feat_size, sel_idx.data_ptr(), sel_cnt.data_ptr(), _s()),
              "murcl_pack_select")
        return sel_idx, sel_cnt

    def gather(self, sel_idx: torch.Tensor, lam: Optional[torch.Tensor] = None, perm: Optional[torch.Tensor] = None,
               out_dtype: torch.dtype = torch.float32, order: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Gather + zero pad (+ mixup) -> ``[S, FS, D]``."""
        return gather_rows_padded(self.feats, sel_idx, lam, perm, out_dtype, order)

    def pack(self, actions, feat_size, lam=None, perm=None, out_dtype=torch.float32, slot_bag=None, order=None):
        sel_idx, _ = self.select(actions, feat_size, slot_bag)
        return self.gather(sel_idx, lam, perm, out_dtype, order)


def perm_cycle_order(perm: torch.Tensor) -> torch.Tensor:
    """``perm`` int32 ``[n, S]`` (or ``[S]``): mixup partners of the S output slots of n gathers -> the slots of each gather
    listed cycle by cycle (``murcl_perm_cycle_order``), the walk order that makes every source row a single DRAM read."""
    if not perm.is_cuda or perm.dtype != torch.int32 or not perm.is_contiguous():
        raise MurclError("perm_cycle_order: perm must be a contiguous int32 CUDA tensor")
    flat = perm.reshape(-1, perm.shape[-1])
    order = torch.empty_like(flat)
    check(_lib.load().murcl_perm_cycle_order(flat.data_ptr(), flat.shape[0], flat.shape[1], order.data_ptr(), _s()),
          "murcl_perm_cycle_order")
    return order.view(perm.shape)


def gather_rows_padded(feats: torch.Tensor, sel_idx: torch.Tensor, lam=None, perm=None, out_dtype=torch.float32, order=None):
    """``order`` (int32 ``[S]``, a permutation of the slots): the sequence in which the kernel walks the output slots;
    default with mixup: the cycle order of ``perm`` (``MURCL_GATHER_ORDER=0``: plain order).  Never changes the result."""
    if not feats.is_cuda or feats.dtype not in (torch.float32, torch.bfloat16) or not feats.is_contiguous():
        raise MurclError("gather: feats must be a contiguous fp32 or bf16 CUDA tensor")
    S, FS = sel_idx.shape
    D = feats.shape[1]
    out = torch.empty((S, FS, D), dtype=out_dtype, device=feats.device)
    if lam is not None:
        lam = lam.detach().reshape(-1).to(torch.float32).contiguous()
        perm = perm.detach().reshape(-1).to(torch.int32).contiguous()
        if lam.numel() != S or perm.numel() != S:
            raise MurclError("gather: lam / perm must have one entry per output slot")
        if order is None and S >= 16 and os.environ.get("MURCL_GATHER_ORDER", "1") != "0":
            order = perm_cycle_order(perm)
    if order is not None:
        if not order.is_cuda or order.dtype != torch.int32 or order.numel() != S or not order.is_contiguous():
            raise MurclError("gather: order must be a contiguous int32 CUDA tensor with one entry per output slot")
    code_of = {torch.float32: F32, torch.bfloat16: BF16}
    code = code_of[out_dtype]
    check(_lib.load().murcl_pack_gather_ordered(feats.data_ptr(), code_of[feats.dtype], D, sel_idx.data_ptr(), S, FS,
                                                None if lam is None else lam.data_ptr(), None if perm is None else perm.data_ptr(),
                                                None if order is None else order.data_ptr(), out.data_ptr(), code, _s()),
          "murcl_pack_gather_ordered")
    return out


class HostBags:
    """A batch of slides staged in PINNED host memory in the CSR layout (what a data loader hands over)."""

    def __init__(self, feat_list, labels_list, num_clusters: int, pin: bool = True, dtype: torch.dtype = torch.float32,
                 ranks_list=None):
        """``labels_list[b][p]`` = cluster of patch p (``features_cluster_indices``); ``ranks_list[b][p]`` = position of
        the patch inside its cluster's list.  Without ``ranks_list`` the lists are taken to be in ascending patch order
        (how ``features_clustering.py:19-25`` writes them): rank = number of earlier patches with the same label."""
        sizes = [int(f.shape[-2]) for f in feat_list]
        d = int(feat_list[0].shape[-1])
        self.offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64).tolist()
        n_rows = self.offsets[-1]
        self.K = int(num_clusters)
        self.feats = torch.empty((n_rows, d), dtype=dtype)         # bf16 in bf16 mode: half the H2D bytes per step
        self.patch_cluster = torch.empty((n_rows,), dtype=torch.int32)
        self.patch_rank = torch.empty((n_rows,), dtype=torch.int32)
        self.cluster_sizes = torch.zeros((len(sizes), self.K), dtype=torch.int32)
        for b, (f, l) in enumerate(zip(feat_list, labels_list)):
            lo, hi = self.offsets[b], self.offsets[b + 1]
            self.feats[lo:hi] = f.reshape(-1, d)
            lab = torch.as_tensor(l).reshape(-1).to(torch.int64)
            self.patch_cluster[lo:hi] = lab.to(torch.int32)
            # rank of a patch inside its cluster = number of earlier patches with the same label
            onehot = torch.nn.functional.one_hot(lab, self.K)
            if ranks_list is not None and ranks_list[b] is not None:
                self.patch_rank[lo:hi] = torch.as_tensor(ranks_list[b]).reshape(-1).to(torch.int32)
            else:
                self.patch_rank[lo:hi] = ((onehot.cumsum(0) - 1) * onehot).sum(1).to(torch.int32)
            self.cluster_sizes[b] = onehot.sum(0).to(torch.int32)
        if pin and torch.cuda.is_available():
            self.feats = self.feats.pin_memory()
            self.patch_cluster = self.patch_cluster.pin_memory()
            self.patch_rank = self.patch_rank.pin_memory()
            self.cluster_sizes = self.cluster_sizes.pin_memory()

    @classmethod
    def from_files(cls, feature_files: Sequence[str], cluster_files: Sequence[str], num_clusters: int, pin: bool = True,
                   dtype: torch.dtype = torch.float32) -> "HostBags":
        """Slides from the reference's on-disk format (README.md:102-136): ``feature_files[b]`` is the ``.npz`` written by
        ``wsi_processing/extract_features.py`` (array ``img_features`` [N, D], read at utils/datasets.py:89,150);
        ``cluster_files[b]`` is either the ``.npz`` of ``features_clustering.py:10-16`` (array ``features_cluster_indices``
        [N, 1]) or the ``.json`` inverted lists of ``features_clustering.py:19-25`` (K lists of patch ids, read at
        utils/datasets.py:151,163).  JSON lists keep their order: a patch's rank is its position in its list."""
        feats, labels, ranks = [], [], []
        for ff, cf in zip(feature_files, cluster_files):
            f, lab, rk = load_slide(ff, cf, num_clusters)
            feats.append(f)
            labels.append(lab)
            ranks.append(rk)
        return cls(feats, labels, num_clusters, pin=pin, dtype=dtype, ranks_list=ranks)

    @property
    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in (self.feats, self.patch_cluster, self.patch_rank, self.cluster_sizes))


class ResidentSlides:
    """A whole dataset (or a rank's shard of it) kept in HBM as ONE CSR arena, filled on first touch.

    The reference copies every slide host -> device each time it is visited (train_MuRCL.py:224-227); Camelyon16's training
    features are ~2.2 GB in bf16 (4.4 GB fp32) against 180 GB of HBM, so here a slide crosses PCIe once: ``ensure(ids)``
    uploads the slides of a batch that are not resident yet (asynchronously, from the pinned ``HostBags`` staging, on the
    current stream) and a step then addresses its bags through ``slot_bag`` - an index list into the arena - which is the
    only per-step host -> device traffic in steady state.  The arena's row ranges are fixed by the dataset's patch counts,
    so its addresses never change (CUDA-graph friendly)."""

    def __init__(self, host: "HostBags", device=None):
        self.host = host
        self.store = BagStore.empty_like_host(host, device)
        self.resident = np.zeros(len(host.offsets) - 1, dtype=bool)
        self.bytes_uploaded = 0

    @property
    def num_slides(self) -> int:
        return len(self.resident)

    def ensure(self, slide_ids) -> int:
        """Upload the not-yet-resident slides among ``slide_ids`` (features, cluster ids, ranks, cluster sizes) on the
        current stream.  Returns the number of bytes queued (0 when everything was resident)."""
        ids = np.unique(np.asarray(slide_ids, dtype=np.int64))
        todo = ids[~self.resident[ids]]
        if todo.size == 0:
            return 0
        h, st, off = self.host, self.store, self.host.offsets
        moved = 0
        # coalesce runs of consecutive slide ids into one copy per array
        runs, start, prev = [], int(todo[0]), int(todo[0])
        for b in todo[1:].tolist():
            if b != prev + 1:
                runs.append((start, prev))
                start = b
            prev = b
        runs.append((start, prev))
        for b0, b1 in runs:
            lo, hi = off[b0], off[b1 + 1]
            for dst, src in ((st.feats, h.feats), (st.patch_cluster, h.patch_cluster), (st.patch_rank, h.patch_rank)):
                dst[lo:hi].copy_(src[lo:hi], non_blocking=True)
                moved += (hi - lo) * src.element_size() * (src.shape[1] if src.dim() == 2 else 1)
            st.cluster_sizes[b0:b1 + 1].copy_(h.cluster_sizes[b0:b1 + 1], non_blocking=True)
            moved += (b1 + 1 - b0) * h.cluster_sizes.shape[1] * 4
        self.resident[todo] = True
        self.bytes_uploaded += moved
        return moved

    def evict_all(self) -> None:
        """Forget what is resident (the next touch uploads again): used to measure a cold epoch."""
        self.resident[:] = False


def load_slide(feature_file: str, cluster_file: str, num_clusters: int):
    """One slide from disk -> (features fp32 [N, D], labels int32 [N], ranks int32 [N] or None).  See
    ``HostBags.from_files`` for the formats.  Raises ``ValueError`` when the two files disagree."""
    import json
    feats = torch.as_tensor(np.load(feature_file)["img_features"], dtype=torch.float32)
    if feats.dim() != 2:
        raise ValueError(f"{feature_file}: img_features must be [N, D], got {tuple(feats.shape)}")
    n = feats.shape[0]
    if str(cluster_file).endswith(".json"):
        with open(cluster_file) as fh:
            lists = json.load(fh)
        if len(lists) != num_clusters:
            raise ValueError(f"{cluster_file}: {len(lists)} cluster lists, expected {num_clusters}")
        labels = np.full((n,), -1, dtype=np.int32)
        ranks = np.zeros((n,), dtype=np.int32)
        for j, ids in enumerate(lists):
            ids = np.asarray(ids, dtype=np.int64)
            if ids.size and (ids.min() < 0 or ids.max() >= n):
                raise ValueError(f"{cluster_file}: patch id out of range for {n} patches")
            labels[ids] = j
            ranks[ids] = np.arange(ids.size, dtype=np.int32)
        if (labels < 0).any() or sum(len(ids) for ids in lists) != n:
            raise ValueError(f"{cluster_file}: the cluster lists are not a partition of the {n} patches")
        return feats, torch.from_numpy(labels), torch.from_numpy(ranks)
    labels = np.load(cluster_file)["features_cluster_indices"].reshape(-1).astype(np.int32)
    if labels.shape[0] != n:
        raise ValueError(f"{cluster_file}: {labels.shape[0]} labels for {n} patches")
    if labels.size and (labels.min() < 0 or labels.max() >= num_clusters):
        raise ValueError(f"{cluster_file}: cluster label outside [0, {num_clusters})")
    return feats, torch.from_numpy(labels), None
