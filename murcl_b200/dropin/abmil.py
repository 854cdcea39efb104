"""``ABMIL`` of models/abmil.py:7-63 on the fused MIL aggregator kernels.

Same constructor, parameter names (``encoder.{0,3,6}``, ``attention.{0,2}``, ``decoder.0``, ``fc``) and
forward contract.  All bags of a call are processed together as CSR rows: the reference's Python loop
over bags (abmil.py:47-51) becomes one batched encoder GEMM chain plus segmented softmax pooling.
"""
import torch
from torch import nn

from .. import ops
from ..bags import to_rows


class ABMIL(nn.Module):
    def __init__(self, dim_in, L=512, D=128, K=1, dim_out=2, dropout=0., precision=None):
        super(ABMIL, self).__init__()
        self.L, self.D, self.K = L, D, K
        if K != 1:
            raise NotImplementedError("ABMIL drop-in supports K=1 attention heads (the reference default)")
        self.dropout_p = float(dropout)
        self.precision = precision          # None -> MURCL_PRECISION env (fp32 | bf16)
        self.shard_rows, self.shard_group = False, None      # see shard_bags()

        # parameter containers only: layout mirrors the reference so checkpoints load unchanged
        self.encoder = nn.Sequential(
            nn.Linear(dim_in, L), nn.ReLU(), nn.Dropout(dropout),
            nn.Linear(L, L), nn.ReLU(), nn.Dropout(dropout),
            nn.Linear(L, L), nn.ReLU(),
        )
        self.attention = nn.Sequential(nn.Linear(L, D), nn.Tanh(), nn.Linear(D, K))
        self.decoder = nn.Sequential(nn.Linear(L, L), nn.ReLU())
        self.fc = nn.Linear(L, dim_out)     # defined but never applied upstream (abmil.py:33)

    def shard_bags(self, enabled=True, group=None):
        """Intra-bag sharding (BASELINE config 5: one 100k-patch bag over 2/4/8 GPUs): every rank of ``group`` passes ITS
        rows of each bag to ``forward`` (same number of bags, in the same order, on every rank; a rank may hold none of a
        bag's rows only if another holds some).  Pooling partials are merged with one all-gather; the outputs are the
        whole-bag results on every rank and the parameter gradients are per-rank partial sums (sum them with
        ``dist.allreduce_grads``, as for data parallelism).  The decoder works on the merged, replicated bag vector: its
        parameter gradients are pre-divided by the group size so that the same sum over ranks is exact for them too.  Any
        layer applied AFTER this module sees replicated inputs as well (average, do not sum, its gradients)."""
        self.shard_rows, self.shard_group = bool(enabled), group
        return self

    # ------------------------------------------------------------------------------------------
    def _meta(self, rows):
        prec = self.precision or ops.default_precision()
        meta = {"B": rows.B, "gated": False, "inv_sqrt_n": True, "dtype": ops.storage_dtype(prec)}
        if self.training and self.dropout_p > 0:
            meta["drop"] = {"enc": [self.dropout_p, self.dropout_p, 0.0], "attn": 0.0}     # abmil.py:15,18
        if self.shard_rows:
            meta.update(shard=True, shard_group=self.shard_group)
        return meta

    def pooled(self, x):
        """Encoder + attention pooling WITHOUT the decoder: ``[B, L]`` bag vectors (abmil.py:36-43).  For callers that run
        the decoder layer themselves (the recurrent-head tape batches its backward over the T patch-steps of a step)."""
        rows = to_rows(x)
        enc = [p for i in (0, 3, 6) for p in (self.encoder[i].weight, self.encoder[i].bias)]
        M, p, _s, _il, _pr = ops.mil_aggregate(rows.rows, rows.offsets, rows.row_seg, self._meta(rows),
                                               self.attention[0].weight, self.attention[0].bias,
                                               self.attention[2].weight, self.attention[2].bias, None, None, enc)
        self.last_attention = p
        return M

    def _aggregate(self, x):
        rows = to_rows(x)
        enc = [p for i in (0, 3, 6) for p in (self.encoder[i].weight, self.encoder[i].bias)]
        M, p, _s, _il, _pr = ops.mil_aggregate(rows.rows, rows.offsets, rows.row_seg, self._meta(rows),
                                               self.attention[0].weight, self.attention[0].bias,
                                               self.attention[2].weight, self.attention[2].bias, None, None, enc)
        dw, db = self.decoder[0].weight, self.decoder[0].bias
        if self.shard_rows:
            # the pooled vectors are identical on every rank of the shard group, so the decoder's parameter gradients are
            # complete on each of them already: pre-divide by the group size so that the usual sum over ranks is exact
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.shard_group) > 1:
                f = 1.0 / dist.get_world_size(self.shard_group)
                dw, db = ops.scale_grad(dw, f), ops.scale_grad(db, f)
        out = ops.linear(M, dw, db, ops.ACT_RELU, self._meta(rows)["dtype"])
        self.last_attention = p             # [n_rows] pooling weights incl. the 1/sqrt(N) post-scale (abmil.py:40-41)
        return out, p

    def bag_forward(self, bag):
        """One bag ``[N, dim_in]`` -> ``[1, L]`` (abmil.py:35-45)."""
        return self._aggregate([bag])[0]

    def batch_forward(self, batch):
        """Sequence of bags -> ``[B, L]`` (abmil.py:47-51), batched instead of looped."""
        return self._aggregate(list(batch) if not isinstance(batch, torch.Tensor) else batch)[0]

    def forward(self, x):  # B x N x dim_in, a bag
        if isinstance(x, list):
            outputs = self.batch_forward(x)
        elif isinstance(x, torch.Tensor):
            if x.shape[0] == 1:
                outputs = self.bag_forward(x.squeeze(0))
            else:
                outputs = self.batch_forward(x)
        else:
            raise TypeError
        return outputs, outputs.detach()
