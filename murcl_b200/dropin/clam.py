"""``CLAM_SB`` (+ ``Attn_Net``, ``Attn_Net_Gated``) of models/clam.py:18-211 on the fused MIL aggregator.

Same constructor, parameter names (``attention_net.0``, ``attention_net.{2|3}.attention_{a.0,b.0,c}``,
``classifiers``, ``instance_classifiers.{i}``), forward signature and return forms: a dict for one bag,
a list of dicts for a batch.  The instance-clustering loss (clam.py:103-132) runs on the segmented
top-k kernel and the fused instance-CE kernel.
"""
import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from .. import ops
from ..bags import to_rows


def initialize_weights(module):
    """Xavier-normal weights, zero biases (clam.py:7-15)."""
    for m in module.modules():
        if isinstance(m, nn.Linear):
            nn.init.xavier_normal_(m.weight)
            m.bias.data.zero_()
        elif isinstance(m, nn.BatchNorm1d):
            nn.init.constant_(m.weight, 1)
            nn.init.constant_(m.bias, 0)


class Attn_Net(nn.Module):
    """Parameter container of the plain attention head (clam.py:18-34); evaluated by CLAM_SB."""

    def __init__(self, L=1024, D=256, dropout=False, n_classes=1):
        super(Attn_Net, self).__init__()
        layers = [nn.Linear(L, D), nn.Tanh()]
        if dropout:
            layers.append(nn.Dropout(0.25))
        layers.append(nn.Linear(D, n_classes))
        self.module = nn.Sequential(*layers)

    def parts(self):
        first, last = self.module[0], self.module[-1]
        return first.weight, first.bias, last.weight, last.bias


class Attn_Net_Gated(nn.Module):
    """Parameter container of the gated attention head (clam.py:37-60); evaluated by CLAM_SB."""

    def __init__(self, L=1024, D=256, dropout=False, n_classes=1):
        super(Attn_Net_Gated, self).__init__()
        a = [nn.Linear(L, D), nn.Tanh()]
        b = [nn.Linear(L, D), nn.Sigmoid()]
        if dropout:
            a.append(nn.Dropout(0.25))
            b.append(nn.Dropout(0.25))
        self.attention_a = nn.Sequential(*a)
        self.attention_b = nn.Sequential(*b)
        self.attention_c = nn.Linear(D, n_classes)

    def parts(self):
        # one [2D, L] projection: tanh branch rows first, sigmoid branch rows second
        w = torch.cat([self.attention_a[0].weight, self.attention_b[0].weight], 0)
        b = torch.cat([self.attention_a[0].bias, self.attention_b[0].bias], 0)
        return w, b, self.attention_c.weight, self.attention_c.bias


class CLAM_SB(nn.Module):
    def __init__(self, gate=True, size_arg="small", dropout=False, k_sample=8, n_classes=2,
                 instance_loss_fn=nn.CrossEntropyLoss(), subtyping=False, in_dim=512, precision=None):
        super(CLAM_SB, self).__init__()
        self.size_dict = {"small": [in_dim, 512, 256], "big": [in_dim, 512, 384]}
        size = self.size_dict[size_arg]
        fc = [nn.Linear(size[0], size[1]), nn.ReLU()]
        if dropout:
            fc.append(nn.Dropout(0.25))
        net = Attn_Net_Gated if gate else Attn_Net
        fc.append(net(L=size[1], D=size[2], dropout=dropout, n_classes=1))
        self.attention_net = nn.Sequential(*fc)
        self.classifiers = nn.Linear(size[1], n_classes)          # never applied upstream (clam.py:171-173)
        self.instance_classifiers = nn.ModuleList([nn.Linear(size[1], 2) for _ in range(n_classes)])
        self.k_sample = k_sample
        self.instance_loss_fn = instance_loss_fn
        if not isinstance(instance_loss_fn, nn.CrossEntropyLoss):
            raise NotImplementedError("the fused instance loss implements nn.CrossEntropyLoss (the reference default)")
        self.n_classes = n_classes
        self.subtyping = subtyping
        self.gate = gate
        self.has_dropout = bool(dropout)
        self.precision = precision
        self.shard_rows, self.shard_group = False, None      # see shard_bags()
        initialize_weights(self)

    def shard_bags(self, enabled=True, group=None):
        """Intra-bag sharding (BASELINE config 5: one 100k-patch bag over 2/4/8 GPUs): every rank of ``group`` passes ITS
        rows of each bag to ``forward``; pooling partials are merged with one all-gather, outputs are the whole-bag
        results on every rank, parameter gradients are per-rank partial sums (``dist.allreduce_grads``).  The instance
        loss needs whole bags and is refused in this mode."""
        self.shard_rows, self.shard_group = bool(enabled), group
        return self

    def relocate(self):
        """Move the sub-modules to the GPU when there is one (clam.py:86-90)."""
        target = torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")
        for name in ("attention_net", "classifiers", "instance_classifiers"):
            setattr(self, name, getattr(self, name).to(target))

    @staticmethod
    def create_positive_targets(length, device):
        return torch.ones(length, dtype=torch.long, device=device)

    @staticmethod
    def create_negative_targets(length, device):
        return torch.zeros(length, dtype=torch.long, device=device)

    # ------------------------------------------------------------------------------------------
    def _meta(self, rows):
        prec = self.precision or ops.default_precision()
        meta = {"B": rows.B, "gated": bool(self.gate), "inv_sqrt_n": False, "dtype": ops.storage_dtype(prec)}
        if self.training and self.has_dropout:
            # Dropout(0.25) after the fc ReLU (clam.py:70-71) and after each attention branch (:26-27,46-48)
            meta["drop"] = {"enc": [0.25], "attn": 0.25}
        if self.shard_rows:
            meta.update(shard=True, shard_group=self.shard_group)
        return meta

    def _check_mode(self):
        pass

    def _instance_plan(self, rows, labels):
        """Host-side layout of the (bag, class) groups the instance loss visits (clam.py:146-166)."""
        k = self.k_sample
        if min(rows.sizes) < k:
            raise RuntimeError(f"selected index k out of range: a bag has fewer than k_sample={k} instances")
        groups, targets, group_off, group_cls = [], [], [0], []
        for b, lab in enumerate(labels):
            for c in range(self.n_classes):
                if c == lab:
                    groups.append((b, c, True))
                    targets += [1] * k + [0] * k
                elif self.subtyping:
                    groups.append((b, c, False))
                    targets += [0] * k
                else:
                    continue
                group_off.append(len(targets))
                group_cls.append(c)
        dev = rows.rows.device
        return {"k": k, "groups": groups,
                "targets": torch.tensor(targets, dtype=torch.int32, device=dev),
                "group_off": torch.tensor(group_off, dtype=torch.int32, device=dev),
                "group_cls": torch.tensor(group_cls, dtype=torch.int32, device=dev)}

    def _run(self, x, label=None, instance_eval=False, return_features=False):
        self._check_mode()
        rows = to_rows(x)
        meta = self._meta(rows)
        wab, bab, wc, bc = self.attention_net[-1].parts()
        enc = [self.attention_net[0].weight, self.attention_net[0].bias]
        inst_w = inst_b = None
        labels = None
        if instance_eval:
            labels = [int(v) for v in torch.as_tensor(label).reshape(-1).tolist()]
            if len(labels) != rows.B:
                raise RuntimeError(f"{len(labels)} labels for {rows.B} bags")
            meta["inst"] = self._instance_plan(rows, labels)
            inst_w = torch.stack([m.weight for m in self.instance_classifiers], 0)
            inst_b = torch.stack([m.bias for m in self.instance_classifiers], 0)
        M, p, s, inst_loss, preds = ops.mil_aggregate(rows.rows, rows.offsets, rows.row_seg, meta, wab, bab, wc, bc,
                                                      inst_w, inst_b, enc)
        self.last_attention = p             # [n_rows] post-softmax attention of all bags of the call, CSR order
        results = [dict() for _ in range(rows.B)]
        if instance_eval:
            plan = meta["inst"]
            preds_h = preds.cpu().numpy()
            targets_h = plan["targets"].cpu().numpy()
            off = plan["group_off"].cpu().numpy()
            per_bag = [[] for _ in range(rows.B)]
            for g, (b, _c, _in) in enumerate(plan["groups"]):
                per_bag[b].append(g)
            for b in range(rows.B):
                gs = per_bag[b]
                if gs:
                    total = inst_loss[gs].sum()
                    sl = [np.arange(off[g], off[g + 1]) for g in gs]
                    sl = np.concatenate(sl)
                    all_preds, all_targets = preds_h[sl].astype(np.int64), targets_h[sl].astype(np.int64)
                else:
                    total, all_preds, all_targets = 0.0, np.array([]), np.array([])
                if self.subtyping:
                    total = total / len(self.instance_classifiers)
                results[b] = {"instance_loss": total, "inst_labels": all_targets, "inst_preds": all_preds}
        if return_features:
            for b in range(rows.B):
                results[b].update({"features": M[b:b + 1]})
        return M, results

    # ------------------------------------------------------------------------------------------
    def bag_forward(self, bag, label=None, instance_eval=False, return_features=False, attention_only=False):
        if len(bag.shape) == 3 and bag.shape[0] == 1:
            bag = bag.squeeze(0)
        assert len(bag.shape) == 2, f"h.shape: {bag.shape}"
        if attention_only:
            self._check_mode()
            rows = to_rows([bag])
            wab, bab, wc, bc = self.attention_net[-1].parts()
            s = ops.mil_attention_scores(rows.rows, self._meta(rows), wab, bab, wc, bc,
                                         [self.attention_net[0].weight, self.attention_net[0].bias])
            return s.reshape(1, -1)            # raw, pre-softmax scores (clam.py:141-142)
        M, results = self._run([bag], label, instance_eval, return_features)
        return M, results[0]

    def batch_forward(self, batch, label=None, instance_eval=False, return_features=False, attention_only=False):
        if attention_only:
            raise ValueError("attention_only is a single-bag call (use bag_forward), as upstream")
        bags = batch if isinstance(batch, torch.Tensor) and batch.dim() == 3 else list(batch)
        return self._run(bags, label, instance_eval, return_features)

    def forward(self, h, label=None, instance_eval=False, return_features=False, attention_only=False):
        if isinstance(h, list):
            outputs, results_dict = self.batch_forward(h, label, instance_eval, return_features, attention_only)
        elif isinstance(h, torch.Tensor):
            if h.shape[0] == 1:
                outputs, results_dict = self.bag_forward(h.squeeze(0), label, instance_eval, return_features,
                                                         attention_only)
            else:
                outputs, results_dict = self.batch_forward(h, label, instance_eval, return_features, attention_only)
        else:
            raise TypeError
        if instance_eval:
            return outputs, outputs.detach(), results_dict
        else:
            return outputs, outputs.detach()
