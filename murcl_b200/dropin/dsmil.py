"""DSMIL (``FCLayer``, ``BClassifier``, ``MILNet``, ``build_dsmil``) of models/dsmil.py:6-119.

Same parameter names (``i_classifier.fc.0``, ``b_classifier.{q, v.1, fcc}``), forward contracts and return
forms (tensors for one bag ``[1,N,D]``, lists / concatenation for batches).  The critical-instance
arg-max, the q.q_max attention, the softmax over instances and the pooling are segmented kernels.
"""
import torch
from torch import nn

from .. import ops
from ..bags import to_rows


def _precision_dtype(precision):
    return ops.storage_dtype(precision or ops.default_precision())


def _as_bag_list(feats):
    """The three input forms of dsmil.py:26-36 -> list of ``[N_i, D]`` tensors, single-bag flag."""
    if isinstance(feats, torch.Tensor) and len(feats.shape) == 3 and feats.shape[0] == 1:
        return [feats[0]], True
    if isinstance(feats, torch.Tensor) and len(feats.shape) == 3 and feats.shape[0] > 1:
        return [feats[i] for i in range(feats.shape[0])], False
    if isinstance(feats, list):
        out = []
        for f in feats:
            assert len(f.shape) == 3 and f.shape[0] == 1, f"feats.shape: {f.shape}"
            out.append(f[0])
        return out, False
    raise TypeError


class FCLayer(nn.Module):
    def __init__(self, in_size, out_size=1, precision=None):
        super(FCLayer, self).__init__()
        self.fc = nn.Sequential(nn.Linear(in_size, out_size))
        self.precision = precision

    def _scores(self, rows):
        return ops.linear(rows, self.fc[0].weight, self.fc[0].bias, ops.ACT_NONE, _precision_dtype(self.precision))

    def bag_forward(self, feats):
        assert len(feats.shape) == 3 and feats.shape[0] == 1, f"feats.shape: {feats.shape}"
        feats = feats.squeeze(0).cuda()
        return feats, self._scores(feats)

    def _batch(self, bags):
        r = to_rows([b.cuda() for b in bags])
        c = self._scores(r.rows)
        return list(bags), list(torch.split(c, r.sizes, 0))

    def batch_forward(self, feats):
        return self._batch(_as_bag_list(feats)[0])

    def forward(self, feats):
        bags, single = _as_bag_list(feats)
        if single:
            return self.bag_forward(feats)
        return self._batch(bags)


class IClassifier(nn.Module):
    """Defined upstream (dsmil.py:39-49) but never used on the path; kept for import compatibility."""

    def __init__(self, feature_extractor, feature_size, output_class):
        super(IClassifier, self).__init__()
        self.feature_extractor = feature_extractor
        self.fc = nn.Linear(feature_size, output_class)

    def forward(self, x):
        feats = self.feature_extractor(x)
        flat = feats.view(feats.shape[0], -1)
        return flat, ops.linear(flat, self.fc.weight, self.fc.bias)


class BClassifier(nn.Module):
    def __init__(self, input_size, output_class, dropout_v=0.0, precision=None):
        super(BClassifier, self).__init__()
        self.q = nn.Linear(input_size, 128)
        self.v = nn.Sequential(nn.Dropout(dropout_v), nn.Linear(input_size, input_size))
        # constructed but unused upstream (dsmil.py:62,80); kept so checkpoints load
        self.fcc = nn.Conv1d(output_class, output_class, kernel_size=input_size)
        self.dropout_v = float(dropout_v)
        self.precision = precision

    def _run(self, bags, cs):
        if self.training and self.dropout_v > 0:
            raise NotImplementedError("dropout_v > 0 in training mode is not implemented (reference default is 0.0)")
        r = to_rows(bags)
        c = cs[0] if len(cs) == 1 else torch.cat(cs, 0)
        meta = {"B": r.B, "dtype": _precision_dtype(self.precision)}
        bag = ops.dsmil_aggregate(r.rows, c, r.offsets, r.row_seg, meta, self.q.weight, self.q.bias,
                                  self.v[1].weight, self.v[1].bias)
        return bag, bag.detach()

    def bag_forward(self, feats, c):
        return self._run([feats], [c])

    def batch_forward(self, feats, c):
        return self._run(list(feats), list(c))

    def forward(self, feats, c):  # N x K, N x C
        if isinstance(feats, torch.Tensor) and isinstance(c, torch.Tensor):
            B, B_detach = self.bag_forward(feats, c)
        elif isinstance(feats, list) and isinstance(c, list):
            B, B_detach = self.batch_forward(feats, c)
        else:
            raise TypeError
        return B, B_detach


class MILNet(nn.Module):
    def __init__(self, i_classifier, b_classifier):
        super(MILNet, self).__init__()
        self.i_classifier = i_classifier
        self.b_classifier = b_classifier

    def forward(self, x):
        feats, classes = self.i_classifier(x)
        prediction_bag, prediction_bag_detach = self.b_classifier(feats, classes)
        return classes, prediction_bag, prediction_bag_detach


def build_dsmil(dim_feat, num_classes, precision=None):
    i_classifier = FCLayer(in_size=dim_feat, out_size=num_classes, precision=precision).cuda()
    b_classifier = BClassifier(input_size=dim_feat, output_class=num_classes, precision=precision).cuda()
    return MILNet(i_classifier, b_classifier).cuda()
