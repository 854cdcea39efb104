"""Contrastive-learning wrapper: the ``CL`` class of the reference (models/cl.py:4-15) keeps an encoder and maps a
list of augmented views to a list of bag embeddings plus their detached copies (the RL states).

The reference encodes view after view.  Bags are independent, so when every view is a dense ``[B, N, D]`` tensor of
the same shape the views are stacked and pushed through the aggregator kernels in one call - half the launches and
twice the rows per GEMM - and split again; ragged / list-shaped views fall back to one call per view.
"""
from typing import List, Sequence, Tuple

import torch
from torch import nn


class CL(nn.Module):
    def __init__(self, encoder: nn.Module, projection_dim: int, n_features: int):
        super().__init__()
        self.encoder = encoder
        # stored but unused upstream as well (cl.py:9-10): the projection lives in rlmil.Full_layer
        self.projection_dim, self.n_features = projection_dim, n_features

    @staticmethod
    def _stackable(views: Sequence) -> bool:
        first = views[0]
        return (all(isinstance(v, torch.Tensor) and v.dim() == 3 for v in views)
                and all(v.shape == first.shape and v.dtype == first.dtype for v in views) and first.shape[0] > 1)

    def forward(self, x_views: List) -> Tuple[List[torch.Tensor], List[torch.Tensor]]:
        assert isinstance(x_views, list), "x_views must be a list of views"
        if len(x_views) > 1 and self._stackable(x_views):
            bags_per_view = x_views[0].shape[0]
            pooled = self.encoder(torch.cat(x_views, 0))[0]
            h_views = list(torch.split(pooled, bags_per_view, 0))
        else:
            h_views = [self.encoder(view)[0] for view in x_views]
        return h_views, [h.detach() for h in h_views]
