"""``NT_Xent`` of utils/losses.py:5-41 on the fused NT-Xent kernel."""
import torch
from torch import nn

from .. import ops


class NT_Xent(nn.Module):
    def __init__(self, batch_size, temperature):
        super(NT_Xent, self).__init__()
        self.batch_size = batch_size
        self.temperature = temperature
        self.last_cosine = None     # cos(z_i[b], z_j[b]) of the last call: the reward of train_MuRCL.py:253,282

    def forward(self, z_i, z_j):
        if z_i.shape[0] != self.batch_size or z_j.shape[0] != self.batch_size:
            # the reference fails here too: its [2B,2B] mask is built for a fixed B (losses.py:11,35)
            raise RuntimeError(f"NT_Xent was built for batch_size={self.batch_size}, got {z_i.shape[0]} and {z_j.shape[0]}")
        loss, cos = ops.ntxent(z_i, z_j, float(self.temperature))
        self.last_cosine = cos
        return loss
