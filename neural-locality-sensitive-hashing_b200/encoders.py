"""MLP trunks of the learned hasher — same constructors / attribute names as the reference's
encoders.py:8-55 (`{i}_linear`, `{i}_relu`, `output_dim`, `fc1`/`fc2`) so that checkpoints
and nlsh.hashings.extract_layers see the same module tree.  These are PyTorch modules: they
are the training side.  The inference forward over the database / query batches runs in
libnlsh_b200.so (nlsh/_native.py: mlp_hash).  The third-party SIREN trunk of
encoders.py:58-79 is not vendored by the reference and is not provided here.
"""
from typing import List

import torch.nn as nn
import torch.nn.functional as F


class TwoLayer256Relu(nn.Module):

    def __init__(self, input_dim: int, with_bias=True):
        super().__init__()
        self._input_dim = input_dim
        self.output_dim = 256
        self.fc1 = nn.Linear(input_dim, 256, bias=with_bias)
        self.fc2 = nn.Linear(256, self.output_dim, bias=with_bias)

    def forward(self, x):
        return F.relu(self.fc2(F.relu(self.fc1(x))))


class MultiLayerRelu(nn.Sequential):

    def __init__(self, input_dim, hidden_dims: List[int], with_batchnorm=False, with_bias=True):
        super().__init__()
        self._input_dim = input_dim
        self._hidden_dims = hidden_dims
        self.output_dim = hidden_dims[-1]
        prev_dim = input_dim
        for layer_idx, dim in enumerate(hidden_dims):
            self.add_module(f"{layer_idx}_linear", nn.Linear(prev_dim, dim, bias=with_bias))
            if with_batchnorm:
                self.add_module(f"{layer_idx}_batch_norm", nn.BatchNorm1d(dim))
            self.add_module(f"{layer_idx}_relu", nn.ReLU())
            prev_dim = dim
