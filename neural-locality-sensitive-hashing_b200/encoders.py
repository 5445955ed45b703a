"""MLP trunks of the learned hasher — same constructors / attribute names as the reference's
encoders.py:8-55 (`{i}_linear`, `{i}_relu`, `output_dim`, `fc1`/`fc2`) so that checkpoints
and nlsh.hashings.extract_layers see the same module tree.  These are PyTorch modules: they
are the training side.  The inference forward over the database / query batches runs in
libnlsh_b200.so (nlsh/_native.py: mlp_hash).

`Siren` (encoders.py:58-79, the trunk main.py:388 constructs) wraps `SIREN` of the third-party package
siren-torch, which the reference neither pins nor vendors (Pipfile:19, absent from Pipfile.lock): when
that package is importable it is used, otherwise `_LocalSIREN` below restates its published structure
(Sitzmann et al. 2020: Linear -> sin(w0_initial x), [Linear -> sin(w0 x)]*, Linear; uniform
+-sqrt(c / fan_in) weights).  PARITY UNPINNED for this trunk: no reference test or golden touches it.
"""
import math
from typing import List

import torch
import torch.nn as nn
import torch.nn.functional as F

try:  # the reference's own dependency, if somebody installed it
    from siren import SIREN as _PackageSIREN
except ImportError:  # not in this image
    _PackageSIREN = None


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


class Sine(nn.Module):
    """sin(w0 * x): the SIREN activation (NLSH_ACT_SIN with act_scale = w0 in the CUDA hasher)."""

    def __init__(self, w0: float = 1.0):
        super().__init__()
        self.w0 = w0

    def forward(self, x):
        return torch.sin(self.w0 * x)


class _LocalSIREN(nn.Module):
    """Restatement of siren-torch's `SIREN(layers, in_features, out_features, w0=1.0, w0_initial=30.0,
    bias=True, initializer='siren', c=6)`: same argument order and defaults, same layer sequence."""

    def __init__(self, layers: List[int], in_features: int, out_features: int, w0: float = 1.0,
                 w0_initial: float = 30.0, bias: bool = True, initializer: str = "siren", c: float = 6):
        super().__init__()
        mods = [nn.Linear(in_features, layers[0], bias=bias), Sine(w0=w0_initial)]
        for index in range(len(layers) - 1):
            mods.extend([nn.Linear(layers[index], layers[index + 1], bias=bias), Sine(w0=w0)])
        mods.append(nn.Linear(layers[-1], out_features, bias=bias))
        self.network = nn.Sequential(*mods)
        if initializer == "siren":
            for m in self.network.modules():
                if isinstance(m, nn.Linear):
                    bound = math.sqrt(c / m.weight.shape[1])
                    with torch.no_grad():
                        m.weight.uniform_(-bound, bound)

    def forward(self, x):
        return self.network(x)


class Siren(nn.Sequential):
    # encoders.py:58-79: hidden_dims[:-1] are the sine layers, hidden_dims[-1] the (linear) output width;
    # with_batchnorm / with_bias are accepted and ignored, as in the reference
    def __init__(self, input_dim, hidden_dims: List[int], with_batchnorm=False, with_bias=True):
        super().__init__()
        self._input_dim = input_dim
        self._hidden_dims = hidden_dims
        self.output_dim = hidden_dims[-1]
        impl = _PackageSIREN if _PackageSIREN is not None else _LocalSIREN
        self.add_module("SIREN", impl(self._hidden_dims[:-1], self._input_dim, self.output_dim))

    def forward(self, x):
        return super().forward(x)
