"""ConcatFusion / SumFusion / MaxFusion / MeanFusion with the reference's call and ``get_output_shape`` contracts
(modules/fusion.py:112-146, 190-221, 258-272; shape pins tests/modules/test_fusion.py:8-47 of the reference)."""
from __future__ import annotations

import torch
from torch import nn

from .. import functional as F


class ConcatFusion:
    def __init__(self, dim=1, **kwargs):
        self.dim = dim

    def __call__(self, *args):
        if self.dim != 1 or any(a.dim() != 3 for a in args):
            # any other rank / dim (the reference accepts them, modules/fusion.py:117): torch.cat on the device tensors -
            # plumbing outside the hot path, which is token concatenation (and zero-copy inside the task modules)
            return torch.cat(args, dim=self.dim)
        return F.concat_tokens(*args)

    def get_output_shape(self, *args, dim=None):
        if dim is not None:
            if not isinstance(args[0], int):
                raise ValueError("The dim argument is only used if the first argument is an int.")
            return sum(args) if dim == self.dim else args[0]
        shape = list(args[0])
        for other in args[1:]:
            shape[self.dim] += other[self.dim]
        return tuple(shape)


class SumFusion:
    def __init__(self, **kwargs):
        pass

    def __call__(self, *args):
        if len(args) != 2:   # torch.add(a, b, alpha) in the reference: a third positional would be `alpha`
            raise TypeError("SumFusion takes exactly two tensors")
        return F.add(args[0], args[1])

    @staticmethod
    def get_output_shape(*args, dim=None, **kwargs):
        if dim is not None and not isinstance(args[0], int):
            raise ValueError("The dim argument is only used if the first argument is an int.")
        if args[0] != args[1]:
            raise ValueError("Input shapes must be equal")
        return args[0]


def _same_shape(args, dim):
    if dim is not None and not isinstance(args[0], int):
        raise ValueError("The dim argument is only used if the first argument is an int.")
    if args[0] != args[1]:
        raise ValueError("Input shapes must be equal")
    return args[0]


class MaxFusion:
    """torch.maximum(a, b) (reference modules/fusion.py:190-204), ties share the gradient evenly as in torch."""

    def __init__(self, **kwargs):
        pass

    def __call__(self, *args):
        if len(args) != 2:
            raise TypeError("MaxFusion takes exactly two tensors")   # torch.maximum(*args) in the reference
        return F.fuse_max(args[0], args[1])

    @staticmethod
    def get_output_shape(*args, dim=None):
        return _same_shape(args, dim)


class MeanFusion:
    """torch.mean(torch.stack(args), 0) (reference modules/fusion.py:258-272); two modalities run as one kernel, more are
    folded pairwise into a running sum and scaled once."""

    def __init__(self, **kwargs):
        pass

    def __call__(self, *args):
        if len(args) == 2:
            return F.fuse_mean(args[0], args[1])
        if len(args) < 2:
            raise TypeError("MeanFusion needs at least two tensors")
        acc = F.add(args[0], args[1])
        for a in args[2:]:
            acc = F.add(acc, a)
        return acc * (1.0 / len(args))

    @staticmethod
    def get_output_shape(*args, dim=None, **kwargs):
        return _same_shape(args, dim)


class BiModalGatedUnit(nn.Module):
    """Gated fusion of two modalities (reference modules/fusion.py:7-55): ``z * tanh(W1 m1) + (1 - z) * tanh(W2 m2)`` with
    ``z = sigmoid(Wz [m1, m2])``.  Same parameters (``mod1_hidden``, ``mod2_hidden``, ``z_hidden``) and
    ``get_output_shape`` contract; the three linears run through ``m2b200_linear_*``, the gate is one kernel each way."""

    def __init__(self, mod1_in, mod2_in, out_size, **kwargs):
        super().__init__()
        self.out_size = out_size
        self.mod1_hidden = nn.Linear(mod1_in, out_size)
        self.mod2_hidden = nn.Linear(mod2_in, out_size)
        self.z_hidden = nn.Linear(mod1_in + mod2_in, out_size)
        self.precision = "fp32"    # exact by default (the gate saturates quickly); "bf16" runs the linears on tcgen05

    def forward(self, mod1, mod2):
        h1 = F.linear(mod1, self.mod1_hidden.weight, self.mod1_hidden.bias, 0, self.precision)
        h2 = F.linear(mod2, self.mod2_hidden.weight, self.mod2_hidden.bias, 0, self.precision)
        zh = F.linear(torch.cat([mod1, mod2], dim=-1), self.z_hidden.weight, self.z_hidden.bias, 0, self.precision)
        return F.gate(h1, h2, zh)

    def get_output_shape(self, *args, dim=None):
        if dim is not None:
            if not isinstance(args[0], int):
                raise ValueError("The dim argument is only used if the first argument is an int.")
            if dim == -1:
                return self.out_size
            return args[0]
        shape1 = list(args[0])
        shape1[-1] = self.out_size
        return tuple(shape1)
