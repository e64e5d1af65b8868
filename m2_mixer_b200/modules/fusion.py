"""ConcatFusion / SumFusion / MaxFusion / MeanFusion with the reference's call and ``get_output_shape`` contracts
(modules/fusion.py:112-146, 190-221, 258-272; shape pins tests/modules/test_fusion.py:8-47 of the reference)."""
from __future__ import annotations

import torch

from .. import functional as F


class ConcatFusion:
    def __init__(self, dim=1, **kwargs):
        self.dim = dim

    def __call__(self, *args):
        if self.dim != 1 or any(a.dim() != 3 for a in args):
            raise NotImplementedError("m2b200 ConcatFusion concatenates [B, N_i, D] token tensors along dim=1")
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
