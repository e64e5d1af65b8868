"""ConcatFusion / SumFusion with the reference's call and ``get_output_shape`` contracts
(modules/fusion.py:112-146, 207-221; shape pins tests/modules/test_fusion.py:8-47 of the reference)."""
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
