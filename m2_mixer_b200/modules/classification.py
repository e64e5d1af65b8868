"""StandardClassifier (reference modules/classification.py:84-90): token mean-pool + Linear.  The attribute keeps the
reference's misspelling ``classifer`` because it is a state-dict key.  In the task modules the three heads and their
summed loss run as ONE kernel (functional.heads_loss); this standalone forward is the single-head form of the same
kernel's logits."""
from __future__ import annotations

import torch
from torch import nn

from .. import functional as F
from .._lib import ACT_NONE


class StandardClassifier(nn.Module):
    def __init__(self, input_shape: tuple, num_classes: int, **kwargs):
        super().__init__()
        self.classifer = nn.Linear(input_shape[-1], num_classes)

    def forward(self, inputs: torch.Tensor) -> torch.Tensor:
        return F.linear(F.mean_pool(inputs), self.classifer.weight, self.classifer.bias, ACT_NONE, "fp32")


def _mlp_chain(layers, x, precision="fp32"):
    """Linear / ReLU chain of the reference's ModuleList classifiers: a ReLU that follows a Linear runs in its GEMM epilogue."""
    from .._lib import ACT_RELU
    mods = list(layers)
    i = 0
    while i < len(mods):
        m = mods[i]
        if isinstance(m, nn.Linear):
            relu = i + 1 < len(mods) and isinstance(mods[i + 1], nn.ReLU)
            x = F.linear(x, m.weight, m.bias, ACT_RELU if relu else ACT_NONE, precision)
            i += 2 if relu else 1
        else:                       # a ReLU that does not follow a Linear (never built by the reference ctors)
            x = torch.relu(x)
            i += 1
    return x


def _build_chain(in_dim, hidden_dims, num_classes):
    layers = nn.ModuleList([nn.Linear(in_dim, hidden_dims[0])])
    for i in range(len(hidden_dims) - 1):     # the reference puts NO ReLU after the first Linear (classification.py:74-78)
        layers.append(nn.Linear(hidden_dims[i], hidden_dims[i + 1]))
        layers.append(nn.ReLU())
    layers.append(nn.Linear(hidden_dims[-1], num_classes))
    return layers


class BasicClassifier(nn.Module):
    """Reference modules/classification.py:69-82 (cfg/avmnist/avmnist_post.yml): Linear / ReLU chain applied to the input as it
    is; state-dict keys ``classifier.<i>.weight / bias``."""

    def __init__(self, input_shape: tuple, hidden_dims: list, num_classes: int, **kwargs):
        super().__init__()
        self.classifier = _build_chain(input_shape[-1], hidden_dims, num_classes)

    def forward(self, inputs: torch.Tensor) -> torch.Tensor:
        return _mlp_chain(self.classifier, inputs)


class MultilayerClassifier(nn.Module):
    """Reference modules/classification.py:33-47: mean over dims 1 and 1 again ([B, a, b, D] -> [B, D]), then the same chain;
    the attribute keeps the reference's ``classifer`` spelling (state-dict key)."""

    def __init__(self, input_shape: tuple, hidden_dims: list, num_classes: int, **kwargs):
        super().__init__()
        self.classifer = _build_chain(input_shape[-1], hidden_dims, num_classes)

    def forward(self, inputs: torch.Tensor) -> torch.Tensor:
        x = inputs.reshape(inputs.shape[0], -1, inputs.shape[-1]) if inputs.dim() == 4 else inputs
        if inputs.dim() != 4:
            raise ValueError("MultilayerClassifier expects [B, a, b, D] inputs (two successive means over dim 1)")
        return _mlp_chain(self.classifer, F.mean_pool(x))
