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
