"""MLP encoder of the MIMIC static modality (reference modules/mlp.py:4-27): ``module_list`` =
[Linear, ReLU, Dropout] * num_blocks (+ output Linear), same indices / state-dict keys.  Each Linear+ReLU is one
GEMM with a fused epilogue."""
from __future__ import annotations

from torch import nn

from .. import functional as F
from .._lib import ACT_NONE, ACT_RELU
from .mixer import _Slot, _check_dropout, get_default_precision


class MLP(nn.Module):
    def __init__(self, input_dim, hidden_dim, num_blocks, output_dim=None, dropout=0., **kwargs):
        super().__init__()
        self.module_list = nn.ModuleList()
        self.output_dim = output_dim
        self.dropout_p = _check_dropout(dropout, "MLP")
        self.precision = "fp32"   # K = 5..64: launch-bound, tensor cores buy nothing; keep exact arithmetic
        for i in range(num_blocks):
            self.module_list.append(nn.Linear(input_dim if i == 0 else hidden_dim, hidden_dim))
            self.module_list.append(_Slot("ReLU (fused into the GEMM epilogue)"))
            self.module_list.append(_Slot(f"Dropout(p={dropout})"))
        if output_dim is not None:
            self.module_list.append(nn.Linear(hidden_dim, output_dim))

    def forward(self, x):
        p = self.dropout_p if self.training else 0.0
        mods = list(self.module_list)
        for i, m in enumerate(mods):
            if isinstance(m, nn.Linear):
                relu = i + 1 < len(mods) and isinstance(mods[i + 1], _Slot)
                # hidden blocks are Linear -> ReLU -> Dropout; the optional output Linear has neither
                x = F.linear(x, m.weight, m.bias, ACT_RELU if relu else ACT_NONE, self.precision, dropout_p=p if relu else 0.0)
        return x
