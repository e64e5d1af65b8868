"""FusedAdam: torch.optim.Adam semantics (as the reference configures it, models/avmnist.py:413-415) executed as ONE
CUDA launch over a flat fp32 parameter buffer.

``FusedAdam(params, lr, betas, eps, weight_decay)`` flattens the given parameters: each ``p.data`` becomes a view into
one contiguous buffer (values preserved, ``load_state_dict`` keeps working in place) and ``p.grad`` a view into one
contiguous gradient buffer - the buffer the data-parallel allreduce buckets slice (see parallel.py).  lr lives in
``param_groups[0]['lr']`` so torch LR schedulers (ReduceLROnPlateau) work unchanged; with ``capturable=True`` lr and
the step counter are mirrored in a 2-float device tensor so that a captured CUDA graph can replay ``step()``.
"""
from __future__ import annotations

from typing import Iterable, List

import torch

from . import ops  # noqa: F401

_O = torch.ops.m2b200


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params: Iterable[torch.nn.Parameter], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0,
                 capturable: bool = False):
        params = [p for p in params]
        super().__init__(params, dict(lr=float(lr), betas=tuple(betas), eps=float(eps), weight_decay=float(weight_decay)))
        ps: List[torch.nn.Parameter] = [p for g in self.param_groups for p in g["params"]]
        if len(self.param_groups) != 1:
            raise ValueError("FusedAdam supports a single param group (the reference uses one)")
        if not ps or not all(p.is_cuda and p.dtype == torch.float32 for p in ps):
            raise RuntimeError("FusedAdam needs float32 CUDA parameters (there is no CPU path)")
        dev = ps[0].device
        # 16-byte aligned slices so every view can be read with 128-bit accesses
        offs, n = [], 0
        for p in ps:
            offs.append(n)
            n += (p.numel() + 3) // 4 * 4
        self.flat_param = torch.zeros(n, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, o in zip(ps, offs):
                view = self.flat_param[o:o + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
                p.grad = self.flat_grad[o:o + p.numel()].view(p.shape)
                # destination for the backward kernels' direct accumulation (functional._direct)
                p._m2_grad = p.grad
        self._params, self._offsets = ps, offs
        self.step_count = 0
        self.capturable = capturable
        self.grad_scale = 1.0
        self._state_dev = torch.tensor([float(lr), 0.0], dtype=torch.float32, device=dev) if capturable else None

    def zero_grad(self, set_to_none: bool = False):
        self.flat_grad.zero_()
        for p, o in zip(self._params, self._offsets):   # re-attach in case autograd replaced .grad
            if p.grad is None or p.grad.data_ptr() != self.flat_grad.data_ptr() + 4 * o:
                p.grad = self.flat_grad[o:o + p.numel()].view(p.shape)

    def sync_lr_to_device(self):
        if self._state_dev is not None:
            self._state_dev[0] = float(self.param_groups[0]["lr"])

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        g = self.param_groups[0]
        self.step_count += 1
        _O.adam_step(self.flat_param, self.flat_grad, self.exp_avg, self.exp_avg_sq, float(g["lr"]), g["betas"][0],
                     g["betas"][1], g["eps"], g["weight_decay"], self.step_count, float(self.grad_scale), self._state_dev)
        from .functional import invalidate_bf16_weights   # the parameters are views of flat_param: their versions did not move
        invalidate_bf16_weights()
        return loss
