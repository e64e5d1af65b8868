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
        self._lr_on_device = float(lr)

    def zero_grad(self, set_to_none: bool = False):
        self.flat_grad.zero_()
        for p, o in zip(self._params, self._offsets):   # re-attach in case autograd replaced .grad
            if p.grad is None or p.grad.data_ptr() != self.flat_grad.data_ptr() + 4 * o:
                p.grad = self.flat_grad[o:o + p.numel()].view(p.shape)

    def sync_lr_to_device(self):
        """Write ``param_groups[0]['lr']`` (what torch LR schedulers such as the reference's ReduceLROnPlateau change,
        models/avmnist.py:416-422) into the device-resident scalar a captured step reads.  ``step()`` and
        ``GraphedTrainStep.replay()`` call this themselves whenever the host value moved, outside any capture."""
        if self._state_dev is not None:
            lr = float(self.param_groups[0]["lr"])
            self._state_dev[0:1].fill_(lr)
            self._lr_on_device = lr

    def lr_changed(self) -> bool:
        return self._state_dev is not None and float(self.param_groups[0]["lr"]) != self._lr_on_device

    def active_ranges(self):
        """Contiguous [start, end) ranges of the flat buffers whose parameters are trainable.  torch.optim.Adam skips a
        parameter whose ``.grad`` is None - what ``requires_grad_(False)`` gives the reference's frozen encoders and heads
        (models/avmnist.py:243-256): their moments and values must not move (no momentum tail, no weight decay)."""
        ranges = []
        for p, o in zip(self._params, self._offsets):
            if not p.requires_grad:
                continue
            end = o + (p.numel() + 3) // 4 * 4
            if ranges and ranges[-1][1] == o:
                ranges[-1][1] = end
            else:
                ranges.append([o, end])
        return ranges

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        g = self.param_groups[0]
        self.step_count += 1
        if self.lr_changed() and not torch.cuda.is_current_stream_capturing():
            self.sync_lr_to_device()
        ranges = self.active_ranges()
        whole = len(ranges) == 1 and ranges[0][0] == 0 and ranges[0][1] == self.flat_param.numel()
        args = (float(g["lr"]), g["betas"][0], g["betas"][1], g["eps"], g["weight_decay"])
        if whole:
            _O.adam_step(self.flat_param, self.flat_grad, self.exp_avg, self.exp_avg_sq, *args, self.step_count,
                         float(self.grad_scale), self._state_dev)
        else:   # frozen parameters: one launch per trainable range; only the first advances the device step counter
            for i, (a, b) in enumerate(ranges):
                _O.adam_step(self.flat_param[a:b], self.flat_grad[a:b], self.exp_avg[a:b], self.exp_avg_sq[a:b], *args,
                             self.step_count if (i == 0 or self._state_dev is None) else -1, float(self.grad_scale),
                             self._state_dev)
            if not ranges and self._state_dev is not None:
                self._state_dev[1:2].add_(1.0)
        from .functional import invalidate_bf16_weights   # the parameters are views of flat_param: their versions did not move
        invalidate_bf16_weights()
        return loss

    # ---- checkpointing in torch.optim.Adam's layout (what Lightning stores under ``optimizer_states``)
    def state_dict(self):
        """{'state': {i: {'step', 'exp_avg', 'exp_avg_sq'}}, 'param_groups': [...]} exactly as torch.optim.Adam writes it, so
        a checkpoint moves between the reference and this optimizer in both directions.  Tensors are detached copies."""
        with torch.no_grad():
            for p, o in zip(self._params, self._offsets):
                n = p.numel()
                self.state[p] = {"step": torch.tensor(float(self.step_count)),
                                 "exp_avg": self.exp_avg[o:o + n].view(p.shape).clone(),
                                 "exp_avg_sq": self.exp_avg_sq[o:o + n].view(p.shape).clone()}
        sd = super().state_dict()
        self.state.clear()
        return sd

    @torch.no_grad()
    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)          # validates the layout, casts / moves tensors, restores param_groups
        steps = []
        for p, o in zip(self._params, self._offsets):
            st = self.state.get(p)
            if not st:
                continue
            n = p.numel()
            self.exp_avg[o:o + n].copy_(st["exp_avg"].reshape(-1))
            self.exp_avg_sq[o:o + n].copy_(st["exp_avg_sq"].reshape(-1))
            steps.append(int(float(st["step"])))
        self.state.clear()
        self.step_count = max(steps) if steps else 0
        if self._state_dev is not None:
            self._state_dev[1:2].fill_(float(self.step_count))
            self.sync_lr_to_device()
