"""Whole-step CUDA graph for the launch-bound configurations.

M2-Mixer-S, MIMIC-H and friends are a few MFLOP per sample: a training step is ~100-300 kernel launches of a few
microseconds each, and the step time is the launch path (Python -> dispatcher -> ctypes -> cudaLaunch), not the GPU.
``GraphedTrainStep`` captures ``zero_grad -> training_step -> backward -> (gradient allreduce) -> optimizer step`` once
and replays it with one ``cudaGraphLaunch`` per step:

* inputs are copied into static tensors (the graph's kernels hold raw pointers),
* the optimizer must be ``FusedAdam(..., capturable=True)`` (lr and the step counter live in device memory),
* data parallel (``grad_sync=parallel.attach(opt)``): the step is captured as TWO graphs - (zero_grad, forward, backward)
  and (optimizer step) - with one eager NCCL sum-allreduce of the flat gradient buffer between them; the collective is
  deliberately not captured (NCCL inside a captured multi-stream step hung on this stack), at the price of not
  overlapping it with the backward: 33 MB over NVLink for M2-Mixer-B, ~0.1-0.2 ms against the ~0.6 ms of launch gaps the
  graphs remove,
* dropout stays random: (p, seed) are launch parameters and would be frozen by the capture, so a device-resident epoch
  counter is registered with the library (``ops.set_dropout_epoch``); every kernel folds it into its mask key at run time
  and the captured step ends by advancing it.

Reference behaviour covered: one iteration of the Lightning loop around ``shared_step`` (modules/train_test_module.py:77-83,
models/avmnist.py:236-312, models/mimic.py:93-142) with ``torch.optim.Adam`` (models/avmnist.py:413-415).
"""
from __future__ import annotations

from typing import Any, Callable, Optional, Sequence

import torch

from . import ops
from .optim import FusedAdam


def _static_like(x: Any) -> Any:
    if torch.is_tensor(x):
        return torch.empty_like(x).copy_(x)
    if isinstance(x, dict):
        return {k: _static_like(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return type(x)(_static_like(v) for v in x)
    return x


def _copy_into(dst: Any, src: Any) -> None:
    if torch.is_tensor(dst):
        if dst.shape != src.shape or dst.dtype != src.dtype:
            raise ValueError(f"GraphedTrainStep was captured for a batch of shape {tuple(dst.shape)} / {dst.dtype}, "
                             f"got {tuple(src.shape)} / {src.dtype}")
        dst.copy_(src, non_blocking=True)
    elif isinstance(dst, dict):
        for k in dst:
            _copy_into(dst[k], src[k])
    elif isinstance(dst, (list, tuple)):
        for d, s in zip(dst, src):
            _copy_into(d, s)


class GraphedTrainStep:
    """``step = GraphedTrainStep(model, opt, example_batch); loss = step(batch)`` - loss is a 0-dim device tensor that a
    later call overwrites (clone it to keep it).

    ``static_batches=[b0, b1, ...]`` (device tensors the caller fills in place, e.g. the buffer sets of a
    ``DevicePrefetcher``) captures one graph per buffer set - they share one memory pool - and ``step.replay(k)`` runs
    the step on set k without any input copy."""

    def __init__(self, model: torch.nn.Module, optimizer: FusedAdam, example_batch: Any = None, warmup: int = 3,
                 grad_sync: Optional[Any] = None, step_fn: Optional[Callable[[Any], torch.Tensor]] = None,
                 static_batches: Optional[Sequence[Any]] = None):
        if not isinstance(optimizer, FusedAdam) or not optimizer.capturable:
            raise ValueError("GraphedTrainStep needs FusedAdam(..., capturable=True): lr / step must live on the device")
        if not torch.cuda.is_available():
            raise RuntimeError("GraphedTrainStep needs a GPU: the hot path has no CPU fallback")
        if (example_batch is None) == (static_batches is None):
            raise ValueError("pass either example_batch or static_batches")
        self.model, self.opt, self.sync = model, optimizer, grad_sync
        self._fn = step_fn or (lambda b: model.training_step(b))
        self.batches = list(static_batches) if static_batches is not None else [_static_like(example_batch)]
        self.batch = self.batches[0]
        dev = optimizer.flat_param.device
        self.epoch = torch.zeros(1, dtype=torch.int32, device=dev)
        self.losses = [torch.zeros((), dtype=torch.float32, device=dev) for _ in self.batches]   # outside the graph pool
        ops.set_dropout_epoch(self.epoch)
        optimizer.sync_lr_to_device()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):          # allocator / lazy-init warm-up outside the capture
                self._one(0)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.split = self.sync is not None and getattr(self.sync, "world", 1) > 1
        if self.split:
            self.sync.enabled = False                 # no per-bucket collectives from the backward hooks
        self.graphs, pool = [], None
        for k in range(len(self.batches)):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=pool):
                self._fwd_bwd(k)
                if not self.split:
                    self._update()
            pool = g.pool()
            self.graphs.append(g)
        self.opt_graph = None
        if self.split:
            self.opt_graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.opt_graph, pool=pool):
                self._update()
            self.opt.step_count -= 1                  # the capture ran step()'s host side without executing the kernels
        else:
            self.opt.step_count -= len(self.batches)
        self.graph, self.loss = self.graphs[0], self.losses[0]
        self.replays = 0

    def _fwd_bwd(self, k: int) -> None:
        self.opt.zero_grad()
        loss = self._fn(self.batches[k])
        loss.backward()
        self.losses[k].copy_(loss.detach())

    def _update(self) -> None:
        self.opt.step()
        ops.dropout_epoch_advance(self.epoch)

    def _one(self, k: int) -> None:                   # eager warm-up step
        self._fwd_bwd(k)
        if self.sync is not None:
            self.sync.finish()
        self._update()

    def replay(self, k: int = 0) -> torch.Tensor:
        self.graphs[k].replay()
        if self.split:
            self.sync.allreduce_all()
            self.opt_graph.replay()
        self.replays += 1
        self.opt.step_count += 1                      # host mirror of the device-resident step counter
        return self.losses[k]

    def __call__(self, batch: Any) -> torch.Tensor:
        _copy_into(self.batches[0], batch)
        return self.replay(0)

    def close(self) -> None:
        """Unregister the dropout epoch (eager calls afterwards use their host seeds only) and hand the gradient hooks back."""
        ops.set_dropout_epoch(None)
        if self.split:
            self.sync.enabled = True
