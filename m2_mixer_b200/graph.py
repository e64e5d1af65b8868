"""Whole-step CUDA graph for the launch-bound configurations.

M2-Mixer-S, MIMIC-H and friends are a few MFLOP per sample: a training step is ~100-300 kernel launches of a few
microseconds each, and the step time is the launch path (Python -> dispatcher -> ctypes -> cudaLaunch), not the GPU.
``GraphedTrainStep`` captures ``zero_grad -> training_step -> backward -> (gradient allreduce) -> optimizer step`` once
and replays it with one ``cudaGraphLaunch`` per step:

* inputs are copied into static tensors (the graph's kernels hold raw pointers),
* the optimizer must be ``FusedAdam(..., capturable=True)`` (lr and the step counter live in device memory),
* data parallel (``grad_sync=parallel.attach(opt)``), ``comm="overlap"`` (or ``M2B200_GRAPH_COMM=overlap``): ONE graph per
  static batch; the bucketed NCCL sum-allreduces of ``parallel.GradSync`` are captured on its communication stream, forked
  from the compute streams as soon as the last gradient of a bucket has been produced and joined before the optimizer step,
  so the collectives run under the remaining backward kernels (the north_star's "bucketed and overlapped with backward").
  Call ``close()`` before ``destroy_process_group``: graphs that hold NCCL kernel nodes must die before the communicator.
  ``comm="split"`` (default) is TWO graphs - (zero_grad, forward, backward) and (optimizer step) - around one eager allreduce
  of the whole flat gradient buffer.  Measured on 2 and 8 x B200 (profiles/r02_multi_gpu.md): since the two encoders run
  on two streams every SM is busy during the backward, and a collective underneath displaces more compute than it hides
  (8 x B200: 2.98 ms overlapped with 16 MiB buckets vs 2.95 ms split; 2 x B200: 2.98 vs 2.87 ms),
* dropout stays random: (p, seed) are launch parameters and would be frozen by the capture, so a device-resident epoch
  counter is registered with the library (``ops.set_dropout_epoch``); every kernel folds it into its mask key at run time
  and the captured step ends by advancing it.

Reference behaviour covered: one iteration of the Lightning loop around ``shared_step`` (modules/train_test_module.py:77-83,
models/avmnist.py:236-312, models/mimic.py:93-142) with ``torch.optim.Adam`` (models/avmnist.py:413-415).
"""
from __future__ import annotations

import gc
import os
from typing import Any, Callable, Optional, Sequence

import torch

from . import functional as F
from . import ops
from .optim import FusedAdam

_EPOCH_OWNER = None      # the live GraphedTrainStep whose epoch counter is registered with the library (process-global)


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
    the step on set k without any input copy.

    Construction runs ``warmup`` real steps on the first batch (allocator / lazy-init warm-up; collectives at world > 1)
    and then RESTORES parameters, Adam moments, step counter and dropout epoch (``restore=True``), so that the first
    ``step(batch)`` is training step 1 of an untouched model.  Scalars captured by value (the loss weights of
    ``shared_step``, the set of frozen parameters) need a new GraphedTrainStep when they change; lr does not - it is
    read from device memory and ``replay`` mirrors ``param_groups[0]['lr']`` there whenever a scheduler moved it.
    The dropout epoch pointer is process-global: the most recently constructed live instance owns it."""

    def __init__(self, model: torch.nn.Module, optimizer: FusedAdam, example_batch: Any = None, warmup: int = 3,
                 grad_sync: Optional[Any] = None, step_fn: Optional[Callable[[Any], torch.Tensor]] = None,
                 static_batches: Optional[Sequence[Any]] = None, comm: Optional[str] = None, restore: bool = True):
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
        global _EPOCH_OWNER
        ops.set_dropout_epoch(self.epoch)
        _EPOCH_OWNER = self
        optimizer.sync_lr_to_device()
        dp = self.sync is not None and getattr(self.sync, "world", 1) > 1
        comm = comm or os.environ.get("M2B200_GRAPH_COMM", "split")
        if comm not in ("overlap", "split"):
            raise ValueError("comm must be 'overlap' or 'split'")
        self.split = dp and comm == "split"
        self.overlap = dp and comm == "overlap"
        saved = None
        if restore:
            saved = (optimizer.flat_param.clone(), optimizer.exp_avg.clone(), optimizer.exp_avg_sq.clone(),
                     optimizer.step_count, optimizer._state_dev.clone())
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):          # allocator / lazy-init warm-up outside the capture
                self._one(0)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        if saved is not None:
            optimizer.flat_param.copy_(saved[0]); optimizer.exp_avg.copy_(saved[1]); optimizer.exp_avg_sq.copy_(saved[2])
            optimizer.step_count = saved[3]
            optimizer._state_dev.copy_(saved[4])
            self.epoch.zero_()
            F.invalidate_bf16_weights()
            F.refresh_bf16_weights()                  # the captured forwards then find every bf16 copy current
            torch.cuda.synchronize(dev)
        if self.split:
            self.sync.enabled = False                 # no per-bucket collectives from the backward hooks
        self.graphs, pool = [], None
        for k in range(len(self.batches)):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=pool):
                self._fwd_bwd(k)
                if self.overlap:
                    self.sync.finish()                # joins the communication stream back into the capture
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
        # The bf16 operand copies are refreshed HERE, right behind the Adam kernel, not lazily at the next forward's first
        # weight lookup: a lazy refresh is captured only by the graph whose capture happened to find the cache stale, and
        # with several static batches (or the split form) the other graphs would replay on weights 1..K-1 steps old.
        F.refresh_bf16_weights()
        ops.dropout_epoch_advance(self.epoch)

    def _one(self, k: int) -> None:                   # eager warm-up step
        self._fwd_bwd(k)
        if self.sync is not None:
            self.sync.finish()
        self._update()

    def replay(self, k: int = 0) -> torch.Tensor:
        if self.opt.lr_changed():                     # an LR scheduler moved param_groups[0]['lr'] (host side, no capture)
            self.opt.sync_lr_to_device()
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
        """Destroy the graphs (before any ``destroy_process_group``: they may hold NCCL kernel nodes), unregister the dropout
        epoch if this instance owns it (eager calls afterwards use their host seeds only) and hand the gradient hooks back."""
        global _EPOCH_OWNER
        torch.cuda.synchronize()
        self.graphs, self.opt_graph, self.graph = [], None, None
        gc.collect()
        if _EPOCH_OWNER is self:
            ops.set_dropout_epoch(None)
            _EPOCH_OWNER = None
        if self.split:
            self.sync.enabled = True
