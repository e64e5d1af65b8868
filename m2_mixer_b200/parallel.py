"""Batch-sharded data parallelism: one process per GPU, NCCL allreduce of the flat gradient buffer, bucketed and
overlapped with backward.

The reference has no distributed code (it delegates to Lightning's default for ``devices=-1``, run.py:59-74); the hot
path shards over the batch only (every op is per-sample), so the single exchange step per iteration is the
sum of weight gradients.  ``GradSync`` slices the optimiser's flat gradient buffer into buckets ordered from the END
of the parameter list to the start (parameters are registered in forward order, backward produces gradients in
reverse), and fires ``all_reduce`` for a bucket on a dedicated communication stream as soon as the last gradient of
that bucket has been accumulated - while the remaining backward kernels keep running on the compute stream.  The
1/world_size scaling is folded into the fused Adam kernel (``optimizer.grad_scale``), saving a pass over the grads.
"""
from __future__ import annotations

import os
from typing import List, Sequence

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None) -> tuple[int, int, int]:
    """(rank, local_rank, world) from torchrun's environment; initialises the default process group if world > 1."""
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world > 1 and not dist.is_initialized():
        if torch.cuda.is_available():
            torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        kw = {}
        if torch.cuda.is_available() and (backend or "nccl") == "nccl":
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend or ("nccl" if torch.cuda.is_available() else "gloo"), rank=rank, world_size=world, **kw)
    return rank, local, world


class GradSync:
    def __init__(self, params: Sequence[torch.nn.Parameter], offsets: Sequence[int], flat_grad: torch.Tensor,
                 bucket_bytes: int = 2 << 20, group=None):
        self.flat_grad, self.group = flat_grad, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.cuda = flat_grad.is_cuda
        # high priority: a collective's CTAs should take the next SMs that free up - a half-scheduled NCCL kernel spins on its
        # peers while it holds SMs the persistent one-CTA-per-SM compute kernels are waiting for
        prio = int(os.environ.get("M2B200_COMM_PRIORITY", "-1"))
        self.comm_stream = torch.cuda.Stream(device=flat_grad.device, priority=prio) if self.cuda else None
        # buckets over contiguous flat ranges, last parameters first
        ends = [o + (p.numel() + 3) // 4 * 4 for p, o in zip(params, offsets)]
        self.buckets: List[list] = []      # [start, end, n_params]
        self.bucket_of = {}
        cur_end, cur_start, count = None, None, 0
        for i in range(len(params) - 1, -1, -1):
            if cur_end is None:
                cur_end = ends[i]
            cur_start, count = offsets[i], count + 1
            self.bucket_of[i] = len(self.buckets)
            if (cur_end - cur_start) * 4 >= bucket_bytes or i == 0:
                self.buckets.append([cur_start, cur_end, count])
                cur_end, count = None, 0
        self._pending = [b[2] for b in self.buckets]
        self._fired = [False] * len(params)          # a parameter counts ONCE per step (see _make_hook)
        self._launched = [False] * len(self.buckets)
        self._handles = []
        self.enabled = True       # False: the hooks do nothing (graph.GraphedTrainStep reduces the whole buffer between graphs)
        if self.world > 1:
            for i, p in enumerate(params):
                if p.requires_grad:
                    hook = self._make_hook(i)
                    p.register_post_accumulate_grad_hook(hook)
                    # kernels that accumulate straight into the flat buffer bypass AccumulateGrad: they call this
                    p._m2_ready = (lambda h=hook, q=p: h(q))
                else:
                    self._pending[self.bucket_of[i]] -= 1
            self._static_pending = list(self._pending)

    def _make_hook(self, i):
        b = self.bucket_of[i]

        def hook(_p):
            # Called through p._m2_ready by the kernels that accumulate straight into the flat buffer AND by autograd's
            # post-accumulate hook: torch runs that hook even when the Function returned None for the parameter (measured
            # on 2 x B200, tools/ddp_debug.py: every parameter reported twice and every bucket was reduced half way through
            # its gradients).  The first report is the one that follows the producing kernels in stream order.
            if not self.enabled or self._fired[i]:
                return
            self._fired[i] = True
            self._pending[b] -= 1
            if self._pending[b] == 0:
                self._launch(b)
        return hook

    def _launch(self, b):
        start, end, _ = self.buckets[b]
        view = self.flat_grad[start:end]
        self._launched[b] = True
        if self.cuda:
            self.comm_stream.wait_stream(torch.cuda.current_stream())
            from .functional import COMPUTE_STREAMS   # forked encoders: this bucket may hold gradients of both streams
            for st in COMPUTE_STREAMS:
                if st.device == self.flat_grad.device:
                    self.comm_stream.wait_stream(st)
            with torch.cuda.stream(self.comm_stream):
                dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group)
        else:
            self._handles.append(dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self):
        """Call after backward, before optimizer.step(): flushes buckets whose hooks did not all fire (unused
        parameters) and makes the compute stream wait for the communication stream."""
        if self.world == 1 or not self.enabled:
            return
        for b in range(len(self.buckets)):
            if not self._launched[b]:
                self._launch(b)
        for h in self._handles:
            h.wait()
        self._handles = []
        if self.cuda:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        self._pending = list(self._static_pending)
        self._fired = [False] * len(self._fired)
        self._launched = [False] * len(self.buckets)


    def allreduce_all(self):
        """One sum-allreduce over the whole flat gradient buffer on the current stream (no overlap): what a graphed step
        uses between its forward/backward graph and its optimizer graph."""
        if self.world > 1:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=self.group)


def attach(optimizer, bucket_bytes: int = 2 << 20, group=None) -> GradSync:
    """Wire a FusedAdam to data-parallel gradient averaging: sum-allreduce buckets + 1/world folded into Adam."""
    sync = GradSync(optimizer._params, optimizer._offsets, optimizer.flat_grad, bucket_bytes, group)
    optimizer.grad_scale = 1.0 / sync.world
    return sync
