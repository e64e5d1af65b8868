"""Host -> device input staging for the training loop.

The reference hands every batch to Lightning, which copies it to the GPU synchronously before ``training_step``
(run.py:59-78, datasets/avmnist.py:113-114).  At M2-Mixer-B's batch 4096 that is 218 MB of fp32 pixels per step - about
4 ms over PCIe, as long as the whole fused forward+backward.  ``DevicePrefetcher`` double-buffers the copy on a side
stream so that the transfer of batch i+1 overlaps the compute of batch i: every batch is still copied from (pinned) host
memory exactly once, only not on the critical path.
"""
from __future__ import annotations

from typing import Dict, Iterable, Iterator, Optional

import torch


def pin_host_batch(batch: Dict[str, torch.Tensor], image_dtype: Optional[torch.dtype] = None) -> Dict[str, torch.Tensor]:
    """Host copy of a batch in pinned memory.  ``image_dtype=torch.bfloat16`` stores the IMAGE tensors (4-D floating
    point: [B, C, H, W]) as bf16: the bf16-mode patch-embedding GEMM rounds every pixel to bf16 as its first operation, so
    the step computes the same bits from half the host-link bytes (218 -> 109 MB per step for M2-Mixer-B at batch 4096,
    which is what bounds the end-to-end rate).  Other tensors (labels, tabular / sequence features) keep their dtype."""
    out = {}
    for k, v in batch.items():
        t = v.detach().cpu()
        if image_dtype is not None and t.is_floating_point() and t.dim() == 4:
            t = t.to(image_dtype)
        out[k] = t.contiguous().pin_memory()
    return out


class DevicePrefetcher:
    """Iterate over host batches (dicts of CPU tensors, ideally pinned) yielding device copies, one batch ahead.

    Two device buffer sets are rotated; a buffer is only overwritten after the compute stream has finished the step
    that read it (tracked with an event recorded by ``__next__`` on the consumer's stream when it hands out the NEXT
    batch - i.e. the consumer must issue all work that reads batch i before asking for batch i+1, which a training
    loop does naturally).
    """

    def __init__(self, batches: Iterable[Dict[str, torch.Tensor]], device: torch.device, depth: int = 2, bufs=None):
        self.src: Iterator[Dict[str, torch.Tensor]] = iter(batches)
        self.device = torch.device(device)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.depth = depth
        # device buffer sets; pass the ``bufs`` of an earlier prefetcher to keep the SAME device tensors (they are the
        # static inputs of captured CUDA graphs, graph.GraphedTrainStep(static_batches=...))
        self.bufs = list(bufs) if bufs is not None else [None] * depth
        self.copied = [torch.cuda.Event() for _ in range(depth)]   # copy of buffer set k finished (copy stream)
        self.released = [None] * depth                  # consumer done with buffer set k (compute stream)
        self.slot = 0
        self.pending = None
        self._issue()

    def _issue(self):
        try:
            host = next(self.src)
        except StopIteration:
            self.pending = None
            return
        k = self.slot
        self.slot = (k + 1) % self.depth
        if self.bufs[k] is None:
            self.bufs[k] = {n: torch.empty(t.shape, dtype=t.dtype, device=self.device) for n, t in host.items()}
        with torch.cuda.stream(self.copy_stream):
            if self.released[k] is not None:
                self.copy_stream.wait_event(self.released[k])
            for n, t in host.items():
                self.bufs[k][n].copy_(t, non_blocking=True)
            self.copied[k].record(self.copy_stream)
        self.pending = k

    def __iter__(self):
        return self

    def index_of(self, batch) -> int:
        """Which buffer set a yielded batch is (the index of the graph captured on it)."""
        for k, b in enumerate(self.bufs):
            if b is batch:
                return k
        raise ValueError("not a batch of this prefetcher")

    def __next__(self) -> Dict[str, torch.Tensor]:
        if self.pending is None:
            raise StopIteration
        k = self.pending
        cur = torch.cuda.current_stream(self.device)
        # everything the consumer queued so far reads older buffers: mark the previous one reusable
        prev = (k - 1) % self.depth
        if self.bufs[prev] is not None:
            ev = torch.cuda.Event()
            ev.record(cur)
            self.released[prev] = ev
        cur.wait_event(self.copied[k])
        out = self.bufs[k]
        self._issue()
        return out
