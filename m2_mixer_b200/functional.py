"""Autograd wiring of the m2b200 ops: each forward op is paired with its hand-written backward kernel sequence.

Backward kernels RECOMPUTE LayerNorm and GELU from the saved block input instead of storing activations
(north_star): the only tensors saved per Mixer block are its two fp32 inputs (x for token mixing, u for
channel mixing) - no [tokens x channel_dim] hidden tensor is ever kept alive between forward and backward.
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence, Tuple

import torch
from torch.autograd import Function

from . import ops  # noqa: F401  (registers torch.ops.m2b200.*)
from ._lib import ACT_NONE, ACT_RELU, BF16, FP32

_O = torch.ops.m2b200

_PRECISION = {"fp32": FP32, "float32": FP32, "bf16": BF16, "bfloat16": BF16, FP32: FP32, BF16: BF16}


def precision_code(p) -> int:
    try:
        return _PRECISION[p]
    except KeyError:
        raise ValueError(f"unknown precision {p!r} (use 'bf16' or 'fp32')") from None


def _up8(v: int) -> int:
    return (v + 7) // 8 * 8


def _direct(params):
    """Gradient destinations registered by FusedAdam (views into its flat gradient buffer).  When EVERY parameter of an
    op has one, the backward kernels accumulate straight into them (no per-parameter zero-fill + AccumulateGrad add
    launches) and autograd gets None for those inputs; data-parallel bucket hooks are notified through _m2_ready."""
    dst = [getattr(p, "_m2_grad", None) if p is not None else None for p in params]
    live = [d for p, d in zip(params, dst) if p is not None and p.requires_grad]
    if live and all(d is not None for d in live) and len(live) == sum(1 for p in params if p is not None):
        return dst
    return None


def _notify(params):
    for p in params:
        cb = getattr(p, "_m2_ready", None) if p is not None else None
        if cb is not None:
            cb()


def _ret(direct, grads):
    return tuple(None for _ in grads) if direct is not None else tuple(grads)


class _TokenMix(Function):
    @staticmethod
    def forward(ctx, x, ln_w, ln_b, w1, b1, w2, b2, precision, p, seed):
        u = _O.token_mix_fwd(x, ln_w, ln_b, w1, b1, w2, b2, precision, p, seed)
        ctx.save_for_backward(x, ln_w, ln_b, w1, b1, w2)
        ctx.precision, ctx.p, ctx.seed = precision, p, seed
        ctx.params = (ln_w, ln_b, w1, b1, w2, b2)
        return u

    @staticmethod
    def backward(ctx, du):
        x, ln_w, ln_b, w1, b1, w2 = ctx.saved_tensors
        direct = _direct(ctx.params)
        dx, *g = _O.token_mix_bwd(du.contiguous(), x, ln_w, ln_b, w1, b1, w2, ctx.precision, ctx.p, ctx.seed, direct)
        if direct is not None:
            _notify(ctx.params)
        return (dx, *_ret(direct, g), None, None, None)


class _ChannelMix(Function):
    @staticmethod
    def forward(ctx, u, ln_w, ln_b, w1, b1, w2, b2, w1b, w2b, precision, p, seed):
        y = _O.channel_mix_fwd(u, ln_w, ln_b, w1, b1, w2, b2, w1b, w2b, precision, p, seed)
        ctx.save_for_backward(u, ln_w, ln_b, w1, b1, w2, w1b, w2b)
        ctx.precision, ctx.p, ctx.seed = precision, p, seed
        ctx.params = (ln_w, ln_b, w1, b1, w2, b2)
        return y

    @staticmethod
    def backward(ctx, dy):
        u, ln_w, ln_b, w1, b1, w2, w1b, w2b = ctx.saved_tensors
        direct = _direct(ctx.params)
        du, *g = _O.channel_mix_bwd(dy.contiguous(), u, ln_w, ln_b, w1, b1, w2, w1b, w2b, ctx.precision, ctx.p, ctx.seed,
                                    direct)
        if direct is not None:
            _notify(ctx.params)
        return (du, *_ret(direct, g), None, None, None, None, None)


class _LayerNorm(Function):
    @staticmethod
    def forward(ctx, x, w, b):
        ctx.save_for_backward(x, w)
        ctx.params = (w, b)
        return _O.layernorm_fwd(x, w, b)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        direct = _direct(ctx.params)
        dx, *g = _O.layernorm_bwd(dy.contiguous(), x, w, direct)
        if direct is not None:
            _notify(ctx.params)
        return (dx, *_ret(direct, g))


class _LayerNormConcat(Function):
    """fused = cat([LN_i(x_i)], dim=1) with every LayerNorm writing (and its backward reading) its token slice of the fused
    buffer in place: the closing LayerNorms of two encoders + ConcatFusion (modules/mixer.py:161, modules/fusion.py:117)."""

    @staticmethod
    def forward(ctx, n, *tensors):
        xs, ws, bs = list(tensors[:n]), list(tensors[n:2 * n]), list(tensors[2 * n:3 * n])
        ctx.save_for_backward(*xs, *ws)
        ctx.n, ctx.params = n, tuple(ws) + tuple(bs)
        return _O.layernorm_concat_fwd(xs, ws, bs)

    @staticmethod
    def backward(ctx, g):
        n = ctx.n
        saved = ctx.saved_tensors
        xs, ws = list(saved[:n]), list(saved[n:2 * n])
        direct = _direct(ctx.params)
        dxs, dws, dbs = _O.layernorm_concat_bwd(g.contiguous(), xs, ws, direct)
        if direct is not None:
            _notify(ctx.params)
            dws, dbs = [None] * n, [None] * n
        return (None, *dxs, *dws, *dbs)


class _Linear(Function):
    @staticmethod
    def forward(ctx, x, w, bias, wb, act, precision, p, seed):
        y = _O.linear_fwd(x, w, wb, bias, act, precision, p, seed)
        ctx.save_for_backward(x, w, wb, y if act == ACT_RELU else None)
        ctx.act, ctx.precision, ctx.has_bias, ctx.p, ctx.seed = act, precision, bias is not None, p, seed
        ctx.params = (w, bias)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, wb, y = ctx.saved_tensors
        direct = _direct(ctx.params) if ctx.has_bias else None
        dx, dw, db = _O.linear_bwd(dy.contiguous(), x, y, w, wb, ctx.act, ctx.needs_input_grad[0], ctx.precision, ctx.p,
                                   ctx.seed, direct)
        if direct is not None:
            _notify(ctx.params)
            dw = db = None
        return (dx if ctx.needs_input_grad[0] else None), dw, (db if ctx.has_bias else None), None, None, None, None, None


class _PatchEmbed(Function):
    """Conv2d(k = stride = patch) + 'b c h w -> b (h w) c' as one GEMM whose image-side operand is gathered from the pixels
    inside the kernel; the backward's weight-gradient GEMM gathers from the same image again (nothing but the input batch
    is kept), there is no input gradient."""

    @staticmethod
    def forward(ctx, img, w, bias, wb, patch, precision):
        y = _O.patch_embed_fwd(img, w, wb, bias, patch, precision)
        ctx.save_for_backward(img, w)
        ctx.precision, ctx.patch, ctx.has_bias, ctx.params = precision, patch, bias is not None, (w, bias)
        return y

    @staticmethod
    def backward(ctx, dy):
        img, w = ctx.saved_tensors
        direct = _direct(ctx.params) if ctx.has_bias else None
        dw, db = _O.patch_embed_bwd(dy.contiguous(), img, w, ctx.patch, ctx.has_bias, ctx.precision, direct)
        if direct is not None:
            _notify(ctx.params)
            dw = db = None
        return None, dw, (db if ctx.has_bias else None), None, None, None


class _Concat(Function):
    @staticmethod
    def forward(ctx, *xs):
        ctx.sizes = [x.shape[1] for x in xs]
        return _O.concat_tokens(list(xs))

    @staticmethod
    def backward(ctx, g):
        return tuple(_O.split_tokens(g.contiguous(), ctx.sizes))


class _Add(Function):
    @staticmethod
    def forward(ctx, a, b):
        return _O.add(a, b)

    @staticmethod
    def backward(ctx, g):
        return g, g


class _Fuse2(Function):
    """MaxFusion (mode 1) / two-input MeanFusion (mode 2), reference modules/fusion.py:190-204, 258-272."""

    @staticmethod
    def forward(ctx, a, b, mode):
        ctx.mode = mode
        if mode == 1:
            ctx.save_for_backward(a, b)
        return _O.fuse2_fwd(a, b, mode)

    @staticmethod
    def backward(ctx, g):
        if ctx.mode == 1:
            a, b = ctx.saved_tensors
            da, db = _O.fuse2_max_bwd(a, b, g.contiguous())
            return da, db, None
        h = g * 0.5
        return h, h, None


class _Gate(Function):
    """z tanh(h1) + (1 - z) tanh(h2), z = sigmoid(zh): the gate of BiModalGatedUnit (reference modules/fusion.py:16-23)."""

    @staticmethod
    def forward(ctx, h1, h2, zh):
        ctx.save_for_backward(h1, h2, zh)
        return _O.gate_fwd(h1, h2, zh)

    @staticmethod
    def backward(ctx, g):
        h1, h2, zh = ctx.saved_tensors
        return _O.gate_bwd(h1, h2, zh, g.contiguous())


class _MeanPool(Function):
    @staticmethod
    def forward(ctx, x):
        ctx.n = x.shape[1]
        return _O.mean_pool_fwd(x)

    @staticmethod
    def backward(ctx, g):
        return _O.mean_pool_bwd(g.contiguous(), ctx.n)


class _HeadsLoss(Function):
    """(losses[4], logits[3,B,K], preds) = heads+loss; only losses[0] (the weighted total) is differentiable."""

    @staticmethod
    def forward(ctx, labels, pos_weight, head_weight, loss_kind, n, slices, *tensors):
        toks, ws, bs = list(tensors[:n]), list(tensors[n:2 * n]), list(tensors[2 * n:3 * n])
        st = None if slices is None else [s for s, _ in slices]
        ln = None if slices is None else [l for _, l in slices]
        losses, logits, preds = _O.heads_loss_fwd(toks, ws, bs, labels, pos_weight, list(head_weight), loss_kind, st, ln)
        ctx.save_for_backward(labels, pos_weight, logits, *tensors)
        ctx.head_weight, ctx.loss_kind, ctx.n, ctx.slices = list(head_weight), loss_kind, n, (st, ln)
        ctx.params = tuple(ws) + tuple(bs)
        ctx.mark_non_differentiable(logits, preds)
        ctx.tok_shapes = [t.shape for t in toks]
        return losses, logits, preds

    @staticmethod
    def backward(ctx, dlosses, _dlogits, _dpreds):
        labels, pos_weight, logits, *tensors = ctx.saved_tensors
        n = ctx.n
        toks, ws, bs = list(tensors[:n]), list(tensors[n:2 * n]), list(tensors[2 * n:3 * n])
        # d(total)/d(.) scaled by the incoming gradient of losses[0]; per-head losses[1:] are reporting-only outputs
        # (passed as a device scalar: no host sync, CUDA-graph capturable)
        direct = _direct(ctx.params)
        dt, dw, db = _O.heads_loss_bwd(toks, ws, bs, labels, pos_weight, ctx.head_weight, ctx.loss_kind, logits, 1.0,
                                       dlosses.contiguous(), direct, ctx.slices[0], ctx.slices[1])
        # heads that pool slices of ONE base tensor share its gradient: returned once (autograd sums the input slots)
        dt = [g.reshape(s) if g.numel() else None for g, s in zip(dt, ctx.tok_shapes)]
        if direct is not None:
            _notify(ctx.params)
            dw, db = [None] * n, [None] * n
        return (None, None, None, None, None, None, *dt, *dw, *db)


# ---------------------------------------------------------------------------------------------------- public API
_DROP_CALLS = 0


def next_dropout_seed() -> int:
    """A fresh 63-bit seed per dropout-carrying op call: deterministic after torch.manual_seed, different per rank."""
    global _DROP_CALLS
    _DROP_CALLS += 1
    rank = int(os.environ.get("RANK", "0"))
    z = (torch.initial_seed() * 0x9E3779B97F4A7C15 + _DROP_CALLS * 0xD1B54A32D192ED03 + rank * 0x94D049BB133111EB)
    return z & 0x7FFFFFFFFFFFFFFF


def _drop_args(p: float, seed):
    p = float(p)
    if p <= 0.0:
        return 0.0, 0
    return p, int(next_dropout_seed() if seed is None else seed)


class _Bf16WeightCache:
    """bf16 operand copies of the GEMM weights (fp32 masters stay in the reference's [out, in] state-dict layout).

    An entry is stale when its parameter's storage pointer or autograd version changed (load_state_dict, torch optimizers and
    any other in-place update of the parameter bump the version) or when ``invalidate_bf16_weights()`` was called since -
    FusedAdam.step() does that, because it updates the flat buffer the parameters are views of, which does NOT touch the
    parameters' own version counters.  Writes through ``.data`` are invisible: call ``invalidate_bf16_weights()`` after them.
    The first stale lookup refreshes EVERY registered copy in one launch (m2b200_cast_bf16_multi) - after an optimiser step
    the weights all changed together - instead of one small launch per matrix and forward."""

    def __init__(self):
        self.entries = {}          # id(param) -> [param weakref, rows, cols, ld, dst, data_ptr, version, epoch]
        self.epoch = 0
        self.table = None
        self.table_key = None

    def get(self, param: torch.Tensor, rows: int, cols: int, ld: int) -> torch.Tensor:
        import weakref
        e = self.entries.get(id(param))
        if e is not None and (e[0]() is not param or e[1:4] != [rows, cols, ld] or e[4].device != param.device):
            e = None
        if e is None:
            if not (param.is_cuda and param.dtype == torch.float32 and param.is_contiguous()):
                raise RuntimeError("bf16 weight cache expects contiguous float32 CUDA parameters")
            dst = torch.empty(rows, ld, dtype=torch.bfloat16, device=param.device)
            key = id(param)
            e = [weakref.ref(param, lambda _r, k=key: self.entries.pop(k, None)), rows, cols, ld, dst, 0, -1, -1]
            self.entries[key] = e
        if e[5] != param.data_ptr() or e[6] != param._version or e[7] != self.epoch:
            self._refresh(param.device)
        return e[4]

    def _refresh(self, device) -> None:
        live = []
        for e in list(self.entries.values()):
            p = e[0]()
            if p is not None and p.device == device:
                live.append((e, p))
        key = tuple((p.data_ptr(), e[4].data_ptr()) for e, p in live)
        if key != self.table_key or self.table is None or self.table.device != device:
            rows = [[p.data_ptr(), e[4].data_ptr(), e[1], e[2], e[2], e[3]] for e, p in live]
            self.table = torch.tensor(rows, dtype=torch.int64).to(device)
            self.table_key = key
        with torch.no_grad():
            ops.cast_bf16_multi(self.table, len(live))
        for e, p in live:
            e[5], e[6], e[7] = p.data_ptr(), p._version, self.epoch


_WCACHE = _Bf16WeightCache()


def invalidate_bf16_weights() -> None:
    """Mark every cached bf16 weight copy stale (the next forward refreshes them all in one launch)."""
    _WCACHE.epoch += 1


def refresh_bf16_weights() -> None:
    """Re-cast every registered bf16 weight copy NOW (one launch per device) instead of at the next stale lookup.  A graphed
    step calls this right behind its optimizer update so that no forward graph depends on having captured the refresh."""
    devices = {e[4].device for e in _WCACHE.entries.values() if e[0]() is not None}
    for dev in devices:
        _WCACHE._refresh(dev)


def ensure_bf16_weights_fresh() -> None:
    """Refresh the bf16 weight copies on the CURRENT stream if any is stale.  Called before the forward forks onto a second
    stream (models/base.py): a lazy refresh issued by whichever branch looks a weight up first would race with the other."""
    stale = set()
    for e in _WCACHE.entries.values():
        p = e[0]()
        if p is not None and (e[5] != p.data_ptr() or e[6] != p._version or e[7] != _WCACHE.epoch):
            stale.add(p.device)
    for dev in stale:
        _WCACHE._refresh(dev)


# The streams the current step's forward / backward run on when the two encoders are forked (models/base.py: the caller's
# stream and the second encoder's).  The data-parallel bucket allreduces wait for all of them (parallel.GradSync._launch):
# a bucket may hold gradients produced on either, whichever stream the last "gradient ready" report came from.
COMPUTE_STREAMS: List["torch.cuda.Stream"] = []


def bf16_weight(param: torch.Tensor, rows: Optional[int] = None, cols: Optional[int] = None) -> torch.Tensor:
    """bf16 [rows][up8(cols)] operand copy of a weight parameter viewed as [rows][cols] (default: its first axis x the rest)."""
    rows = param.shape[0] if rows is None else rows
    cols = param.numel() // rows if cols is None else cols
    return _WCACHE.get(param, rows, cols, _up8(cols))


def token_mix(x, ln_w, ln_b, w1, b1, w2, b2, precision, dropout_p: float = 0.0, seed=None) -> torch.Tensor:
    p, seed = _drop_args(dropout_p, seed)
    return _TokenMix.apply(x, ln_w, ln_b, w1, b1, w2, b2, precision_code(precision), p, seed)


def channel_mix(u, ln_w, ln_b, w1, b1, w2, b2, precision, w1b=None, w2b=None, dropout_p: float = 0.0,
                seed=None) -> torch.Tensor:
    prec = precision_code(precision)
    p, seed = _drop_args(dropout_p, seed)
    if prec == BF16 and (w1b is None or w2b is None):
        if isinstance(w1, torch.nn.Parameter) and isinstance(w2, torch.nn.Parameter) and w1.shape[1] % 8 == 0:
            w1b, w2b = bf16_weight(w1), bf16_weight(w2)          # cached, refreshed together after an optimiser step
        else:
            with torch.no_grad():
                w1b = _O.cast_bf16(w1, w1.shape[1])
                w2b = _O.cast_bf16(w2, _up8(w2.shape[1]))
    return _ChannelMix.apply(u, ln_w, ln_b, w1, b1, w2, b2, w1b, w2b, prec, p, seed)


def layer_norm_concat(xs: Sequence[torch.Tensor], ws: Sequence[torch.Tensor], bs: Sequence[torch.Tensor]) -> torch.Tensor:
    """cat([LayerNorm_i(xs[i])], dim=1) without the copy: each LayerNorm writes its token slice of the result."""
    return _LayerNormConcat.apply(len(xs), *xs, *ws, *bs)


def layer_norm(x, w, b) -> torch.Tensor:
    return _LayerNorm.apply(x, w, b)


def linear(x, w, bias=None, act: int = ACT_NONE, precision="bf16", wb=None, dropout_p: float = 0.0, seed=None) -> torch.Tensor:
    prec = precision_code(precision)
    p, seed = _drop_args(dropout_p, seed)
    if prec == BF16 and wb is None:
        if isinstance(w, torch.nn.Parameter):
            wb = bf16_weight(w)
        else:
            with torch.no_grad():
                wb = _O.cast_bf16(w.reshape(w.shape[0], -1), _up8(w[0].numel()))
    return _Linear.apply(x, w.reshape(w.shape[0], -1), bias, wb, act, prec, p, seed)


def patch_embed(img, conv_w, conv_b, patch: int, precision="bf16") -> torch.Tensor:
    """Conv2d(k = stride = patch) + 'b c h w -> b (h w) c' as gather + GEMM (reference modules/mixer.py:143-146)."""
    prec = precision_code(precision)
    wb = None
    if prec == BF16:
        if isinstance(conv_w, torch.nn.Parameter):
            wb = bf16_weight(conv_w)
        else:
            with torch.no_grad():
                wb = _O.cast_bf16(conv_w.reshape(conv_w.shape[0], -1), _up8(conv_w[0].numel()))
    return _PatchEmbed.apply(img, conv_w, conv_b, wb, patch, prec)


def mean_pool(x) -> torch.Tensor:
    """[B, ..., D] -> [B, D]: mean over every token axis."""
    return _MeanPool.apply(x.reshape(x.shape[0], -1, x.shape[-1]))


def concat_tokens(*xs) -> torch.Tensor:
    return _Concat.apply(*xs)


def gate(h1, h2, zh) -> torch.Tensor:
    return _Gate.apply(h1, h2, zh)


def fuse_max(a, b) -> torch.Tensor:
    return _Fuse2.apply(a, b, 1)


def fuse_mean(a, b) -> torch.Tensor:
    return _Fuse2.apply(a, b, 2)


def add(a, b) -> torch.Tensor:
    return _Add.apply(a, b)


def heads_loss(toks: Sequence[torch.Tensor], ws: Sequence[torch.Tensor], bs: Sequence[torch.Tensor], labels,
               head_weight: Sequence[float], loss_kind: int = 0, pos_weight: Optional[torch.Tensor] = None,
               slices: Optional[Sequence[Tuple[int, int]]] = None):
    """Returns (losses[4] = total, L_0, L_1, L_2; logits [3,B,K]; preds).  ``slices[h] = (start, len)``: head h pools that
    token range of toks[h] in place (per-modality heads on the fused-token buffer of ``layer_norm_concat``)."""
    n = len(toks)
    if slices is not None and any(t.dim() == 3 and l * t.shape[2] >= 16384 for t, (_, l) in zip(toks, slices)):
        toks = [t[:, s0:s0 + l] for t, (s0, l) in zip(toks, slices)]     # wide token sets: pooled by the batch-parallel kernel below
        slices = None
    # The heads kernels give one WARP a sample (mean-pool, Linear, loss in one pass): right for the shipped configs (4-49
    # tokens of 32-128 floats).  With hundreds of wide tokens per sample and a small batch (the Scaled config: 392 x 768
    # floats per sample, batch 64 = 64 warps on the whole GPU) the pooling is done first by the batch x dim parallel kernel.
    toks = [mean_pool(t).unsqueeze(1) if (t.dim() == 3 and t.shape[1] * t.shape[2] >= 16384) else t for t in toks]
    if slices is not None:
        slices = tuple((int(s0), int(l)) for s0, l in slices)
    return _HeadsLoss.apply(labels, pos_weight, tuple(float(h) for h in head_weight), loss_kind, n, slices, *toks, *ws, *bs)
