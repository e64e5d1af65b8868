"""Autograd wiring of the m2b200 ops: each forward op is paired with its hand-written backward kernel sequence.

Backward kernels RECOMPUTE LayerNorm and GELU from the saved block input instead of storing activations
(north_star): the only tensors saved per Mixer block are its two fp32 inputs (x for token mixing, u for
channel mixing) - no [tokens x channel_dim] hidden tensor is ever kept alive between forward and backward.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

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


class _TokenMix(Function):
    @staticmethod
    def forward(ctx, x, ln_w, ln_b, w1, b1, w2, b2, precision):
        u = _O.token_mix_fwd(x, ln_w, ln_b, w1, b1, w2, b2, precision)
        ctx.save_for_backward(x, ln_w, ln_b, w1, b1, w2)
        ctx.precision = precision
        return u

    @staticmethod
    def backward(ctx, du):
        x, ln_w, ln_b, w1, b1, w2 = ctx.saved_tensors
        dx, dln_w, dln_b, dw1, db1, dw2, db2 = _O.token_mix_bwd(du.contiguous(), x, ln_w, ln_b, w1, b1, w2, ctx.precision)
        return dx, dln_w, dln_b, dw1, db1, dw2, db2, None


class _ChannelMix(Function):
    @staticmethod
    def forward(ctx, u, ln_w, ln_b, w1, b1, w2, b2, w1b, w2b, precision):
        y = _O.channel_mix_fwd(u, ln_w, ln_b, w1, b1, w2, b2, w1b, w2b, precision)
        ctx.save_for_backward(u, ln_w, ln_b, w1, b1, w2, w1b, w2b)
        ctx.precision = precision
        return y

    @staticmethod
    def backward(ctx, dy):
        u, ln_w, ln_b, w1, b1, w2, w1b, w2b = ctx.saved_tensors
        du, dln_w, dln_b, dw1, db1, dw2, db2 = _O.channel_mix_bwd(dy.contiguous(), u, ln_w, ln_b, w1, b1, w2, w1b, w2b,
                                                                 ctx.precision)
        return du, dln_w, dln_b, dw1, db1, dw2, db2, None, None, None


class _LayerNorm(Function):
    @staticmethod
    def forward(ctx, x, w, b):
        ctx.save_for_backward(x, w)
        return _O.layernorm_fwd(x, w, b)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        return _O.layernorm_bwd(dy.contiguous(), x, w)


class _Linear(Function):
    @staticmethod
    def forward(ctx, x, w, bias, wb, act, precision):
        y = _O.linear_fwd(x, w, wb, bias, act, precision)
        ctx.save_for_backward(x, w, wb, y if act == ACT_RELU else None)
        ctx.act, ctx.precision, ctx.has_bias = act, precision, bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, wb, y = ctx.saved_tensors
        dx, dw, db = _O.linear_bwd(dy.contiguous(), x, y, w, wb, ctx.act, ctx.needs_input_grad[0], ctx.precision)
        return (dx if ctx.needs_input_grad[0] else None), dw, (db if ctx.has_bias else None), None, None, None


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
    def forward(ctx, labels, pos_weight, head_weight, loss_kind, n, *tensors):
        toks, ws, bs = list(tensors[:n]), list(tensors[n:2 * n]), list(tensors[2 * n:3 * n])
        losses, logits, preds = _O.heads_loss_fwd(toks, ws, bs, labels, pos_weight, list(head_weight), loss_kind)
        ctx.save_for_backward(labels, pos_weight, logits, *tensors)
        ctx.head_weight, ctx.loss_kind, ctx.n = list(head_weight), loss_kind, n
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
        dt, dw, db = _O.heads_loss_bwd(toks, ws, bs, labels, pos_weight, ctx.head_weight, ctx.loss_kind, logits, 1.0,
                                       dlosses.contiguous())
        dt = [g.reshape(s) for g, s in zip(dt, ctx.tok_shapes)]
        return (None, None, None, None, None, *dt, *dw, *db)


# ---------------------------------------------------------------------------------------------------- public API
def token_mix(x, ln_w, ln_b, w1, b1, w2, b2, precision) -> torch.Tensor:
    return _TokenMix.apply(x, ln_w, ln_b, w1, b1, w2, b2, precision_code(precision))


def channel_mix(u, ln_w, ln_b, w1, b1, w2, b2, precision, w1b=None, w2b=None) -> torch.Tensor:
    prec = precision_code(precision)
    if prec == BF16 and (w1b is None or w2b is None):
        with torch.no_grad():
            w1b = _O.cast_bf16(w1, w1.shape[1])
            w2b = _O.cast_bf16(w2, _up8(w2.shape[1]))
    return _ChannelMix.apply(u, ln_w, ln_b, w1, b1, w2, b2, w1b, w2b, prec)


def layer_norm(x, w, b) -> torch.Tensor:
    return _LayerNorm.apply(x, w, b)


def linear(x, w, bias=None, act: int = ACT_NONE, precision="bf16", wb=None) -> torch.Tensor:
    prec = precision_code(precision)
    if prec == BF16 and wb is None:
        with torch.no_grad():
            wb = _O.cast_bf16(w.reshape(w.shape[0], -1), _up8(w[0].numel()))
    return _Linear.apply(x, w.reshape(w.shape[0], -1), bias, wb, act, prec)


def patch_embed(img, conv_w, conv_b, patch: int, precision="bf16") -> torch.Tensor:
    """Conv2d(k = stride = patch) + 'b c h w -> b (h w) c' as gather + GEMM (reference modules/mixer.py:143-146)."""
    with torch.no_grad():
        cols = _O.patch_gather(img, patch)     # the input image needs no gradient
    return linear(cols, conv_w, conv_b, ACT_NONE, precision)


def mean_pool(x) -> torch.Tensor:
    """[B, ..., D] -> [B, D]: mean over every token axis."""
    return _MeanPool.apply(x.reshape(x.shape[0], -1, x.shape[-1]))


def concat_tokens(*xs) -> torch.Tensor:
    return _Concat.apply(*xs)


def add(a, b) -> torch.Tensor:
    return _Add.apply(a, b)


def heads_loss(toks: Sequence[torch.Tensor], ws: Sequence[torch.Tensor], bs: Sequence[torch.Tensor], labels,
               head_weight: Sequence[float], loss_kind: int = 0, pos_weight: Optional[torch.Tensor] = None):
    """Returns (losses[4] = total, L_0, L_1, L_2; logits [3,B,K]; preds)."""
    n = len(toks)
    return _HeadsLoss.apply(labels, pos_weight, tuple(float(h) for h in head_weight), loss_kind, n, *toks, *ws, *bs)
