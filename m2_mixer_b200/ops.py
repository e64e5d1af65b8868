"""torch.library registration of the m2b200 C-ABI kernels (namespace ``m2b200``), CUDA-only.

Thin by design: each op checks device / dtype / contiguity, allocates outputs and workspace with torch's caching
allocator, passes ``torch.cuda.current_stream()`` and turns a non-zero status into an exception.  There is no CPU
implementation and no composite fallback: calling an op with CPU tensors raises NotImplementedError from the
dispatcher, a missing library raises at import of this module's first use.

Autograd wiring lives in ``functional.py`` (``torch.autograd.Function`` wrappers that call the *_fwd / *_bwd ops).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import BF16, FP32, check

_LIB = torch.library.Library("m2b200", "DEF")


def _L():
    return _lib.load()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"m2b200: {name} must be a CUDA tensor (there is no CPU path)")
    if t.dtype != torch.float32:
        raise RuntimeError(f"m2b200: {name} must be float32, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def _ws(nbytes: int, like: torch.Tensor) -> Tuple[Optional[torch.Tensor], Optional[int]]:
    if nbytes == 0:
        return None, None
    w = torch.empty(nbytes, dtype=torch.uint8, device=like.device)
    return w, w.data_ptr()


def _grad_dst(grads, i: int, like: torch.Tensor) -> torch.Tensor:
    """Destination of a parameter gradient: the caller's running buffer (accumulated into) or fresh zeros."""
    if grads is not None and grads[i] is not None:
        g = grads[i]
        if not (g.is_cuda and g.dtype == torch.float32 and g.is_contiguous() and g.numel() == like.numel()):
            raise RuntimeError("m2b200: gradient destination must be a contiguous float32 CUDA tensor of the parameter's size")
        return g
    return torch.zeros_like(like, memory_format=torch.contiguous_format)


def _define(name: str, schema: str, fn):
    _LIB.define(f"{name}{schema}")
    _LIB.impl(name, fn, "CUDA")


# ------------------------------------------------------------------------------------------------ bf16 cache
def cast_bf16(w: torch.Tensor, ld: int) -> torch.Tensor:
    w = _f32c(w, "w")
    rows, cols = w.shape
    out = torch.empty(rows, ld, dtype=torch.bfloat16, device=w.device)
    check(_L().m2b200_cast_bf16(w.data_ptr(), cols, out.data_ptr(), ld, rows, cols, _stream()), "cast_bf16")
    return out


_define("cast_bf16", "(Tensor w, int ld) -> Tensor", cast_bf16)


def cast_bf16_multi(table: torch.Tensor, n: int) -> None:
    """table: int64 CUDA tensor [n, 6] of {src ptr, dst ptr, rows, cols, lds, ldd}: every bf16 weight copy in one launch."""
    if not (table.is_cuda and table.dtype == torch.int64 and table.is_contiguous() and table.numel() >= 6 * n):
        raise RuntimeError("m2b200::cast_bf16_multi expects a contiguous int64 CUDA table [n, 6]")
    check(_L().m2b200_cast_bf16_multi(table.data_ptr(), n, _stream()), "cast_bf16_multi")


# ------------------------------------------------------------------------------------------------ token mixing
def token_mix_fwd(x, ln_w, ln_b, w1, b1, w2, b2, precision: int, dropout_p: float = 0.0, seed: int = 0):
    x = _f32c(x, "x")
    B, N, D = x.shape
    T = w1.shape[0]
    u = torch.empty_like(x)
    nbytes = _L().m2b200_token_mix_fwd_workspace_bytes(B, N, D, T, precision)
    ws, wsp = _ws(nbytes, x)
    check(_L().m2b200_token_mix_fwd(x.data_ptr(), _f32c(ln_w, "ln_w").data_ptr(), _f32c(ln_b, "ln_b").data_ptr(),
                                    _f32c(w1, "w1").data_ptr(), _f32c(b1, "b1").data_ptr(), _f32c(w2, "w2").data_ptr(),
                                    _f32c(b2, "b2").data_ptr(), u.data_ptr(), B, N, D, T, precision, dropout_p, seed,
                                    wsp, nbytes, _stream()),
          "token_mix_fwd")
    return u


def token_mix_bwd(du, x, ln_w, ln_b, w1, b1, w2, precision: int, dropout_p: float = 0.0, seed: int = 0, grads=None):
    du, x = _f32c(du, "du"), _f32c(x, "x")
    B, N, D = x.shape
    T = w1.shape[0]
    dx = torch.empty_like(x)
    dln_w, dln_b, dw1, db1, dw2 = (_grad_dst(grads, i, t) for i, t in enumerate((ln_w, ln_b, w1, b1, w2)))
    db2 = _grad_dst(grads, 5, w2[:, 0])
    nbytes = _L().m2b200_token_mix_bwd_workspace_bytes(B, N, D, T, precision)
    ws, wsp = _ws(nbytes, x)
    check(_L().m2b200_token_mix_bwd(du.data_ptr(), x.data_ptr(), _f32c(ln_w, "ln_w").data_ptr(),
                                    _f32c(ln_b, "ln_b").data_ptr(), _f32c(w1, "w1").data_ptr(), _f32c(b1, "b1").data_ptr(),
                                    _f32c(w2, "w2").data_ptr(), dx.data_ptr(), dln_w.data_ptr(), dln_b.data_ptr(),
                                    dw1.data_ptr(), db1.data_ptr(), dw2.data_ptr(), db2.data_ptr(), B, N, D, T, precision,
                                    dropout_p, seed, wsp, nbytes, _stream()), "token_mix_bwd")
    return dx, dln_w, dln_b, dw1, db1, dw2, db2


_define("token_mix_fwd", "(Tensor x, Tensor ln_w, Tensor ln_b, Tensor w1, Tensor b1, Tensor w2, Tensor b2, int precision, "
        "float dropout_p=0.0, int seed=0) -> Tensor",
        token_mix_fwd)
_define("token_mix_bwd", "(Tensor du, Tensor x, Tensor ln_w, Tensor ln_b, Tensor w1, Tensor b1, Tensor w2, int precision, "
        "float dropout_p=0.0, int seed=0, Tensor?[]? grads=None) -> "
        "(Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor)", token_mix_bwd)


# ------------------------------------------------------------------------------------------------ channel mixing
def channel_mix_fwd(u, ln_w, ln_b, w1, b1, w2, b2, w1b, w2b, precision: int, dropout_p: float = 0.0, seed: int = 0):
    u = _f32c(u, "u")
    D = u.shape[-1]
    M = u.numel() // D
    Cc = w1.shape[0]
    y = torch.empty_like(u)
    nbytes = _L().m2b200_channel_mix_workspace_bytes(M, D, Cc, precision, 0)
    ws, wsp = _ws(nbytes, u)
    ldw2 = 0 if w2b is None else w2b.shape[1]
    check(_L().m2b200_channel_mix_fwd(u.data_ptr(), _f32c(ln_w, "ln_w").data_ptr(), _f32c(ln_b, "ln_b").data_ptr(),
                                      _f32c(w1, "w1").data_ptr(), _f32c(b1, "b1").data_ptr(), _f32c(w2, "w2").data_ptr(),
                                      _f32c(b2, "b2").data_ptr(), _ptr(w1b), _ptr(w2b), ldw2, y.data_ptr(), M, D, Cc,
                                      precision, dropout_p, seed, wsp, nbytes, _stream()), "channel_mix_fwd")
    return y


def channel_mix_bwd(dy, u, ln_w, ln_b, w1, b1, w2, w1b, w2b, precision: int, dropout_p: float = 0.0, seed: int = 0,
                    grads=None):
    dy, u = _f32c(dy, "dy"), _f32c(u, "u")
    D = u.shape[-1]
    M = u.numel() // D
    Cc = w1.shape[0]
    du = torch.empty_like(u)
    dln_w, dln_b, dw1, db1, dw2 = (_grad_dst(grads, i, t) for i, t in enumerate((ln_w, ln_b, w1, b1, w2)))
    db2 = _grad_dst(grads, 5, ln_w)
    nbytes = _L().m2b200_channel_mix_workspace_bytes(M, D, Cc, precision, 1)
    ws, wsp = _ws(nbytes, u)
    ldw2 = 0 if w2b is None else w2b.shape[1]
    check(_L().m2b200_channel_mix_bwd(dy.data_ptr(), u.data_ptr(), _f32c(ln_w, "ln_w").data_ptr(),
                                      _f32c(ln_b, "ln_b").data_ptr(), _f32c(w1, "w1").data_ptr(), _f32c(b1, "b1").data_ptr(),
                                      _f32c(w2, "w2").data_ptr(), _ptr(w1b), _ptr(w2b), ldw2, du.data_ptr(),
                                      dln_w.data_ptr(), dln_b.data_ptr(), dw1.data_ptr(), db1.data_ptr(), dw2.data_ptr(),
                                      db2.data_ptr(), M, D, Cc, precision, dropout_p, seed, wsp, nbytes, _stream()),
          "channel_mix_bwd")
    return du, dln_w, dln_b, dw1, db1, dw2, db2


_define("channel_mix_fwd", "(Tensor u, Tensor ln_w, Tensor ln_b, Tensor w1, Tensor b1, Tensor w2, Tensor b2, Tensor? w1b, "
        "Tensor? w2b, int precision, float dropout_p=0.0, int seed=0) -> Tensor", channel_mix_fwd)
_define("channel_mix_bwd", "(Tensor dy, Tensor u, Tensor ln_w, Tensor ln_b, Tensor w1, Tensor b1, Tensor w2, Tensor? w1b, "
        "Tensor? w2b, int precision, float dropout_p=0.0, int seed=0, Tensor?[]? grads=None) -> "
        "(Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor)", channel_mix_bwd)


# ------------------------------------------------------------------------------------------------ LayerNorm
def layernorm_fwd(x, w, b):
    x = _f32c(x, "x")
    D = x.shape[-1]
    rows = x.numel() // D
    out = torch.empty_like(x)
    check(_L().m2b200_layernorm_fwd(x.data_ptr(), _f32c(w, "w").data_ptr(), _f32c(b, "b").data_ptr(), out.data_ptr(), 1,
                                    rows, D, 0, _stream()), "layernorm_fwd")
    return out


def layernorm_bwd(dy, x, w, grads=None):
    dy, x = _f32c(dy, "dy"), _f32c(x, "x")
    D = x.shape[-1]
    rows = x.numel() // D
    dx = torch.empty_like(x)
    dw, db = _grad_dst(grads, 0, w), _grad_dst(grads, 1, w)
    check(_L().m2b200_layernorm_bwd(dy.data_ptr(), 0, x.data_ptr(), _f32c(w, "w").data_ptr(), None, dx.data_ptr(),
                                    dw.data_ptr(), db.data_ptr(), 1, rows, D, _stream()), "layernorm_bwd")
    return dx, dw, db


def layernorm_concat_fwd(xs: Sequence[torch.Tensor], ws: Sequence[torch.Tensor], bs: Sequence[torch.Tensor]):
    """fused[:, off_i : off_i + N_i, :] = LayerNorm_i(xs[i]): every encoder normalises straight into its slice of the
    fused-token buffer (ConcatFusion(dim=1) of the closing LayerNorms without a copy, modules/fusion.py:117)."""
    xs = [_f32c(x, "x") for x in xs]
    B, D = xs[0].shape[0], xs[0].shape[-1]
    if any(x.dim() != 3 or x.shape[0] != B or x.shape[2] != D for x in xs):
        raise ValueError("layernorm_concat expects [B, N_i, D] token tensors with equal B and D")
    ntot = sum(x.shape[1] for x in xs)
    out = torch.empty(B, ntot, D, dtype=torch.float32, device=xs[0].device)
    off = 0
    for x, w, b in zip(xs, ws, bs):
        check(_L().m2b200_layernorm_fwd(x.data_ptr(), _f32c(w, "w").data_ptr(), _f32c(b, "b").data_ptr(),
                                        out.data_ptr() + off * D * 4, B, x.shape[1], D, ntot * D, _stream()), "layernorm_concat_fwd")
        off += x.shape[1]
    return out


def layernorm_concat_bwd(g, xs: Sequence[torch.Tensor], ws: Sequence[torch.Tensor], grads=None):
    """Backward of layernorm_concat_fwd: every LayerNorm backward reads its slice of the fused-token gradient in place."""
    g = _f32c(g, "g")
    xs = [_f32c(x, "x") for x in xs]
    B, ntot, D = g.shape
    n = len(xs)
    dxs, dws, dbs, off = [], [], [], 0
    for i, (x, w) in enumerate(zip(xs, ws)):
        dx = torch.empty_like(x)
        dw, db = _grad_dst(grads, i, w), _grad_dst(grads, n + i, w)
        check(_L().m2b200_layernorm_bwd(g.data_ptr() + off * D * 4, ntot * D, x.data_ptr(), _f32c(w, "w").data_ptr(), None,
                                        dx.data_ptr(), dw.data_ptr(), db.data_ptr(), B, x.shape[1], D, _stream()),
              "layernorm_concat_bwd")
        off += x.shape[1]
        dxs.append(dx); dws.append(dw); dbs.append(db)
    return dxs, dws, dbs


_define("layernorm_fwd", "(Tensor x, Tensor w, Tensor b) -> Tensor", layernorm_fwd)
_define("layernorm_concat_fwd", "(Tensor[] xs, Tensor[] ws, Tensor[] bs) -> Tensor", layernorm_concat_fwd)
_define("layernorm_concat_bwd", "(Tensor g, Tensor[] xs, Tensor[] ws, Tensor?[]? grads=None) -> (Tensor[], Tensor[], Tensor[])",
        layernorm_concat_bwd)
_define("layernorm_bwd", "(Tensor dy, Tensor x, Tensor w, Tensor?[]? grads=None) -> (Tensor, Tensor, Tensor)", layernorm_bwd)


# ------------------------------------------------------------------------------------------------ linear / patches
def linear_fwd(x, w, wb, bias, act: int, precision: int, dropout_p: float = 0.0, seed: int = 0):
    x = _f32c(x, "x")
    K = x.shape[-1]
    M = x.numel() // K
    N = w.shape[0]
    y = torch.empty(*x.shape[:-1], N, dtype=torch.float32, device=x.device)
    nbytes = _L().m2b200_linear_workspace_bytes(M, N, K, precision, 0)
    ws, wsp = _ws(nbytes, x)
    check(_L().m2b200_linear_fwd(x.data_ptr(), _f32c(w, "w").data_ptr(), _ptr(wb), 0 if wb is None else wb.shape[1],
                                 None if bias is None else _f32c(bias, "bias").data_ptr(), act, y.data_ptr(), M, N, K,
                                 precision, dropout_p, seed, wsp, nbytes, _stream()), "linear_fwd")
    return y


def linear_bwd(dy, x, y, w, wb, act: int, need_dx: bool, precision: int, dropout_p: float = 0.0, seed: int = 0, grads=None):
    x = _f32c(x, "x")
    dy = _f32c(dy, "dy")
    if act == _lib.ACT_RELU or dropout_p > 0.0:
        dy = dy.clone()   # masked in place by the kernel
    K = x.shape[-1]
    M = x.numel() // K
    N = w.shape[0]
    dx = torch.empty_like(x) if need_dx else None
    dw = _grad_dst(grads, 0, w)
    db = _grad_dst(grads, 1, w[:, 0])
    nbytes = _L().m2b200_linear_workspace_bytes(M, N, K, precision, 1)
    ws, wsp = _ws(nbytes, x)
    check(_L().m2b200_linear_bwd(dy.data_ptr(), x.data_ptr(), _ptr(y), _f32c(w, "w").data_ptr(), _ptr(wb),
                                 0 if wb is None else wb.shape[1], act, _ptr(dx), dw.data_ptr(), db.data_ptr(), M, N, K,
                                 precision, dropout_p, seed, wsp, nbytes, _stream()), "linear_bwd")
    if dx is None:
        dx = torch.empty(0, dtype=torch.float32, device=x.device)
    return dx, dw, db


def patch_gather(img, patch: int):
    img = _f32c(img, "img")
    B, cin, H, W = img.shape
    if H % patch or W % patch:
        raise AssertionError("Image dimensions must be divisible by the patch size.")
    cols = torch.empty(B, (H // patch) * (W // patch), cin * patch * patch, dtype=torch.float32, device=img.device)
    check(_L().m2b200_patch_gather(img.data_ptr(), cols.data_ptr(), B, cin, H, W, patch, _stream()), "patch_gather")
    return cols


_define("linear_fwd", "(Tensor x, Tensor w, Tensor? wb, Tensor? bias, int act, int precision, float dropout_p=0.0, int seed=0) "
        "-> Tensor", linear_fwd)
_define("linear_bwd", "(Tensor dy, Tensor x, Tensor? y, Tensor w, Tensor? wb, int act, bool need_dx, int precision, "
        "float dropout_p=0.0, int seed=0, Tensor?[]? grads=None) -> "
        "(Tensor, Tensor, Tensor)", linear_bwd)
def _img(img: torch.Tensor, precision: int) -> torch.Tensor:
    """Pixels as the kernels take them: contiguous CUDA fp32, or bf16 in BF16 mode (what DevicePrefetcher stages when the
    host batches are bf16 - the GEMM operand is the pixel rounded to bf16 either way)."""
    if img.dtype == torch.bfloat16 and precision == BF16 and img.is_cuda:
        return img if img.is_contiguous() else img.contiguous()
    return _f32c(img, "img")


def patch_embed_fwd(img, w, wb, bias, patch: int, precision: int):
    img = _img(img, precision)
    B, cin, H, W = img.shape
    if H % patch or W % patch:
        raise AssertionError("Image dimensions must be divisible by the patch size.")
    D = w.shape[0]
    n = (H // patch) * (W // patch)
    is_bf16 = int(img.dtype == torch.bfloat16)
    nbytes = _L().m2b200_patch_embed_workspace_bytes(img.data_ptr(), is_bf16, B, cin, H, W, patch, D, precision, 0)
    ws, wsp = _ws(nbytes, img)
    y = torch.empty(B, n, D, dtype=torch.float32, device=img.device)
    w2d = _f32c(w, "w").reshape(D, -1)
    check(_L().m2b200_patch_embed_fwd(img.data_ptr(), is_bf16, w2d.data_ptr(), _ptr(wb), 0 if wb is None else wb.shape[1],
                                      None if bias is None else _f32c(bias, "bias").data_ptr(), y.data_ptr(),
                                      B, cin, H, W, patch, D, precision, wsp, nbytes, _stream()), "patch_embed_fwd")
    return y


def patch_embed_bwd(dy, img, w, patch: int, has_bias: bool, precision: int, grads=None):
    dy = _f32c(dy, "dy")
    img = _img(img, precision)
    B, cin, H, W = img.shape
    D = dy.shape[-1]
    is_bf16 = int(img.dtype == torch.bfloat16)
    dw = _grad_dst(grads, 0, w)
    db = _grad_dst(grads, 1, w.reshape(D, -1)[:, 0]) if has_bias else None
    nbytes = _L().m2b200_patch_embed_workspace_bytes(img.data_ptr(), is_bf16, B, cin, H, W, patch, D, precision, 1)
    ws, wsp = _ws(nbytes, dy)
    check(_L().m2b200_patch_embed_bwd(dy.data_ptr(), img.data_ptr(), is_bf16, dw.data_ptr(), _ptr(db), B, cin, H, W, patch, D,
                                      precision, wsp, nbytes, _stream()), "patch_embed_bwd")
    return dw, (db if db is not None else torch.empty(0, dtype=torch.float32, device=dy.device))


_define("patch_embed_fwd", "(Tensor img, Tensor w, Tensor? wb, Tensor? bias, int patch, int precision) -> Tensor",
        patch_embed_fwd)
_define("patch_embed_bwd", "(Tensor dy, Tensor img, Tensor w, int patch, bool has_bias, int precision, Tensor?[]? grads=None) -> "
        "(Tensor, Tensor)", patch_embed_bwd)
_define("patch_gather", "(Tensor img, int patch) -> Tensor", patch_gather)


def dropout_mask(rows: int, cols: int, ld: int, dropout_p: float, seed: int, site: int, device="cuda") -> torch.Tensor:
    """Keep-mask (1/0) the kernels use for a dropout site (see include/m2b200.h); test helper."""
    out = torch.empty(rows, cols, dtype=torch.float32, device=device)
    check(_L().m2b200_dropout_mask(out.data_ptr(), rows, cols, ld, dropout_p, seed, site, _stream()), "dropout_mask")
    return out


def dropout_threshold(p: float) -> int:
    """round(p * 128) clamped to [1, 127] (0 when dropout is off): the kernels quantise the drop probability to 1/128."""
    return min(max(int(p * 128.0 + 0.5), 1), 127) if p > 0 else 0


def dropout_scale(p: float) -> float:
    """Inverse of the REALISED keep probability (csrc/common.cuh make_drop)."""
    return 128.0 / (128.0 - dropout_threshold(p))


# ------------------------------------------------------------------------------------------------ fusion
def concat_tokens(xs: Sequence[torch.Tensor]):
    xs = [_f32c(x, "x") for x in xs]
    B, D = xs[0].shape[0], xs[0].shape[-1]
    ntot = sum(x.shape[1] for x in xs)
    out = torch.empty(B, ntot, D, dtype=torch.float32, device=xs[0].device)
    off = 0
    for x in xs:
        per = x.shape[1] * D
        check(_L().m2b200_copy_tokens(x.data_ptr(), per, out.data_ptr() + off * 4, ntot * D, B, per, 0, _stream()),
              "concat_tokens")
        off += per
    return out


def split_tokens(g, sizes: Sequence[int]):
    g = _f32c(g, "g")
    B, ntot, D = g.shape
    outs, off = [], 0
    for n in sizes:
        o = torch.empty(B, n, D, dtype=torch.float32, device=g.device)
        check(_L().m2b200_copy_tokens(g.data_ptr() + off * 4, ntot * D, o.data_ptr(), n * D, B, n * D, 0, _stream()),
              "split_tokens")
        outs.append(o)
        off += n * D
    return outs


def add(a, b):
    a, b = _f32c(a, "a"), _f32c(b, "b")
    if a.shape != b.shape:
        raise RuntimeError("m2b200::add expects equal shapes")
    out = torch.empty_like(a)
    check(_L().m2b200_add(a.data_ptr(), b.data_ptr(), out.data_ptr(), a.numel(), _stream()), "add")
    return out


def mean_pool_fwd(x):
    x = _f32c(x, "x")
    B, N, D = x.shape
    out = torch.empty(B, D, dtype=torch.float32, device=x.device)
    check(_L().m2b200_mean_pool_fwd(x.data_ptr(), out.data_ptr(), B, N, D, _stream()), "mean_pool_fwd")
    return out


def mean_pool_bwd(dp, n: int):
    dp = _f32c(dp, "dp")
    B, D = dp.shape
    dx = torch.empty(B, n, D, dtype=torch.float32, device=dp.device)
    check(_L().m2b200_mean_pool_bwd(dp.data_ptr(), dx.data_ptr(), B, n, D, _stream()), "mean_pool_bwd")
    return dx


_define("mean_pool_fwd", "(Tensor x) -> Tensor", mean_pool_fwd)
_define("mean_pool_bwd", "(Tensor dp, int n) -> Tensor", mean_pool_bwd)
_define("concat_tokens", "(Tensor[] xs) -> Tensor", concat_tokens)
_define("split_tokens", "(Tensor g, int[] sizes) -> Tensor[]", split_tokens)
_define("add", "(Tensor a, Tensor b) -> Tensor", add)


def fuse2_fwd(a, b, mode: int):
    a, b = _f32c(a, "a"), _f32c(b, "b")
    if a.shape != b.shape:
        raise RuntimeError("m2b200::fuse2_fwd expects equal shapes")
    out = torch.empty_like(a)
    check(_L().m2b200_fuse2_fwd(a.data_ptr(), b.data_ptr(), out.data_ptr(), a.numel(), mode, _stream()), "fuse2_fwd")
    return out


def fuse2_max_bwd(a, b, g):
    a, b, g = _f32c(a, "a"), _f32c(b, "b"), _f32c(g, "g")
    da, db = torch.empty_like(a), torch.empty_like(a)
    check(_L().m2b200_fuse2_max_bwd(a.data_ptr(), b.data_ptr(), g.data_ptr(), da.data_ptr(), db.data_ptr(), a.numel(),
                                    _stream()), "fuse2_max_bwd")
    return da, db


def gate_fwd(h1, h2, zh):
    h1, h2, zh = _f32c(h1, "h1"), _f32c(h2, "h2"), _f32c(zh, "zh")
    if not (h1.shape == h2.shape == zh.shape):
        raise RuntimeError("m2b200::gate_fwd expects equal shapes")
    out = torch.empty_like(h1)
    check(_L().m2b200_gate_fwd(h1.data_ptr(), h2.data_ptr(), zh.data_ptr(), out.data_ptr(), h1.numel(), _stream()), "gate_fwd")
    return out


def gate_bwd(h1, h2, zh, g):
    h1, h2, zh, g = _f32c(h1, "h1"), _f32c(h2, "h2"), _f32c(zh, "zh"), _f32c(g, "g")
    d1, d2, dz = torch.empty_like(h1), torch.empty_like(h1), torch.empty_like(h1)
    check(_L().m2b200_gate_bwd(h1.data_ptr(), h2.data_ptr(), zh.data_ptr(), g.data_ptr(), d1.data_ptr(), d2.data_ptr(),
                               dz.data_ptr(), h1.numel(), _stream()), "gate_bwd")
    return d1, d2, dz


_define("gate_fwd", "(Tensor h1, Tensor h2, Tensor zh) -> Tensor", gate_fwd)
_define("gate_bwd", "(Tensor h1, Tensor h2, Tensor zh, Tensor g) -> (Tensor, Tensor, Tensor)", gate_bwd)
_define("fuse2_fwd", "(Tensor a, Tensor b, int mode) -> Tensor", fuse2_fwd)
_define("fuse2_max_bwd", "(Tensor a, Tensor b, Tensor g) -> (Tensor, Tensor)", fuse2_max_bwd)


# ------------------------------------------------------------------------------------------------ heads + loss
def _tok_slices(toks, tok_start, tok_len):
    """Per head (start, len) token range of its [B, N, D] base tensor (default: all of it)."""
    if tok_start is None:
        return [(0, t.shape[1]) for t in toks]
    sl = [(int(s), int(l)) for s, l in zip(tok_start, tok_len)]
    for t, (s0, l0) in zip(toks, sl):
        if s0 < 0 or l0 <= 0 or s0 + l0 > t.shape[1]:
            raise ValueError("token slice outside its base tensor")
    return sl


def _heads_args(toks, ws, bs, head_weight, slices=None):
    n = len(toks)
    arr_p = (C.c_void_p * 3)
    if slices is not None:
        tok = arr_p(*[t.data_ptr() + s0 * t.shape[2] * 4 for t, (s0, _) in zip(toks, slices)] + [None] * (3 - n))
        w = arr_p(*[t.data_ptr() for t in ws] + [None] * (3 - n))
        b = arr_p(*[t.data_ptr() for t in bs] + [None] * (3 - n))
        bstride = (C.c_int64 * 3)(*[t.shape[1] * t.shape[2] for t in toks] + [0] * (3 - n))
        ntok = (C.c_int * 3)(*[l0 for _, l0 in slices] + [0] * (3 - n))
        dim = (C.c_int * 3)(*[t.shape[2] for t in toks] + [0] * (3 - n))
        hw = (C.c_float * 3)(*list(head_weight) + [0.0] * (3 - n))
        return tok, bstride, ntok, dim, w, b, hw
    tok = arr_p(*[t.data_ptr() for t in toks] + [None] * (3 - n))
    w = arr_p(*[t.data_ptr() for t in ws] + [None] * (3 - n))
    b = arr_p(*[t.data_ptr() for t in bs] + [None] * (3 - n))
    bstride = (C.c_int64 * 3)(*[t.shape[1] * t.shape[2] for t in toks] + [0] * (3 - n))
    ntok = (C.c_int * 3)(*[t.shape[1] for t in toks] + [0] * (3 - n))
    dim = (C.c_int * 3)(*[t.shape[2] for t in toks] + [0] * (3 - n))
    hw = (C.c_float * 3)(*list(head_weight) + [0.0] * (3 - n))
    return tok, bstride, ntok, dim, w, b, hw


def _as_tokens(t: torch.Tensor) -> torch.Tensor:
    t = _f32c(t, "tokens")
    return t.reshape(t.shape[0], -1, t.shape[-1])


def heads_loss_fwd(toks: Sequence[torch.Tensor], ws: Sequence[torch.Tensor], bs: Sequence[torch.Tensor], labels,
                   pos_weight, head_weight: Sequence[float], loss_kind: int, tok_start=None, tok_len=None):
    """``tok_start`` / ``tok_len``: head h pools tokens [start, start + len) of toks[h] - the per-modality heads read their
    slices of the fused-token buffer in place (the same tensor may be given for several heads)."""
    toks = [_as_tokens(t) for t in toks]
    slices = _tok_slices(toks, tok_start, tok_len)
    ws = [_f32c(w, "w") for w in ws]
    bs = [_f32c(b, "b") for b in bs]
    n, B, K = len(toks), toks[0].shape[0], ws[0].shape[0]
    dev = toks[0].device
    if loss_kind == 0:
        labels = labels.to(torch.int64).contiguous()
    else:
        labels = labels.to(torch.float32).contiguous()
    logits = torch.empty(3, B, K, dtype=torch.float32, device=dev)
    losses = torch.empty(4, dtype=torch.float32, device=dev)
    preds = torch.zeros((3, B) if loss_kind == 0 else (3, B, K), dtype=torch.int64, device=dev)
    tok, bstride, ntok, dim, w, b, hw = _heads_args(toks, ws, bs, head_weight, slices)
    check(_L().m2b200_heads_loss_fwd(tok, bstride, ntok, dim, w, b, n, B, K, loss_kind, labels.data_ptr(),
                                     _ptr(pos_weight), hw, logits.data_ptr(), losses.data_ptr(), preds.data_ptr(),
                                     _stream()), "heads_loss_fwd")
    return losses, logits, preds


def heads_loss_bwd(toks, ws, bs, labels, pos_weight, head_weight, loss_kind: int, logits, grad_scale: float, grad_scale_dev=None,
                   grads=None, tok_start=None, tok_len=None):
    """Returns (dtoks, dws, dbs); heads that share one base tensor (token slices) share ONE gradient tensor: it is returned
    at the first head's position and the later positions hold an empty placeholder."""
    toks = [_as_tokens(t) for t in toks]
    slices = _tok_slices(toks, tok_start, tok_len)
    ws = [_f32c(w, "w") for w in ws]
    bs = [_f32c(b, "b") for b in bs]
    n, B, K = len(toks), toks[0].shape[0], ws[0].shape[0]
    labels = labels.to(torch.int64).contiguous() if loss_kind == 0 else labels.to(torch.float32).contiguous()
    owner, dbase = {}, []
    for i, t in enumerate(toks):                      # one gradient tensor per distinct base
        key = (t.data_ptr(), tuple(t.shape))
        if key not in owner:
            covered = sum(l0 for u, (_, l0) in zip(toks, slices) if (u.data_ptr(), tuple(u.shape)) == key)
            owner[key] = (torch.empty_like(t) if covered >= t.shape[1] else torch.zeros_like(t))
            dbase.append(owner[key])
        else:
            dbase.append(None)
    dfull = [owner[(t.data_ptr(), tuple(t.shape))] for t in toks]
    dws = [_grad_dst(grads, i, w) for i, w in enumerate(ws)]
    dbs = [_grad_dst(grads, n + i, b) for i, b in enumerate(bs)]
    tok, bstride, ntok, dim, w, b, hw = _heads_args(toks, ws, bs, head_weight, slices)
    arr_p = (C.c_void_p * 3)
    dtok = arr_p(*[t.data_ptr() + s0 * t.shape[2] * 4 for t, (s0, _) in zip(dfull, slices)] + [None] * (3 - n))
    dstride = (C.c_int64 * 3)(*[t.shape[1] * t.shape[2] for t in dfull] + [0] * (3 - n))
    acc = (C.c_int * 3)(0, 0, 0)
    dw = arr_p(*[t.data_ptr() for t in dws] + [None] * (3 - n))
    db = arr_p(*[t.data_ptr() for t in dbs] + [None] * (3 - n))
    check(_L().m2b200_heads_loss_bwd(tok, bstride, ntok, dim, w, b, n, B, K, loss_kind, labels.data_ptr(),
                                     _ptr(pos_weight), hw, _f32c(logits, "logits").data_ptr(), float(grad_scale),
                                     None if grad_scale_dev is None else _f32c(grad_scale_dev, "grad_scale_dev").data_ptr(), dtok,
                                     dstride, acc, dw, db, _stream()), "heads_loss_bwd")
    dtoks = [d if d is not None else torch.empty(0, dtype=torch.float32, device=toks[0].device) for d in dbase]
    return dtoks, dws, dbs


_define("heads_loss_fwd", "(Tensor[] toks, Tensor[] ws, Tensor[] bs, Tensor labels, Tensor? pos_weight, float[] head_weight, "
        "int loss_kind, int[]? tok_start=None, int[]? tok_len=None) -> (Tensor, Tensor, Tensor)", heads_loss_fwd)
_define("heads_loss_bwd", "(Tensor[] toks, Tensor[] ws, Tensor[] bs, Tensor labels, Tensor? pos_weight, float[] head_weight, "
        "int loss_kind, Tensor logits, float grad_scale, Tensor? grad_scale_dev, Tensor?[]? grads=None, int[]? tok_start=None, "
        "int[]? tok_len=None) -> (Tensor[], Tensor[], Tensor[])", heads_loss_bwd)


# ------------------------------------------------------------------------------------------------ optimiser / gemm
def adam_step(param, grad, exp_avg, exp_avg_sq, lr: float, beta1: float, beta2: float, eps: float, weight_decay: float,
              step: int, grad_scale: float, state: Optional[torch.Tensor] = None) -> None:
    for t in (param, grad, exp_avg, exp_avg_sq):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise RuntimeError("m2b200::adam_step expects contiguous float32 CUDA buffers")
    check(_L().m2b200_adam_step(param.data_ptr(), grad.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(), param.numel(),
                                lr, beta1, beta2, eps, weight_decay, step, grad_scale, _ptr(state), _stream()), "adam_step")


_define("adam_step", "(Tensor(a!) param, Tensor grad, Tensor(b!) exp_avg, Tensor(c!) exp_avg_sq, float lr, float beta1, "
        "float beta2, float eps, float weight_decay, int step, float grad_scale, Tensor? state) -> ()", adam_step)


def gemm(precision: int, A, a_mn: bool, B, b_mn: bool, M: int, N: int, K: int, *, batch: int = 1, a_batch_rows: int = 0,
         b_batch_rows: int = 0, bias=None, bias_mode: int = 0, act: int = 0, residual=None, out_bf16: bool = False,
         out: Optional[torch.Tensor] = None, accumulate: bool = False, splitk: int = 1) -> torch.Tensor:
    """Direct access to the generic GEMM (tests / benchmarks).  A, B: 2-D row-major (fp32 for FP32, bf16 for BF16)."""
    if out is None:
        shape = (batch, M, N) if batch > 1 else (M, N)
        mk = torch.zeros if (splitk > 1 or accumulate) else torch.empty
        out = mk(shape, dtype=torch.bfloat16 if out_bf16 else torch.float32, device=A.device)
    check(_L().m2b200_gemm(precision, A.data_ptr(), int(a_mn), A.stride(0), B.data_ptr(), int(b_mn), B.stride(0), M, N, K,
                           batch, a_batch_rows, b_batch_rows, _ptr(bias), bias_mode, act, _ptr(residual), N, M * N,
                           out.data_ptr(), int(out_bf16), N, M * N, int(accumulate), splitk, _stream()), "gemm")
    return out


# ------------------------------------------------------------------------------------------------ graph-safe dropout
def set_dropout_epoch(t: Optional[torch.Tensor]) -> None:
    """Register (or, with None, clear) the device-resident uint32/int32 counter every dropout kernel folds into its mask key
    (include/m2b200.h: m2b200_set_dropout_epoch_ptr).  The caller keeps the tensor alive."""
    if t is not None and not (t.is_cuda and t.numel() >= 1 and t.element_size() == 4):
        raise RuntimeError("m2b200: the dropout epoch must be a 4-byte CUDA tensor")
    _L().m2b200_set_dropout_epoch_ptr(None if t is None else t.data_ptr())


def dropout_epoch_advance(t: torch.Tensor) -> None:
    check(_L().m2b200_dropout_epoch_advance(t.data_ptr(), _stream()), "dropout_epoch_advance")


# ------------------------------------------------------------------------------------------------ fake (meta) kernels
# Shape / dtype propagation for FakeTensor and meta tracing (torch.compile, torch.export, make_fx): the ops stay opaque
# CUDA kernels, but a tracer can see what they return without launching them.  (register_fake.)
def _f32_like(t, shape=None):
    return torch.empty(t.shape if shape is None else shape, dtype=torch.float32, device=t.device)


def _register_fakes() -> None:
    fake = torch.library.register_fake
    ns = "m2b200::"

    fake(ns + "token_mix_fwd")(lambda x, *a, **k: _f32_like(x))
    fake(ns + "token_mix_bwd")(lambda du, x, ln_w, ln_b, w1, b1, w2, *a, **k: (
        _f32_like(x), _f32_like(ln_w), _f32_like(ln_b), _f32_like(w1), _f32_like(b1), _f32_like(w2), _f32_like(w2, (w2.shape[0],))))
    fake(ns + "channel_mix_fwd")(lambda u, *a, **k: _f32_like(u))
    fake(ns + "channel_mix_bwd")(lambda dy, u, ln_w, ln_b, w1, b1, w2, *a, **k: (
        _f32_like(u), _f32_like(ln_w), _f32_like(ln_b), _f32_like(w1), _f32_like(b1), _f32_like(w2), _f32_like(ln_w)))
    fake(ns + "layernorm_fwd")(lambda x, w, b: _f32_like(x))
    fake(ns + "layernorm_bwd")(lambda dy, x, w, grads=None: (_f32_like(x), _f32_like(w), _f32_like(w)))
    fake(ns + "layernorm_concat_fwd")(lambda xs, ws, bs: _f32_like(xs[0], (xs[0].shape[0], sum(x.shape[1] for x in xs), xs[0].shape[2])))
    fake(ns + "layernorm_concat_bwd")(lambda g, xs, ws, grads=None: (
        [_f32_like(x) for x in xs], [_f32_like(w) for w in ws], [_f32_like(w) for w in ws]))
    fake(ns + "linear_fwd")(lambda x, w, wb, bias, act, precision, dropout_p=0.0, seed=0: _f32_like(x, (*x.shape[:-1], w.shape[0])))
    fake(ns + "linear_bwd")(lambda dy, x, y, w, wb, act, need_dx, precision, dropout_p=0.0, seed=0, grads=None: (
        _f32_like(x) if need_dx else _f32_like(x, (0,)), _f32_like(w), _f32_like(w, (w.shape[0],))))

    def _patch_fwd(img, w, wb, bias, patch, precision):
        B, _, H, W = img.shape
        return _f32_like(img, (B, (H // patch) * (W // patch), w.shape[0]))

    fake(ns + "patch_embed_fwd")(_patch_fwd)
    fake(ns + "patch_embed_bwd")(lambda dy, img, w, patch, has_bias, precision, grads=None: (
        _f32_like(w), _f32_like(w, (w.shape[0] if has_bias else 0,))))
    fake(ns + "patch_gather")(lambda img, patch: _f32_like(img, (img.shape[0], (img.shape[2] // patch) * (img.shape[3] // patch),
                                                                   img.shape[1] * patch * patch)))
    fake(ns + "mean_pool_fwd")(lambda x: _f32_like(x, (x.shape[0], x.shape[-1])))
    fake(ns + "mean_pool_bwd")(lambda dp, n: _f32_like(dp, (dp.shape[0], n, dp.shape[-1])))
    fake(ns + "concat_tokens")(lambda xs: _f32_like(xs[0], (xs[0].shape[0], sum(x.shape[1] for x in xs), xs[0].shape[2])))
    fake(ns + "split_tokens")(lambda g, sizes: [_f32_like(g, (g.shape[0], n, g.shape[2])) for n in sizes])
    fake(ns + "add")(lambda a, b: _f32_like(a))
    fake(ns + "gate_fwd")(lambda h1, h2, zh: _f32_like(h1))
    fake(ns + "gate_bwd")(lambda h1, h2, zh, g: (_f32_like(h1), _f32_like(h1), _f32_like(h1)))
    fake(ns + "fuse2_fwd")(lambda a, b, mode: _f32_like(a))
    fake(ns + "fuse2_max_bwd")(lambda a, b, g: (_f32_like(a), _f32_like(a)))
    fake(ns + "cast_bf16")(lambda w, ld: torch.empty((w.shape[0], ld), dtype=torch.bfloat16, device=w.device))

    def _heads_fwd(toks, ws, bs, labels, pos_weight, head_weight, loss_kind, tok_start=None, tok_len=None):
        B, K = toks[0].shape[0], ws[0].shape[0]
        dev = toks[0].device
        return (torch.empty(4, dtype=torch.float32, device=dev), torch.empty(3, B, K, dtype=torch.float32, device=dev),
                torch.empty((3, B) if loss_kind == 0 else (3, B, K), dtype=torch.int64, device=dev))

    fake(ns + "heads_loss_fwd")(_heads_fwd)
    fake(ns + "heads_loss_bwd")(lambda toks, ws, bs, *a, **k: ([_f32_like(t) for t in toks], [_f32_like(w) for w in ws],
                                                               [_f32_like(b) for b in bs]))
    fake(ns + "adam_step")(lambda *a, **k: None)


_register_fakes()
