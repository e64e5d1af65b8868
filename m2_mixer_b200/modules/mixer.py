"""Drop-in Mixer building blocks backed by the m2b200 CUDA ops.

Same class names, constructor signatures (incl. the ``**kwargs`` swallow and ``dropout=`` keyword), public
attributes (``num_patch``, ``mixer_blocks``, ``layer_norm``) and state-dict keys / shapes / fp32 ``[out, in]`` layout
as the reference (modules/mixer.py:9-47, 112-186, 232-264; key layout SURVEY 3.3), so ``load_state_dict(strict=True)``
works in both directions with a published ``.ckpt``.  The nn.Linear / nn.LayerNorm children are *parameter
containers* created in the reference's construction order (identical default init under the same seed); their
``forward`` is never called - ``MixerBlock.forward`` hands the parameters to two fused kernels chains:

    u = x + Wt2 . GELU(Wt1 . LN1(x) + bt1) + bt2      token mixing, no materialised transposes   (token_mix)
    y = u + Wc2 . GELU(Wc1 . LN2(u) + bc1) + bc2      channel mixing on tcgen05 / TMEM / TMA      (channel_mix)
"""
from __future__ import annotations

import os

import torch
from torch import nn

from .. import functional as F
from .._lib import ACT_NONE

_DEFAULT_PRECISION = os.environ.get("M2B200_PRECISION", "bf16")


def set_default_precision(p: str) -> None:
    """'bf16' (tcgen05 tensor cores, default) or 'fp32' (CUDA-core parity mode) for modules built afterwards."""
    global _DEFAULT_PRECISION
    F.precision_code(p)
    _DEFAULT_PRECISION = p


def get_default_precision() -> str:
    return _DEFAULT_PRECISION


def _check_dropout(p: float, owner: str) -> float:
    p = float(p)
    if not 0.0 <= p < 1.0:
        raise ValueError(f"{owner}: dropout must be in [0, 1), got {p}")
    return p


class _Slot(nn.Module):
    """Parameter-free placeholder keeping nn.Sequential indices identical to the reference's (GELU, Dropout,
    Rearrange positions), so that parameter names such as ``token_mix.2.net.3.weight`` come out the same."""

    def __init__(self, what: str):
        super().__init__()
        self.what = what

    def extra_repr(self) -> str:
        return self.what

    def forward(self, x):  # pragma: no cover - containers are never executed
        raise RuntimeError("m2b200 parameter container: the fused MixerBlock kernels replace this layer")


class FeedForward(nn.Module):
    """Parameter container with the reference layout ``net = [Linear, GELU, Dropout, Linear, Dropout]``
    (modules/mixer.py:9-22).  Inside a MixerBlock it is consumed by the fused kernels, never called."""

    def __init__(self, dim, hidden_dim, dropout=0., out_dim=None):
        super().__init__()
        out_dim = out_dim or dim
        self.dropout_p = _check_dropout(dropout, "FeedForward")
        self.net = nn.Sequential(nn.Linear(dim, hidden_dim), _Slot("GELU (fused)"), _Slot(f"Dropout(p={dropout}) (fused)"),
                                 nn.Linear(hidden_dim, out_dim), _Slot(f"Dropout(p={dropout}) (fused)"))

    @property
    def fc1(self) -> nn.Linear:
        return self.net[0]

    @property
    def fc2(self) -> nn.Linear:
        return self.net[3]

    def forward(self, x):
        raise RuntimeError("FeedForward is fused into MixerBlock in m2_mixer_b200; call the MixerBlock instead")


class MixerBlock(nn.Module):
    def __init__(self, hidden_dim, num_patch, token_dim, channel_dim, dropout=0.):
        super().__init__()
        self.hidden_dim, self.num_patch, self.token_dim, self.channel_dim = hidden_dim, num_patch, token_dim, channel_dim
        self.dropout_p = _check_dropout(dropout, "MixerBlock")
        self.precision = _DEFAULT_PRECISION
        self.token_mix = nn.Sequential(nn.LayerNorm(hidden_dim), _Slot("b n d -> b d n (operand-major swap)"),
                                       FeedForward(num_patch, token_dim, dropout), _Slot("b d n -> b n d (operand-major swap)"))
        self.channel_mix = nn.Sequential(nn.LayerNorm(hidden_dim), FeedForward(hidden_dim, channel_dim, dropout))

    def forward(self, x):
        p = self.dropout_p if self.training else 0.0     # nn.Dropout semantics: identity in eval mode
        if x.dim() != 3 or x.shape[1] != self.num_patch or x.shape[2] != self.hidden_dim:
            raise ValueError(f"MixerBlock expects [B, {self.num_patch}, {self.hidden_dim}], got {tuple(x.shape)}")
        ln1, tff = self.token_mix[0], self.token_mix[2]
        ln2, cff = self.channel_mix[0], self.channel_mix[1]
        u = F.token_mix(x, ln1.weight, ln1.bias, tff.fc1.weight, tff.fc1.bias, tff.fc2.weight, tff.fc2.bias, self.precision,
                        dropout_p=p)
        return F.channel_mix(u, ln2.weight, ln2.bias, cff.fc1.weight, cff.fc1.bias, cff.fc2.weight, cff.fc2.bias,
                             self.precision, dropout_p=p)


class _Stack(nn.Module):
    """mixer_blocks + closing LayerNorm shared by every encoder below."""

    def _build_stack(self, hidden_dim, num_patch, num_mixers, token_dim, channel_dim, dropout):
        self.mixer_blocks = nn.ModuleList([])
        for _ in range(num_mixers):
            self.mixer_blocks.append(MixerBlock(hidden_dim, num_patch, token_dim, channel_dim, dropout=dropout))

    def _run_blocks(self, x):
        for blk in self.mixer_blocks:
            x = blk(x)
        return x

    def _run_stack(self, x):
        return F.layer_norm(self._run_blocks(x), self.layer_norm.weight, self.layer_norm.bias)

    def forward(self, x):
        return F.layer_norm(self.forward_features(x), self.layer_norm.weight, self.layer_norm.bias)

    @property
    def precision(self) -> str:
        return self.mixer_blocks[0].precision if len(self.mixer_blocks) else self._precision

    @precision.setter
    def precision(self, p: str) -> None:
        F.precision_code(p)
        self._precision = p
        for blk in self.mixer_blocks:
            blk.precision = p


class FusionMixer(_Stack):
    def __init__(self, hidden_dim, num_patches, num_mixers, token_dim, channel_dim, dropout=0., **kwargs):
        super().__init__()
        self._precision = _DEFAULT_PRECISION
        self.num_patch = num_patches
        self._build_stack(hidden_dim, self.num_patch, num_mixers, token_dim, channel_dim, dropout)
        self.layer_norm = nn.LayerNorm(hidden_dim)

    def forward_features(self, x):
        """The stack WITHOUT its closing LayerNorm (the task modules normalise two encoders straight into the fused-token
        buffer with F.layer_norm_concat: zero-copy ConcatFusion)."""
        return self._run_blocks(x)


class MLPMixer(_Stack):
    def __init__(self, in_channels, hidden_dim, patch_size, image_size, num_mixers, token_dim, channel_dim, dropout=0.,
                 **kwargs):
        super().__init__()
        self._precision = _DEFAULT_PRECISION
        assert (image_size[0] % patch_size == 0) and (image_size[1] % patch_size == 0), \
            'Image dimensions must be divisible by the patch size.'
        self.patch_size = patch_size
        self.num_patch = (image_size[0] // patch_size) * (image_size[1] // patch_size)
        # index 0 = Conv2d parameter container (weight [D, cin, p, p], bias [D]); index 1 = the rearrange slot
        self.to_patch_embedding = nn.Sequential(nn.Conv2d(in_channels, hidden_dim, patch_size, patch_size),
                                                _Slot("b c h w -> b (h w) c (folded into the patch GEMM)"))
        self._build_stack(hidden_dim, self.num_patch, num_mixers, token_dim, channel_dim, dropout)
        self.layer_norm = nn.LayerNorm(hidden_dim)

    def forward_features(self, x):
        conv = self.to_patch_embedding[0]
        x = F.patch_embed(x, conv.weight, conv.bias, self.patch_size, self.precision)
        return self._run_blocks(x)


class MLPMixerNoPatching(_Stack):
    def __init__(self, hidden_dim, num_patch, num_mixers, token_dim, channel_dim, embedding_dim, proj_dim, dropout=0.,
                 **kwargs):
        super().__init__()
        self._precision = _DEFAULT_PRECISION
        self.num_patch = num_patch
        self.proj = nn.Linear(embedding_dim, proj_dim)
        self._build_stack(hidden_dim, self.num_patch, num_mixers, token_dim, channel_dim, dropout)
        self.layer_norm = nn.LayerNorm(hidden_dim)

    def forward_features(self, x):
        x = F.linear(x, self.proj.weight, self.proj.bias, ACT_NONE, self.precision)
        return self._run_blocks(x)


class PNLPMixer(_Stack):
    def __init__(self, max_seq_len, hidden_dim, num_mixers, mlp_hidden_dim, bottleneck_window_size,
                 bottleneck_features_size, dropout=0., **kwargs):
        super().__init__()
        self._precision = _DEFAULT_PRECISION
        self.num_patch = max_seq_len
        self.mixer_blocks = nn.ModuleList([])
        self.bottleneck = nn.Linear((2 * bottleneck_window_size + 1) * bottleneck_features_size, hidden_dim)
        for _ in range(num_mixers):   # token_dim == channel_dim == mlp_hidden_dim (reference modules/mixer.py:249)
            self.mixer_blocks.append(MixerBlock(hidden_dim, max_seq_len, mlp_hidden_dim, mlp_hidden_dim, dropout=dropout))
        self.layer_norm = nn.LayerNorm(hidden_dim)

    def forward_features(self, x):
        x = F.linear(x, self.bottleneck.weight, self.bottleneck.bias, ACT_NONE, self.precision)
        return self._run_blocks(x)
