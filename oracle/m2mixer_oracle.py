"""CPU oracle for the M2-Mixer training hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this file.  The product package
(``m2_mixer_b200``) never imports it and has no CPU fallback.

Parity status: the reference ships NO numeric test for this path (its only
tests are shape checks in ``tests/modules/test_fusion.py:8-92``), so the oracle
is pinned against *outputs of the reference itself*: ``tests/golden/make_golden.py``
imports the unmodified ``/root/reference/modules`` classes in the build
container, runs them on seeded inputs and commits inputs / state dict / logits /
losses / gradients under ``tests/golden/``.  ``tests/test_oracle.py`` checks this
restatement against those vectors (and against the live reference when
``/root/reference`` is present).

The restatement is functional (a flat ``state_dict`` with the reference's key
names, no nn.Module) and transpose-free: token mixing is written as two einsums
over the patch axis instead of permute -> Linear -> permute.

Reference lines restated here
  FeedForward            modules/mixer.py:9-22
  MixerBlock             modules/mixer.py:25-47
  FusionMixer            modules/mixer.py:112-132
  MLPMixer               modules/mixer.py:135-162
  MLPMixerNoPatching     modules/mixer.py:165-186
  PNLPMixer              modules/mixer.py:232-264
  MLP                    modules/mlp.py:4-27
  ConcatFusion/SumFusion modules/fusion.py:112-146, 207-221
  StandardClassifier     modules/classification.py:84-90
  AV-MNIST shared_step   models/avmnist.py:236-312
  MIMIC shared_step      models/mimic.py:93-142
  MM-IMDB shared_step    models/mmimdb.py:65-147
The arithmetic underneath is PyTorch's (aten linear / native_layer_norm / erf
GELU / log_softmax+nll / bce_with_logits); the reference pins torch==1.13.1,
this image has 2.11 - fp32 semantics of these ops are unchanged.
"""
from __future__ import annotations

import math
from typing import Dict, Mapping, Sequence

import torch
import torch.nn.functional as F

SD = Mapping[str, torch.Tensor]
LN_EPS = 1e-5  # nn.LayerNorm default, modules/mixer.py:31


# --------------------------------------------------------------------------- blocks
def layer_norm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)  # biased, as native_layer_norm
    return (x - mu) * torch.rsqrt(var + LN_EPS) * w + b


def gelu_erf(x: torch.Tensor) -> torch.Tensor:
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


def _drop(x: torch.Tensor, p: float, training: bool) -> torch.Tensor:
    return F.dropout(x, p, training) if (training and p > 0.0) else x


def mixer_block(x: torch.Tensor, sd: SD, pre: str, p: float = 0.0, training: bool = False) -> torch.Tensor:
    """modules/mixer.py:42-47.  x [B,N,D]."""
    # token mixing (modules/mixer.py:30-35): contraction over the patch axis n
    xn = layer_norm(x, sd[pre + "token_mix.0.weight"], sd[pre + "token_mix.0.bias"])
    wt1, bt1 = sd[pre + "token_mix.2.net.0.weight"], sd[pre + "token_mix.2.net.0.bias"]   # [T,N],[T]
    wt2, bt2 = sd[pre + "token_mix.2.net.3.weight"], sd[pre + "token_mix.2.net.3.bias"]   # [N,T],[N]
    h = torch.einsum("tn,bnd->btd", wt1, xn) + bt1[None, :, None]
    h = _drop(gelu_erf(h), p, training)
    u = x + _drop(torch.einsum("nt,btd->bnd", wt2, h) + bt2[None, :, None], p, training)
    # channel mixing (modules/mixer.py:37-40): contraction over the hidden axis d
    un = layer_norm(u, sd[pre + "channel_mix.0.weight"], sd[pre + "channel_mix.0.bias"])
    wc1, bc1 = sd[pre + "channel_mix.1.net.0.weight"], sd[pre + "channel_mix.1.net.0.bias"]  # [C,D],[C]
    wc2, bc2 = sd[pre + "channel_mix.1.net.3.weight"], sd[pre + "channel_mix.1.net.3.bias"]  # [D,C],[D]
    g = _drop(gelu_erf(un @ wc1.t() + bc1), p, training)
    return u + _drop(g @ wc2.t() + bc2, p, training)


def _num_blocks(sd: SD, pre: str) -> int:
    n = 0
    while f"{pre}mixer_blocks.{n}.token_mix.0.weight" in sd:
        n += 1
    return n


def _stack(x: torch.Tensor, sd: SD, pre: str, p: float, training: bool) -> torch.Tensor:
    for i in range(_num_blocks(sd, pre)):
        x = mixer_block(x, sd, f"{pre}mixer_blocks.{i}.", p, training)
    return layer_norm(x, sd[pre + "layer_norm.weight"], sd[pre + "layer_norm.bias"])


def patch_embed(img: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Conv2d(k=stride=p) + 'b c h w -> b (h w) c' (modules/mixer.py:143-146) as one GEMM."""
    bsz, cin, hh, ww = img.shape
    dd, _, ph, pw = w.shape
    gh, gw = hh // ph, ww // pw
    cols = img.reshape(bsz, cin, gh, ph, gw, pw).permute(0, 2, 4, 1, 3, 5).reshape(bsz, gh * gw, cin * ph * pw)
    return cols @ w.reshape(dd, -1).t() + b


def mlp_mixer(img: torch.Tensor, sd: SD, pre: str, p: float = 0.0, training: bool = False) -> torch.Tensor:
    x = patch_embed(img, sd[pre + "to_patch_embedding.0.weight"], sd[pre + "to_patch_embedding.0.bias"])
    return _stack(x, sd, pre, p, training)


def fusion_mixer(x: torch.Tensor, sd: SD, pre: str, p: float = 0.0, training: bool = False) -> torch.Tensor:
    return _stack(x, sd, pre, p, training)


def mlp_mixer_no_patching(x: torch.Tensor, sd: SD, pre: str, p: float = 0.0, training: bool = False) -> torch.Tensor:
    x = x @ sd[pre + "proj.weight"].t() + sd[pre + "proj.bias"]
    return _stack(x, sd, pre, p, training)


def pnlp_mixer(x: torch.Tensor, sd: SD, pre: str, p: float = 0.0, training: bool = False) -> torch.Tensor:
    x = x @ sd[pre + "bottleneck.weight"].t() + sd[pre + "bottleneck.bias"]
    return _stack(x, sd, pre, p, training)


def mlp_encoder(x: torch.Tensor, sd: SD, pre: str, p: float = 0.0, training: bool = False,
                has_output: bool = True) -> torch.Tensor:
    """modules/mlp.py:4-27: module_list = [Linear, ReLU, Dropout]*num_blocks (+ output Linear).

    Linear layers sit at positions 0,3,...; with ``output_dim`` set (the MIMIC cfg always sets it,
    cfg/mimic/mimic_m2-mixer_H.yml:45) the last one is the output layer and has no activation.
    """
    idx = sorted({int(k[len(pre + "module_list."):].split(".")[0]) for k in sd if k.startswith(pre + "module_list.")})
    for j, i in enumerate(idx):
        x = x @ sd[f"{pre}module_list.{i}.weight"].t() + sd[f"{pre}module_list.{i}.bias"]
        if not (has_output and j == len(idx) - 1):
            x = _drop(torch.relu(x), p, training)
    return x


# --------------------------------------------------------------------------- fusion / heads
def concat_fusion(*xs: torch.Tensor, dim: int = 1) -> torch.Tensor:
    return torch.cat(xs, dim=dim)


def sum_fusion(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    return a + b


def concat_out_shape(*args, dim=None, self_dim: int = 1):
    """ConcatFusion.get_output_shape, modules/fusion.py:119-146."""
    if dim is not None:
        if not isinstance(args[0], int):
            raise ValueError("The dim argument is only used if the first argument is an int.")
        return sum(args) if dim == self_dim else args[0]
    shape = list(args[0])
    for a in args[1:]:
        shape[self_dim] += a[self_dim]
    return tuple(shape)


def pooled_linear(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """mean over all token axes then Linear (models/avmnist.py:267-272, classification.py:90)."""
    return x.reshape(x.shape[0], -1, x.shape[-1]).mean(dim=1) @ w.t() + b


def cross_entropy(logits: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    lse = torch.logsumexp(logits, dim=1)
    return (lse - logits.gather(1, y[:, None].long())[:, 0]).mean()


def bce_pos_weight(logits: torch.Tensor, y: torch.Tensor, pw: torch.Tensor) -> torch.Tensor:
    # -[pw*y*log sig(x) + (1-y)*log(1-sig(x))], mean over every element (mmimdb.py:48-50)
    return (pw * y * F.softplus(-logits) + (1.0 - y) * F.softplus(logits)).mean()


# --------------------------------------------------------------------------- shared_step
def avmnist_shared_step(sd: SD, batch: Dict[str, torch.Tensor], fusion_loss_weight: float = 1.0 / 3,
                        fusion: str = "ConcatFusion", p: float = 0.0, training: bool = False) -> Dict[str, torch.Tensor]:
    """models/avmnist.py:236-312 (default branch: no softadapt / gradblend / freezing)."""
    y = batch["label"]
    img = mlp_mixer(batch["image"], sd, "image_mixer.", p, training)
    aud = mlp_mixer(batch["audio"], sd, "audio_mixer.", p, training)
    fused = concat_fusion(img, aud) if fusion == "ConcatFusion" else sum_fusion(img, aud)
    fus = fusion_mixer(fused, sd, "fusion_mixer.", p, training)
    li = pooled_linear(img, sd["classifier_image.weight"], sd["classifier_image.bias"])
    la = pooled_linear(aud, sd["classifier_audio.weight"], sd["classifier_audio.bias"])
    lf = pooled_linear(fus, sd["classifier_fusion.classifer.weight"], sd["classifier_fusion.classifer.bias"])
    loss_i, loss_a, loss_f = cross_entropy(li, y), cross_entropy(la, y), cross_entropy(lf, y)
    ow = (1 - fusion_loss_weight) / 2
    loss = (fusion_loss_weight * loss_f + ow * loss_i + ow * loss_a) * 3
    return {"preds": lf.argmax(1), "preds_image": li.argmax(1), "preds_audio": la.argmax(1), "labels": y,
            "loss": loss, "loss_image": loss_i, "loss_audio": loss_a, "loss_fusion": loss_f,
            "image_logits": li, "audio_logits": la, "logits": lf}


def mimic_shared_step(sd: SD, batch, fusion_loss_weight: float = 1.0 / 3, p: float = 0.0,
                      training: bool = False) -> Dict[str, torch.Tensor]:
    """models/mimic.py:93-142 (no gradblend).  Note: no `*3` here, unlike AV-MNIST."""
    static, time, y = batch
    s = mlp_encoder(static, sd, "static_extractor.", p, training)
    t = mlp_mixer_no_patching(time, sd, "time_mixer.", p, training)
    fus = fusion_mixer(concat_fusion(s[:, None, :], t), sd, "fusion_mixer.", p, training)
    ls = s @ sd["classifier_static.weight"].t() + sd["classifier_static.bias"]
    lt = pooled_linear(t, sd["classifier_time.weight"], sd["classifier_time.bias"])
    lf = pooled_linear(fus, sd["classifier_fusion.classifer.weight"], sd["classifier_fusion.classifer.bias"])
    loss_f, loss_s, loss_t = cross_entropy(lf, y), cross_entropy(ls, y), cross_entropy(lt, y)
    ow = (1 - fusion_loss_weight) / 2
    loss = fusion_loss_weight * loss_f + ow * loss_s + ow * loss_t
    return {"preds": torch.softmax(lf, 1), "preds_static": torch.softmax(ls, 1), "preds_time": torch.softmax(lt, 1),
            "labels": y.long(), "loss": loss, "loss_fusion": loss_f, "loss_static": loss_s, "loss_time": loss_t,
            "logits": lf, "logits_static": ls, "logits_time": lt}


def mmimdb_shared_step(sd: SD, batch: Dict[str, torch.Tensor], pos_weight: torch.Tensor, text_encoder: str = "MLPMixer",
                       p: float = 0.0, training: bool = False) -> Dict[str, torch.Tensor]:
    """models/mmimdb.py:65-147 (BCE-with-pos-weight x3, plain sum)."""
    y = batch["label"].to(pos_weight.dtype)
    img = mlp_mixer(batch["image"], sd, "image_mixer.", p, training)
    enc = mlp_mixer if text_encoder == "MLPMixer" else pnlp_mixer
    txt = enc(batch["text"], sd, "text_mixer.", p, training)
    fus = fusion_mixer(concat_fusion(img, txt), sd, "fusion_mixer.", p, training)
    li = pooled_linear(img, sd["classifier_image.weight"], sd["classifier_image.bias"])
    lt = pooled_linear(txt, sd["classifier_text.weight"], sd["classifier_text.bias"])
    lf = pooled_linear(fus, sd["classifier_fusion.classifer.weight"], sd["classifier_fusion.classifer.bias"])
    loss_i, loss_t, loss_f = (bce_pos_weight(l, y, pos_weight) for l in (li, lt, lf))
    return {"preds": (lf > 0).long(), "preds_image": (li > 0).long(), "preds_text": (lt > 0).long(),
            "labels": batch["label"], "loss": loss_i + loss_t + loss_f, "loss_image": loss_i, "loss_text": loss_t,
            "loss_fusion": loss_f, "image_logits": li, "text_logits": lt, "logits": lf}


# --------------------------------------------------------------------------- optimiser
def adam_step(params: Sequence[torch.Tensor], grads: Sequence[torch.Tensor], exp_avg, exp_avg_sq, step: int,
              lr: float, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0) -> None:
    """torch.optim.Adam (non-amsgrad, L2 weight decay) as used at models/avmnist.py:413-415."""
    b1, b2 = betas
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    for p_, g, m, v in zip(params, grads, exp_avg, exp_avg_sq):
        if weight_decay != 0.0:
            g = g + weight_decay * p_
        m.mul_(b1).add_(g, alpha=1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
        p_.addcdiv_(m, denom, value=-lr / bc1)


def grads_of(loss: torch.Tensor, sd: SD) -> Dict[str, torch.Tensor]:
    names = [k for k, v in sd.items() if v.requires_grad]
    gs = torch.autograd.grad(loss, [sd[k] for k in names], allow_unused=True)
    return {k: (g if g is not None else torch.zeros_like(sd[k])) for k, g in zip(names, gs)}
