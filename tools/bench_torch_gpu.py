"""Same-box stock-PyTorch GPU comparator (SURVEY 8(d) "Timing method"): the reference's training step as the aten ops its
modules call - F.conv2d(k = stride = patch) + permute, F.layer_norm, permute -> F.linear -> F.gelu -> F.dropout -> F.linear
-> permute (modules/mixer.py:9-47, 135-162), torch.cat fusion, mean-pool + Linear heads, three cross entropies
(models/avmnist.py:259-312) - eager, under torch.autocast(bf16) (and in fp32), with torch.optim.Adam (models/avmnist.py:413-415),
on the same synthetic batch shapes and step count as bench.py.

COMPARATOR ONLY: nothing in m2_mixer_b200/ imports this file.  /root/reference does not exist on the GPU box, so the modules
cannot be imported there; this restates their op sequence (cuDNN conv, cuBLAS GEMMs, torch's LayerNorm / GELU / dropout
kernels, materialised transposes) - the path a user of the reference runs today.

    python tools/bench_torch_gpu.py [--config avmnist_B] [--batch 4096] [--steps 20] [--warmup 5] [--fused-adam]
"""
import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def block(x, sd, pre, p):
    h = F.layer_norm(x, x.shape[-1:], sd[pre + "token_mix.0.weight"], sd[pre + "token_mix.0.bias"])
    h = h.permute(0, 2, 1)                                                 # Rearrange('b n d -> b d n')
    h = F.dropout(F.gelu(F.linear(h, sd[pre + "token_mix.2.net.0.weight"], sd[pre + "token_mix.2.net.0.bias"])), p, True)
    h = F.dropout(F.linear(h, sd[pre + "token_mix.2.net.3.weight"], sd[pre + "token_mix.2.net.3.bias"]), p, True)
    x = x + h.permute(0, 2, 1)                                             # Rearrange('b d n -> b n d')
    h = F.layer_norm(x, x.shape[-1:], sd[pre + "channel_mix.0.weight"], sd[pre + "channel_mix.0.bias"])
    h = F.dropout(F.gelu(F.linear(h, sd[pre + "channel_mix.1.net.0.weight"], sd[pre + "channel_mix.1.net.0.bias"])), p, True)
    h = F.dropout(F.linear(h, sd[pre + "channel_mix.1.net.3.weight"], sd[pre + "channel_mix.1.net.3.bias"]), p, True)
    return x + h


def stack(x, sd, pre, p):
    i = 0
    while f"{pre}mixer_blocks.{i}.token_mix.0.weight" in sd:
        x = block(x, sd, f"{pre}mixer_blocks.{i}.", p)
        i += 1
    return F.layer_norm(x, x.shape[-1:], sd[pre + "layer_norm.weight"], sd[pre + "layer_norm.bias"])


def mlp_mixer(img, sd, pre, p):
    w = sd[pre + "to_patch_embedding.0.weight"]
    x = F.conv2d(img, w, sd[pre + "to_patch_embedding.0.bias"], stride=w.shape[-1])
    return stack(x.flatten(2).permute(0, 2, 1), sd, pre, p)                # 'b c h w -> b (h w) c'


def avmnist_step(sd, batch, p):
    it = mlp_mixer(batch["image"], sd, "image_mixer.", p)
    at = mlp_mixer(batch["audio"], sd, "audio_mixer.", p)
    ft = stack(torch.cat([it, at], dim=1), sd, "fusion_mixer.", p)
    li = F.linear(it.mean(1), sd["classifier_image.weight"], sd["classifier_image.bias"])
    la = F.linear(at.mean(1), sd["classifier_audio.weight"], sd["classifier_audio.bias"])
    lf = F.linear(ft.mean(1), sd["classifier_fusion.classifer.weight"], sd["classifier_fusion.classifer.bias"])
    y = batch["label"]
    w, ow = 1.0 / 3, 1.0 / 3
    return (w * F.cross_entropy(lf.float(), y) + ow * F.cross_entropy(li.float(), y) + ow * F.cross_entropy(la.float(), y)) * 3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="avmnist_B")
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--fused-adam", action="store_true", help="torch.optim.Adam(fused=True) instead of the reference's default")
    args = ap.parse_args()
    import bench
    from m2_mixer_b200 import models, presets
    from oracle.seeding import seeded_state_dict
    cfg = presets.get(args.config)
    assert cfg["type"] == "AVMnistMixerMultiLoss", "comparator covers the AV-MNIST task module (C1, C2, C5)"
    B = args.batch or bench.DEFAULT_BATCH[args.config]
    dev = torch.device("cuda", 0)
    shapes = {k: tuple(v.shape) for k, v in models.get_model(cfg["type"])(dict(cfg, dropout=0.0), {}).state_dict().items()}
    p = cfg.get("dropout", 0.0)
    g = torch.Generator(device=dev).manual_seed(1234)
    batches = [bench.make_batch(cfg, B, dev, g) for _ in range(4)]
    fl = bench.model_flops(cfg)["fwd_bwd"]
    out = []
    for mode in ("autocast_bf16", "fp32"):
        sd = {k: v.to(dev).requires_grad_(True) for k, v in seeded_state_dict(shapes, 42).items()}
        opt = torch.optim.Adam(list(sd.values()), lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, fused=args.fused_adam)

        def step(b):
            opt.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "autocast_bf16")):
                loss = avmnist_step(sd, b, p)
            loss.backward()
            opt.step()
            return loss

        for i in range(args.warmup):
            step(batches[i % 4])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            loss = step(batches[i % 4])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        out.append({"impl": "stock PyTorch eager (" + mode + ", torch.optim.Adam" + (" fused" if args.fused_adam else "") + ")",
                    "config": args.config, "batch": B, "dropout": p, "steps": args.steps, "ms_per_step": ms,
                    "samples_per_s": B / ms * 1e3, "model_tflops": fl * B / ms / 1e9, "loss_last": float(loss),
                    "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30, "torch": torch.__version__})
        print(json.dumps(out[-1]), flush=True)
        del sd, opt
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()


if __name__ == "__main__":
    main()
