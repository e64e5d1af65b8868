"""The two large synthetic configurations of BASELINE.json: C4 MM-IMDB-shaped image + text Mixer (BCE multi-label)
and C5 Scaled M2-Mixer (Mixer-B/16 width x 12 blocks per modality).  A few training steps each, kernel by kernel, with the
per-kernel device-time table: these shapes run on the generic paths (batched tcgen05 GEMM token mixing for large N / T,
unfused channel mixing for D > 256) - the point is that they RUN and where their time goes.

    python tools/bench_configs.py [--c4-batch 256] [--c5-batch 32] [--steps 5]
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from m2_mixer_b200 import _lib, models, presets  # noqa: E402
from m2_mixer_b200.optim import FusedAdam  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--c4-batch", type=int, default=256)
    ap.add_argument("--c5-batch", type=int, default=32)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(0)
    rn = lambda *s: torch.randn(*s, device=dev, generator=g)
    jobs = []
    if args.only in ("", "c4"):
        B = args.c4_batch
        jobs.append(("mmimdb_C4", B, {"image": rn(B, 3, 224, 224), "text": rn(B, 512, 1280),
                                      "label": (torch.rand(B, 23, device=dev, generator=g) < 0.1).long()}, 7391.8e6))
    if args.only in ("", "c5"):
        B = args.c5_batch
        jobs.append(("scaled_C5", B, {"image": rn(B, 3, 224, 224), "audio": rn(B, 3, 224, 224),
                                      "label": torch.randint(0, 10, (B,), device=dev, generator=g)}, 301036.9e6))
    for name, B, batch, flops_per_sample in jobs:
        cfg = presets.get(name)
        torch.manual_seed(42)
        m = models.get_model(cfg["type"])(cfg, {}).to(dev).set_precision("bf16").train()
        nparam = sum(p.numel() for p in m.parameters())
        opt = FusedAdam(m.parameters(), lr=1e-3)

        def step():
            opt.zero_grad()
            loss = m.training_step(batch)
            loss.backward()
            opt.step()
            return loss

        t0 = time.time()
        l0 = float(step())
        torch.cuda.synchronize()
        first = time.time() - t0
        step()
        torch.cuda.synchronize()
        lc = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            loss = step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        launches = (_lib.launch_count() - lc) / args.steps
        with _lib.profile() as prof:
            step()
            torch.cuda.synchronize()
        top = sorted(prof.table.items(), key=lambda kv: -kv[1][1])[:8]
        rec = {"config": name, "batch": B, "params": nparam, "ms_per_step": ms, "samples_per_s": B / ms * 1e3,
               "model_tflops": flops_per_sample * B / ms / 1e9, "launches_per_step": launches, "first_step_s": first,
               "loss_first": l0, "loss_last": float(loss), "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
               "top_kernels_ms": {k: round(v[1], 3) for k, v in top}}
        print(json.dumps(rec), flush=True)
        del m, opt
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
