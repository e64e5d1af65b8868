"""Launch-bound configurations (BASELINE config 3: MIMIC-III M2-Mixer-H; config 1's shape: M2-Mixer-S): step latency and
kernel launches per step, eager vs one CUDA graph per step (m2_mixer_b200.graph.GraphedTrainStep).

    python tools/bench_small.py [--steps 200] [--json out.json]
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from m2_mixer_b200 import _lib, models, ops, presets  # noqa: E402
from m2_mixer_b200.graph import GraphedTrainStep  # noqa: E402
from m2_mixer_b200.optim import FusedAdam  # noqa: E402


def batch_for(kind, B, dev, seed=0):
    g = torch.Generator(device=dev).manual_seed(seed)
    rn = lambda *s: torch.randn(*s, device=dev, generator=g)
    if kind == "mimic":
        return (rn(B, 5), rn(B, 24, 12), torch.randint(0, 6, (B,), device=dev, generator=g))
    return {"image": rn(B, 1, 28, 28), "audio": rn(B, 1, 112, 112), "label": torch.randint(0, 10, (B,), device=dev, generator=g)}


def timed(fn, steps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps * 1e3, (time.perf_counter() - t0) / steps * 1e6   # device us, wall us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    out = []
    for name, kind, batches in (("mimic_H", "mimic", (128, 4096)), ("avmnist_S", "avmnist", (32, 4096))):
        cfg = presets.get(name)
        for B in batches:
            torch.manual_seed(42)
            m = models.get_model(cfg["type"])(cfg, {}).to(dev).train()
            opt = FusedAdam(m.parameters(), lr=1e-3, capturable=True)
            bts = [batch_for(kind, B, dev, s) for s in range(4)]

            def eager(i):
                opt.zero_grad()
                loss = m.training_step(bts[i % 4])
                loss.backward()
                opt.step()

            for i in range(10):
                eager(i)
            l0 = _lib.launch_count()
            us_e, wall_e = timed(eager, args.steps)
            launches = (_lib.launch_count() - l0) / args.steps
            step = GraphedTrainStep(m, opt, bts[0], warmup=3)
            for i in range(10):
                step(bts[i % 4])
            us_g, wall_g = timed(lambda i: step(bts[i % 4]), args.steps)
            step.close()
            rec = {"config": name, "batch": B, "dropout": cfg.get("dropout", 0.0), "launches_per_step": launches,
                   "eager_us_per_step": us_e, "graph_us_per_step": us_g, "eager_samples_per_s": B / us_e * 1e6,
                   "graph_samples_per_s": B / us_g * 1e6, "speedup": us_e / us_g}
            out.append(rec)
            print(json.dumps(rec), flush=True)
    ops.set_dropout_epoch(None)
    if args.json:
        json.dump(out, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
