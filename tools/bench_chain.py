"""Micro-benchmark of the channel-mix chain ops alone (CUDA events), M2-Mixer-B shapes.

    python tools/bench_chain.py [M D C [iters]]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from m2_mixer_b200 import _lib, ops  # noqa: E402
from m2_mixer_b200._lib import BF16  # noqa: E402


def main():
    M, D, C = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (16384, 128, 3072)
    iters = int(sys.argv[4]) if len(sys.argv) > 4 else 20
    dev = "cuda"
    torch.manual_seed(0)
    u = torch.randn(M, D, device=dev)
    dy = torch.randn(M, D, device=dev)
    ln_w, ln_b = torch.ones(D, device=dev), torch.zeros(D, device=dev)
    w1, b1 = torch.randn(C, D, device=dev) / D ** 0.5, torch.randn(C, device=dev) * 0.1
    w2, b2 = torch.randn(D, C, device=dev) / C ** 0.5, torch.randn(D, device=dev) * 0.1
    w1b, w2b = ops.cast_bf16(w1, D), ops.cast_bf16(w2, (C + 7) // 8 * 8)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def run(fn, name, flops):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        med = ts[len(ts) // 2]
        print(f"{name:28s} median {med * 1e3:8.1f} us  min {ts[0] * 1e3:8.1f} us  {flops / med / 1e9:8.1f} TFLOP/s (algorithmic)")

    f = 4.0 * M * D * C
    run(lambda: ops.channel_mix_fwd(u, ln_w, ln_b, w1, b1, w2, b2, w1b, w2b, BF16), "channel_mix_fwd", f)
    run(lambda: ops.channel_mix_fwd(u, ln_w, ln_b, w1, b1, w2, b2, w1b, w2b, BF16, 0.5, 1234), "channel_mix_fwd dropout .5", f)
    run(lambda: ops.channel_mix_bwd(dy, u, ln_w, ln_b, w1, b1, w2, w1b, w2b, BF16), "channel_mix_bwd (all)", 2 * f)
    run(lambda: ops.channel_mix_bwd(dy, u, ln_w, ln_b, w1, b1, w2, w1b, w2b, BF16, 0.5, 1234), "channel_mix_bwd dropout .5", 2 * f)
    for pdrop in (0.0, 0.5):
        with _lib.profile() as p:
            for _ in range(5):
                ops.channel_mix_fwd(u, ln_w, ln_b, w1, b1, w2, b2, w1b, w2b, BF16, pdrop, 1234)
                ops.channel_mix_bwd(dy, u, ln_w, ln_b, w1, b1, w2, w1b, w2b, BF16, pdrop, 1234)
            torch.cuda.synchronize()
        for k, (n, ms) in sorted(p.table.items(), key=lambda kv: -kv[1][1]):
            print(f"   p={pdrop} {k:24s} {ms / n * 1e3:8.1f} us/launch x{n // 5}")


if __name__ == "__main__":
    main()
