"""CI-sized run of the mbarrier / TMEM-aliasing kernels for compute-sanitizer.

    compute-sanitizer --tool racecheck python tools/sanitize_chain.py
    compute-sanitizer --tool memcheck  python tools/sanitize_chain.py

One channel-mixing forward, dgrad chain and weight-gradient launch (chain_fwd_ts, chain_bwd_ts, wgrad_dh; wgrad_fused with M2B200_CHAIN_GEN=2) plus one
token-mixing forward/backward at a shape small enough for the sanitizer's serialised execution.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from m2_mixer_b200 import ops  # noqa: E402
from m2_mixer_b200._lib import BF16  # noqa: E402


def main():
    M, D, C = 256, 128, 256
    B, N, T = 64, 4, 32
    dev = "cuda"
    torch.manual_seed(0)
    u, dy = torch.randn(M, D, device=dev), torch.randn(M, D, device=dev)
    ln_w, ln_b = torch.ones(D, device=dev), torch.zeros(D, device=dev)
    w1, b1 = torch.randn(C, D, device=dev) / D ** 0.5, torch.randn(C, device=dev) * 0.1
    w2, b2 = torch.randn(D, C, device=dev) / C ** 0.5, torch.randn(D, device=dev) * 0.1
    w1b, w2b = ops.cast_bf16(w1, D), ops.cast_bf16(w2, (C + 7) // 8 * 8)
    for p in (0.0, 0.5):
        y = ops.channel_mix_fwd(u, ln_w, ln_b, w1, b1, w2, b2, w1b, w2b, BF16, p, 5)
        g = ops.channel_mix_bwd(dy, u, ln_w, ln_b, w1, b1, w2, w1b, w2b, BF16, p, 5)
        torch.cuda.synchronize()
        print("channel mix p=%.1f" % p, float(y.abs().sum()), float(g[0].abs().sum()))
    x, du = torch.randn(B, N, D, device=dev), torch.randn(B, N, D, device=dev)
    tw1, tb1 = torch.randn(T, N, device=dev) / N ** 0.5, torch.randn(T, device=dev) * 0.1
    tw2, tb2 = torch.randn(N, T, device=dev) / T ** 0.5, torch.randn(N, device=dev) * 0.1
    o = ops.token_mix_fwd(x, ln_w, ln_b, tw1, tb1, tw2, tb2, BF16, 0.5, 7)
    gt = ops.token_mix_bwd(du, x, ln_w, ln_b, tw1, tb1, tw2, BF16, 0.5, 7)
    torch.cuda.synchronize()
    print("token mix", float(o.abs().sum()), float(gt[0].abs().sum()))


if __name__ == "__main__":
    main()
