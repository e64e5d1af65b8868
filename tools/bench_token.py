"""Micro-benchmark of the token-mixing ops alone (CUDA events), M2-Mixer-B shapes.   python tools/bench_token.py [B N D T]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from m2_mixer_b200 import _lib, ops  # noqa: E402
from m2_mixer_b200._lib import BF16  # noqa: E402


def main():
    B, N, D, T = (int(v) for v in sys.argv[1:5]) if len(sys.argv) >= 5 else (4096, 4, 128, 32)
    dev = "cuda"
    torch.manual_seed(0)
    x, du = torch.randn(B, N, D, device=dev), torch.randn(B, N, D, device=dev)
    ln_w, ln_b = torch.ones(D, device=dev), torch.zeros(D, device=dev)
    w1, b1 = torch.randn(T, N, device=dev) / N ** 0.5, torch.randn(T, device=dev) * 0.1
    w2, b2 = torch.randn(N, T, device=dev) / T ** 0.5, torch.randn(N, device=dev) * 0.1
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def run(fn, name):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(20):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        print(f"{name:34s} median {ts[len(ts) // 2] * 1e3:8.1f} us  min {ts[0] * 1e3:8.1f} us")

    for p in (0.0, 0.5):
        run(lambda: ops.token_mix_fwd(x, ln_w, ln_b, w1, b1, w2, b2, BF16, p, 7), f"token_mix_fwd  B{B} N{N} p={p}")
        run(lambda: ops.token_mix_bwd(du, x, ln_w, ln_b, w1, b1, w2, BF16, p, 7), f"token_mix_bwd  B{B} N{N} p={p}")
    with _lib.profile() as pr:
        for _ in range(5):
            ops.token_mix_bwd(du, x, ln_w, ln_b, w1, b1, w2, BF16, 0.5, 7)
        torch.cuda.synchronize()
    for k, (n, ms) in sorted(pr.table.items(), key=lambda kv: -kv[1][1]):
        print(f"   {k:24s} {ms / n * 1e3:8.1f} us/launch x{n // 5}")


if __name__ == "__main__":
    main()
