"""Micro-benchmark of the generic bf16 GEMM (m2b200_gemm) on the layer shapes of BASELINE configs 4 / 5 (CUDA events, L2
flushed between iterations).   python tools/bench_gemm.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from m2_mixer_b200 import ops  # noqa: E402
from m2_mixer_b200._lib import BF16  # noqa: E402


def main():
    dev = "cuda"
    torch.manual_seed(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rn = lambda *s: torch.randn(*s, device=dev).bfloat16()

    def run(name, M, N, K, a_mn, b_mn, **kw):
        A = rn(K, M) if a_mn else rn(M, K)
        B = rn(K, N) if b_mn else rn(N, K)
        batch = kw.get("batch", 1)
        out = None
        if kw.get("splitk", 1) > 1:
            out = torch.zeros(M, N, device=dev)
            kw["out"] = out
        fn = lambda: ops.gemm(BF16, A, a_mn, B, b_mn, M, N, K, **kw)
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        med = ts[len(ts) // 2]
        print(f"{name:58s} {med * 1e3:8.1f} us  {2.0 * M * N * K * batch / med / 1e9:7.1f} TFLOP/s", flush=True)

    M = 64 * 196
    bias_n = torch.randn(3072, device=dev)
    bias_d = torch.randn(768, device=dev)
    res = torch.randn(M, 768, device=dev)
    run("C5 fwd GEMM1  [M x 3072] = Xn W1^T +b, GELU, bf16 out", M, 3072, 768, 0, 0, bias=bias_n, bias_mode=1, act=1, out_bf16=True)
    run("C5 fwd GEMM1  same, no epilogue work, bf16 out", M, 3072, 768, 0, 0, out_bf16=True)
    run("C5 bwd H      [M x 3072] fp32 out + bias", M, 3072, 768, 0, 0, bias=bias_n, bias_mode=1)
    run("C5 fwd GEMM2  [M x 768] = G W2^T + b + u, fp32 out", M, 768, 3072, 0, 0, bias=bias_d, bias_mode=1, residual=res)
    run("C5 bwd dG     [M x 3072] = dY W2 (B MN-major), fp32 out", M, 3072, 768, 0, 1)
    run("C5 bwd dXn    [M x 768] = dH W1 (B MN-major), fp32 out", M, 768, 3072, 0, 1)
    run("C5 bwd dW2    [768 x 3072] = dY^T G, split-K 2", 768, 3072, M, 1, 1, splitk=2)
    run("C5 bwd dW1    [3072 x 768] = dH^T Xn, split-K 2", 3072, 768, M, 1, 1, splitk=2)
    run("big square 8192^3 bf16 out", 8192, 8192, 8192, 0, 0, out_bf16=True)


if __name__ == "__main__":
    main()
