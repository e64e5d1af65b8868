"""One launch of every hot kernel at the M2-Mixer-B encoder shapes (and one wide GEMM at the Scaled config's shape), twice:
the command ncu captures (`-k regex:"patch_embed|token_mix_mma|chain_.*_ts|wgrad_dh|wgrad_fused|umma_gemm2" -s 8 -c 8`)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from m2_mixer_b200 import functional as F  # noqa: E402
from m2_mixer_b200 import modules as M  # noqa: E402
from m2_mixer_b200 import ops  # noqa: E402
from m2_mixer_b200._lib import BF16  # noqa: E402

torch.manual_seed(0)
dev = "cuda"
blk = M.MixerBlock(128, 4, 32, 3072, dropout=0.5).cuda().train()
conv = torch.nn.Conv2d(1, 128, 56, 56).cuda()
audio = torch.randn(4096, 1, 112, 112, device=dev)
A = torch.randn(12544, 768, device=dev).bfloat16()
W = torch.randn(3072, 768, device=dev).bfloat16()
bias = torch.randn(3072, device=dev)
for _ in range(2):
    x = F.patch_embed(audio, conv.weight, conv.bias, 56, "bf16")
    y = blk(x)
    y.backward(torch.ones_like(y))
    ops.gemm(BF16, A, False, W, False, 12544, 3072, 768, bias=bias, bias_mode=1, act=1, out_bf16=True)
    ops.gemm(BF16, A, False, W, False, 12544, 3072, 768)
    torch.cuda.synchronize()
print("ok")
