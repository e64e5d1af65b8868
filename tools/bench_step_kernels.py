"""One M2-Mixer-B block fwd+bwd at B=4096 (for ncu captures of the non-chain kernels)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from m2_mixer_b200 import modules as M  # noqa: E402

torch.manual_seed(0)
blk = M.MixerBlock(128, 4, 32, 3072).cuda()
x = torch.randn(4096, 4, 128, device="cuda", requires_grad=True)
for _ in range(3):
    y = blk(x)
    y.backward(torch.ones_like(y))
torch.cuda.synchronize()
print("ok")
