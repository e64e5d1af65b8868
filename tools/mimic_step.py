"""A few MIMIC-H training steps at the cfg's batch (for `ncu --metrics gpu__time_duration.sum`: which of the 46 launches cost what)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from m2_mixer_b200 import models, presets
from m2_mixer_b200.optim import FusedAdam
from bench_small import batch_for
dev = torch.device("cuda", 0)
cfg = presets.get("mimic_H")
torch.manual_seed(42)
m = models.get_model(cfg["type"])(cfg, {}).to(dev).train()
opt = FusedAdam(m.parameters(), lr=1e-3)
bt = batch_for("mimic", int(sys.argv[1]) if len(sys.argv) > 1 else 128, dev)
for _ in range(4):
    opt.zero_grad(); loss = m.training_step(bt); loss.backward(); opt.step()
torch.cuda.synchronize()
print("ok", float(loss))
