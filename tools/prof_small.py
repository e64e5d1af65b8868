import sys, torch
sys.path.insert(0, '/root/repo')
from m2_mixer_b200 import _lib, models, presets
from m2_mixer_b200.optim import FusedAdam
sys.path.insert(0, '/root/repo/tools')
from bench_small import batch_for
dev = torch.device('cuda', 0)
for name, kind, B in (("mimic_H", "mimic", 128), ("mimic_H", "mimic", 4096), ("avmnist_S", "avmnist", 4096)):
    cfg = presets.get(name)
    torch.manual_seed(42)
    m = models.get_model(cfg["type"])(cfg, {}).to(dev).train()
    opt = FusedAdam(m.parameters(), lr=1e-3)
    bt = batch_for(kind, B, dev)
    def step():
        opt.zero_grad(); loss = m.training_step(bt); loss.backward(); opt.step()
    for _ in range(5): step()
    torch.cuda.synchronize()
    with _lib.profile() as prof:
        for _ in range(5): step()
        torch.cuda.synchronize()
    print(name, B)
    for k, (n, ms) in sorted(prof.table.items(), key=lambda kv: -kv[1][1]):
        print(f"   {k:28s} x{n/5:4.0f}  {ms/5*1e3:9.1f} us/step")
