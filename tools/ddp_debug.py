"""Debug aid: order of the gradient-ready notifications and bucket launches of parallel.GradSync on ONE GPU (the
collectives are replaced by a logger)."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from m2_mixer_b200 import models, parallel, presets
from m2_mixer_b200.optim import FusedAdam
from oracle.seeding import seeded_state_dict, synthetic_batch
dev = torch.device("cuda", 0)
cfg = dict(presets.get("avmnist_S"), dropout=0.0)
m = models.AVMnistMixerMultiLoss(cfg, {}).to(dev).set_precision(sys.argv[1] if len(sys.argv) > 1 else "fp32").train()
m.load_state_dict(seeded_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, 3))
opt = FusedAdam(m.parameters(), lr=1e-2)
names = {id(p): n for n, p in m.named_parameters()}
log = []
class Spy(parallel.GradSync):
    def _launch(self, b):
        self._launched[b] = True
        log.append(f"LAUNCH bucket {b}")
real = dist.get_world_size
dist.is_initialized = lambda: True
dist.get_world_size = lambda group=None: 2
sync = Spy(opt._params, opt._offsets, opt.flat_grad, 64 << 10)
for i, p in enumerate(opt._params):
    old = p._m2_ready
    p._m2_ready = (lambda o=old, q=p, i=i: (log.append(f"  direct {names[id(q)]} (bucket {sync.bucket_of[i]})"), o())[1])
    p.register_post_accumulate_grad_hook(lambda q, i=i: log.append(f"  autograd-hook {names[id(q)]} (bucket {sync.bucket_of[i]})"))
print("buckets", [(b, c) for b, (_, _, c) in enumerate(sync.buckets)])
batch = {k: v.to(dev) for k, v in synthetic_batch("avmnist", 16, 77).items()}
opt.zero_grad(); m.training_step(batch).backward(); torch.cuda.synchronize()
print("\n".join(log))
print("pending after backward", sync._pending)
