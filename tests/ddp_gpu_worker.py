"""Two-rank NCCL worker of tests/test_parity_gpu_r02.py::test_nccl_world2_* (launched with torch.distributed.run).

Checks, on real GPUs over NCCL (the DDP parity oracle):
  1. the averaged shard gradients of the bucketed, overlapped allreduce equal the single-process gradients of the
     concatenated global batch;
  2. eager, graphed-overlap (collectives captured on the communication stream) and graphed-split steps give the same
     per-rank loss curve with several static batches (stale bf16 weights in the graphed path);
  3. the processes tear down (graphs destroyed before the communicator).
Prints DDP_GPU_OK from rank 0."""
import os
import sys
import threading

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from m2_mixer_b200 import models, parallel, presets
    from m2_mixer_b200.graph import GraphedTrainStep
    from m2_mixer_b200.optim import FusedAdam
    from oracle.seeding import seeded_state_dict, synthetic_batch
    rank, local, world = parallel.init_from_env()
    assert world == 2
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    cfg = dict(presets.get("avmnist_S"), dropout=0.0)

    def make(precision, capturable=False, seed=3):
        m = models.AVMnistMixerMultiLoss(cfg, {}).to(dev).set_precision(precision).train()
        m.load_state_dict(seeded_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed))
        return m, FusedAdam(m.parameters(), lr=1e-2, capturable=capturable)

    # ---- 1. gradient parity (fp32 mode): shards of a global batch of 32
    full = synthetic_batch("avmnist", 32, 77)
    shard = {k: v[rank * 16:(rank + 1) * 16].to(dev) for k, v in full.items()}
    m, opt = make("fp32")
    sync = parallel.attach(opt, bucket_bytes=64 << 10)
    assert len(sync.buckets) >= 3
    opt.zero_grad()
    m.training_step(shard).backward()
    sync.finish()
    torch.cuda.synchronize()
    avg = opt.flat_grad / world
    m1, o1 = make("fp32")
    o1.zero_grad()
    m1.training_step({k: v.to(dev) for k, v in full.items()}).backward()
    torch.cuda.synchronize()
    err = float((avg - o1.flat_grad).norm() / o1.flat_grad.norm())
    assert err < 1e-5, f"rank {rank}: shard-gradient parity {err}"

    # ---- 2. loss curves: eager vs graphed (overlap / split), three static batches, bf16 mode
    batches = [{k: v[rank * 16:(rank + 1) * 16].to(dev) for k, v in synthetic_batch("avmnist", 32, 90 + i).items()} for i in range(3)]
    m, opt = make("bf16", capturable=True)
    sync = parallel.attach(opt)
    eager = []
    for i in range(9):
        opt.zero_grad(); loss = m.training_step(batches[i % 3]); loss.backward(); sync.finish(); opt.step()
        eager.append(float(loss))
    curves = {}
    for comm in ("overlap", "split"):
        m, opt = make("bf16", capturable=True)
        sync = parallel.attach(opt)
        step = GraphedTrainStep(m, opt, static_batches=batches, warmup=2, grad_sync=sync, comm=comm)
        curves[comm] = [float(step.replay(i % 3)) for i in range(9)]
        step.close()
    for comm, c in curves.items():
        d = max(abs(a - b) for a, b in zip(eager, c))
        assert d < 2e-3 * max(eager), f"rank {rank}: graphed[{comm}] vs eager {d}: {eager} {c}"

    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        print("DDP_GPU_OK", flush=True)
    # ---- 3. teardown; a hang here must not take the test (or the box) with it
    t = threading.Timer(60.0, lambda: (print("TEARDOWN_HANG", flush=True), os._exit(3)))
    t.daemon = True
    t.start()
    dist.destroy_process_group()
    t.cancel()


if __name__ == "__main__":
    main()
