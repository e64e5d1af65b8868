"""world_size-2 gloo test of the data-parallel host logic (bucketing, hook-driven allreduce, 1/world scaling)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from m2_mixer_b200.parallel import GradSync
    torch.manual_seed(0)                      # identical replicas
    shapes = [(7, 5), (5,), (3, 7), (3,), (11,)]
    offs, n = [], 0
    for s in shapes:
        offs.append(n)
        n += (int(torch.tensor(s).prod()) + 3) // 4 * 4
    flat_p, flat_g = torch.randn(n), torch.zeros(n)
    params = []
    for s, o in zip(shapes, offs):
        numel = int(torch.tensor(s).prod())
        p = torch.nn.Parameter(flat_p[o:o + numel].view(s))
        p.grad = flat_g[o:o + numel].view(s)
        params.append(p)
    sync = GradSync(params, offs, flat_g, bucket_bytes=64)       # tiny buckets -> several of them
    assert len(sync.buckets) >= 3
    # shard-local loss on this rank's half of a global batch of 8
    g = torch.Generator().manual_seed(123)
    X = torch.randn(8, 5, generator=g)
    xs = X[rank * 4:(rank + 1) * 4]
    h = torch.tanh(xs @ params[0].t() + params[1][None, :5].sum() * 0 + params[1].sum())
    loss = ((h @ params[2].t() + params[3]) ** 2).mean() + params[4].sum() * 0.0
    loss.backward()
    sync.finish()
    avg = flat_g / world
    # oracle: single process, global batch mean
    ps = [p.detach().clone().requires_grad_(True) for p in params]
    hh = torch.tanh(X @ ps[0].t() + ps[1].sum())
    full = ((hh @ ps[2].t() + ps[3]) ** 2).mean()
    full.backward()
    ok = True
    for p, r, s, o in zip(params, ps, shapes, offs):
        numel = r.numel()
        got = avg[o:o + numel].view(s)
        want = r.grad if r.grad is not None else torch.zeros_like(r)
        ok &= bool(torch.allclose(got, want, atol=1e-6))
    # second iteration works too (state reset)
    flat_g.zero_()
    (params[0].sum() * (rank + 1)).backward()
    sync.finish()
    ok &= bool(torch.allclose(flat_g[offs[0]:offs[0] + 35], torch.full((35,), 3.0)))
    # graphed-step mode (graph.GraphedTrainStep with a process group): the hooks are switched off, finish() is a no-op and
    # ONE allreduce over the whole flat buffer runs between the two captured graphs
    sync.enabled = False
    flat_g.zero_()
    (params[0].sum() * (rank + 1) + params[4].sum() * (2 * rank + 1)).backward()
    sync.finish()
    ok &= bool(torch.allclose(flat_g[offs[0]:offs[0] + 35], torch.full((35,), float(rank + 1))))     # nothing reduced yet
    sync.allreduce_all()
    ok &= bool(torch.allclose(flat_g[offs[0]:offs[0] + 35], torch.full((35,), 3.0)))
    ok &= bool(torch.allclose(flat_g[offs[4]:offs[4] + 11], torch.full((11,), 4.0)))
    sync.enabled = True                       # and back: the bucketed path works again
    flat_g.zero_()
    (params[2].sum() * (rank + 1)).backward()
    sync.finish()
    ok &= bool(torch.allclose(flat_g[offs[2]:offs[2] + 21], torch.full((21,), 3.0)))
    # the CUDA kernels that accumulate straight into the flat buffer report a parameter through p._m2_ready, and torch
    # ALSO runs the post-accumulate hook of that parameter (seen on 2 x B200): a parameter must count once per step, or
    # its bucket is reduced before the later gradients of the bucket exist.  Here: parameter 3 (the last one of its bucket
    # in backward order) reports twice before parameter 2's gradient has been produced.
    flat_g.zero_()
    b3 = sync.bucket_of[3]
    assert sync.bucket_of[4] == b3 and sync.buckets[b3][2] == 2
    params[4]._m2_ready(); params[4]._m2_ready()
    ok &= not sync._launched[b3]
    h = torch.tanh(xs @ params[0].t() + params[1].sum())
    ((h @ params[2].t() + params[3]) ** 2).mean().backward()
    ok &= all(v >= 0 for v in sync._pending)
    sync.finish()
    ok &= bool(torch.allclose(flat_g[offs[3]:offs[3] + 3] / world, ps[3].grad, atol=1e-6))
    ok &= bool(torch.allclose(flat_g[offs[2]:offs[2] + 21].view(3, 7) / world, ps[2].grad, atol=1e-6))
    q.put((rank, ok))
    dist.destroy_process_group()


def test_gradsync_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res
