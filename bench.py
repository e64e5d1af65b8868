"""Headline benchmark: train samples/s (fwd + bwd + optimizer step) of M2-Mixer on synthetic AV-MNIST-shaped data.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config avmnist_B] [--batch 4096]

Contract (see the task statement): one process per GPU (torchrun sets RANK/LOCAL_RANK/WORLD_SIZE), W untimed warm-up
steps, EXACTLY K timed steps bracketed by barrier + synchronize, timed with CUDA events on the launching stream, MAX
over ranks, rank 0 prints ONE JSON line.  `value` has inputs resident in HBM; `e2e` goes through the public API with
pinned host buffers (H2D of the batch and D2H of the loss inside the timed region).  `--impl reference` times the
reference's algorithm on the host CPU (the oracle port - the reference is pure Python/PyTorch and cannot travel to
the GPU box as source; kind = "port").
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "train samples/s (fwd+bwd+step)"
UNIT = "samples/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="avmnist_B")
    ap.add_argument("--batch", type=int, default=None,
                    help="per-GPU batch (weak scaling); default per config: avmnist_* 4096, mimic_H 4096, mmimdb_C4 256, scaled_C5 64")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-batch", type=int, default=None, help="bounded sample for the CPU legs (default per config)")
    ap.add_argument("--graph-comm", default="split", choices=["overlap", "split"],
                    help="N > 1, graphed step: 'split' (default: the faster one since both encoders keep every SM busy, "
                         "profiles/r02_multi_gpu.md) is one allreduce between two graphs, 'overlap' captures the bucketed NCCL "
                         "allreduces on the communication stream under the backward")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--bucket-mb", type=float, default=16.0, help="N > 1: size of the gradient allreduce buckets (MiB)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-host-dtype", default="bf16", choices=["bf16", "fp32"],
                    help="dtype of the IMAGE tensors in the pinned host batches of the e2e leg (bf16 halves the host link "
                         "bytes; the bf16 GEMM operand is the pixel rounded to bf16 either way, results are bit-identical)")
    ap.add_argument("--dropout", type=float, default=None, help="override the cfg's dropout (default: the reference cfg's value)")
    ap.add_argument("--no-graph", action="store_true", help="launch the step kernel by kernel instead of replaying one CUDA graph per step")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------- model FLOPs
DEFAULT_BATCH = {"avmnist_S": 4096, "avmnist_M": 4096, "avmnist_B": 4096, "mimic_H": 4096, "mmimdb_C4": 256, "scaled_C5": 64,
                 "mmimdb_tiny": 256}
DEFAULT_CPU_BATCH = {"avmnist_B": 128, "avmnist_S": 512, "avmnist_M": 256, "mimic_H": 512, "mmimdb_C4": 4, "scaled_C5": 1, "mmimdb_tiny": 64}


def _encoder_flops(e: dict):
    """(forward GEMM FLOPs per sample, of which channel mixing, number of tokens) of one encoder dict."""
    bt = e["block_type"]
    if bt == "MLPMixer":
        n = (e["image_size"][0] // e["patch_size"]) * (e["image_size"][1] // e["patch_size"])
        d, t, c, L = e["hidden_dim"], e["token_dim"], e["channel_dim"], e["num_mixers"]
        return 2.0 * e["in_channels"] * e["patch_size"] ** 2 * d * n + L * (4.0 * n * d * t + 4.0 * n * d * c), L * 4.0 * n * d * c, n
    if bt == "PNLPMixer":       # token_dim = channel_dim = mlp_hidden_dim (reference modules/mixer.py:249)
        n, d, h, L = e["max_seq_len"], e["hidden_dim"], e["mlp_hidden_dim"], e["num_mixers"]
        kin = (2 * e["bottleneck_window_size"] + 1) * e["bottleneck_features_size"]
        return 2.0 * kin * d * n + L * (4.0 * n * d * h + 4.0 * n * d * h), L * 4.0 * n * d * h, n
    if bt == "MLPMixerNoPatching":
        n, d, t, c, L = e["num_patch"], e["hidden_dim"], e["token_dim"], e["channel_dim"], e["num_mixers"]
        return 2.0 * e["embedding_dim"] * e["proj_dim"] * n + L * (4.0 * n * d * t + 4.0 * n * d * c), L * 4.0 * n * d * c, n
    if bt == "MLP":             # reference modules/mlp.py:4-27: Linear in, num_blocks hidden Linears, Linear out
        h = e["hidden_dim"]
        return 2.0 * (e["input_dim"] * h + e["num_blocks"] * h * h + h * e["output_dim"]), 0.0, 1
    raise ValueError(f"no FLOP model for block_type {bt}")


def model_flops(cfg: dict) -> dict:
    """Algorithmic GEMM FLOPs per sample (2*MACs), SURVEY 8(d): block fwd = 4NDT + 4NDC, patch embed fwd = 2 K D N,
    heads = 3*2*D*K; fwd+bwd = 3 x fwd."""
    m = cfg["modalities"]
    tot, chan, ntok = 0.0, 0.0, 0
    for name, e in m.items():
        if name in ("classification", "multimodal"):
            continue
        f, c, n = _encoder_flops(e)
        tot, chan, ntok = tot + f, chan + c, ntok + n
    f = m["multimodal"]
    tot += f["num_mixers"] * (4.0 * ntok * f["hidden_dim"] * f["token_dim"] + 4.0 * ntok * f["hidden_dim"] * f["channel_dim"])
    chan += f["num_mixers"] * 4.0 * ntok * f["hidden_dim"] * f["channel_dim"]
    tot += 3 * 2.0 * f["hidden_dim"] * m["classification"]["num_classes"]
    return {"fwd": tot, "fwd_bwd": 3 * tot, "channel_mix_fwd": chan}


def batch_shapes(cfg: dict, bsz: int):
    """{name: (shape, kind)} of one synthetic batch in the layout the reference data modules produce (datasets/avmnist.py:15-23,
    datasets/mimic.py:77, models/mmimdb.py:68-70); kind = 'normal' | 'label<K>' | 'multilabel'."""
    m = cfg["modalities"]
    typ = cfg["type"]
    k = m["classification"]["num_classes"]

    def enc(e):
        if e["block_type"] == "MLPMixer":
            return (bsz, e["in_channels"], *e["image_size"])
        if e["block_type"] == "PNLPMixer":
            return (bsz, e["max_seq_len"], (2 * e["bottleneck_window_size"] + 1) * e["bottleneck_features_size"])
        if e["block_type"] == "MLPMixerNoPatching":
            return (bsz, e["num_patch"], e["embedding_dim"])
        if e["block_type"] == "MLP":
            return (bsz, e["input_dim"])
        raise ValueError(e["block_type"])

    if typ == "AVMnistMixerMultiLoss":
        return {"image": (enc(m["image"]), "normal"), "audio": (enc(m["audio"]), "normal"), "label": ((bsz,), f"label{k}")}
    if typ == "MMIMDBMixerMultiLoss":
        return {"image": (enc(m["image"]), "normal"), "text": (enc(m["text"]), "normal"), "label": ((bsz, k), "multilabel")}
    if typ == "MimicMixerMultiLoss":
        return {"static": (enc(m["static"]), "normal"), "time": (enc(m["time"]), "normal"), "label": ((bsz,), f"label{k}")}
    raise ValueError(typ)


def make_batch(cfg: dict, bsz: int, device, gen):
    import torch
    out = {}
    for name, (shape, kind) in batch_shapes(cfg, bsz).items():
        if kind == "normal":
            out[name] = torch.randn(*shape, device=device, generator=gen)
        elif kind == "multilabel":
            out[name] = (torch.rand(*shape, device=device, generator=gen) < 0.1).long()
        else:
            out[name] = torch.randint(0, int(kind[5:]), shape, device=device, generator=gen)
    if cfg["type"] == "MimicMixerMultiLoss":       # the reference's MIMIC batches are tuples (datasets/mimic.py:77)
        return (out["static"], out["time"], out["label"])
    return out


def batch_bytes(batch) -> int:
    vals = batch.values() if isinstance(batch, dict) else batch
    return sum(v.numel() * v.element_size() for v in vals)


# ----------------------------------------------------------------------------------------------- CPU legs
def cpu_port_run(cfg: dict, bsz: int, steps: int, warmup: int, dropout: float):
    """The reference's algorithm (oracle port) on the host cores: fwd + bwd + Adam step, fp32, all threads."""
    import torch
    from m2_mixer_b200 import models
    from oracle import m2mixer_oracle as O
    from oracle.seeding import seeded_state_dict
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    shapes = {k: tuple(v.shape) for k, v in models.get_model(cfg["type"])(dict(cfg, dropout=0.0), {}).state_dict().items()}
    shapes.pop("pos_weight", None)
    sd = {k: v.requires_grad_(True) for k, v in seeded_state_dict(shapes, 42).items()}
    names = list(sd)
    m_ = [torch.zeros_like(sd[k]) for k in names]
    v_ = [torch.zeros_like(sd[k]) for k in names]
    batch = make_batch(cfg, bsz, "cpu", torch.Generator().manual_seed(43))
    typ = cfg["type"]
    if typ == "AVMnistMixerMultiLoss":
        fwd = lambda: O.avmnist_shared_step(sd, batch, p=dropout, training=True)                        # noqa: E731
    elif typ == "MimicMixerMultiLoss":
        fwd = lambda: O.mimic_shared_step(sd, batch, p=dropout, training=True)                          # noqa: E731
    else:
        pw = torch.tensor(cfg["pos_weight"])
        sd.pop("pos_weight", None)
        fwd = lambda: O.mmimdb_shared_step(sd, batch, pw, text_encoder=cfg["modalities"]["text"]["block_type"], p=dropout,   # noqa: E731
                                           training=True)
    names = list(sd)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        out = fwd()
        grads = O.grads_of(out["loss"], sd)
        with torch.no_grad():
            O.adam_step([sd[k] for k in names], [grads[k] for k in names], m_, v_, it + 1, lr=1e-2)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    total = sum(times)
    return {"value": bsz * steps / total, "ms_per_step": 1e3 * total / steps, "cores": cores,
            "sample": f"{steps} steps of batch {bsz} (fp32, dropout {dropout}, torch {torch.__version__} CPU, {cores} threads)"}


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    r = cpu_port_run(cfg, args.cpu_batch, args.steps, args.warmup, cfg.get("dropout", 0.0))
    note = ("BASELINE.md section 4 describes this arm as the reference modules at batch 512; /root/reference does not exist on the "
            "GPU box, so it is the oracle restatement (kind 'port', pinned to the reference by tests/golden) at the bounded batch below")
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.config} (reference algorithm on host CPU, bounded sample batch {args.cpu_batch})", "note": note},
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi polled every 20 ms from BEFORE the warm-up (its start-up alone takes ~100 ms, longer than a short timed
    region); stop(t0, t1) keeps the samples whose timestamps fall inside the timed region [t0, t1] (host wall clock) and
    falls back to every sample taken while the GPU was busy (warm-up included) if the region was too short to catch one."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    @staticmethod
    def _ts(txt: str):
        import datetime
        try:
            return datetime.datetime.strptime(txt.strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
        except ValueError:
            return None

    def stop(self, t0: float = None, t1: float = None) -> dict:
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ""
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = []
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                rows.append((self._ts(f[0]), float(f[1]), float(f[2]), [n for n, v in zip(names, f[4:8]) if v.lower().startswith("active")]))
            except ValueError:
                continue
        inside = [r for r in rows if t0 is not None and r[0] is not None and t0 <= r[0] <= t1]
        window = "timed region"
        if not inside:
            inside, window = [r for r in rows if t1 is None or r[0] is None or r[0] <= t1], "warm-up + timed region (region shorter than the sampling period)"
        sm, mx = [r[1] for r in inside], [r[2] for r in inside]
        reasons = sorted({n for r in inside for n in r[3]})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "window": window, "reasons": reasons}


# ----------------------------------------------------------------------------------------------- our arm
def run_ours(args, cfg):
    import torch
    import torch.distributed as dist
    from m2_mixer_b200 import _lib, models, parallel, presets
    from m2_mixer_b200.optim import FusedAdam

    # NCCL writes its banner ("NCCL version ...") to STDOUT when the communicator is created: keep stdout for the one
    # JSON line by pointing fd 1 at stderr while the process group comes up.
    sys.stdout.flush()
    saved_fd = os.dup(1)
    os.dup2(2, 1)
    try:
        # Captured allreduces (--graph-comm overlap) run UNDER the backward kernels, which are one persistent CTA per SM with
        # all of its shared memory and tensor memory: an NCCL kernel that spreads over many SMs delays them by more than the
        # collective takes, so its CTAs are capped (measured on 2 and 8 x B200: profiles/r02_multi_gpu.md); an explicit
        # NCCL_MAX_CTAS in the environment wins.  The split form (default) keeps NCCL's own choice.
        if int(os.environ.get("WORLD_SIZE", "1")) > 1 and args.graph_comm == "overlap" and not args.no_graph:
            os.environ.setdefault("NCCL_MAX_CTAS", "32")
        rank, local, world = parallel.init_from_env()
        if world > 1 and torch.cuda.is_available():
            t = torch.zeros(1, device=torch.device("cuda", local))
            dist.all_reduce(t)
            torch.cuda.synchronize()
    finally:
        sys.stdout.flush()
        os.dup2(saved_fd, 1)
        os.close(saved_fd)
    assert torch.cuda.is_available(), "bench.py (impl=ours) needs a GPU: the hot path has no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    B = args.batch
    ref_dropout = cfg.get("dropout", 0.0)
    run_dropout = ref_dropout if args.dropout is None else args.dropout
    cfg = dict(cfg, dropout=run_dropout)
    torch.manual_seed(42)                                         # cfg seed (reference cfg train.seed)
    model = models.get_model(cfg["type"])(cfg, dict(presets.AVMNIST_OPTIM)).to(dev).set_precision(args.precision)
    model.train()
    use_graph = not args.no_graph                    # N > 1: the bucketed NCCL allreduces are captured inside the graph (graph.py)
    lr = 1e-2 if args.config.startswith("avmnist") else (1e-4 if args.config == "scaled_C5" else 1e-3)   # cfg values; C5: tests
    opt = FusedAdam(model.parameters(), lr=lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, capturable=use_graph)
    sync = parallel.attach(opt, bucket_bytes=int(args.bucket_mb * (1 << 20))) if world > 1 else None

    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    NB = 4   # rotate distinct batches: M2-Mixer-B's is 218 MB of fp32 input, larger than the 126 MB L2
    batches = [make_batch(cfg, B, dev, g) for _ in range(NB)]
    in_bytes = batch_bytes(batches[0])

    def step(batch):
        opt.zero_grad()
        loss = model.training_step(batch)
        loss.backward()
        if sync is not None:
            sync.finish()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        barrier()
        return float(ms)

    clocks = ClockSampler(local) if rank == 0 else None
    gstep = None
    if use_graph:
        # one captured graph per resident batch (no input copies); the launch count of a step is taken from the capture
        from m2_mixer_b200.graph import GraphedTrainStep
        step(batches[0])
        lc = _lib.launch_count()
        step(batches[0])
        per_step_launches = _lib.launch_count() - lc
        try:
            if os.environ.get("M2B200_BENCH_FAIL_CAPTURE"):       # test hook for the fallback below
                raise RuntimeError("forced by M2B200_BENCH_FAIL_CAPTURE")
            gstep = GraphedTrainStep(model, opt, static_batches=batches, warmup=2, grad_sync=sync, comm=args.graph_comm)
        except Exception as e:                                    # never lose the run to the launch mode: fall back, and say so
            print(f"[bench] CUDA graph capture failed ({type(e).__name__}: {e}); running kernel by kernel", file=sys.stderr, flush=True)
            use_graph, gstep = False, None
            if sync is not None:
                sync.enabled = True
            from m2_mixer_b200 import ops as _ops
            _ops.set_dropout_epoch(None)
            torch.cuda.synchronize()
    if gstep is not None:
        run = lambda i: gstep.replay(i % NB)
    else:
        run = lambda i: step(batches[i % NB])
    for i in range(args.warmup):
        run(i)
    l0 = _lib.launch_count()
    torch.cuda.synchronize()
    w0 = time.time()
    ms = timed(run, args.steps)
    w1 = time.time()
    launches = (per_step_launches + 1) * args.steps if use_graph else _lib.launch_count() - l0   # + the epoch-advance launch
    clk = clocks.stop(w0, w1) if clocks else None
    gstep_overlap = bool(gstep is not None and gstep.overlap)
    if gstep is not None:
        gstep.close()

    # ---- end to end through the public API from pinned host memory: the batch of EVERY step is copied host -> device
    # inside the timed region (m2_mixer_b200.data.DevicePrefetcher double-buffers it on a side stream, one batch ahead)
    # and the loss of every step is read back to the host.
    e2e = None
    if not args.no_e2e:
        from m2_mixer_b200.data import DevicePrefetcher
        as_dict = lambda b: b if isinstance(b, dict) else {str(i): v for i, v in enumerate(b)}   # noqa: E731  (MIMIC: tuples)
        from_dict = lambda d: d if isinstance(batches[0], dict) else tuple(d[str(i)] for i in range(len(d)))   # noqa: E731
        from m2_mixer_b200.data import pin_host_batch
        host = [pin_host_batch(as_dict(b), image_dtype=torch.bfloat16 if args.e2e_host_dtype == "bf16" else None)
                for b in batches[:2]]
        sink = []

        def host_stream(n):
            for i in range(n):
                yield host[i % 2]

        pre = DevicePrefetcher(host_stream(2), dev)
        estep = None
        for b in pre:                                             # warm the copy path / allocator
            sink.append(float(step(from_dict(b)).detach()))
        if use_graph:
            # the prefetcher's two device buffer sets are the static inputs of two captured graphs
            from m2_mixer_b200.graph import GraphedTrainStep
            try:
                estep = GraphedTrainStep(model, opt, static_batches=[from_dict(b) for b in pre.bufs], warmup=1, grad_sync=sync,
                                         comm=args.graph_comm)
            except Exception as e:
                print(f"[bench] CUDA graph capture failed in the e2e leg ({type(e).__name__}: {e}); running kernel by kernel",
                      file=sys.stderr, flush=True)
                estep = None
                if sync is not None:
                    sync.enabled = True
                from m2_mixer_b200 import ops as _ops
                _ops.set_dropout_epoch(None)
                torch.cuda.synchronize()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pre = DevicePrefetcher(host_stream(args.steps), dev, bufs=pre.bufs if estep is not None else None)   # first copy is issued (and waited for) inside the region
        pending = None
        for b in pre:
            if estep is not None:
                loss = estep.replay(pre.index_of(b)).clone()      # the graph's loss slot is overwritten two steps later
            else:
                loss = step(from_dict(b)).detach()
            if pending is not None:
                sink.append(float(pending))                       # D2H read of the previous step's loss: no pipeline bubble
            pending = loss
        sink.append(float(pending))
        e1.record()
        torch.cuda.synchronize()
        ems_t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ems_t, op=dist.ReduceOp.MAX)
        barrier()
        ems = float(ems_t)
        if estep is not None:
            estep.close()
        assert len(sink) == 2 + args.steps
        h2d = sum(v.numel() * v.element_size() for v in host[0].values())
        e2e = {"value": world * B * args.steps / (ems / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
               "ms_per_step": ems / args.steps,
               "host_image_dtype": args.e2e_host_dtype,
               "note": "H2D of each step's batch from pinned memory (double-buffered on a copy stream) + D2H of each step's loss; "
                       "host image tensors are " + ("bf16 (the patch-embedding GEMM rounds pixels to bf16 anyway: same bits as "
                       "fp32 host batches, half the link bytes; --e2e-host-dtype fp32 for the other form)"
                       if args.e2e_host_dtype == "bf16" else "fp32")}

    # ---- per-kernel device time (CUDA events on the launching stream) for the roofline of the dominant kernel
    roof = None
    # the per-kernel times are CUDA-event pairs around each launch: they only mean "this kernel's duration" when nothing else
    # runs beside it, so the profiled steps keep both encoders on one stream (the timed steps above fork them)
    two_streams = os.environ.get("M2B200_BRANCH_STREAMS", "1") != "0"
    os.environ["M2B200_BRANCH_STREAMS"] = "0"
    if rank != 0:
        for i in range(3):                                        # keep the collectives of rank 0's profiled steps matched
            step(batches[i % NB])
        torch.cuda.synchronize()
    if rank == 0:
        with _lib.profile() as prof:
            for i in range(3):
                step(batches[i % NB])
            torch.cuda.synchronize()
        table = prof.table
        mm = cfg["modalities"]
        per_sample_chain = model_flops(cfg)["channel_mix_fwd"]          # 4*N*D*C summed over every block
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = peaks.get("bf16_tflops_sustained")
        peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)"
        if peak is None:
            peak, peak_src = 1400.0, "fallback (B200_PROFILING.md sustained figure)"
        burst = peaks.get("bf16_tflops", 1644.4)
        # the three tcgen05 kernels of a channel-mixing block; each carries 4*M*D*C ALGORITHMIC FLOPs per launch:
        #   chain_fwd   : the two forward GEMMs
        #   chain_bwd   : dG = dY W2 and dXn = dH W1            (its recomputed H GEMM is not counted)
        #   wgrad_dh    : dW1 = dH^T LN(u) and dW2 = dY^T G      (its recomputed H GEMM and the ones-GEMM of db1 are not counted;
        #                 wgrad_fused with M2B200_CHAIN_GEN=2: recomputes H and dG)
        cands = {k: v for k, v in table.items() if k in ("chain_fwd", "chain_bwd", "wgrad_fused", "wgrad_dh")}
        share = {k: round(v[1] / 3, 4) for k, v in sorted(table.items(), key=lambda kv: -kv[1][1])}
        if cands:
            name = max(cands, key=lambda k: cands[k][1])
            n, tot_ms = cands[name]
            flops = per_sample_chain * B * 3                            # 3 profiled steps
            ach = flops / (tot_ms / 1e3) / 1e12
            traffic = None
            tpath = os.path.join(ROOT, "profiles", "traffic.json")
            if os.path.exists(tpath):
                traffic = json.load(open(tpath)).get(name)
            chain_ms = sum(v[1] for v in cands.values())
            roof = {"kernel": name, "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                    "frac_of_burst_peak": ach / burst, "burst_peak": burst,
                    "traffic": traffic, "launches": n, "avg_launch_ms": tot_ms / n, "peak_source": peak_src,
                    "algorithmic_flops_per_launch": "4*M*D*C (recomputed GEMMs not counted)",
                    "all_chain_kernels": {"achieved": 3 * flops / (chain_ms / 1e3) / 1e12, "frac": 3 * flops / (chain_ms / 1e3) / 1e12 / peak,
                                          "frac_of_burst_peak": 3 * flops / (chain_ms / 1e3) / 1e12 / burst,
                                          "note": "fwd + dgrad + wgrad together: 12*M*D*C algorithmic FLOPs over their summed time"},
                    "step_share": share}
        else:
            # configurations whose channel mixing runs on the generic GEMM (D > 128): the dominant kernel is the GEMM itself;
            # achieved = the step's algorithmic GEMM FLOPs over the summed time of every umma_gemm launch
            gemm = {k: v for k, v in table.items() if k.startswith("umma_gemm")}
            if gemm:
                n = sum(v[0] for v in gemm.values())
                tot_ms = sum(v[1] for v in gemm.values())
                flops = model_flops(cfg)["fwd_bwd"] * B * 3
                ach = flops / (tot_ms / 1e3) / 1e12
                roof = {"kernel": "umma_gemm (all instances)", "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                        "frac": ach / peak, "frac_of_burst_peak": ach / burst, "burst_peak": burst, "traffic": None, "launches": n,
                        "avg_launch_ms": tot_ms / n, "peak_source": peak_src,
                        "algorithmic_flops_per_launch": "the step's GEMM FLOPs (SURVEY 8d, fwd+bwd = 3 x fwd) over all GEMM launches",
                        "step_share": share}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_port_run(dict(cfg, dropout=ref_dropout), args.cpu_batch, 2, 1, ref_dropout)
        cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]}

    if rank == 0:
        fl = model_flops(cfg)
        value = world * B * args.steps / (ms / 1e3)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": args.precision, "data": "synthetic",
                "config": {"workload": f"{args.config} ({cfg['type']}), per-GPU batch {B}, fwd+bwd+FusedAdam",
                           "global_batch": world * B, "parallelism": f"dp{world}",
                           "l2": f"{NB} rotating input batches of {in_bytes >> 20} MiB each" +
                                 (" (> 126 MB L2)" if in_bytes > (126 << 20) else
                                  f"; parameters + Adam state + activations of a step are {4 * 4 * sum(p.numel() for p in model.parameters()) >> 20}+ MiB, beyond the 126 MB L2"),
                           "dropout": f"{run_dropout} (reference cfg: {ref_dropout}; fused counter-based masks, regenerated in backward)",
                           "launch": (("one CUDA graph replay per step (m2_mixer_b200.graph.GraphedTrainStep)" +
                                       ("" if world == 1 else ("; gradient allreduce: bucketed NCCL collectives captured on the communication stream, "
                                                               "overlapped with backward" if (gstep_overlap) else
                                                               "; gradient allreduce: one NCCL call between two graphs")))
                                      if use_graph else "kernel by kernel" + ("" if world == 1 else "; bucketed NCCL allreduce overlapped with backward")) +
                                     ("; the two encoders run on two streams (fork / join inside the graph)" if two_streams else ""),
                           "model_tflops_per_gpu": fl["fwd_bwd"] * value / world / 1e12,
                           "frac_of_bf16_peak_burst": fl["fwd_bwd"] * value / world / 1e12 / 1644.4,
                           "frac_of_bf16_peak_sustained": fl["fwd_bwd"] * value / world / 1e12 / 1402.3},
                "gpu_launches": int(launches), "clocks": clk, "e2e": e2e, "roofline": roof, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    if world > 1:
        # the JSON line is out: a teardown that hangs (NCCL communicator destruction) must not hold the run hostage
        import threading
        t = threading.Timer(30.0, lambda: os._exit(0))
        t.daemon = True
        t.start()
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()
        t.cancel()


def main():
    args = parse()
    from m2_mixer_b200 import presets
    cfg = presets.get(args.config)
    if args.batch is None:
        args.batch = DEFAULT_BATCH.get(args.config, 4096)
    if args.cpu_batch is None:
        args.cpu_batch = DEFAULT_CPU_BATCH.get(args.config, 128)
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
