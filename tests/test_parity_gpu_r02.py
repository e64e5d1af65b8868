"""Round-2 parity cases (needs a B200, run with -m gpu): the layer shapes of BASELINE configs 4 / 5, every gradient element
of M2-Mixer-B, the loss-curve criterion with dropout ON against a band of oracle seeds, optimizer checkpointing / frozen
parameters, and the graphed step with several static batches.

Tolerances (BASELINE.json north_star): fp32 mode <= 1e-4 relative on outputs and gradients; bf16 mode <= 2e-2 on outputs
(gradients: 3e-2, no stated bar)."""
import os

import pytest
import torch

from tests.golden_util import BLOCK_KEYS, block_inputs, full_grad, load, rebuild, rel_err

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name", ["block_c5", "block_c4text", "block_c4fus"])
@pytest.mark.parametrize("precision,tol,gtol", [("fp32", 1e-4, 1e-4), ("bf16", 2e-2, 3e-2)])
def test_large_config_blocks_golden(name, precision, tol, gtol):
    """One MixerBlock at the C5 encoder shape (N=196, D=768, T=384, C=3072) and at C4's text (N=T=C=512, D=256) and fusion
    (N=708, T=16) shapes against the unmodified reference (modules/mixer.py:25-47, 232-264): forward, input gradient and
    EVERY parameter gradient (norm + a strided sample of 4096 elements each)."""
    from m2_mixer_b200 import modules as M
    from oracle.seeding import seeded_state_dict
    z = load(name)
    B, N, D, T, C = (int(v) for v in z["meta.dims"])
    blk = M.MixerBlock(D, N, T, C).cuda()
    blk.precision = precision
    blk.load_state_dict(seeded_state_dict(BLOCK_KEYS(N, D, T, C), 77))
    x, dy = block_inputs(z, torch.float32)
    x = x.cuda().requires_grad_(True)
    y = blk(x)
    y.backward(dy.cuda())
    torch.cuda.synchronize()
    assert rel_err(y, z["y"]) < tol
    assert rel_err(x.grad, z["dx"]) < gtol
    for k, p in blk.named_parameters():
        gn = float(z["gnorm." + k])
        if gn < 1e-9:   # token_mix.2.net.3.bias: exactly zero in exact arithmetic (removed by the next LayerNorm)
            continue
        assert abs(float(p.grad.norm()) - gn) < gtol * gn, (k, float(p.grad.norm()), gn)
        samp = p.grad.flatten()[::int(z["gstride." + k])][:4096]
        assert rel_err(samp, z["gsamp." + k]) < (gtol if precision == "fp32" else 5e-2), k


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-3), ("bf16", 3e-2)])
def test_m2_mixer_b_every_gradient_element(precision, tol):
    """M2-Mixer-B (cfg/avmnist/avmnist_m2-mixer_B.yml) at batch 2: EVERY element of EVERY parameter gradient against the
    reference run (stored as scaled float16, 2^-11 steps: the fp32 bar of 1e-4 is checked on the exact norms + heads here
    and in test_fp32_mode_matches_reference_golden[avmnist_B_b16])."""
    from m2_mixer_b200 import models, presets
    cfg = dict(presets.get("avmnist_B"), dropout=0.0)
    m = models.AVMnistMixerMultiLoss(cfg, {}).cuda().set_precision(precision).train()
    z, sd, batch = rebuild("avmnist_B_b2", torch.float32, "cuda", requires_grad=False)
    m.load_state_dict(sd, strict=True)
    out = m.shared_step(batch, mode="train")
    out["loss"].backward()
    torch.cuda.synchronize()
    for k in ("logits", "image_logits", "audio_logits", "loss"):
        assert rel_err(out[k], z["out." + k]) < (1e-4 if precision == "fp32" else 2e-2), k
    num = den = 0.0
    for k, p in m.named_parameters():
        ref = full_grad(z, k)
        gn = float(z["gnorm." + k])
        if gn < 1e-9:
            assert float(p.grad.norm()) < 1e-5, k
            continue
        if precision == "fp32":
            assert abs(float(p.grad.norm()) - gn) < 1e-4 * gn, k
            assert rel_err(p.grad.flatten()[:16], z["ghead." + k]) < 1e-3, k
        e = rel_err(p.grad, ref)
        assert e < (tol if precision == "fp32" else 6e-2), (k, e)      # per tensor
        num += float((p.grad.double().cpu() - ref).pow(2).sum())
        den += float(ref.pow(2).sum())
    assert (num / den) ** 0.5 < tol                                      # all 8.3 M elements together


def _c5_small():
    from m2_mixer_b200 import presets
    cfg = presets.get("scaled_C5")
    for k in ("image", "audio", "multimodal"):
        cfg["modalities"][k]["num_mixers"] = 2      # the full 12 + 12 + 12 blocks are run by the training test below
    return cfg


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 2e-2)])
def test_c5_shaped_model_matches_oracle(precision, tol):
    """The Scaled config's layer shapes end to end (two 3x224x224 / p16 encoders -> 392 fused tokens, D=768, T=384, C=3072,
    two blocks per stack) against the CPU oracle: logits, losses and the gradients of the first / last layers."""
    from m2_mixer_b200 import models
    from oracle import m2mixer_oracle as O
    from oracle.seeding import seeded_state_dict
    cfg = _c5_small()
    m = models.AVMnistMixerMultiLoss(cfg, {}).cuda().set_precision(precision).train()
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    sd32 = seeded_state_dict(shapes, 31)
    m.load_state_dict(sd32)
    g = torch.Generator().manual_seed(32)
    batch = {"image": torch.randn(2, 3, 224, 224, generator=g), "audio": torch.randn(2, 3, 224, 224, generator=g),
             "label": torch.randint(0, 10, (2,), generator=g)}
    out = m.shared_step({k: v.cuda() for k, v in batch.items()}, mode="train")
    out["loss"].backward()
    torch.cuda.synchronize()
    sd64 = {k: v.double().requires_grad_(True) for k, v in sd32.items()}
    ref = O.avmnist_shared_step(sd64, {k: (v.double() if v.is_floating_point() else v) for k, v in batch.items()})
    gref = O.grads_of(ref["loss"], sd64)
    for k in ("loss", "loss_image", "loss_audio", "loss_fusion", "logits", "image_logits", "audio_logits"):
        assert rel_err(out[k], ref[k]) < tol, k
    gtol = tol if precision == "fp32" else 5e-2
    for k, p in m.named_parameters():
        if float(gref[k].norm()) > 1e-9:
            assert rel_err(p.grad, gref[k]) < gtol, (k, rel_err(p.grad, gref[k]))


def test_c4_model_matches_oracle_bf16_and_fp32():
    """BASELINE config 4 (MM-IMDB-shaped) at its real layer shapes, batch 2: image 196 tokens (T=16), PNLPMixer
    text 512 tokens with T = C = 512, 708 fused tokens, BCE-with-pos-weight x 3 (reference models/mmimdb.py:97-147)."""
    from m2_mixer_b200 import models, presets
    from oracle import m2mixer_oracle as O
    from oracle.seeding import seeded_state_dict, synthetic_batch
    cfg = presets.get("mmimdb_C4")
    mm = cfg["modalities"]
    batch = synthetic_batch(("mmimdb", mm["image"], mm["text"]), 2, 5)
    pw = torch.tensor(cfg["pos_weight"], dtype=torch.float64)
    ref = gref = None
    for precision, tol in (("fp32", 1e-4), ("bf16", 2e-2)):
        m = models.MMIMDBMixerMultiLoss(cfg, {}).cuda().set_precision(precision).train()
        shapes = {k: tuple(v.shape) for k, v in m.state_dict().items() if k != "pos_weight"}
        sd32 = seeded_state_dict(shapes, 41)
        m.load_state_dict(sd32, strict=False)
        if ref is None:
            sd64 = {k: v.double().requires_grad_(True) for k, v in sd32.items()}
            ref = O.mmimdb_shared_step(sd64, {k: (v.double() if v.is_floating_point() else v) for k, v in batch.items()}, pw,
                                       text_encoder="PNLPMixer")
            gref = O.grads_of(ref["loss"], sd64)
        out = m.shared_step({k: v.cuda() for k, v in batch.items()}, mode="train")
        out["loss"].backward()
        torch.cuda.synchronize()
        for k in ("loss", "logits", "image_logits", "text_logits"):
            assert rel_err(out[k], ref[k]) < tol, (precision, k, rel_err(out[k], ref[k]))
        gtol = tol if precision == "fp32" else 5e-2
        for k, p in m.named_parameters():
            if float(gref[k].norm()) > 1e-9:
                assert rel_err(p.grad, gref[k]) < gtol, (precision, k, rel_err(p.grad, gref[k]))


def test_scaled_c5_full_depth_trains():
    """The full Scaled C5 model (12 + 12 + 12 blocks, 178.6 M parameters), bf16, batch 8, a fixed batch, Adam: the loss falls.
    Round 1's tool printed a RISING loss (6.91 -> 8.67 over 7 steps, profiles/r01_large_configs.json): that run used Adam at
    lr 1e-3 on 178 M freshly initialised parameters - every element moves by ~lr per step whatever its gradient, which at
    this depth overshoots; the gradients themselves are pinned by test_c5_shaped_model_matches_oracle and
    test_large_config_blocks_golden.  At lr 1e-4 the same model trains."""
    from m2_mixer_b200 import models, presets
    from m2_mixer_b200.optim import FusedAdam
    cfg = presets.get("scaled_C5")
    torch.manual_seed(0)
    m = models.AVMnistMixerMultiLoss(cfg, {}).cuda().set_precision("bf16").train()
    opt = FusedAdam(m.parameters(), lr=1e-4)
    g = torch.Generator(device="cuda").manual_seed(1)
    batch = {"image": torch.randn(8, 3, 224, 224, device="cuda", generator=g),
             "audio": torch.randn(8, 3, 224, 224, device="cuda", generator=g),
             "label": torch.randint(0, 10, (8,), device="cuda", generator=g)}
    losses = []
    for _ in range(12):
        opt.zero_grad()
        loss = m.training_step(batch)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert all(l == l for l in losses), losses
    assert losses[-1] < 0.5 * losses[0], losses
    del m, opt
    torch.cuda.empty_cache()


def test_loss_curve_with_dropout_inside_oracle_seed_band():
    """north_star / SURVEY 8(c): loss curves over 200 Adam steps within run-to-run noise, the band taken from >= 5 oracle
    seeds with DROPOUT ON (the cfg's p = 0.1 for M2-Mixer-S, cfg/avmnist/avmnist_m2-mixer_S.yml).  The oracle restatement
    (torch dropout, 5 different mask seeds, same weights / batches / lr) runs on this GPU as the checker; the CUDA path's
    masks are a sixth draw.  Compared on 20-step window means: |ours - band mean| <= 4 sigma + 2 % of the mean."""
    from m2_mixer_b200 import models, presets
    from m2_mixer_b200.optim import FusedAdam
    from oracle import m2mixer_oracle as O
    from oracle.seeding import seeded_state_dict, synthetic_batch
    cfg = presets.get("avmnist_S")
    p = cfg["dropout"]
    assert p > 0
    batches = [{k: v.cuda() for k, v in synthetic_batch("avmnist", 64, 200 + i).items()} for i in range(8)]
    shapes = None
    ours = {}
    for prec in ("bf16", "fp32"):
        m = models.AVMnistMixerMultiLoss(cfg, {}).cuda().set_precision(prec).train()
        shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        m.load_state_dict(seeded_state_dict(shapes, 7))
        opt = FusedAdam(m.parameters(), lr=1e-2)
        torch.manual_seed(1000)
        ls = []
        for step in range(200):
            opt.zero_grad()
            loss = m.training_step(batches[step % 8])
            loss.backward()
            opt.step()
            ls.append(loss.detach())
        ours[prec] = torch.stack(ls).cpu().double()
    band = []
    for seed in range(5):
        sd = {k: v.cuda().requires_grad_(True) for k, v in seeded_state_dict(shapes, 7).items()}
        ropt = torch.optim.Adam(list(sd.values()), lr=1e-2)
        torch.manual_seed(seed)
        ls = []
        for step in range(200):
            ropt.zero_grad()
            loss = O.avmnist_shared_step(sd, batches[step % 8], p=p, training=True)["loss"]
            loss.backward()
            ropt.step()
            ls.append(loss.detach())
        band.append(torch.stack(ls).cpu().double())
    band = torch.stack(band)                                  # [5, 200]
    win = lambda c: c.reshape(*c.shape[:-1], 10, 20).mean(-1)  # noqa: E731  20-step window means
    bw = win(band)
    mu, sd_ = bw.mean(0), bw.std(0)
    assert float(mu[-1]) < 0.6 * float(mu[0])                 # the oracle itself trains
    for prec, c in ours.items():
        dev = (win(c) - mu).abs()
        lim = 4 * sd_ + 0.02 * mu + 0.01
        assert bool((dev <= lim).all()), (prec, dev.tolist(), lim.tolist())


def test_fused_adam_state_dict_round_trip_and_frozen_parameters():
    """torch.optim.Adam's checkpoint layout in both directions (the reference's optimizer, models/avmnist.py:413-415), and
    frozen parameters (requires_grad False, models/avmnist.py:243-256) stay EXACTLY where they are."""
    from m2_mixer_b200 import models, presets
    from m2_mixer_b200.optim import FusedAdam
    from oracle.seeding import synthetic_batch
    cfg = dict(presets.get("avmnist_S"), dropout=0.0)
    bt = {k: v.cuda() for k, v in synthetic_batch("avmnist", 16, 3).items()}

    def make_model():
        torch.manual_seed(0)
        return models.AVMnistMixerMultiLoss(cfg, {}).cuda().set_precision("fp32").train()

    def make():
        m = make_model()
        return m, FusedAdam(m.parameters(), lr=1e-2, weight_decay=0.01)

    def run(m, opt, n):
        out = []
        for _ in range(n):
            opt.zero_grad(); loss = m.training_step(bt); loss.backward(); opt.step()
            out.append(float(loss))
        return out

    m1, o1 = make()
    a = run(m1, o1, 6)
    m2, o2 = make()
    run(m2, o2, 3)
    ck_model, ck_opt = {k: v.clone() for k, v in m2.state_dict().items()}, o2.state_dict()
    assert len(ck_opt["state"]) == len(list(m2.parameters())) and float(ck_opt["state"][0]["step"]) == 3
    m3, o3 = make()
    m3.load_state_dict(ck_model)
    o3.load_state_dict(ck_opt)
    b = run(m3, o3, 3)
    assert max(abs(x - y) for x, y in zip(a[3:], b)) < 1e-6 * max(a), (a, b)
    # a torch.optim.Adam checkpoint loads too, and the resumed curves agree
    m4 = make_model()       # plain parameters: no FusedAdam has flattened them
    t4 = torch.optim.Adam(m4.parameters(), lr=1e-2, weight_decay=0.01)
    for _ in range(3):
        t4.zero_grad(); m4.training_step(bt).backward(); t4.step()
    m5, o5 = make()
    m5.load_state_dict(m4.state_dict())
    o5.load_state_dict(t4.state_dict())
    c = run(m5, o5, 3)
    assert max(abs(x - y) for x, y in zip(a[3:], c)) < 1e-4 * max(a), (a, c)
    # frozen encoder: bit-identical after further steps, the rest keeps training
    for prm in m5.image_mixer.parameters():
        prm.requires_grad_(False)
    frozen = {k: v.clone() for k, v in m5.image_mixer.state_dict().items()}
    other = m5.audio_mixer.mixer_blocks[0].channel_mix[1].net[0].weight.clone()
    run(m5, o5, 3)
    for k, v in m5.image_mixer.state_dict().items():
        assert torch.equal(v, frozen[k]), k
    assert not torch.equal(other, m5.audio_mixer.mixer_blocks[0].channel_mix[1].net[0].weight)


def test_graphed_step_with_several_static_batches_uses_current_weights():
    """with more than one static batch every captured graph must see the weights of the previous optimizer
    step - the bf16 operand copies are refreshed right behind the Adam kernel, not lazily by whichever graph captured it."""
    from m2_mixer_b200 import models, presets
    from m2_mixer_b200.graph import GraphedTrainStep
    from m2_mixer_b200.optim import FusedAdam
    from oracle.seeding import synthetic_batch
    cfg = dict(presets.get("avmnist_S"), dropout=0.0)
    batches = [{k: v.cuda() for k, v in synthetic_batch("avmnist", 32, 50 + i).items()} for i in range(3)]

    def make():
        torch.manual_seed(0)
        m = models.AVMnistMixerMultiLoss(cfg, {}).cuda().set_precision("bf16").train()
        return m, FusedAdam(m.parameters(), lr=1e-2, capturable=True)

    m, opt = make()
    eager = []
    for i in range(12):
        opt.zero_grad(); loss = m.training_step(batches[i % 3]); loss.backward(); opt.step()
        eager.append(float(loss))
    m, opt = make()
    step = GraphedTrainStep(m, opt, static_batches=batches, warmup=2)
    graphed = [float(step.replay(i % 3)) for i in range(12)]
    step.close()
    assert max(abs(a - b) for a, b in zip(eager, graphed)) < 2e-3 * max(eager), (eager, graphed)
    # lr changed by a scheduler on the host reaches the captured step without a re-capture
    m, opt = make()
    step = GraphedTrainStep(m, opt, static_batches=batches[:1], warmup=1)
    w_init = opt.flat_param.clone()
    step.replay(0)
    torch.cuda.synchronize()
    w0 = opt.flat_param.clone()
    assert not torch.equal(w_init, w0)
    opt.param_groups[0]["lr"] = 0.0
    step.replay(0)
    step.replay(0)
    torch.cuda.synchronize()
    assert torch.equal(w0, opt.flat_param)       # lr = 0 from the next replay on: the parameters no longer move
    step.close()


def _run_world2(script_args, timeout=600):
    import subprocess
    import sys
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(ROOT, "tests", "ddp_gpu_worker.py"), *script_args]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (run under gpurun --gpus 2)")
def test_nccl_world2_shard_gradients_equal_single_process_large_batch():
    """the DDP parity oracle on real GPUs over NCCL - the averaged shard gradients of a 2-rank run equal the
    single-process gradients of the concatenated batch; eager (bucketed, overlapped) and graphed (overlap and split) steps
    produce the same loss curve, with several static batches."""
    r = _run_world2([])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "DDP_GPU_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_bf16_host_images_give_the_same_bits_as_fp32_images():
    """data.pin_host_batch(image_dtype=bf16) + DevicePrefetcher: the bf16-mode patch embedding (reference Conv2d,
    modules/mixer.py:143-146) rounds every pixel to bf16 as its first operation, so a training step fed bf16 images must
    reproduce the fp32-fed step bit for bit - logits, loss and the patch-embedding weight gradients (M2-Mixer-B: the audio
    branch takes the gather-inside-the-GEMM kernels of patch_gemm.cu, the 28x28 image branch the gather + GEMM fallback)."""
    from m2_mixer_b200 import models, presets
    from m2_mixer_b200.data import DevicePrefetcher, pin_host_batch
    from oracle.seeding import seeded_state_dict, synthetic_batch
    cfg = dict(presets.get("avmnist_B"), dropout=0.0)
    m = models.AVMnistMixerMultiLoss(cfg, {}).cuda().set_precision("bf16").train()
    m.load_state_dict(seeded_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, 5))
    batch = synthetic_batch("avmnist", 24, 5)
    res = []
    for dt in (None, torch.bfloat16):
        host = pin_host_batch(batch, image_dtype=dt)
        dev = next(iter(DevicePrefetcher([host], torch.device("cuda"))))
        assert dev["audio"].dtype == (torch.float32 if dt is None else torch.bfloat16)
        m.zero_grad()
        out = m.shared_step(dev, mode="train")
        out["loss"].backward()
        torch.cuda.synchronize()
        res.append((out["logits"].clone(), out["loss"].clone(),
                    m.audio_mixer.to_patch_embedding[0].weight.grad.clone(), m.image_mixer.to_patch_embedding[0].weight.grad.clone()))
    assert torch.equal(res[0][0], res[1][0])
    assert abs(float(res[0][1]) - float(res[1][1])) < 1e-6 * float(res[0][1])   # the loss mean is an atomic reduction
    # the weight-gradient GEMMs reduce their row splits with fp32 atomics: equal up to summation order
    assert rel_err(res[1][2], res[0][2]) < 1e-5 and rel_err(res[1][3], res[0][3]) < 1e-5



@pytest.mark.parametrize("gen", ["2", "4"])
def test_channel_mix_backward_generations_agree_with_the_fp64_reference(gen):
    """The default channel-mixing backward (generation 4: dH spilled through TMA stores, wgrad_dh recomputes only G) and the
    generation-2 A/B form (G and dH recomputed on chip, no spill) both against the fp64 restatement of
    modules/mixer.py:37-40 at the encoder shape (M = 16384, C = 3072), the ragged fusion width (C = 3078) and D = 32 / 64:
    tools/gpu_selfcheck.py chain_bwd in a subprocess (the generation is read once per process).  bf16: <= 2e-2 on every
    gradient."""
    import subprocess
    import sys
    env = dict(os.environ, M2B200_CHAIN_GEN=gen)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gpu_selfcheck.py"), "chain_bwd"], capture_output=True,
                       text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "FAIL" not in r.stdout and r.stdout.count("[ok]") >= 42, r.stdout[-3000:]


def test_heads_kernels_vectorised_and_generic_paths():
    """Heads + losses (models/avmnist.py:267-293, models/mmimdb.py:115-125) against the fp64 restatement on both kernel
    families: K <= 16 classes with 16-byte-aligned token rows (vectorised kernels) and K = 23 (generic kernels), CE and BCE,
    forward (logits, four losses, preds) and backward (token, weight and bias gradients) to 1e-5."""
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gpu_selfcheck.py"), "heads"], capture_output=True,
                       text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "FAIL" not in r.stdout and r.stdout.count("[ok]") >= 48, r.stdout[-3000:]
