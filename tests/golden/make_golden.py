"""Generate golden vectors from the UNMODIFIED reference modules (build container only).

Run:  PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py
Needs /root/reference (read-only mount).  Writes tests/golden/*.npz.  The reference
task modules (models/*.py) cannot be imported here (pytorch_lightning / omegaconf are
not installed), so the model is assembled from the reference's own ``modules.*``
classes through its name registry exactly as the ctors do (models/avmnist.py:178-196,
models/mimic.py:36-52) and the shared_step arithmetic uses torch's own
CrossEntropyLoss / BCEWithLogitsLoss as the reference does.

Weights are NOT the nn default init: every parameter is overwritten by
``seeded_state_dict`` (U(-1/sqrt(fan_in), 1/sqrt(fan_in)) from a torch.Generator in
sorted-key order; LayerNorm gamma = 1 + 0.1 U, beta = 0.1 U) so that the big configs can be
regenerated from a seed instead of being committed.
"""
import os
import sys

import numpy as np
import torch
import yaml

REF = "/root/reference"
sys.path.insert(0, REF)
sys.dont_write_bytecode = True
import modules  # noqa: E402  (the reference package)

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle.seeding import seeded_state_dict, synthetic_batch  # noqa: E402


class _NS(dict):
    __getattr__ = dict.get


def _ns(d):
    return _NS({k: _ns(v) for k, v in d.items()}) if isinstance(d, dict) else d


class RefAVMnist(torch.nn.Module):
    def __init__(self, mc, fusion="ConcatFusion"):
        super().__init__()
        m = mc["modalities"]
        self.image_mixer = modules.get_block_by_name(**m["image"], dropout=0.0)
        self.audio_mixer = modules.get_block_by_name(**m["audio"], dropout=0.0)
        mm = dict(m["multimodal"], fusion_function=fusion)
        self.fusion_function = modules.get_fusion_by_name(**mm)
        npatch = self.fusion_function.get_output_shape(self.image_mixer.num_patch, self.audio_mixer.num_patch, dim=1)
        self.fusion_mixer = modules.get_block_by_name(**mm, num_patches=npatch, dropout=0.0)
        k = m["classification"]["num_classes"]
        self.classifier_image = torch.nn.Linear(m["image"]["hidden_dim"], k)
        self.classifier_audio = torch.nn.Linear(m["audio"]["hidden_dim"], k)
        self.classifier_fusion = modules.get_classifier_by_name(**m["classification"])
        self.w = mc.get("fusion_loss_weight", 1.0 / 3)

    def forward(self, batch):
        ce = torch.nn.CrossEntropyLoss()
        i = self.image_mixer(batch["image"])
        a = self.audio_mixer(batch["audio"])
        f = self.fusion_mixer(self.fusion_function(i, a))
        a = a.reshape(a.shape[0], -1, a.shape[-1])
        i = i.reshape(i.shape[0], -1, i.shape[-1])
        li, la, lf = self.classifier_image(i.mean(dim=1)), self.classifier_audio(a.mean(dim=1)), self.classifier_fusion(f)
        Li, La, Lf = ce(li, batch["label"]), ce(la, batch["label"]), ce(lf, batch["label"])
        ow = (1 - self.w) / 2
        return dict(loss=(self.w * Lf + ow * Li + ow * La) * 3, loss_image=Li, loss_audio=La, loss_fusion=Lf,
                    image_logits=li, audio_logits=la, logits=lf)


class RefMimic(torch.nn.Module):
    def __init__(self, mc):
        super().__init__()
        m = mc["modalities"]
        self.time_mixer = modules.get_block_by_name(**m["time"], dropout=0.0)
        self.static_extractor = modules.get_block_by_name(**m["static"], dropout=0.0)
        self.fusion_function = modules.get_fusion_by_name(**m["multimodal"])
        npatch = self.fusion_function.get_output_shape(1, self.time_mixer.num_patch, dim=1)
        self.fusion_mixer = modules.get_block_by_name(**m["multimodal"], num_patches=npatch, dropout=0.0)
        k = m["classification"]["num_classes"]
        self.classifier_static = torch.nn.Linear(m["static"]["output_dim"], k)
        self.classifier_time = torch.nn.Linear(m["time"]["hidden_dim"], k)
        self.classifier_fusion = modules.get_classifier_by_name(**m["classification"])
        self.w = mc.get("fusion_loss_weight", 1.0 / 3)

    def forward(self, batch):
        ce = torch.nn.CrossEntropyLoss()
        static, time, y = batch
        s = self.static_extractor(static)
        t = self.time_mixer(time)
        f = self.fusion_mixer(self.fusion_function(s.unsqueeze(1), t))
        ls, lt, lf = self.classifier_static(s), self.classifier_time(t.mean(1)), self.classifier_fusion(f)
        Lf, Ls, Lt = ce(lf, y), ce(ls, y), ce(lt, y)
        ow = (1 - self.w) / 2
        return dict(loss=self.w * Lf + ow * Ls + ow * Lt, loss_fusion=Lf, loss_static=Ls, loss_time=Lt,
                    logits=lf, logits_static=ls, logits_time=lt)


class RefMMIMDB(torch.nn.Module):
    """Synthetic C4-shaped model at reduced size: MLPMixer image + PNLPMixer text, BCE x3."""

    def __init__(self, image_cfg, text_cfg, mm_cfg, k, pos_weight):
        super().__init__()
        self.image_mixer = modules.get_block_by_name(**image_cfg, dropout=0.0)
        self.text_mixer = modules.get_block_by_name(**text_cfg, dropout=0.0)
        self.fusion_function = modules.get_fusion_by_name(**mm_cfg)
        npatch = self.fusion_function.get_output_shape(self.image_mixer.num_patch, self.text_mixer.num_patch, dim=1)
        self.fusion_mixer = modules.get_block_by_name(**mm_cfg, num_patches=npatch, dropout=0.0)
        self.classifier_image = torch.nn.Linear(image_cfg["hidden_dim"], k)
        self.classifier_text = torch.nn.Linear(text_cfg["hidden_dim"], k)
        self.classifier_fusion = modules.get_classifier_by_name(classifier="StandardClassifier",
                                                                input_shape=[1, 1, mm_cfg["hidden_dim"]], num_classes=k)
        self.register_buffer("pos_weight", pos_weight, persistent=False)

    def forward(self, batch):
        bce = torch.nn.BCEWithLogitsLoss(pos_weight=self.pos_weight)
        i = self.image_mixer(batch["image"])
        t = self.text_mixer(batch["text"])
        f = self.fusion_mixer(self.fusion_function(i, t))
        li = self.classifier_image(i.reshape(i.shape[0], -1, i.shape[-1]).mean(dim=1))
        lt = self.classifier_text(t.reshape(t.shape[0], -1, t.shape[-1]).mean(dim=1))
        lf = self.classifier_fusion(f)
        y = batch["label"].to(li.dtype)
        Li, Lt, Lf = bce(li, y), bce(lt, y), bce(lf, y)
        return dict(loss=Li + Lt + Lf, loss_image=Li, loss_text=Lt, loss_fusion=Lf, image_logits=li, text_logits=lt,
                    logits=lf)


def load_cfg(rel):
    with open(os.path.join(REF, rel)) as f:
        return yaml.safe_load(f)


def run(model, batch, seed, dtype):
    model = model.to(dtype)
    sd = seeded_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed, dtype)
    model.load_state_dict(sd, strict=True)
    model.train()  # dropout is 0.0; exercises the training branch
    out = model(batch)
    out["loss"].backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    return sd, {k: v.detach() for k, v in out.items()}, grads


def pack(out, grads, full_grads):
    """full_grads: False = norm + first 16 elements; True = every element as float32; "f16" = every element as float16
    after a per-tensor power-of-two scale (8.3 M gradient elements of M2-Mixer-B in ~15 MB: 2^-11 relative steps, finer
    than the bf16 bar they are compared under; the fp32 bar uses the exact norms + heads next to them)."""
    d = {"out." + k: v.double().numpy() for k, v in out.items()}
    for k, g in grads.items():
        g = g.double()
        d["gnorm." + k] = np.array(float(g.norm()))
        d["ghead." + k] = g.flatten()[:16].numpy()
        if full_grads == "f16":
            amax = float(g.abs().max())
            e = 0 if amax == 0.0 else int(np.floor(np.log2(amax)))
            d["gexp." + k] = np.array(e)
            d["grad16." + k] = (g * 2.0 ** (-e)).numpy().astype(np.float16)
        elif full_grads:
            d["grad." + k] = g.numpy().astype(np.float32)
    return d


def cast_batch(b, dtype):
    f = lambda t: t.to(dtype) if t.is_floating_point() else t
    return {k: f(v) for k, v in b.items()} if isinstance(b, dict) else tuple(f(v) for v in b)


def main(only=None):
    torch.set_num_threads(8)
    jobs = []
    # name, builder, batch kind, B, full grads?
    cfgS = load_cfg("cfg/avmnist/avmnist_m2-mixer_S.yml")["model"]
    cfgM = load_cfg("cfg/avmnist/avmnist_m2-mixer_M.yml")["model"]
    cfgB = load_cfg("cfg/avmnist/avmnist_m2-mixer_B.yml")["model"]
    cfgH = load_cfg("cfg/mimic/mimic_m2-mixer_H.yml")["model"]
    jobs.append(("avmnist_S_b8", lambda: RefAVMnist(cfgS), "avmnist", 8, True))
    jobs.append(("avmnist_S_sum_b8", lambda: RefAVMnist(cfgS, "SumFusion"), "avmnist", 8, False))
    jobs.append(("avmnist_M_b4", lambda: RefAVMnist(cfgM), "avmnist", 4, False))
    jobs.append(("avmnist_B_b16", lambda: RefAVMnist(cfgB), "avmnist", 16, False))
    jobs.append(("avmnist_B_b2", lambda: RefAVMnist(cfgB), "avmnist", 2, "f16"))   # EVERY gradient element of M2-Mixer-B
    jobs.append(("mimic_H_b16", lambda: RefMimic(cfgH), "mimic", 16, True))
    img = dict(block_type="MLPMixer", in_channels=3, hidden_dim=64, patch_size=16, image_size=[64, 48], token_dim=16,
               channel_dim=96, num_mixers=1)
    txt = dict(block_type="PNLPMixer", max_seq_len=24, hidden_dim=64, num_mixers=1, mlp_hidden_dim=48,
               bottleneck_window_size=1, bottleneck_features_size=40)
    mm = dict(block_type="FusionMixer", fusion_function="ConcatFusion", hidden_dim=64, token_dim=16, channel_dim=96,
              num_mixers=1)
    pw = torch.tensor(load_cfg("cfg/mmimdb/mmimdb_3loss.yml")["model"]["pos_weight"], dtype=torch.float64)
    jobs.append(("mmimdb_tiny_b6", lambda: RefMMIMDB(img, txt, mm, 23, pw), ("mmimdb", img, txt), 6, True))

    for name, build, kind, bsz, full in jobs:
        if only is not None and name != only:
            continue
        seed = 1234
        res = {}
        for dtype, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
            torch.manual_seed(0)
            model = build()
            batch = cast_batch(synthetic_batch(kind, bsz, seed), dtype)
            sd, out, grads = run(model, batch, seed, dtype)
            res[tag] = (out, grads)
        d = pack(*res["f64"], full_grads=full)
        # fp32-vs-fp64 noise floor of the reference itself, for the record
        for k in res["f64"][0]:
            a, b = res["f64"][0][k].double(), res["f32"][0][k].double()
            d["f32err." + k] = np.array(float((a - b).norm() / (a.norm() + 1e-30)))
        d["meta.seed"] = np.array(seed)
        d["meta.batch"] = np.array(bsz)
        shapes = {k: tuple(v.shape) for k, v in sd.items()}
        d["meta.keys"] = np.array(sorted(shapes.keys()))
        d["meta.shapes"] = np.array([",".join(map(str, shapes[k])) for k in sorted(shapes.keys())])
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
        print(name, "loss", float(res["f64"][0]["loss"]), "f32 logit err", float(d["f32err.logits"]),
              "bytes", os.path.getsize(os.path.join(HERE, name + ".npz")))

    if only is not None:
        return
    # single MixerBlock at awkward sizes (C not a multiple of anything, N odd) incl. input grad
    for name, (B, N, D, T, C) in {"block_odd": (5, 7, 48, 10, 70), "block_b": (6, 8, 128, 32, 3078)}.items():
        blk = modules.MixerBlock(D, N, T, C, dropout=0.0).double()
        sd = seeded_state_dict({k: tuple(v.shape) for k, v in blk.state_dict().items()}, 77, torch.float64)
        blk.load_state_dict(sd)
        g = torch.Generator().manual_seed(78)
        x = torch.randn(B, N, D, generator=g, dtype=torch.float64).requires_grad_(True)
        dy = torch.randn(B, N, D, generator=g, dtype=torch.float64)
        y = blk(x)
        y.backward(dy)
        d = {"x": x.detach().numpy(), "dy": dy.numpy(), "y": y.detach().numpy(), "dx": x.grad.numpy()}
        big = C > 1000
        for k, p in blk.named_parameters():
            d["gnorm." + k] = np.array(float(p.grad.norm()))
            if not big:
                d["grad." + k] = p.grad.numpy()
        d["meta.dims"] = np.array([B, N, D, T, C])
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
        print(name, "bytes", os.path.getsize(os.path.join(HERE, name + ".npz")))


    big_blocks()


def big_blocks():
    """One MixerBlock at the layer shapes of BASELINE configs 4 and 5: these are the shapes that leave the
    fused small-N kernels - token mixing as a real [T x N] contraction, channel mixing at D = 256 / 768.
    x / dy are regenerated from the seed by the tests (not stored); y, dx as float32; per-parameter gradient norms
    (fp64) + a strided sample of 4096 elements of every gradient."""
    shapes = {"block_c5": (1, 196, 768, 384, 3072),        # Scaled C5 encoder block (Mixer-B/16 width)
              "block_c4text": (1, 512, 256, 512, 512),     # C4 PNLPMixer text block: T = C = 512, 512 tokens
              "block_c4fus": (1, 708, 256, 16, 512)}       # C4 fusion block: 708 tokens, T = 16
    for name, (B, N, D, T, C) in shapes.items():
        blk = modules.MixerBlock(D, N, T, C, dropout=0.0).double()
        sd = seeded_state_dict({k: tuple(v.shape) for k, v in blk.state_dict().items()}, 77, torch.float64)
        blk.load_state_dict(sd)
        g = torch.Generator().manual_seed(78)
        x = torch.randn(B, N, D, generator=g, dtype=torch.float64).requires_grad_(True)
        dy = torch.randn(B, N, D, generator=g, dtype=torch.float64)
        y = blk(x)
        y.backward(dy)
        d = {"y": y.detach().numpy().astype(np.float32), "dx": x.grad.numpy().astype(np.float32),
             "ynorm": np.array(float(y.norm())), "dxnorm": np.array(float(x.grad.norm()))}
        for k, p in blk.named_parameters():
            gflat = p.grad.flatten()
            stride = max(1, gflat.numel() // 4096)
            d["gnorm." + k] = np.array(float(p.grad.norm()))
            d["gsamp." + k] = gflat[::stride][:4096].numpy()
            d["gstride." + k] = np.array(stride)
        d["meta.dims"] = np.array([B, N, D, T, C])
        d["meta.regen"] = np.array(1)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
        print(name, "bytes", os.path.getsize(os.path.join(HERE, name + ".npz")))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "big_blocks":
        big_blocks()
    elif len(sys.argv) > 1 and sys.argv[1] == "b_full":
        main(only="avmnist_B_b2")
    else:
        main()
