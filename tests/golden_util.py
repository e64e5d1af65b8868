"""Helpers to load tests/golden/*.npz and rebuild the seeded state dict / batch they were made with."""
import os

import numpy as np
import torch

from oracle.seeding import seeded_state_dict, synthetic_batch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

IMG_TINY = dict(block_type="MLPMixer", in_channels=3, hidden_dim=64, patch_size=16, image_size=[64, 48], token_dim=16,
                channel_dim=96, num_mixers=1)
TXT_TINY = dict(block_type="PNLPMixer", max_seq_len=24, hidden_dim=64, num_mixers=1, mlp_hidden_dim=48,
                bottleneck_window_size=1, bottleneck_features_size=40)
MM_TINY = dict(block_type="FusionMixer", fusion_function="ConcatFusion", hidden_dim=64, token_dim=16, channel_dim=96,
               num_mixers=1)
POS_WEIGHT = [4.57642832, 7.38544978, 10.79846869, 13.23391421, 15.59020924, 18.62735849, 22.48861048, 25.21711367,
              74.50943396, 31.31641554, 31.79549114, 32.90833333, 39.64859438, 56.90201729, 40.46106557, 58.24483776,
              67.3890785, 84.92473118, 58.33087149, 62.68253968, 114.13294798, 141.54121864, 116.83431953]

KIND = {"avmnist_S_b8": "avmnist", "avmnist_S_sum_b8": "avmnist", "avmnist_M_b4": "avmnist",
        "avmnist_B_b16": "avmnist", "avmnist_B_b2": "avmnist", "mimic_H_b16": "mimic", "mmimdb_tiny_b6": ("mmimdb", IMG_TINY, TXT_TINY)}


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def rebuild(name, dtype=torch.float32, device="cpu", requires_grad=True):
    z = load(name)
    shapes = {k: tuple(int(t) for t in s.split(",")) for k, s in zip(z["meta.keys"].tolist(), z["meta.shapes"].tolist())}
    seed, bsz = int(z["meta.seed"]), int(z["meta.batch"])
    sd = {k: v.to(device).requires_grad_(requires_grad) for k, v in seeded_state_dict(shapes, seed, dtype).items()}
    batch = synthetic_batch(KIND[name], bsz, seed, dtype)
    mv = lambda t: t.to(device)
    batch = {k: mv(v) for k, v in batch.items()} if isinstance(batch, dict) else tuple(mv(v) for v in batch)
    return z, sd, batch


def rel_err(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def full_grad(z, name):
    """Every element of the gradient `name` of a golden record: stored as float32 ("grad.") or as scaled float16 ("grad16." +
    power-of-two exponent "gexp.", tests/golden/make_golden.py pack()).  None when the record only has norm + head."""
    if "grad." + name in z:
        return torch.as_tensor(z["grad." + name]).double()
    if "grad16." + name in z:
        return torch.as_tensor(z["grad16." + name].astype("float32")).double() * 2.0 ** int(z["gexp." + name])
    return None


BLOCK_KEYS = lambda N, D, T, C: {  # noqa: E731  state-dict layout of one reference MixerBlock (modules/mixer.py:25-40)
    "token_mix.0.weight": (D,), "token_mix.0.bias": (D,), "token_mix.2.net.0.weight": (T, N),
    "token_mix.2.net.0.bias": (T,), "token_mix.2.net.3.weight": (N, T), "token_mix.2.net.3.bias": (N,),
    "channel_mix.0.weight": (D,), "channel_mix.0.bias": (D,), "channel_mix.1.net.0.weight": (C, D),
    "channel_mix.1.net.0.bias": (C,), "channel_mix.1.net.3.weight": (D, C), "channel_mix.1.net.3.bias": (D,)}


def block_inputs(z, dtype=torch.float64):
    """(x, dy) of a block golden: stored, or regenerated from the generator seed the record was made with."""
    B, N, D, T, C = (int(v) for v in z["meta.dims"])
    if "x" in z:
        return torch.tensor(z["x"]).to(dtype), torch.tensor(z["dy"]).to(dtype)
    g = torch.Generator().manual_seed(78)
    x = torch.randn(B, N, D, generator=g, dtype=torch.float64)
    dy = torch.randn(B, N, D, generator=g, dtype=torch.float64)
    return x.to(dtype), dy.to(dtype)
