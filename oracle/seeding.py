"""Seeded weights and synthetic batches shared by the golden generator, the tests and bench.py.

TEST / BENCH INFRASTRUCTURE (CPU torch only).  Deterministic for a given torch build: everything is
drawn on the CPU from a ``torch.Generator`` in a fixed order, then moved to whatever device the
caller wants.  Batch layouts follow the reference data modules (datasets/avmnist.py:15-23,113-114;
datasets/mimic.py:77; models/mmimdb.py:68-70).
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch


def _fan_in(key: str, shape: Tuple[int, ...], shapes: Dict[str, Tuple[int, ...]]) -> int:
    if len(shape) >= 2:
        return int(math.prod(shape[1:]))
    w = shapes.get(key[: -len("bias")] + "weight")
    return int(math.prod(w[1:])) if (w is not None and len(w) >= 2) else 1


def _is_norm(key: str, shape, shapes) -> bool:
    if len(shape) != 1:
        return False
    w = shapes.get(key.rsplit(".", 1)[0] + ".weight")
    return w is not None and len(w) == 1


def seeded_state_dict(shapes: Dict[str, Tuple[int, ...]], seed: int, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """U(+-1/sqrt(fan_in)) for Linear/Conv weight+bias; LayerNorm gamma = 1+0.1U(-1,1), beta = 0.1U(-1,1).

    Drawn in float64 in sorted key order, rounded to fp32, then cast: fp32 and fp64 runs share bit-identical weights."""
    g = torch.Generator().manual_seed(int(seed))
    sd = {}
    for k in sorted(shapes):
        shp = tuple(shapes[k])
        u = torch.rand(shp, generator=g, dtype=torch.float64) * 2 - 1
        if _is_norm(k, shp, shapes):
            v = (1.0 + 0.1 * u) if k.endswith("weight") else 0.1 * u
        else:
            v = u / math.sqrt(_fan_in(k, shp, shapes))
        sd[k] = v.to(torch.float32).to(dtype)  # fp32-representable in every dtype
    return sd


def synthetic_batch(kind, bsz: int, seed: int, dtype=torch.float32):
    """N(0,1) inputs + uniform labels, SURVEY 8(d) shapes."""
    g = torch.Generator().manual_seed(int(seed) + 1)
    rn = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64).to(torch.float32).to(dtype)
    if kind == "avmnist":
        return {"image": rn(bsz, 1, 28, 28), "audio": rn(bsz, 1, 112, 112),
                "label": torch.randint(0, 10, (bsz,), generator=g)}
    if kind == "mimic":
        return (rn(bsz, 5), rn(bsz, 24, 12), torch.randint(0, 6, (bsz,), generator=g))
    if isinstance(kind, tuple) and kind[0] == "mmimdb":
        _, img, txt = kind
        image = rn(bsz, img["in_channels"], *img["image_size"])
        if txt.get("block_type") == "PNLPMixer":
            text = rn(bsz, txt["max_seq_len"], (2 * txt["bottleneck_window_size"] + 1) * txt["bottleneck_features_size"])
        else:
            text = rn(bsz, txt["in_channels"], *txt["image_size"])
        label = (torch.rand(bsz, 23, generator=g) < 0.1).long()
        return {"image": image, "text": text, "label": label}
    raise ValueError(kind)
