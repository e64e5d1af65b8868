"""Config plumbing compatible with the reference's cfg/*.yml files (three sections: train / dataset / model).

The reference loads them with OmegaConf (run.py:28) and reads values both as attributes and with ``.get``.
OmegaConf is not a dependency here: ``Cfg`` is a dict with attribute access, ``load_yaml`` uses PyYAML and coerces
YAML-1.1 strings such as ``1e-2`` (which PyYAML leaves as str, OmegaConf parses as float) to floats.
"""
from __future__ import annotations

import re
from typing import Any, Mapping

_FLOAT = re.compile(r"^[+-]?(\d+\.?\d*|\.\d+)([eE][+-]?\d+)?$")


class Cfg(dict):
    """dict with attribute access (read/write) and OmegaConf-style ``.get`` / ``.pop``; nested mappings are wrapped."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        for k, v in dict(*args, **kwargs).items():
            self[k] = v

    def __setitem__(self, k, v):
        super().__setitem__(k, wrap(v))

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k) from None

    def __setattr__(self, k, v):
        self[k] = v


def wrap(v: Any) -> Any:
    if isinstance(v, Cfg):
        return v
    if isinstance(v, Mapping):
        return Cfg(v)
    if isinstance(v, (list, tuple)):
        return [wrap(x) for x in v]
    if isinstance(v, str) and _FLOAT.match(v) and not v.isdigit():
        return float(v)
    return v


def load_yaml(path: str) -> Cfg:
    import yaml
    with open(path) as f:
        return Cfg(yaml.safe_load(f))


def deep_update(base: Cfg, dotted: str, value: Any) -> None:
    """``--model.modalities.image.hidden_dim=64`` style override (reference run.py:33-40, utils/utils.py:9-18)."""
    keys = dotted.split(".")
    node = base
    for k in keys[:-1]:
        if k not in node:
            node[k] = Cfg()
        node = node[k]
    node[keys[-1]] = value
