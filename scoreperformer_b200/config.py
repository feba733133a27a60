"""OmegaConf-free configuration plumbing.

The reference resolves its YAML recipes with OmegaConf (experiments/components.py:30-63, utils/config.py:36-45),
which is not installed on the target boxes.  This module provides the small subset the model constructors need:
attribute-style dict nodes, recursive merge, `base:` inheritance, `${a.b}` interpolation and `_disable_` pruning,
so the UNMODIFIED `recipes/*.yaml` files load.  A real omegaconf DictConfig is accepted wherever a config is.
"""
from __future__ import annotations

import copy
import dataclasses
import os
import re
from typing import Any, Dict, Optional

MISSING = "???"


class DictConfig(dict):
    """dict with attribute access (the part of omegaconf.DictConfig the constructors use)."""

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError as e:
            raise AttributeError(key) from e

    def __setattr__(self, key, value):
        self[key] = value

    def __delattr__(self, key):
        del self[key]

    def _get_flag(self, name):
        return False

    def __deepcopy__(self, memo):
        return DictConfig({k: copy.deepcopy(v, memo) for k, v in self.items()})


class ListConfig(list):
    pass


def to_plain(obj: Any) -> Any:
    """Any supported container (dict, DictConfig, omegaconf node, dataclass config) -> plain python containers."""
    try:  # real omegaconf, when present
        from omegaconf import OmegaConf  # type: ignore
        from omegaconf.basecontainer import BaseContainer  # type: ignore
        if isinstance(obj, BaseContainer):
            return OmegaConf.to_container(obj, resolve=True)
    except Exception:
        pass
    if dataclasses.is_dataclass(obj) and not isinstance(obj, type):
        return {k: to_plain(v) for k, v in obj.__dict__.items()}
    if isinstance(obj, dict):
        return {k: to_plain(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [to_plain(v) for v in obj]
    return obj


def wrap(obj: Any) -> Any:
    if isinstance(obj, dict):
        return DictConfig({k: wrap(v) for k, v in obj.items()})
    if isinstance(obj, (list, tuple)) and not isinstance(obj, ListConfig):
        return ListConfig([wrap(v) for v in obj])
    return obj


def deep_merge(a: Dict, b: Dict) -> Dict:
    out = dict(a)
    for k, v in b.items():
        if k in out and isinstance(out[k], dict) and isinstance(v, dict):
            out[k] = deep_merge(out[k], v)
        else:
            out[k] = v
    return out


def merge(*containers) -> DictConfig:
    """Recursive merge, later containers win (OmegaConf.merge semantics for the cases the constructors hit)."""
    import torch
    out: Dict = {}
    for c in containers:
        if c is None:
            continue
        modules = {}
        if isinstance(c, dict):  # keep nn.Module values out of the deep copy / conversion
            modules = {k: v for k, v in c.items() if isinstance(v, torch.nn.Module)}
            c = {k: v for k, v in c.items() if k not in modules}
        out = deep_merge(out, to_plain(c))
        out.update(modules)
    return wrap(out)


# ----------------------------------------------------------------------------- recipes
_INTERP = re.compile(r"^\$\{([^}:]+)\}$")


def _load_yaml_with_base(path: str, recipes_root: str) -> Dict:
    import yaml
    with open(path) as f:
        cfg = yaml.safe_load(f) or {}
    base = cfg.pop("base", None)
    if base:
        for b in ([base] if isinstance(base, str) else base):
            cand = os.path.join(os.path.dirname(path), b)
            if not os.path.exists(cand):
                cand = os.path.join(recipes_root, b)
            cfg = deep_merge(_load_yaml_with_base(cand, recipes_root), cfg)
    return cfg


def _resolve(node, root, depth=0):
    if isinstance(node, dict):
        return {k: _resolve(v, root, depth) for k, v in node.items()}
    if isinstance(node, list):
        return [_resolve(v, root, depth) for v in node]
    if isinstance(node, str):
        m = _INTERP.match(node)
        if m and depth < 16:
            cur = root
            for part in m.group(1).split("."):
                cur = cur[part]
            return _resolve(copy.deepcopy(cur), root, depth + 1)
    return node


def disable_nodes(node):
    """`_disable_: true` ejects a node (utils/config.py:36-45)."""
    if isinstance(node, dict):
        out = {}
        for k, v in node.items():
            if isinstance(v, dict) and v.get("_disable_", False):
                out[k] = None
            else:
                out[k] = disable_nodes(v)
        out.pop("_disable_", None)
        return out
    return node


def load_recipe(config_name: str, config_root: str) -> DictConfig:
    """Load `<config_root>/<config_name>` the way ExperimentComponents does (experiments/components.py:50-76)."""
    cfg = _load_yaml_with_base(os.path.join(config_root, config_name), config_root)
    cfg = _resolve(cfg, cfg)
    return wrap(disable_nodes(cfg))
