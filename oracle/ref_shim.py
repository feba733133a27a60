"""TEST INFRASTRUCTURE ONLY -- makes the *reference* ScorePerformer importable in the authoring
container so that golden vectors can be generated from it (see oracle/gen_golden.py).

The reference (/root/reference, read-only, absent on the GPU box) imports `omegaconf`, `miditok`
and `miditoolkit`, none of which are installed here.  None of them touch hot-path arithmetic
(SURVEY.md §8c): omegaconf is config plumbing (modules/constructor.py:9,68-86), miditok/miditoolkit
are MIDI tokenisation (data/tokenizers/*).  We register minimal stand-ins in `sys.modules`
*before* importing `scoreperformer`, then build the default recipe by hand the way
`ScorePerformer.inject_data_config` (models/scoreperformer/model.py:374-394) would.

Nothing under scoreperformer_b200/ may import this file.
"""
from __future__ import annotations

import copy
import dataclasses
import os
import re
import sys
import types

import yaml

REFERENCE_ROOT = os.environ.get("SPB200_REFERENCE_ROOT", "/root/reference")


# --------------------------------------------------------------------------- omegaconf stand-in
class DictConfig(dict):
    """dict with attribute access, enough for Constructor.init / merge()."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def __delattr__(self, k):
        del self[k]

    def _get_flag(self, name):
        return False

    def get(self, k, default=None):
        return super().get(k, default)


class ListConfig(list):
    pass


def _wrap(obj):
    if isinstance(obj, dict) and not isinstance(obj, DictConfig):
        return DictConfig({k: _wrap(v) for k, v in obj.items()})
    if isinstance(obj, DictConfig):
        return DictConfig({k: _wrap(v) for k, v in obj.items()})
    if isinstance(obj, (list, tuple)) and not isinstance(obj, ListConfig):
        return ListConfig([_wrap(v) for v in obj])
    return obj


def _to_plain(obj):
    if dataclasses.is_dataclass(obj) and not isinstance(obj, type):
        obj = {k: v for k, v in obj.__dict__.items()}
    if isinstance(obj, dict):
        return {k: _to_plain(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [_to_plain(v) for v in obj]
    return obj


def _merge2(a, b):
    out = DictConfig(a)
    for k, v in b.items():
        if k in out and isinstance(out[k], dict) and isinstance(v, dict):
            out[k] = _merge2(out[k], v)
        else:
            out[k] = _wrap(v) if isinstance(v, (dict, list, tuple)) else v
    return out


class OmegaConf:
    @staticmethod
    def merge(*containers):
        out = DictConfig()
        for c in containers:
            if dataclasses.is_dataclass(c) and not isinstance(c, type):
                c = c.__dict__
            out = _merge2(out, _wrap(dict(c)))
        return out

    @staticmethod
    def create(obj=None):
        return _wrap(obj or {})

    @staticmethod
    def load(path):
        with open(path) as f:
            return _wrap(yaml.safe_load(f))

    @staticmethod
    def resolve(cfg):
        return cfg

    @staticmethod
    def set_readonly(cfg, flag):
        return None

    @staticmethod
    def to_container(cfg, resolve=True):
        return _to_plain(cfg)

    @staticmethod
    def register_new_resolver(*a, **k):
        return None


class _Anything:
    """Import-time placeholder for names we never call."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, k):
        return _Anything()

    def __mro_entries__(self, bases):
        return (object,)


def _stub_module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)

    def __getattr__(attr):
        if attr.startswith("__"):
            raise AttributeError(attr)
        return type(attr, (), {"__init__": lambda self, *a, **k: None})

    m.__getattr__ = __getattr__
    sys.modules[name] = m
    return m


def install_stubs():
    if "omegaconf" not in sys.modules:
        _stub_module("omegaconf", DictConfig=DictConfig, ListConfig=ListConfig, OmegaConf=OmegaConf, MISSING="???")
    if "miditok" not in sys.modules:
        @dataclasses.dataclass
        class TokSequence:
            tokens: object = None
            ids: object = None
            bytes: object = None
            events: object = None
            ids_bpe_encoded: bool = False
            _ids_no_bpe: object = None

        class TokenizerConfig:
            def __init__(self, *a, **k):
                self.__dict__.update(k)

        class MIDITokenizer:
            def __init__(self, *a, **k):
                pass

        def _in_as_seq(*a, **k):
            def deco(fn):
                return fn
            return deco

        _stub_module("miditok", MIDITokenizer=MIDITokenizer, Event=type("Event", (), {}))
        _stub_module("miditok.classes", TokSequence=TokSequence, TokenizerConfig=TokenizerConfig)
        _stub_module("miditok.constants", TIME_SIGNATURE=(4, 4), TEMPO=120, MIDI_INSTRUMENTS=[{}] * 128)
        _stub_module("miditok.midi_tokenizer", _in_as_seq=_in_as_seq, MIDITokenizer=MIDITokenizer)
        _stub_module("miditok.utils")
    if "miditoolkit" not in sys.modules:
        _stub_module("miditoolkit")
    for name in ("note_seq", "matplotlib", "matplotlib.pyplot", "librosa", "librosa.display", "pretty_midi"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                _stub_module(name)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


# --------------------------------------------------------------------------- recipe loading
_INTERP = re.compile(r"^\$\{([^}:]+)\}$")


def _deep_merge(a, b):
    out = copy.deepcopy(a)
    for k, v in b.items():
        if k in out and isinstance(out[k], dict) and isinstance(v, dict):
            out[k] = _deep_merge(out[k], v)
        else:
            out[k] = copy.deepcopy(v)
    return out


def _load_with_base(path):
    with open(path) as f:
        cfg = yaml.safe_load(f) or {}
    base = cfg.pop("base", None)
    if base:
        for b in ([base] if isinstance(base, str) else base):
            cand = os.path.join(os.path.dirname(path), b)
            if not os.path.exists(cand):
                cand = os.path.join(REFERENCE_ROOT, "recipes", b)
            cfg = _deep_merge(_load_with_base(cand), cfg)
    return cfg


def _lookup(root, dotted):
    node = root
    for part in dotted.split("."):
        node = node[part]
    return node


def _resolve(node, root, depth=0):
    if isinstance(node, dict):
        return {k: _resolve(v, root, depth) for k, v in node.items()}
    if isinstance(node, list):
        return [_resolve(v, root, depth) for v in node]
    if isinstance(node, str):
        m = _INTERP.match(node)
        if m and depth < 16:
            return _resolve(copy.deepcopy(_lookup(root, m.group(1))), root, depth + 1)
    return node


def _disable(node):
    if isinstance(node, dict):
        out = {}
        for k, v in node.items():
            if isinstance(v, dict) and v.get("_disable_", False):
                out[k] = None
                continue
            out[k] = _disable(v)
        out.pop("_disable_", None)
        return out
    return node


def load_recipe(name="scoreperformer/base.yaml"):
    """YAML -> plain dict with `base:` inheritance, `${a.b}` interpolation and `_disable_` applied
    (mirrors experiments/components.py:30-63 + utils/config.py:36-45)."""
    cfg = _load_with_base(os.path.join(REFERENCE_ROOT, "recipes", name))
    cfg = _resolve(cfg, cfg)
    return _disable(cfg)


PERF_SIZES = {  # SURVEY.md Appendix A.1 (SPMupleWindow)
    "Bar": 260, "Position": 132, "Pitch": 92, "Velocity": 132, "Duration": 133, "Tempo": 125,
    "TimeSig": 26, "PositionShift": 69, "NotesInOnset": 16, "PositionInOnset": 16,
    "RelOnsetDev": 165, "RelPerfDuration": 85,
}
SCORE_KEYS = list(PERF_SIZES)[:10]
DIRECTION_CLASSES = {  # data/directions/direction_classes.json group sizes (+1 "none" class each)
    "dynamic/absolute": 10, "dynamic/hairpin": 3, "dynamic/accent": 3, "tempo/absolute": 14,
    "tempo/relative": 11, "articulation/arpeggiate": 2, "articulation/fermata": 2,
    "articulation/staccato": 2, "articulation/tenuto": 2,
}


def default_model_config(recipe="scoreperformer/base.yaml", num_tokens=None, direction_classes=None):
    """What build_model + inject_data_config produce for the default recipe (model.py:374-394)."""
    import numpy as np
    num_tokens = dict(num_tokens or PERF_SIZES)
    direction_classes = dict(direction_classes or DIRECTION_CLASSES)
    cfg = load_recipe(recipe)["model"]
    cfg["num_tokens"] = num_tokens
    cfg["num_score_tokens"] = {k: v for k, v in num_tokens.items() if k in SCORE_KEYS}
    token_values = {
        k: [0.0, 0.0, 0.0, 0.0] + np.linspace(0.0, 1.0, v - 4).tolist() for k, v in num_tokens.items()
    }
    for key in ("score_encoder", "perf_encoder", "perf_decoder"):
        if cfg.get(key) is not None:
            tv = token_values if key != "score_encoder" else {k: token_values[k] for k in cfg["num_score_tokens"]}
            cfg[key]["token_embeddings"]["token_values"] = copy.deepcopy(tv)
    if cfg.get("classifiers") is not None:
        cfg["classifiers"]["num_classes"] = direction_classes
        cfg["classifiers"]["class_samples"] = {k: [1.0 / v] * v for k, v in direction_classes.items()}
    return cfg


def build_reference_model(recipe="scoreperformer/base.yaml", seed=23, **kw):
    install_stubs()
    import torch
    from scoreperformer.models import ScorePerformer
    cfg = _wrap(default_model_config(recipe, **kw))
    torch.manual_seed(seed)
    model = ScorePerformer.init(cfg)
    return model, cfg
