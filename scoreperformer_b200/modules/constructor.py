"""Config-based module construction: `Cls.init(config, **overrides)`, registries keyed by `_target_`.

Mirrors the reference interface (scoreperformer/modules/constructor.py:13-130): same names, same argument meaning,
same error behaviour (unknown keys are dropped with a warning, `???` raises RuntimeError).
"""
from __future__ import annotations

import copy
import logging
from dataclasses import dataclass
from inspect import signature
from typing import Callable, Optional, Union

import torch

from ..config import DictConfig, MISSING, merge as _merge

logger = logging.getLogger("scoreperformer_b200")


@dataclass
class ModuleConfig:
    def update(self, **kwargs):
        kwargs = {k: v for k, v in kwargs.items() if not k.startswith("_")}
        invalid = [k for k in kwargs if k not in self.__dict__]
        if invalid:
            logger.warning("The following params are incompatible with the config %s, so they will be ignored: %s.",
                           type(self).__name__, invalid)
        for k, v in kwargs.items():
            if k not in invalid:
                setattr(self, k, v)
        return self

    def to_dict(self, check_missing=False, make_copy=True):
        if check_missing:
            missing = [k for k, v in self.__dict__.items() if isinstance(v, str) and v == MISSING]
            if missing:
                raise RuntimeError(f"The following params are mandatory to set: {missing}")
        return copy.deepcopy(self.__dict__) if make_copy else dict(self.__dict__)

    def get(self, key, default=None):
        return self.__dict__.get(key, default)


@dataclass
class VariableModuleConfig(ModuleConfig):
    _target_: str = "default"


def merge(*containers, as_omega=False):
    merged = _merge(*containers)
    return merged if as_omega else dict(merged)


class Constructor:
    @classmethod
    def _pre_init(cls, config=None, **parameters):
        modules = {k: v for k, v in parameters.items() if isinstance(v, torch.nn.Module)}
        parameters = {k: v for k, v in parameters.items() if k not in modules}
        config = merge(config or {}, parameters)
        config.update(modules)
        return {k: v for k, v in config.items() if not k.startswith("_")}

    @classmethod
    def init(cls, config=None, **parameters):
        config = cls._pre_init(config, **parameters)
        sig = dict(signature(cls.__init__).parameters)
        if "kwargs" not in sig:
            invalid = [k for k in config if k not in sig]
            if invalid:
                logger.warning("The following params are incompatible with the %s constructor, so they will be ignored: %s.",
                               cls.__name__, invalid)
                config = {k: v for k, v in config.items() if k not in invalid}
        missing = [k for k, v in config.items() if isinstance(v, str) and v == MISSING]
        if missing:
            raise RuntimeError(f"The following params are mandatory to set: {missing}")
        return cls(**config)


class Registry:
    def __init__(self):
        self._objects = {}

    def register(self, name: str, module: Optional[Callable] = None):
        if not isinstance(name, str):
            raise TypeError(f"`name` must be a str, got {name}")

        def _register(obj):
            self._objects[name] = obj
            return obj

        return _register if module is None else _register(module)

    def instantiate(self, config, **kwargs):
        target = config["_target_"] if isinstance(config, dict) else config._target_
        return self.get(target).init(config, **kwargs)

    def get(self, key: str):
        try:
            return self._objects[key]
        except KeyError:
            raise KeyError(f"'{key}' not found in registry. Available names: {self.available_names}")

    def remove(self, name):
        self._objects.pop(name)

    @property
    def objects(self):
        return self._objects

    @property
    def available_names(self):
        return tuple(self._objects.keys())

    def __str__(self):
        return f"Objects={self.available_names}"
