"""Embedding classifier heads (reference: scoreperformer/models/classifiers/model.py:18-226).

The nine direction heads of MultiHeadEmbeddingClassifier are evaluated by ONE fused kernel (dropout -> Linear -> class-
weighted CE, forward and backward): no boolean-mask gather, no per-head launches, no host sync.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Union

import numpy as np
import torch
import torch.nn as nn
from torch import Tensor

from ... import fused, kernels as K
from ...config import MISSING
from ...modules.constructor import Registry, VariableModuleConfig
from ..base import Model


@dataclass
class EmbeddingClassifierOutput:
    logits: Tensor = None
    loss: Optional[Tensor] = None
    losses: Optional[Dict[str, Tensor]] = None


EmbeddingClassifiersRegistry = type("_EmbeddingClassifiersRegistry", (Registry,), {})()


@dataclass
class EmbeddingClassifierConfig(VariableModuleConfig):
    input_dim: int = MISSING
    num_classes: int = MISSING
    dropout: float = 0.
    weight: Optional[List[float]] = None


@dataclass
class LinearEmbeddingClassifierConfig(EmbeddingClassifierConfig):
    _target_: str = "linear"
    hidden_dims: Optional[Sequence[int]] = field(default_factory=lambda: (32,))


@EmbeddingClassifiersRegistry.register("linear")
class LinearEmbeddingClassifier(Model):
    def __init__(self, input_dim: int, num_classes: int, hidden_dims: Optional[Sequence[int]] = (32,), dropout: float = 0.,
                 class_weights: Optional[List[float]] = None):
        super().__init__()
        self.num_classes = num_classes
        class_weights = torch.ones(num_classes) if class_weights is None else torch.tensor(class_weights)
        self.register_buffer("class_weights", class_weights.float())
        hidden_dims = hidden_dims or []
        hidden_dims = [hidden_dims] if isinstance(hidden_dims, int) else list(hidden_dims)
        in_dims, out_dims = [input_dim] + hidden_dims, hidden_dims + [num_classes]
        layers = []
        for i, (i_d, o_d) in enumerate(zip(in_dims, out_dims)):
            layers.append(nn.Linear(i_d, o_d))
            if i < len(in_dims) - 1:
                layers.append(nn.ReLU())
        self.layers = nn.Sequential(*layers)
        self.dropout_p = float(dropout)
        self.dropout = nn.Dropout(dropout) if dropout > 0. else nn.Identity()

    @property
    def is_single_linear(self) -> bool:
        return len(self.layers) == 1

    def forward(self, embeddings: Tensor, labels: Optional[Tensor] = None):
        """Stand-alone head: one-head instance of the fused kernel."""
        if not self.is_single_linear:
            raise NotImplementedError("hidden classifier layers are not used by any recipe (`hidden_dims: []`)")
        x = embeddings.squeeze(1) if embeddings.ndim == 3 else embeddings
        x = x.float().contiguous()
        logits = K.clf_logits(x, self.layers[0].weight.detach().contiguous(), self.layers[0].bias.detach().contiguous())
        loss = None
        if labels is not None:
            rowmask = torch.ones(x.shape[0], dtype=torch.bool, device=x.device)
            p = self.dropout_p if self.training else 0.0
            loss, _ = fused.ClassifierHeadsFn.apply(x, rowmask, labels.reshape(-1, 1).contiguous(), self.layers[0].weight,
                                                    self.layers[0].bias, self.class_weights, (self.num_classes,), p,
                                                    K.seed_from_torch() if p > 0 else 0, 1.0)
        return EmbeddingClassifierOutput(logits=logits, loss=loss, losses=None)

    def prepare_inputs(self, inputs):
        return inputs


@dataclass
class SequentialEmbeddingClassifierConfig(EmbeddingClassifierConfig):
    _target_: str = "sequential"
    hidden_dim: int = 32


@EmbeddingClassifiersRegistry.register("sequential")
class SequentialEmbeddingClassifier(Model):
    """GRU head; constructed for checkpoint compatibility, unused by the recipes (out of scope)."""

    def __init__(self, input_dim: int, num_classes: int, hidden_dim: int = 32, dropout: float = 0., class_weights=None):
        super().__init__()
        self.num_classes = num_classes
        class_weights = torch.ones(num_classes) if class_weights is None else torch.tensor(class_weights)
        self.register_buffer("class_weights", class_weights.float())
        self.gru = nn.GRU(input_size=input_dim, hidden_size=hidden_dim, batch_first=True, dropout=dropout)
        self.output = nn.Linear(hidden_dim, num_classes)

    def forward(self, embeddings: Tensor, labels: Optional[Tensor] = None):
        raise NotImplementedError("SequentialEmbeddingClassifier (GRU) is out of scope: no recipe uses it")

    def prepare_inputs(self, inputs):
        return inputs


@dataclass
class MultiHeadEmbeddingClassifierOutput:
    logits: Dict[str, Tensor] = None
    loss: Optional[Tensor] = None
    losses: Optional[Dict[str, Tensor]] = None


@dataclass
class MultiHeadEmbeddingClassifierConfig(VariableModuleConfig):
    _target_: str = "multi-head"
    input_dim: int = MISSING
    num_classes: Dict[str, int] = MISSING
    classifier: LinearEmbeddingClassifierConfig = MISSING
    class_samples: Optional[Dict[str, List[int]]] = None
    weighted_classes: bool = False
    loss_weight: float = 1.
    detach_inputs: Union[bool, float] = False


@EmbeddingClassifiersRegistry.register("multi-head")
class MultiHeadEmbeddingClassifier(Model):
    def __init__(self, input_dim: int, num_classes: Dict[str, int], classifier, class_samples: Optional[Dict[str, List[int]]] = None,
                 loss_weight: float = 1., weighted_classes: bool = False, detach_inputs: Union[bool, float] = False):
        super().__init__()
        self.num_classes = dict(num_classes)
        self.heads = nn.ModuleDict({})
        for key, num in num_classes.items():
            num_samples = class_samples.get(key, None) if class_samples is not None else None
            class_weights = self._class_weights(num_samples) if weighted_classes and num_samples is not None else None
            self.heads[key] = LinearEmbeddingClassifier.init(config=classifier, input_dim=input_dim, num_classes=num,
                                                             class_weights=class_weights)
        self.loss_weight = loss_weight
        self.detach_inputs = float(detach_inputs)

    @staticmethod
    def _class_weights(num_samples: List[int], beta: float = 0.999, mult: int = 1e4):
        num_samples = np.maximum(num_samples, 1e-6)
        effective_num = 1.0 - np.power(beta, np.array(num_samples) * mult)
        weights = (1.0 - beta) / np.array(effective_num)
        weights = weights / np.sum(weights) * len(num_samples)
        return weights.tolist()

    def _packed(self):
        heads = list(self.heads.values())
        W = torch.cat([h.layers[0].weight for h in heads], dim=0)
        b = torch.cat([h.layers[0].bias for h in heads], dim=0)
        cw = torch.cat([h.class_weights for h in heads], dim=0)
        return W, b, cw, tuple(int(h.num_classes) for h in heads)

    def forward(self, embeddings: Tensor, labels: Optional[Tensor] = None, rowmask: Optional[Tensor] = None, with_logits: bool = True):
        """embeddings [n, input_dim] (or [B, T, input_dim] with `rowmask` [B, T] selecting the classified notes -- the fused
        replacement of the reference's `full_embeddings[clf_mask]` gather, models/scoreperformer/model.py:323-329)."""
        if self.detach_inputs != 1.0:
            raise NotImplementedError("only `detach_inputs: true` (every recipe) is implemented")
        if not all(h.is_single_linear for h in self.heads.values()):
            raise NotImplementedError("hidden classifier layers are not used by any recipe (`hidden_dims: []`)")
        x = embeddings.reshape(-1, embeddings.shape[-1]).float().contiguous()
        n = x.shape[0]
        rm = torch.ones(n, dtype=torch.bool, device=x.device) if rowmask is None else rowmask.reshape(-1).contiguous()
        W, b, cw, ncls = self._packed()
        logits = None
        if with_logits:
            all_logits = K.clf_logits(x, W.detach().contiguous(), b.detach().contiguous())
            if rowmask is not None:
                all_logits = all_logits[rm]          # reference shape [n_selected, C]; evaluation-side only (host sync)
            logits, off = {}, 0
            for key, c in zip(self.heads.keys(), ncls):
                logits[key] = all_logits[:, off:off + c]
                off += c
        loss = losses = None
        if labels is not None:
            head0 = next(iter(self.heads.values()))
            p = head0.dropout_p if self.training else 0.0
            loss, per_head = fused.ClassifierHeadsFn.apply(x, rm, labels.reshape(n, -1).contiguous(), W, b, cw, ncls, p,
                                                           K.seed_from_torch() if p > 0 else 0, float(self.loss_weight))
            losses = {"clf/" + key: per_head[i] for i, key in enumerate(self.heads.keys())}
            losses["clf"] = loss
        return MultiHeadEmbeddingClassifierOutput(logits=logits, loss=loss, losses=losses)

    def prepare_inputs(self, inputs):
        return inputs
