"""Transformer / Encoder / Decoder stacks with hidden/KV caching.

Reference: scoreperformer/modules/transformer/transformer.py:25-257.  Module tree, parameter names and the
`TransformerIntermediates` cache contract are the reference's; the forward of the whole stack is ONE fused autograd
node (scoreperformer_b200.fused.TransformerStackFn) over hand-written sm_100a kernels.
"""
from __future__ import annotations

import copy
from dataclasses import dataclass
from functools import partial
from typing import List, Optional, Union

import torch
import torch.nn as nn
from torch import Tensor

from ... import fused, kernels as K
from ...utils import equals
from ..constructor import Constructor, Registry, VariableModuleConfig
from ..layers import AdaptiveLayerNorm, LayerNorm, Residual
from .attend import AttentionIntermediates
from .attention import Attention, AttentionConfig
from .feedforward import FeedForward, FeedForwardConfig


@dataclass
class TransformerIntermediates:
    hiddens: Optional[List[Tensor]] = None
    attention: Optional[List[AttentionIntermediates]] = None


TransformerRegistry = type("_TransformerRegistry", (Registry,), {})()


@dataclass
class TransformerConfig(VariableModuleConfig):
    _target_: str = "default"
    dim: int = 512
    depth: int = 4
    heads: int = 8
    attention: Optional[AttentionConfig] = None
    feed_forward: Optional[FeedForwardConfig] = None
    causal: bool = False
    cross_attend: bool = False
    only_cross: bool = False
    pre_norm: bool = True
    use_adanorm: bool = False
    style_emb_dim: Optional[int] = None


@TransformerRegistry.register("default")
class Transformer(nn.Module, Constructor):
    def __init__(self, dim: int = 512, depth: int = 4, heads: int = 8, attention=None, feed_forward=None, causal: bool = False,
                 cross_attend: bool = False, only_cross: bool = False, pre_norm: bool = True, use_adanorm: bool = False,
                 style_emb_dim: Optional[int] = None):
        super().__init__()
        attention = attention if attention else AttentionConfig()
        feed_forward = feed_forward if feed_forward else FeedForwardConfig()

        self.dim = dim
        self.depth = depth
        self.heads = heads
        self.causal = causal
        self.layers = nn.ModuleList([])
        self.pre_norm = pre_norm
        self.ada_norm = use_adanorm
        assert not use_adanorm or style_emb_dim is not None, "condition_dim should be provided with adanorm"
        norm_fn = partial(AdaptiveLayerNorm, dim, style_emb_dim) if use_adanorm else partial(LayerNorm, dim)

        self.cross_attend = cross_attend
        if cross_attend and not only_cross:
            default_block = ("a", "c", "f")
        elif cross_attend and only_cross:
            default_block = ("c", "f")
        else:
            default_block = ("a", "f")
        self.layer_types = default_block * depth
        self.num_attn_layers = len(list(filter(equals("a"), self.layer_types)))

        self.final_norm = norm_fn() if pre_norm else nn.Identity()

        for layer_type in self.layer_types:
            if layer_type == "a":
                layer = Attention.init(config=attention, dim=dim, heads=heads, causal=causal)
            elif layer_type == "c":
                layer = Attention.init(config=attention, dim=dim, heads=heads)
            elif layer_type == "f":
                layer = FeedForward.init(config=feed_forward, dim=dim)
            else:
                raise Exception(f"invalid layer type {layer_type}")
            norms = nn.ModuleList([norm_fn() if pre_norm else None, None, norm_fn() if not pre_norm else None])
            self.layers.append(nn.ModuleList([norms, layer, Residual(dim)]))

    # ------------------------------------------------------------------ fused path plumbing
    def _check_fused(self):
        ok = self.pre_norm and self.layer_types == ("a", "f") * self.depth
        for (_, block, _), kind in zip(self.layers, self.layer_types):
            ok = ok and block.fused_supported
        if not ok:
            raise NotImplementedError(
                "scoreperformer_b200.Transformer: the sm_100a path implements the recipes' configuration (pre-norm ('a','f') "
                "blocks, MQA + learned ALiBi, GLU-SiLU FFN, `context_emb_mode: cat`); cross-attention blocks "
                "(`context_emb_mode: attention`) and post-norm are not implemented yet (DESIGN.md, out of scope for round 1)")

    def _norm_params(self, norm):
        return (norm.linear.weight, norm.linear.bias) if self.ada_norm else (norm.weight, norm.bias)

    def flat_params(self):
        ps = []
        for (norms, block, _), kind in zip(self.layers, self.layer_types):
            ps.extend(self._norm_params(norms[0]))
            if kind == "a":
                ps.extend([block.to_q.weight, block.to_k.weight, block.to_v.weight, block.to_out.weight, block.logslopes()])
            else:
                ps.extend([block.ff[0].proj.weight, block.ff[0].proj.bias, block.ff[3].weight])
        ps.extend(self._norm_params(self.final_norm))
        return ps

    def stack_spec(self, return_hiddens: bool) -> fused.StackSpec:
        attn, ff = self.layers[0][1], self.layers[1][1]
        return fused.StackSpec(depth=self.depth, heads=self.heads, dim=self.dim, dim_head=attn.dim_head, ff_inner=ff.inner_dim,
                               causal=self.causal, ada=self.ada_norm, attn_dropout=attn.dropout, ff_dropout=ff.dropout,
                               training=self.training, eps=1e-5, return_hiddens=return_hiddens)

    def forward(self, x: Tensor, mask: Optional[Tensor] = None, context: Optional[Tensor] = None, context_mask: Optional[Tensor] = None,
                attn_mask: Optional[Tensor] = None, style_embeddings: Optional[Tensor] = None, mems: Optional[List[Tensor]] = None,
                intermediates_cache: Optional[TransformerIntermediates] = None, return_hiddens: bool = False):
        assert not (self.cross_attend ^ (context is not None)), "context must be passed in if cross_attend is set to True"
        assert not self.ada_norm or style_embeddings is not None, "style_embeddings must be passed for AdaLayerNorm"
        if not x.is_cuda:
            raise RuntimeError("scoreperformer_b200 runs on CUDA only: there is no CPU fallback")
        if mems is not None or attn_mask is not None:
            raise NotImplementedError("memory tokens / explicit attn_mask are not used by any recipe")
        self._check_fused()
        if intermediates_cache is not None:
            from ...decode import cached_stack_step
            return cached_stack_step(self, x, mask, style_embeddings, intermediates_cache, return_hiddens)

        spec = self.stack_spec(return_hiddens)
        need_seeds = self.training and (spec.attn_dropout > 0 or spec.ff_dropout > 0)
        seeds = tuple(K.seed_from_torch() for _ in range(2 * self.depth)) if need_seeds else (0,) * (2 * self.depth)
        x = x.float() if x.dtype != torch.float32 else x
        mask = mask.contiguous() if mask is not None else None
        style = style_embeddings.float() if (style_embeddings is not None and self.ada_norm) else None
        out, hiddens, kvs = fused.TransformerStackFn.apply(spec, x, mask, style, seeds, *self.flat_params())
        if return_hiddens:
            b, t, _ = x.shape
            dh = spec.dim_head
            hid = [hiddens[l] for l in range(self.depth)] + [out]
            att = [AttentionIntermediates(keys=kvs[l][:, :dh].view(b, t, dh), values=kvs[l][:, dh:].view(b, t, dh))
                   for l in range(self.depth)]
            return out, TransformerIntermediates(hiddens=hid, attention=att)
        return out


@dataclass
class EncoderConfig(TransformerConfig):
    _target_: str = "encoder"
    causal: bool = False


@TransformerRegistry.register("encoder")
class Encoder(Transformer):
    def __init__(self, **kwargs):
        super().__init__(causal=False, **kwargs)


@dataclass
class DecoderConfig(TransformerConfig):
    _target_: str = "decoder"
    causal: bool = True


@TransformerRegistry.register("decoder")
class Decoder(Transformer):
    def __init__(self, **kwargs):
        super().__init__(causal=True, **kwargs)
