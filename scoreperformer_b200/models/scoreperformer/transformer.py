"""TupleTransformer: transformer over tuple-token sequences (reference: models/scoreperformer/transformer.py:23-222)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Union

import torch
import torch.nn as nn
from torch import Tensor

from ... import fused
from ...config import MISSING, DictConfig
from ...modules.constructor import Constructor, ModuleConfig
from ...modules.layers import LayerNorm
from ...modules.transformer import (AbsolutePositionalEmbedding, TransformerConfig, TransformerIntermediates, TransformerRegistry)
from ...utils import ExplicitEnum
from .embeddings import (TupleTokenEmbeddingsConfig, TupleTokenEmbeddingsRegistry, TupleTokenHeadsConfig, TupleTokenHeadsRegistry,
                         TupleTokenRegressionHead, TupleTokenRegressionHeadConfig)


class EmbeddingModes(ExplicitEnum):
    SUM = "mean"
    CONCAT = "cat"
    ATTENTION = "attention"
    ADANORM = "adanorm"


@dataclass
class TupleTransformerCaches:
    token_emb: Optional[Tensor] = None
    transformer: Optional[TransformerIntermediates] = None


@dataclass
class TupleTransformerOutput:
    hidden_state: Tensor
    logits: Optional[Dict[str, Tensor]] = None
    attentions: Optional[List[Tensor]] = None
    caches: Optional[TupleTransformerCaches] = None
    reg_values: Optional[Dict[str, Tensor]] = None


@dataclass
class TupleTransformerConfig(ModuleConfig):
    num_tokens: Dict[str, int] = MISSING
    dim: int = 512
    max_seq_len: int = 1024
    transformer: Union[DictConfig, TransformerConfig] = field(default_factory=lambda: TransformerConfig(_target_="default"))
    token_embeddings: Union[DictConfig, TupleTokenEmbeddingsConfig] = field(default_factory=TupleTokenEmbeddingsConfig)
    use_abs_pos_emb: bool = True
    emb_norm: bool = False
    emb_dropout: float = 0.0
    context_emb_dim: Optional[int] = None
    context_emb_mode: str = EmbeddingModes.ATTENTION
    style_emb_dim: Optional[int] = None
    style_emb_mode: str = EmbeddingModes.CONCAT
    lm_head: Optional[Union[DictConfig, TupleTokenHeadsConfig]] = None
    regression_head: Optional[Union[DictConfig, TupleTokenRegressionHeadConfig]] = None


def _cfg_get(cfg, key, default=None):
    return cfg.get(key, default) if hasattr(cfg, "get") else getattr(cfg, key, default)


class TupleTransformer(nn.Module, Constructor):
    def __init__(self, num_tokens: Dict[str, int], dim: int = 512, max_seq_len: int = 1024, transformer=None, token_embeddings=None,
                 use_abs_pos_emb: bool = True, emb_norm: bool = False, emb_dropout: float = 0.0, context_emb_dim: Optional[int] = None,
                 context_emb_mode: str = EmbeddingModes.ATTENTION, style_emb_dim: Optional[int] = None,
                 style_emb_mode: str = EmbeddingModes.CONCAT, lm_head=None, regression_head=None):
        super().__init__()
        transformer = transformer if transformer is not None else TransformerConfig(_target_="default")
        token_embeddings = token_embeddings if token_embeddings is not None else TupleTokenEmbeddingsConfig()

        self.dim = dim
        self.max_seq_len = max_seq_len
        emb_dim = dim
        self.context_emb_dim = context_emb_dim or 0
        self.context_emb_mode = context_emb_mode
        self.style_emb_dim = style_emb_dim or 0
        self.style_emb_mode = style_emb_mode

        self.token_emb = TupleTokenEmbeddingsRegistry.instantiate(
            config=token_embeddings, num_tokens=num_tokens, emb_dims=_cfg_get(token_embeddings, "emb_dims", emb_dim),
            project_emb_dim=emb_dim)

        if self.context_emb_mode != EmbeddingModes.ATTENTION:
            if isinstance(transformer, dict):
                transformer["cross_attend"] = False
            else:
                transformer.cross_attend = False

        self.transformer = TransformerRegistry.instantiate(
            transformer, dim=dim, use_adanorm=self.style_emb_mode == EmbeddingModes.ADANORM, style_emb_dim=self.style_emb_dim)

        self.pos_emb = None
        if use_abs_pos_emb:
            self.pos_emb = AbsolutePositionalEmbedding(emb_dim, self.max_seq_len)
            nn.init.kaiming_normal_(self.pos_emb.emb.weight)

        self.emb_norm = LayerNorm(emb_dim) if emb_norm else nn.Identity()
        self.emb_dropout = nn.Dropout(emb_dropout) if emb_dropout > 0. else nn.Identity()

        self.project_emb = nn.Identity()
        total_emb_dim = (emb_dim + int(context_emb_mode == EmbeddingModes.CONCAT) * self.context_emb_dim
                         + int(style_emb_mode == EmbeddingModes.CONCAT) * self.style_emb_dim)
        if total_emb_dim != dim:
            self.project_emb = nn.Linear(total_emb_dim, dim)

        self.lm_head = None
        if lm_head is not None:
            self.lm_head = TupleTokenHeadsRegistry.instantiate(config=lm_head, dim=dim, embeddings=self.token_emb)

        self.regression_head = None
        if regression_head is not None:
            assert self.token_emb.continuous, "TupleTokenRegressionHead depends on `continuous` token embeddings."
            self.regression_head = TupleTokenRegressionHead.init(config=regression_head, dim=dim)

    def embed_inputs(self, x, x_extra=None, style_embeddings=None, context=None, token_emb_cache=None, table_cache=None):
        """token_emb -> (+pos) -> emb_norm -> cat(context / style) -> project_emb; transformer.py:158-185.
        Returns (stream input fp32 [B,T,D], token_emb, remaining style, remaining context)."""
        table = self.token_emb.table(table_cache)
        if hasattr(self.token_emb, "multiseq_mode") and x_extra is not None:
            x_extra = [x_extra] if isinstance(x_extra, Tensor) else x_extra
            token_emb = self.token_emb([x] + x_extra, cache=token_emb_cache, table=table)
        else:
            token_emb = self.token_emb(x, cache=token_emb_cache, table=table)
        h = token_emb
        if self.pos_emb is not None:
            h = h.float() + self.pos_emb(h)
        cat_parts = []
        if context is not None and self.context_emb_mode == EmbeddingModes.CONCAT:
            cat_parts.append(context[:, :h.shape[1]])
            context = None
        if style_embeddings is not None:
            style_embeddings = style_embeddings[:, :h.shape[1]]
            if self.style_emb_mode == EmbeddingModes.CONCAT:
                cat_parts.append(style_embeddings)
                style_embeddings = None
        if isinstance(self.emb_norm, nn.Identity):
            h = h.float()
        else:
            h = self.emb_norm(h, out_fp32=not cat_parts)      # bf16 when it feeds the concat GEMM, fp32 residual stream otherwise
        if cat_parts:
            h = torch.cat([h] + [c.to(h.dtype) for c in cat_parts], dim=-1)
        h = self.emb_dropout(h)
        if not isinstance(self.project_emb, nn.Identity):
            h = fused.linear(h, self.project_emb.weight, self.project_emb.bias, out_fp32=True)
        return h.float(), token_emb, style_embeddings, context

    def forward(self, x: Tensor, mask: Optional[Tensor] = None, x_extra=None, style_embeddings: Optional[Tensor] = None,
                context: Optional[Tensor] = None, context_mask: Optional[Tensor] = None, caches: Optional[TupleTransformerCaches] = None,
                logits_keys: Optional[List] = None, return_embeddings: bool = False, return_attn: bool = False,
                return_caches: bool = False, table_cache: Optional[dict] = None, pre_embedded: Optional[tuple] = None, **kwargs):
        if return_attn:
            raise NotImplementedError("attention maps are never materialised by the fused attention kernel")
        table_cache = {} if table_cache is None else table_cache
        if pre_embedded is not None:
            # embed_inputs() was evaluated ahead of time (ScorePerformer.forward runs it next to the performance encoder, before
            # the style embeddings exist); only possible when the style does not enter through the concatenation
            assert self.style_emb_mode != EmbeddingModes.CONCAT and caches is None
            h, token_emb, context = pre_embedded
            if style_embeddings is not None:
                style_embeddings = style_embeddings[:, :h.shape[1]]
        else:
            h, token_emb, style_embeddings, context = self.embed_inputs(
                x, x_extra, style_embeddings, context, caches.token_emb if caches is not None else None, table_cache)
        res = self.transformer(h, mask=mask, context=context, context_mask=context_mask, style_embeddings=style_embeddings,
                               intermediates_cache=caches.transformer if caches is not None else None, return_hiddens=return_caches)
        out, intermediates = res if return_caches else (res, None)

        logits = None
        if not return_embeddings and self.lm_head is not None:
            logits = self.lm_head(out, keys=logits_keys)
        reg_values = None
        if not return_embeddings and self.regression_head is not None:
            reg_values = self.regression_head(out, keys=logits_keys)
        out_caches = TupleTransformerCaches(token_emb=token_emb, transformer=intermediates) if return_caches else None
        return TupleTransformerOutput(hidden_state=out, logits=logits, attentions=None, caches=out_caches, reg_values=reg_values)
