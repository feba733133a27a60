"""Multi-query attention block (reference: scoreperformer/modules/transformer/attention.py:27-222)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn as nn
from torch import Tensor

from ... import fused, kernels as K
from ...utils import default
from ..constructor import Constructor, ModuleConfig
from .attend import AttentionIntermediates
from .embeddings import ALiBiPositionalBias, LearnedALiBiPositionalBias


@dataclass
class AttentionSharedIntermediates:
    rel_pos_bias: Optional[Tensor] = None


@dataclass
class AttentionConfig(ModuleConfig):
    dim: int = 512
    dim_head: int = 64
    heads: int = 8
    causal: bool = False
    dropout: float = 0.
    one_kv_head: bool = False
    num_mem_kv: int = 0
    shared_kv: bool = False
    value_dim_head: Optional[int] = None
    max_attend_past: Optional[int] = None
    alibi_pos_bias: bool = False
    alibi_num_heads: Optional[int] = None
    alibi_symmetric: bool = True
    alibi_learned: bool = False


class Attention(nn.Module, Constructor):
    def __init__(self, dim: int, dim_head: int = 64, heads: int = 8, causal: bool = False, dropout: float = 0.,
                 one_kv_head: bool = False, num_mem_kv: int = 0, max_attend: Optional[int] = None, alibi_pos_bias: bool = False,
                 alibi_num_heads: Optional[int] = None, alibi_symmetric: bool = True, alibi_learned: bool = False):
        super().__init__()
        self.scale = dim_head ** -0.5
        self.heads = heads
        self.dim_head = dim_head
        self.causal = causal
        self.max_attend = max_attend
        self.dropout = dropout
        self.one_kv_head = one_kv_head
        out_dim = q_dim = dim_head * heads
        kv_dim = dim_head if one_kv_head else dim_head * heads

        self.to_q = nn.Linear(dim, q_dim, bias=False)
        self.to_k = nn.Linear(dim, kv_dim, bias=False)
        self.to_v = nn.Linear(dim, kv_dim, bias=False)

        self.rel_pos = None
        if alibi_pos_bias:
            alibi_num_heads = default(alibi_num_heads, heads)
            assert alibi_num_heads <= heads, "number of ALiBi heads must be less than the total number of heads"
            klass = LearnedALiBiPositionalBias if alibi_learned else ALiBiPositionalBias
            self.rel_pos = klass(heads=alibi_num_heads, total_heads=heads, symmetric=alibi_symmetric or causal)

        self.num_mem_kv = num_mem_kv
        if num_mem_kv > 0:
            self.mem_k = nn.Parameter(torch.randn(heads, num_mem_kv, dim_head))
            self.mem_v = nn.Parameter(torch.randn(heads, num_mem_kv, dim_head))

        self.to_out = nn.Linear(out_dim, dim, bias=False)

    @property
    def fused_supported(self) -> bool:
        """The sm_100a attention kernel covers the recipes' configuration: MQA, dim_head 64, symmetric (learned) ALiBi."""
        return (self.one_kv_head and self.dim_head == 64 and self.num_mem_kv == 0 and self.max_attend is None
                and self.rel_pos is not None and self.rel_pos.symmetric and self.rel_pos.heads == self.heads)

    def logslopes(self) -> Tensor:
        return self.rel_pos.get_logslopes()

    def forward(self, x: Tensor, context: Optional[Tensor] = None, mask: Optional[Tensor] = None, context_mask: Optional[Tensor] = None,
                attn_mask: Optional[Tensor] = None, prev_attn: Optional[Tensor] = None, mem: Optional[Tensor] = None,
                cache: Optional[AttentionIntermediates] = None, shared_cache: Optional[AttentionSharedIntermediates] = None):
        """Stand-alone block (the training path goes through the fused stack in Transformer.forward)."""
        if context is not None or mem is not None or attn_mask is not None or cache is not None or not self.fused_supported:
            raise NotImplementedError("scoreperformer_b200.Attention: only self-attention in the recipes' MQA/ALiBi configuration "
                                      "is implemented on sm_100a (cross-attention / memory / attn_mask are listed in DESIGN.md)")
        b, n, _ = x.shape
        w = torch.cat([self.to_q.weight, self.to_k.weight, self.to_v.weight], dim=0)
        qkv = fused.linear(x.reshape(b * n, -1), w)
        p = self.dropout if self.training else 0.0
        seed = K.seed_from_torch() if p > 0 else 0
        o = fused.AttentionCoreFn.apply(qkv, mask, self.logslopes(), b, n, self.heads, self.causal, p, seed)
        out = fused.linear(o, self.to_out.weight, out_fp32=True).view(b, n, -1)
        if mask is not None:
            out = out * mask[..., None]
        hq = self.heads * self.dim_head
        inter = AttentionIntermediates(keys=qkv[:, hq:hq + self.dim_head].view(b, n, -1), values=qkv[:, hq + self.dim_head:].view(b, n, -1))
        return out, inter, AttentionSharedIntermediates(rel_pos_bias=None)
