"""Tuple-token embeddings and language-modelling heads.

Reference: scoreperformer/models/scoreperformer/embeddings.py:30-462.  Same registries / constructor arguments /
parameter names; the forward of the recipes' configuration (mode `cat`, `emb_norm`, uniform 128-wide fields, continuous
dense tables) is the fused gather -> LayerNorm kernel + tcgen05 projection (scoreperformer_b200.fused.TupleEmbedFn).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple, Union

import torch
import torch.nn as nn
from torch import Tensor

from ... import fused
from ...config import MISSING
from ...modules.constructor import Constructor, Registry, VariableModuleConfig
from ...modules.transformer import DiscreteContinuousEmbedding, DiscreteDenseContinuousEmbedding

TupleTokenEmbeddingsRegistry = type("_TupleTokenEmbeddingsRegistry", (Registry,), {})()


def _fused_table_ok(embs: nn.ModuleDict) -> bool:
    for emb in embs.values():
        if not isinstance(emb, DiscreteDenseContinuousEmbedding) or emb.embedding_dim != 128 or emb.discrete or not emb.continuous:
            return False
        if emb.discrete_ids is None or emb.token_values is None or len(emb.value_layer) != 2 or not emb.index_weight.is_cuda:
            return False
    return len(embs) <= 16


def build_table(embs: nn.ModuleDict, cache: Optional[dict] = None) -> Tensor:
    """Concatenated computed tables of the field modules.  The recipes' configuration (dense continuous embeddings, 128 wide)
    is built by ONE fused kernel (fused.TableBuildFn); anything else falls back to the per-field module arithmetic."""
    key = ("cat",) + tuple(id(e) for e in embs.values())
    if cache is not None:
        if key in cache:
            return cache[key]
        for other, tab in cache.items():        # tied score fields are a prefix of the performance fields: slice, don't rebuild
            if isinstance(other, tuple) and other and other[0] == "cat" and len(other) > len(key) and other[:len(key)] == key:
                rows = sum(int(e.num_embeddings) for e in embs.values())
                cache[key] = tab[:rows]
                return cache[key]
    if _fused_table_ok(embs):
        sizes = tuple(int(e.num_embeddings) for e in embs.values())
        consts, params = [], []
        for e in embs.values():
            consts += [e.token_values, e._discrete_mask]
            params += [e.index_weight, e.value_layer[0][0].weight, e.value_layer[0][0].bias, e.value_layer[1][0].weight,
                       e.value_layer[1][0].bias]
        table = fused.TableBuildFn.apply(sizes, tuple(consts), *params)
    else:
        table = torch.cat([e.weight for e in embs.values()], dim=0)
    if cache is not None:
        cache[key] = table
    return table


@dataclass
class TupleTokenEmbeddingsConfig(VariableModuleConfig):
    _target_: str = "simple"
    num_tokens: Dict[str, int] = MISSING
    emb_dims: Union[Dict[str, int], int] = MISSING
    mode: str = "cat"
    project_emb_dim: int = 512
    emb_norm: bool = False
    discrete: bool = True
    continuous: Union[bool, List[str]] = False
    continuous_dense: bool = False
    token_values: Optional[Dict[str, list]] = None
    discrete_ids: Optional[List[int]] = None
    tie_keys: Optional[Dict[str, str]] = None


@TupleTokenEmbeddingsRegistry.register("simple")
class TupleTokenEmbeddings(nn.Module, Constructor):
    def __init__(self, num_tokens: Dict[str, int], emb_dims: Union[Dict[str, int], int], mode: str = "cat", project_emb_dim: int = 512,
                 emb_norm: bool = False, discrete: bool = True, continuous: Union[bool, List[str]] = False,
                 continuous_dense: bool = False, token_values: Optional[Dict[str, list]] = None,
                 discrete_ids: Optional[List[int]] = None, tie_keys: Optional[Dict[str, str]] = None):
        super().__init__()
        self.mode = mode
        if self.mode == "sum":
            assert isinstance(emb_dims, int) or all(d == list(emb_dims.values())[0] for d in emb_dims.values()), \
                "`emb_dims` in TupleTokenEmbeddings' `sum` mode should be the same for all keys."

        continuous_keys = continuous
        if isinstance(continuous, bool):
            continuous_keys = [key for key in num_tokens] if continuous else []
        else:
            continuous = len(continuous) > 0

        total_emb_dim = 0
        embeddings = {}
        token_values = token_values or {}
        for key, num in num_tokens.items():
            emb_dim = emb_dims if isinstance(emb_dims, int) else emb_dims[key]
            if tie_keys and key in tie_keys:
                embeddings[key] = embeddings[tie_keys[key]]
                emb_dim = emb_dims if isinstance(emb_dims, int) else emb_dims[tie_keys[key]]
            elif key in continuous_keys:
                cls = DiscreteDenseContinuousEmbedding if continuous_dense else DiscreteContinuousEmbedding
                embeddings[key] = cls(num_embeddings=num, embedding_dim=emb_dim, discrete=discrete, continuous=True,
                                      discrete_ids=discrete_ids, token_values=token_values.get(key, None), padding_idx=0)
            else:
                embeddings[key] = nn.Embedding(num, emb_dim, padding_idx=0)
            total_emb_dim += emb_dim if self.mode == "cat" else emb_dim - total_emb_dim

        self.embs = nn.ModuleDict(embeddings)
        self.norm = nn.LayerNorm(total_emb_dim) if emb_norm else nn.Identity()
        if total_emb_dim != project_emb_dim:
            self.project_emb = nn.Linear(total_emb_dim, project_emb_dim)

        self.num_tokens = dict(num_tokens)
        self.emb_dims = emb_dims
        self.total_emb_dim = total_emb_dim
        self.continuous = continuous
        self.continuous_keys = continuous_keys
        self.token_values = token_values
        self.init_()

    def init_(self):
        if not self.continuous:
            for key, emb in self.embs.items():
                weight_attr = "index_weight" if key in self.continuous_keys else "weight"
                nn.init.kaiming_normal_(getattr(emb, weight_attr))

    # ------------------------------------------------------------------ fused path
    @property
    def field_sizes(self) -> Tuple[int, ...]:
        return tuple(int(v) for v in self.num_tokens.values())

    @property
    def fused_supported(self) -> bool:
        dims = {e.embedding_dim for e in self.embs.values()}
        return (self.mode == "cat" and isinstance(self.norm, nn.LayerNorm) and dims == {128} and hasattr(self, "project_emb")
                and len(self.embs) in (10, 12))

    def table(self, cache: Optional[dict] = None) -> Tensor:
        """Concatenated per-field tables [sum V_f, 128] (fp32, differentiable).  `cache` (one dict per model forward) makes
        tied field modules evaluate their value MLP once per step instead of once per stack."""
        return build_table(self.embs, cache)

    def embed(self, tokens: Tensor, table: Optional[Tensor] = None) -> Tensor:
        """tokens int64 [..., F] -> bf16 [..., project_emb_dim]."""
        if not self.fused_supported:
            raise NotImplementedError(
                "scoreperformer_b200.TupleTokenEmbeddings: the sm_100a gather kernel implements the recipes' layout (mode `cat`, "
                "emb_norm, 128-wide fields, 10 or 12 fields); `sum` mode / mixed widths / discrete-only tables are not implemented")
        table = self.table() if table is None else table
        shape = tokens.shape
        y = fused.TupleEmbedFn.apply(tokens.reshape(-1, shape[-1]), table, self.norm.weight, self.norm.bias, self.project_emb.weight,
                                     self.project_emb.bias, self.field_sizes)
        return y.view(*shape[:-1], y.shape[-1])

    def forward(self, x: Tensor, values: Optional[Tensor] = None, cache: Optional[Tensor] = None, return_embeddings: bool = False,
                table: Optional[Tensor] = None):
        if values is not None or return_embeddings:
            raise NotImplementedError("explicit `values` / `return_embeddings` are not used by the training or rendering path")
        if cache is not None:
            x = x[:, cache.shape[1]:]
        token_emb = self.embed(x, table)
        if cache is not None:
            token_emb = torch.cat([cache.to(token_emb.dtype), token_emb], dim=1)
        return token_emb


@dataclass
class MultiSeqTupleTokenEmbeddingsConfig(TupleTokenEmbeddingsConfig):
    _target_: str = "multi-seq"
    multiseq_mode: str = "pre-sum"
    num_sequences: int = 2


@TupleTokenEmbeddingsRegistry.register("multi-seq")
class MultiSeqTupleTokenEmbeddings(TupleTokenEmbeddings):
    def __init__(self, num_tokens: Dict[str, int], emb_dims: Union[Dict[str, int], int], mode: str = "cat", project_emb_dim: int = 512,
                 emb_norm: bool = False, discrete: bool = True, continuous: Union[bool, List[str]] = False,
                 continuous_dense: bool = False, token_values: Optional[Dict[str, list]] = None,
                 discrete_ids: Optional[List[int]] = None, tie_keys: Optional[Dict[str, str]] = None,
                 multiseq_mode: str = "pre-sum", num_sequences: int = 2):
        super().__init__(num_tokens=num_tokens, emb_dims=emb_dims, mode=mode, project_emb_dim=project_emb_dim, emb_norm=emb_norm,
                         discrete=discrete, continuous=continuous, continuous_dense=continuous_dense, token_values=token_values,
                         discrete_ids=discrete_ids, tie_keys=tie_keys)
        self.multiseq_mode = multiseq_mode
        self.num_sequences = num_sequences
        if self.multiseq_mode == "post-cat":
            self.project_multiemb = nn.Linear(num_sequences * project_emb_dim, project_emb_dim)

    def forward(self, tokens: Union[Tensor, List[Tensor]], values=None, cache: Optional[Tensor] = None, return_embeddings: bool = False,
                table: Optional[Tensor] = None):
        if isinstance(tokens, list) and len(tokens) == 1:
            tokens = tokens[0]
        if isinstance(tokens, Tensor):
            return super().forward(tokens, values=values, cache=cache, return_embeddings=return_embeddings, table=table)
        if values is not None or return_embeddings:
            raise NotImplementedError("explicit `values` / `return_embeddings` are not used by the training or rendering path")
        if cache is not None:
            tokens = [t[:, cache.shape[1]:] for t in tokens]
        table = self.table() if table is None else table
        if self.multiseq_mode == "post-cat":
            assert len(tokens) == self.num_sequences
            projected = [self.embed(t, table) for t in tokens]
            token_emb = fused.linear(torch.cat(projected, dim=-1), self.project_multiemb.weight, self.project_multiemb.bias,
                                     out_fp32=True)
        elif self.multiseq_mode == "post-sum":
            token_emb = sum(self.embed(t, table).float() for t in tokens)
        else:
            raise NotImplementedError("multiseq_mode `pre-sum` is not implemented on the sm_100a path (no recipe uses it)")
        if cache is not None:
            token_emb = torch.cat([cache.to(token_emb.dtype), token_emb], dim=1)
        return token_emb


TupleTokenHeadsRegistry = type("_TupleTokenHeadsRegistry", (Registry,), {})()


@dataclass
class TupleTokenHeadsConfig(VariableModuleConfig):
    dim: int = MISSING


@dataclass
class TupleTokenLMHeadConfig(TupleTokenHeadsConfig):
    _target_: str = "lm"
    num_tokens: Optional[Dict[str, int]] = None
    embeddings: Optional[TupleTokenEmbeddings] = None
    filter_keys: Optional[List[str]] = None


def _select(keys, i, key):
    return keys is None or i in keys or key in keys


@TupleTokenHeadsRegistry.register("lm")
class TupleTokenLMHead(nn.Module, Constructor):
    def __init__(self, dim: int, num_tokens: Optional[Dict[str, int]] = None, embeddings: Optional[TupleTokenEmbeddings] = None,
                 filter_keys: Optional[List[str]] = None):
        assert num_tokens is not None or embeddings is not None
        super().__init__()
        num_tokens = num_tokens or embeddings.num_tokens
        self.heads = nn.ModuleDict({key: nn.Linear(dim, num) for key, num in num_tokens.items() if not filter_keys or key in filter_keys})

    def forward(self, x: Tensor, keys=None):
        return {key: fused.linear(x, head.weight, head.bias, out_fp32=True)
                for i, (key, head) in enumerate(self.heads.items()) if _select(keys, i, key)}


@dataclass
class TupleTokenTiedLMHeadConfig(TupleTokenHeadsConfig):
    _target_: str = "lm-tied"
    embeddings: TupleTokenEmbeddings = MISSING
    reuse_projection: bool = True


@TupleTokenHeadsRegistry.register("lm-tied")
class TupleTokenTiedLMHead(nn.Module, Constructor):
    def __init__(self, dim: int, embeddings: TupleTokenEmbeddings, reuse_projection: bool = True):
        super().__init__()
        self.embs = embeddings.embs
        self.total_emb_dim = embeddings.total_emb_dim
        self.split_dims = [e.embedding_dim for e in embeddings.embs.values()]
        self.num_tokens = dict(embeddings.num_tokens)
        if reuse_projection:
            assert dim == embeddings.project_emb.out_features, \
                f"Projection layer could be reused only if last input tensor dimension is equal to projection layer's " \
                f"`out_features = {embeddings.project_emb.out_features}`"
            self.project_emb = embeddings.project_emb
            self.proj_is_kn = True            # x @ W with W = project_emb.weight [dim, total]
        else:
            self.project_emb = nn.Linear(dim, self.total_emb_dim, bias=False)
            self.proj_is_kn = False
        self.norm = nn.LayerNorm(self.total_emb_dim)

    @property
    def field_sizes(self) -> Tuple[int, ...]:
        return tuple(int(v) for v in self.num_tokens.values())

    def proj_weight_kn(self) -> Tensor:
        """[dim, total_emb_dim] view of the head projection (embeddings.py:346: `x @ self.project_emb.weight`)."""
        return self.project_emb.weight if self.proj_is_kn else self.project_emb.weight.t()

    def table(self, cache: Optional[dict] = None) -> Tensor:
        return build_table(self.embs, cache)

    def forward(self, x: Tensor, keys=None, table: Optional[Tensor] = None):
        """Logits dict (fp32), ordered by field; inference / evaluation path (training uses the fused CE)."""
        names = list(self.embs.keys())
        fields = [i for i, key in enumerate(names) if _select(keys, i, key)]
        shape = x.shape
        table = self.table() if table is None else table
        outs = fused.tied_head_logits(x.reshape(-1, shape[-1]), self.proj_weight_kn().contiguous(), self.norm.weight, self.norm.bias,
                                      table, self.field_sizes, fields, self.split_dims[0])
        return {names[f]: o.view(*shape[:-1], o.shape[-1]) for f, o in zip(fields, outs)}


@dataclass
class TupleTokenTiedSplitLMHeadConfig(TupleTokenHeadsConfig):
    _target_: str = "lm-tied-split"
    embeddings: TupleTokenEmbeddings = MISSING
    filter_keys: Optional[List[str]] = None


@TupleTokenHeadsRegistry.register("lm-tied-split")
class TupleTokenTiedSplitLMHead(nn.Module, Constructor):
    def __init__(self, dim: int, embeddings: TupleTokenEmbeddings, filter_keys: Optional[List[str]] = None):
        super().__init__()
        self.to_embs = nn.ModuleDict({
            key: nn.Sequential(nn.Linear(dim, emb.embedding_dim), nn.LayerNorm(emb.embedding_dim))
            for key, emb in embeddings.embs.items() if not filter_keys or key in filter_keys})
        self.embs = embeddings.embs

    def forward(self, x: Tensor, keys=None):
        raise NotImplementedError("`lm-tied-split` heads are constructed for checkpoint compatibility; their sm_100a forward is not "
                                  "implemented (no recipe uses them)")


@dataclass
class TupleTokenRegressionHeadConfig(TupleTokenHeadsConfig):
    _target_: str = "regression"
    regression_keys: List[str] = MISSING


@TupleTokenHeadsRegistry.register("regression")
class TupleTokenRegressionHead(nn.Module, Constructor):
    def __init__(self, dim: int, regression_keys: List[str]):
        super().__init__()
        self.layers = nn.ModuleDict({key: nn.Linear(dim, 1) for key in regression_keys})

    def forward(self, x: Tensor, keys=None):
        return {key: fused.linear(x, layer.weight, layer.bias, out_fp32=True)
                for i, (key, layer) in enumerate(self.layers.items()) if _select(keys, i, key)}


@dataclass
class TupleTokenEmbeddingHeadConfig(TupleTokenHeadsConfig):
    _target_: str = "embedding"
    emb_dim: int = MISSING
    hidden_dim: Optional[int] = None
    depth: int = 2
    detach_inputs: Union[bool, float] = True


@TupleTokenHeadsRegistry.register("embedding")
class TupleTokenEmbeddingHead(nn.Module, Constructor):
    def __init__(self, dim: int, emb_dim: int, hidden_dim: Optional[int] = None, depth: int = 2, detach_inputs: Union[bool, float] = True):
        super().__init__()
        hidden_dim = hidden_dim or emb_dim
        input_dims = [dim] + [hidden_dim] * (depth - 1)
        output_dims = [hidden_dim] * (depth - 1) + [emb_dim]
        layers = []
        for i, (in_dim, out_dim) in enumerate(zip(input_dims, output_dims)):
            layers.append(nn.Linear(in_dim, out_dim))
            if i < depth - 1:
                layers.append(nn.Mish())
        self.layers = nn.Sequential(*layers)
        self.detach_inputs = detach_inputs

    def forward(self, x: Tensor):
        raise NotImplementedError("`embedding` heads are constructed for checkpoint compatibility only (no recipe uses them)")
