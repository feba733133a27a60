"""Language-modelling wrappers (reference: scoreperformer/models/scoreperformer/wrappers.py:19-444).

Training: the per-field cross-entropy of the tied head is the fused kernel pair of fused.TiedHeadCEFn; logits are
materialised lazily, only when a caller (the evaluator) reads `output.logits`.
Rendering: `unmask_tokens` / `generate` keep the reference's argument lists and cache contract.
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass
from typing import Callable, Dict, Optional

import torch
import torch.nn as nn
from torch import Tensor

from ... import fused
from ...modules.sampling import top_k
from ...utils import ExplicitEnum, exists
from .embeddings import TupleTokenLMHead, TupleTokenTiedLMHead
from .transformer import TupleTransformer, TupleTransformerCaches, TupleTransformerOutput


class LMWrapper(nn.Module):
    def __init__(self, model):
        super().__init__()
        self.model = model
        self.max_seq_len = self.model.max_seq_len

    def forward(self, seq, labels=None, **kwargs):
        ...


class LazyLogits(OrderedDict):
    """Field-ordered logits dict whose tensors are computed on first access (the training step never needs them)."""

    def __init__(self, keys, compute):
        super().__init__((k, None) for k in keys)
        self._compute = compute
        self._done = False

    def _materialise(self):
        if not self._done:
            self._done = True
            for k, v in self._compute().items():
                super().__setitem__(k, v)

    def __getitem__(self, key):
        self._materialise()
        return super().__getitem__(key)

    def values(self):
        self._materialise()
        return super().values()

    def items(self):
        self._materialise()
        return super().items()


@dataclass
class ScorePerformerLMOutput(TupleTransformerOutput):
    loss: Optional[Tensor] = None
    losses: Optional[Dict[str, Tensor]] = None
    eval_stats: Optional[Dict[str, Tensor]] = None


class ScorePerformerLMWrapper(LMWrapper):
    def __init__(self, model: TupleTransformer, ignore_index: int = -100):
        super().__init__(model=model)
        self.ignore_index = ignore_index
        # Fields that ever carry labels.  The reference decides per batch with `torch.any(labels[..., i] != ignore)` (a host
        # sync per field, wrappers.py:56); here it is decided once (first batch) and fields without labels contribute
        # nothing on the device.  Set to a tuple of field indices to skip the probe entirely.
        self.label_fields = None
        # name -> fp32 [V] token values; set by ScorePerformerEvaluator.attach(): the head kernel then accumulates the evaluator's
        # statistics while the logits are in tensor memory, and `forward` returns them as `eval_stats`
        self.eval_token_values = None
        self._unexpected_labels = None
        self._excluded_index = {}

    def _probe_label_fields(self, labels: Tensor):
        """Fields that carry labels.  The reference asks every batch (wrappers.py:49-59, a host sync per field); here the set is
        probed on the first batch (or given: `label_fields`) and then frozen, so the head only evaluates those fields and the
        step can be captured.  Labels that later show up in an excluded field are COUNTED on the device
        (`unexpected_label_count()`, read by TrainStep.unexpected_label_count) instead of silently ignored."""
        n_fields = labels.shape[-1]
        if self.label_fields is None:
            has = (labels != self.ignore_index).reshape(-1, n_fields).any(dim=0)
            self.label_fields = tuple(int(i) for i in torch.nonzero(has).flatten().tolist())
        excluded = [i for i in range(n_fields) if i not in self.label_fields]
        if excluded and labels.is_cuda and self.training:
            if self._unexpected_labels is None or self._unexpected_labels.device != labels.device:
                self._unexpected_labels = torch.zeros((), dtype=torch.int64, device=labels.device)
            idx = self._excluded_index.get((labels.device, tuple(excluded)))
            if idx is None:
                idx = self._excluded_index[(labels.device, tuple(excluded))] = torch.tensor(excluded, device=labels.device)
            self._unexpected_labels += (labels.reshape(-1, n_fields).index_select(1, idx) != self.ignore_index).sum()
        return self.label_fields

    def unexpected_label_count(self) -> int:
        """Labels seen so far in fields outside `label_fields` (they did not enter the loss).  Non-zero means the frozen field set
        is too small for this data: set `label_fields` explicitly (or to None to re-probe).  Reading it synchronises."""
        return 0 if self._unexpected_labels is None else int(self._unexpected_labels)

    def forward(self, seq: Tensor, labels: Optional[Tensor] = None, **kwargs):
        head = self.model.lm_head
        fused_head = isinstance(head, TupleTokenTiedLMHead) and self.model.regression_head is None and exists(labels)
        if not fused_head:
            if exists(labels) and isinstance(head, TupleTokenLMHead) and self.model.regression_head is None:
                # untied `lm` head (ablation recipe no_io_tie): one GEMM + one masked-CE kernel per labelled field, the logits of a
                # field live only between the two (wrappers.py:45-59 of the reference)
                kwargs.pop("return_embeddings", None)
                out = self.model(seq, return_embeddings=True, **kwargs)
                hidden = out.hidden_state
                b, t, d = hidden.shape
                names = list(head.heads.keys())
                fields = self._probe_label_fields(labels)
                params = [p for f in fields for p in (head.heads[names[f]].weight, head.heads[names[f]].bias)]
                loss, per_field, _ = fused.UntiedHeadCEFn.apply(hidden.reshape(b * t, d), labels.reshape(b * t, -1).contiguous(),
                                                                self.ignore_index, tuple(fields), *params)
                out.logits = LazyLogits(names, lambda: head(hidden.detach()))
                return ScorePerformerLMOutput(loss=loss, losses={names[f]: per_field[i] for i, f in enumerate(fields)}, **out.__dict__)
            out = self.model(seq, **kwargs)
            loss = losses = None
            if exists(labels):
                raise NotImplementedError("training with a regression head is not implemented on the sm_100a path (no recipe of the "
                                          "reference uses it); see DESIGN.md")
            return ScorePerformerLMOutput(loss=loss, losses=losses, **out.__dict__)

        table_cache = kwargs.get("table_cache")
        table_cache = {} if table_cache is None else table_cache
        kwargs["table_cache"] = table_cache
        kwargs.pop("return_embeddings", None)
        out = self.model(seq, return_embeddings=True, **kwargs)
        hidden = out.hidden_state
        b, t, d = hidden.shape
        names = list(head.embs.keys())
        fields = self._probe_label_fields(labels)
        table = head.table(table_cache)
        tvs = None
        if self.eval_token_values is not None:
            tvs = tuple(self._token_values_on(names[f], hidden.device) if f in fields else None for f in range(len(names)))
        loss, per_field, count, stats = fused.TiedHeadCEFn.apply(
            hidden.reshape(b * t, d), head.proj_weight_kn(), head.norm.weight, head.norm.bias, table,
            labels.reshape(b * t, -1).contiguous(), head.field_sizes, fields, head.split_dims[0], self.ignore_index, tvs)
        losses = {names[f]: per_field[f] for f in fields}
        logits = LazyLogits(names, lambda: head(hidden.detach(), table=table.detach()))
        out.logits = logits
        res = ScorePerformerLMOutput(loss=loss, losses=losses, **out.__dict__)
        if tvs is not None:
            # per labelled field: (hits, labelled rows, sum |tv[argmax]-tv[label]|, sum_v p_v |tv[label]-tv[v]|)
            res.eval_stats = {names[f]: torch.cat([stats[f, :1], count[f:f + 1], stats[f, 1:]]) for f in fields}
        return res

    def _token_values_on(self, name: str, device) -> Optional[Tensor]:
        tv = self.eval_token_values.get(name)
        if tv is None:
            return None
        if tv.device != device or tv.dtype != torch.float32 or not tv.is_contiguous():
            tv = self.eval_token_values[name] = tv.to(device=device, dtype=torch.float32).contiguous()
        return tv


class ScorePerformerMLMWrapper(ScorePerformerLMWrapper):
    def __init__(self, model: TupleTransformer, mask_token_id: int = 1, num_special_tokens: int = 4, ignore_index: int = -100):
        super().__init__(model=model, ignore_index=ignore_index)
        self.mask_token_id = mask_token_id
        self.num_special_tokens = num_special_tokens

    def unmask_tokens(self, tokens: Tensor, single_run: bool = True, temperature: float = 1., filter_logits_fn: Callable = top_k,
                      filter_kwargs=None, filter_key_ids=None, disable_tqdm: bool = False, **kwargs):
        """Reference signature (wrappers.py:131-198); the work is decode.unmask_mlm."""
        from ... import decode
        return decode.unmask_mlm(self, tokens, single_run, temperature, filter_logits_fn, filter_kwargs, filter_key_ids, **kwargs)


class ScorePerformerARWrapper(ScorePerformerLMWrapper):
    def __init__(self, model: TupleTransformer, pad_token_id: int = 0, eos_token_id: int = 3, num_special_tokens: int = 4,
                 ignore_index: int = -100):
        super().__init__(model=model, ignore_index=ignore_index)
        self.pad_token_id = pad_token_id
        self.eos_token_id = eos_token_id
        self.num_special_tokens = num_special_tokens

    def generate(self, start_tokens: Tensor, seq_len: int, max_bar: Optional[int] = None, temperature: float = 1.,
                 filter_logits_fn: Callable = top_k, filter_kwargs=None, caches: Optional[TupleTransformerCaches] = None,
                 return_caches: bool = False, tokenizer=None, fix_errors: bool = True, disable_tqdm: bool = False, **kwargs):
        """Reference signature (wrappers.py:200-307); the work is decode.generate_ar (cached stack steps)."""
        from ... import decode
        return decode.generate_ar(self, start_tokens, seq_len, max_bar, temperature, filter_logits_fn, filter_kwargs, caches,
                                  return_caches, tokenizer, fix_errors, **kwargs)

    def forward(self, seq: Tensor, labels: Optional[Tensor] = None, **kwargs):
        seq = seq[:, :-1]
        labels = labels[:, 1:] if exists(labels) else None
        context = kwargs.get("context", None)
        if exists(context) and self.model.context_emb_mode == "cat":
            kwargs["context"] = context[:, 1:]
        style_embeddings = kwargs.get("style_embeddings", None)
        if exists(style_embeddings):
            kwargs["style_embeddings"] = style_embeddings[:, 1:]
        mask = kwargs.get("mask", None)
        if exists(mask) and mask.shape[1] == seq.shape[1] + 1:
            kwargs["mask"] = mask[:, :-1]
        return super().forward(seq, labels=labels, **kwargs)


class ScorePerformerMixedLMWrapper(ScorePerformerLMWrapper):
    def __init__(self, model: TupleTransformer, pad_token_id: int = 0, mask_token_id: int = 1, num_special_tokens: int = 4,
                 ignore_index: int = -100):
        super().__init__(model=model, ignore_index=ignore_index)
        self.pad_token_id = pad_token_id
        self.mask_token_id = mask_token_id
        self.num_special_tokens = num_special_tokens

    def unmask_tokens(self, tokens: Tensor, tokens_masked, temperature: float = 1., filter_logits_fn: Callable = top_k,
                      filter_kwargs=None, filter_key_ids=None, caches: Optional[TupleTransformerCaches] = None,
                      return_caches: bool = False, disable_tqdm: bool = False, **kwargs):
        """Reference signature and cache contract (wrappers.py:324-407).  A whole window with top-k / greedy sampling and no incoming
        caches runs in decode.render_decoder's device-resident loop (any batch size: one persistent kernel per note for the
        stack, one for heads + sampling); anything else takes decode's general cached stepper."""
        from ... import decode
        return decode.unmask_mixlm(self, tokens, tokens_masked, temperature, filter_logits_fn, filter_kwargs, filter_key_ids, caches,
                                   return_caches, **kwargs)

    def pre_embed(self, seq: Tensor, seq_masked: Optional[Tensor] = None, context: Optional[Tensor] = None,
                  table_cache: Optional[dict] = None):
        """The input embedding of forward() (same shifts, TupleTransformer.embed_inputs) for callers that can evaluate it before
        the style embeddings exist; hand the result to forward(pre_embedded=...).  None when the style enters by concatenation."""
        if self.model.style_emb_mode == "cat":
            return None
        seq = seq[:, :-1]
        if exists(seq_masked):
            seq_masked = seq_masked[:, 1:]
        if exists(context) and self.model.context_emb_mode == "cat":
            context = context[:, 1:]
        h, token_emb, _, context = self.model.embed_inputs(seq, seq_masked, None, context, None, table_cache)
        return h, token_emb, context

    def forward(self, seq: Tensor, labels: Optional[Tensor] = None, **kwargs):
        seq = seq[:, :-1]
        labels = labels[:, 1:] if exists(labels) else None
        seq_masked = kwargs.pop("seq_masked", None)
        if exists(seq_masked):
            seq_masked = seq_masked[:, 1:]
        context = kwargs.get("context", None)
        if exists(context) and self.model.context_emb_mode == "cat":
            kwargs["context"] = context[:, 1:]
        style_embeddings = kwargs.get("style_embeddings", None)
        if exists(style_embeddings):
            kwargs["style_embeddings"] = style_embeddings[:, 1:]
        mask = kwargs.get("mask", None)
        if exists(mask) and mask.shape[1] == seq.shape[1] + 1:
            kwargs["mask"] = mask[:, :-1]
        return super().forward(seq, labels=labels, x_extra=seq_masked, **kwargs)


class ScorePerformerLMModes(ExplicitEnum):
    MLM = "mlm"
    CLM = "clm"
    MixedLM = "mixlm"


ScorePerformerLMWrappers = {
    ScorePerformerLMModes.MLM: ScorePerformerMLMWrapper,
    ScorePerformerLMModes.CLM: ScorePerformerARWrapper,
    ScorePerformerLMModes.MixedLM: ScorePerformerMixedLMWrapper,
}
