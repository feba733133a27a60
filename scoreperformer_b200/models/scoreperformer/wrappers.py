"""Language-modelling wrappers (reference: scoreperformer/models/scoreperformer/wrappers.py:19-444).

Training: the per-field cross-entropy of the tied head is the fused kernel pair of fused.TiedHeadCEFn; logits are
materialised lazily, only when a caller (the evaluator) reads `output.logits`.
Rendering: `unmask_tokens` / `generate` keep the reference's argument lists and cache contract.
"""
from __future__ import annotations

import warnings
from collections import OrderedDict
from dataclasses import dataclass
from typing import Callable, Dict, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor

from ... import fused
from ...modules.sampling import filter_logits_and_sample, top_k
from ...utils import ExplicitEnum, exists
from .embeddings import TupleTokenTiedLMHead
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

    def _probe_label_fields(self, labels: Tensor):
        if self.label_fields is None:
            has = (labels != self.ignore_index).reshape(-1, labels.shape[-1]).any(dim=0)
            self.label_fields = tuple(int(i) for i in torch.nonzero(has).flatten().tolist())
        return self.label_fields

    def forward(self, seq: Tensor, labels: Optional[Tensor] = None, **kwargs):
        head = self.model.lm_head
        fused_head = isinstance(head, TupleTokenTiedLMHead) and self.model.regression_head is None and exists(labels)
        if not fused_head:
            out = self.model(seq, **kwargs)
            loss = losses = None
            if exists(labels):
                # generic (untied `lm` head) path: logits come from the head, CE per field through the fused CE kernel
                raise NotImplementedError("training with an untied `lm` head / regression head is not implemented on the sm_100a "
                                          "path yet (ablation recipes no_io_tie); see DESIGN.md")
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


def _sample_fields(logits: Dict[str, Tensor], banned, filter_key_ids, filter_logits_fn, filter_kwargs, temperature):
    samples = []
    for key, lg in logits.items():
        for tok in banned:
            lg[:, tok] = -float("Inf")
        ids = filter_key_ids.get(key, None)
        if ids is not None:
            lg[:, ids] = -float("Inf")
        samples.append(filter_logits_and_sample(lg, filter_logits_fn, filter_kwargs=filter_kwargs, temperature=temperature))
    return torch.cat(samples, dim=-1)[None]


class ScorePerformerMLMWrapper(ScorePerformerLMWrapper):
    def __init__(self, model: TupleTransformer, mask_token_id: int = 1, num_special_tokens: int = 4, ignore_index: int = -100):
        super().__init__(model=model, ignore_index=ignore_index)
        self.mask_token_id = mask_token_id
        self.num_special_tokens = num_special_tokens

    @torch.inference_mode()
    def unmask_tokens(self, tokens: Tensor, single_run: bool = True, temperature: float = 1., filter_logits_fn: Callable = top_k,
                      filter_kwargs=None, filter_key_ids=None, disable_tqdm: bool = False, **kwargs):
        assert callable(filter_logits_fn)
        was_training = self.model.training
        if was_training:
            self.model.eval()
        num_dims = len(tokens.shape)
        if num_dims == 2:
            tokens = tokens[None, :]
        out = tokens.clone().detach()
        mask = kwargs.pop("mask", None)
        if mask is None:
            mask = torch.full_like(out[..., 0], True, dtype=torch.bool, device=out.device)
        filter_key_ids = filter_key_ids or dict()
        unmask_mask = out == self.mask_token_id
        if single_run:
            warnings.warn("`single_run` unmasking with sampling is not yet implemented, using argmax.")
            outputs = self.model(out, mask=mask, **kwargs)
            samples = torch.cat([torch.argmax(l, dim=-1, keepdim=True) for l in outputs.logits.values()], dim=-1)
            out[unmask_mask] = samples[unmask_mask]
        else:
            unmask_ids = torch.where(torch.any(unmask_mask, dim=2))[1]
            for idx in unmask_ids.tolist():
                type_mask = unmask_mask[:, idx][0]
                logits_keys = torch.where(type_mask)[0].tolist()
                outputs = self(out[:, :idx + 1], mask=mask[:, :idx + 1], return_embeddings=True, **kwargs)
                logits = self.model.lm_head(outputs.hidden_state[:, idx - 1], keys=logits_keys)
                out[:, idx, type_mask] = _sample_fields(logits, range(self.num_special_tokens), filter_key_ids, filter_logits_fn,
                                                        filter_kwargs, temperature)
        if num_dims == 2:
            out = out.squeeze(0)
        if was_training:
            self.model.train(was_training)
        return out


class ScorePerformerARWrapper(ScorePerformerLMWrapper):
    def __init__(self, model: TupleTransformer, pad_token_id: int = 0, eos_token_id: int = 3, num_special_tokens: int = 4,
                 ignore_index: int = -100):
        super().__init__(model=model, ignore_index=ignore_index)
        self.pad_token_id = pad_token_id
        self.eos_token_id = eos_token_id
        self.num_special_tokens = num_special_tokens

    @torch.inference_mode()
    def generate(self, start_tokens: Tensor, seq_len: int, max_bar: Optional[int] = None, temperature: float = 1.,
                 filter_logits_fn: Callable = top_k, filter_kwargs=None, caches: Optional[TupleTransformerCaches] = None,
                 return_caches: bool = False, tokenizer=None, fix_errors: bool = True, disable_tqdm: bool = False, **kwargs):
        assert callable(filter_logits_fn)
        was_training = self.model.training
        if was_training:
            self.model.eval()
        num_dims = len(start_tokens.shape)
        if num_dims == 2:
            start_tokens = start_tokens[None, :]
        b, t = start_tokens.shape[:2]
        out = start_tokens
        mask = kwargs.pop("mask", None)
        if mask is None:
            mask = torch.full_like(out[..., 0], True, dtype=torch.bool, device=out.device)
        for _ in range(t, seq_len + 1):
            x = out[:, -self.max_seq_len:]
            mask = mask[:, -self.max_seq_len:]
            outputs = self(x, mask=mask, caches=caches, return_embeddings=True, return_caches=True, **kwargs)
            logits = self.model.lm_head(outputs.hidden_state[:, -1])
            caches = outputs.caches
            samples = {}
            for key, logits_i in logits.items():
                do_sample = True
                if fix_errors and exists(tokenizer):
                    if key == "Bar":
                        last_bar = out[:, -1, tokenizer.vocab_types_idx["Bar"]]
                        logits_i[:, 4:last_bar] = -float("Inf")
                    same_bar = samples.get("Bar", -1) == out[:, -1, tokenizer.vocab_types_idx["Bar"]]
                    if (key == "Tempo" and same_bar) or key == "TimeSig":
                        sample = out[:, -1, tokenizer.vocab_types_idx[key]][None]
                        do_sample = False
                if do_sample:
                    logits_i[:, :2] = -float("Inf")
                    sample = filter_logits_and_sample(logits_i, filter_logits_fn, filter_kwargs=filter_kwargs, temperature=temperature)
                samples[key] = sample
            samples = torch.cat(list(samples.values()), dim=-1)[None]
            out = torch.cat((out, samples), dim=1)
            mask = F.pad(mask, (0, 1), value=True)
            if exists(self.eos_token_id):
                if (out[..., -1, 0] == self.eos_token_id).any(dim=-1):
                    out[:, -1, 1:] = self.pad_token_id
                    break
            elif exists(max_bar):
                if (out[..., -1, 0] > max_bar).any(dim=-1):
                    out = out[:, :-1, :]
                    break
        out = out[:, t:]
        if num_dims == 2:
            out = out.squeeze(0)
        if was_training:
            self.model.train(was_training)
        if return_caches:
            return out, caches
        return out

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

    @torch.inference_mode()
    def unmask_tokens(self, tokens: Tensor, tokens_masked, temperature: float = 1., filter_logits_fn: Callable = top_k,
                      filter_kwargs=None, filter_key_ids=None, caches: Optional[TupleTransformerCaches] = None,
                      return_caches: bool = False, disable_tqdm: bool = False, **kwargs):
        """Note-by-note unmasking with hidden/KV caches (wrappers.py:324-407); batch-1 like the reference.
        The batched, device-resident renderer is scoreperformer_b200.decode.render_batch."""
        assert callable(filter_logits_fn)
        was_training = self.model.training
        if was_training:
            self.model.eval()
        num_dims = len(tokens.shape)
        if num_dims == 2:
            tokens = tokens[None, :]
            tokens_masked = tokens_masked[None, :]
        out = tokens.clone().detach()
        mask = kwargs.pop("mask", None)
        if mask is None:
            mask = torch.full_like(out[..., 0], True, dtype=torch.bool, device=out.device)
        filter_key_ids = filter_key_ids or dict()
        unmask_mask = out == self.mask_token_id
        unmask_ids = torch.where(torch.any(unmask_mask, dim=2))[1]
        for idx in unmask_ids.tolist():
            type_mask = unmask_mask[:, idx][0]
            logits_keys = torch.where(type_mask)[0].tolist()
            outputs = self(out[:, :idx + 1], seq_masked=tokens_masked[:, :idx + 1], mask=mask[:, :idx + 1], return_embeddings=True,
                           return_caches=True, caches=caches, **kwargs)
            caches = outputs.caches
            logits = self.model.lm_head(outputs.hidden_state[:, idx - 1], keys=logits_keys)
            out[:, idx, type_mask] = _sample_fields(logits, (self.pad_token_id, self.mask_token_id), filter_key_ids, filter_logits_fn,
                                                    filter_kwargs, temperature)
        if num_dims == 2:
            out = out.squeeze(0)
        if was_training:
            self.model.train(was_training)
        if return_caches:
            return out, caches
        return out

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
