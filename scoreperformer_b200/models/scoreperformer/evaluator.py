"""ScorePerformer metric evaluator, drop-in for scoreperformer/models/scoreperformer/evaluator.py:15-106 (same constructor, same
call, same metric names and values) -- but the numbers come out of the output head.

The reference recomputes everything from the [B, T, V] logits of all twelve fields after every training step
(experiments/trainer.py:462-464): twelve argmaxes, a softmax per evaluated field, value look-ups.  Here `attach()` hands the token
values to the LM wrapper of the model, whose fused head + cross-entropy kernel (csrc/head_ce.cu) then accumulates, per labelled
field and while the logits sit in tensor memory,

    hits = #(argmax == label)      rows = #labelled      dist = sum |tv[argmax] - tv[label]|      wdist = sum_v p_v |tv[label] - tv[v]|

and the wrapper returns them as `outputs.eval_stats`.  Every metric of the reference is a ratio of those sums, so calling the
evaluator costs a handful of scalar divisions and the training logits are never materialised.  Outputs that do not carry the
statistics (inference outputs, a model the evaluator was not attached to) are reduced from their logits by `_stats_from_logits`,
which states the same sums in plain PyTorch -- that path is also what the tests compare the kernel against.
"""
from typing import Dict, List, Optional

import torch
from torch import Tensor

from .model import ScorePerformerOutputs
from .wrappers import ScorePerformerLMModes

_HITS, _ROWS, _DIST, _WDIST = range(4)


class ScorePerformerEvaluator:
    def __init__(self, model, tokenizer=None, label_pad_token_id: int = -100, weighted_distance: bool = False,
                 ignore_keys: Optional[List[str]] = None):
        self.model = model
        self.tokenizer = tokenizer
        self.label_pad_token_id = label_pad_token_id
        self.weighted_distance = weighted_distance
        self.ignore_keys = ignore_keys
        self.token_values = None
        if tokenizer is not None:       # evaluator.py:30-35: one column of values per field
            self.token_values = {key: torch.as_tensor(values)[:, None] for key, values in tokenizer.token_values(normalize=False).items()}
        self.attach(model)

    # ------------------------------------------------------------------ fused path
    def attach(self, model) -> None:
        """Let the model's LM wrapper accumulate the statistics in its head kernel (no-op for models without that wrapper)."""
        wrapper = getattr(model, getattr(model, "_lm_attr", "perf_decoder"), None)
        if wrapper is None or not hasattr(wrapper, "eval_token_values"):
            return
        if self.token_values is None:
            wrapper.eval_token_values = {}              # hit counts only
        else:
            wrapper.eval_token_values = {k: v.reshape(-1).float() for k, v in self.token_values.items()}

    # ------------------------------------------------------------------ statement of the statistics on materialised logits
    def _stats_from_logits(self, logits: Dict[str, Tensor], labels: Tensor) -> Dict[str, Tensor]:
        stats = {}
        for i, (key, lg) in enumerate(logits.items()):
            lab = labels[..., i].reshape(-1)
            use = lab != self.label_pad_token_id
            lg = lg.reshape(-1, lg.shape[-1])
            pred = lg.argmax(dim=-1)
            safe = lab.clamp(min=0)
            row = [((pred == safe) & use).sum().float(), use.sum().float()]
            tv = None if self.token_values is None or key not in self.token_values else self.token_values[key].reshape(-1).to(lg.device).float()
            if tv is None:
                row += [row[0].new_zeros(()), row[0].new_zeros(())]
            else:
                row.append(((tv[pred] - tv[safe]).abs() * use).sum())
                if self.weighted_distance:
                    spread = (tv[safe][:, None] - tv[None, :]).abs()
                    row.append(((lg.float().softmax(dim=-1) * spread).sum(dim=-1) * use).sum())
                else:
                    row.append(row[0].new_zeros(()))
            stats[key] = torch.stack(row)
        return stats

    # ------------------------------------------------------------------ the call of trainer.py:462-464
    @torch.no_grad()
    def __call__(self, inputs, outputs, ignore_keys: Optional[List[str]] = None):
        ignore = set(ignore_keys or self.ignore_keys or ())
        if isinstance(outputs, ScorePerformerOutputs):
            outputs = outputs.perf_decoder
        stats = getattr(outputs, "eval_stats", None)
        from_head = stats is not None       # the wrapper only reports fields that carry labels: no per-field host check needed
        if stats is None:
            labels = inputs["labels"] if isinstance(inputs, dict) else inputs.labels.tokens.to(outputs.hidden_state.device)
            if self.model.mode in (ScorePerformerLMModes.CLM, ScorePerformerLMModes.MixedLM):
                labels = labels[:, 1:]
            stats = self._stats_from_logits(dict(outputs.logits.items()), labels)
        # fields without a single label contribute nothing anywhere (their masks are empty in the reference)
        metrics = {}
        total = torch.stack(list(stats.values())).sum(dim=0) if stats else None
        if total is not None:
            metrics["accuracy"] = total[_HITS] / total[_ROWS]
        if ignore:
            kept = [v for k, v in stats.items() if k not in ignore]
            if kept:
                sub = torch.stack(kept).sum(dim=0)
                metrics["accuracy/pred"] = sub[_HITS] / sub[_ROWS]
        # the per-field entries exist only for fields that have labels in this batch (a host decision in the reference too)
        have = {k: from_head or bool(v[_ROWS] > 0) for k, v in stats.items() if k not in ignore} if stats else {}
        for key, s in stats.items():
            if key in ignore or not have[key]:
                continue
            metrics[f"accuracy/{key}"] = s[_HITS] / s[_ROWS]
        if self.token_values is not None:
            for key, s in stats.items():
                if key in ignore or not have[key] or key not in self.token_values:
                    continue
                metrics[f"distance/{key}"] = s[_WDIST if self.weighted_distance else _DIST] / s[_ROWS]
        return metrics
