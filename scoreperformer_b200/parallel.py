"""Data-parallel training over the GPUs of one box: one process per GPU, bucketed NCCL all-reduce of gradients over
NVLink overlapped with the rest of backward.

The reference has no distributed code at all (experiments/trainer.py:122-130); this is the insertion point a
DDP-enabled `Optimizer.step` (experiments/optimizers.py:151-169) needs: call `sync_gradients()` between `backward()`
and `clip_grad_norm_`.

Buckets follow reverse execution order (decoder + heads -> perf encoder -> score encoder -> tied embedding tables last,
because the tied tables receive gradient from all three stacks).  Each bucket is one flat fp32 buffer; gradients are
copied in by a post-accumulate hook, the all-reduce of a bucket starts on a side stream as soon as its last gradient
arrived, and `sync_gradients()` waits and scatters the averaged values back.

The model runs backward on several streams (encoder branches, weight-gradient side streams): every hook records an event
on the stream its gradient was produced on, and the communication stream waits for ALL events of a bucket before the
bucket is copied and reduced.  With gradient accumulation (several backward passes per optimiser step) call
`sync_gradients()` only after the last pass: hooks that fire again for an already complete bucket withdraw its
in-flight reduction, which is redone on the accumulated values.

`train_step.TrainStep` does not use this class: it owns flat gradient buffers and reduces their slices in place, inside
the captured step.  This class serves a reference-style `Trainer` + `Optimizer` that keeps per-parameter `.grad`s.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.distributed as dist


def default_buckets(model: torch.nn.Module) -> List[List[str]]:
    """Parameter names per bucket, reverse execution order; tied tensors are listed once (under their first name)."""
    seen, order = set(), {"dec": [], "perf": [], "score": [], "clf": [], "emb": []}
    for name, p in model.named_parameters():
        if id(p) in seen or not p.requires_grad:
            continue
        seen.add(id(p))
        if ".token_emb.embs." in name:
            order["emb"].append(name)
        elif name.startswith("perf_decoder"):
            order["dec"].append(name)
        elif name.startswith("perf_encoder"):
            order["perf"].append(name)
        elif name.startswith("score_encoder"):
            order["score"].append(name)
        else:
            order["clf"].append(name)
    return [b for b in (order["dec"] + order["clf"], order["perf"], order["score"], order["emb"]) if b]


class GradientBuckets:
    def __init__(self, model: torch.nn.Module, process_group=None, buckets: Optional[List[List[str]]] = None, overlap: bool = True):
        self.model = model
        self.group = process_group
        self.world_size = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.overlap = overlap
        params = dict(model.named_parameters())
        self.buckets: List[Dict] = []
        for names in (buckets or default_buckets(model)):
            ps = [params[n] for n in names]
            numel = sum(p.numel() for p in ps)
            flat = torch.zeros(numel, dtype=torch.float32, device=ps[0].device)
            views, off = [], 0
            for p in ps:
                views.append(flat[off:off + p.numel()].view_as(p))
                off += p.numel()
            self.buckets.append(dict(params=ps, flat=flat, views=views, pending=len(ps), work=None, events=[None] * len(ps), dirty=False))
        self._owner = {id(p): (bi, pi) for bi, b in enumerate(self.buckets) for pi, p in enumerate(b["params"])}
        self._stream = torch.cuda.Stream() if (overlap and ps[0].is_cuda) else None
        self._hooks = []
        if self.world_size > 1:
            for b in self.buckets:
                for p in b["params"]:
                    self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))

    # one gradient finished accumulating
    def _on_grad(self, p: torch.Tensor):
        bi, pi = self._owner[id(p)]
        b = self.buckets[bi]
        if p.is_cuda:
            ev = b["events"][pi]
            if ev is None:
                ev = b["events"][pi] = torch.cuda.Event()
            ev.record()                     # on the stream that produced this gradient (the hook runs on it)
        if b["pending"] <= 0:
            # another backward pass of a gradient-accumulation step: the reduction in flight saw partial sums, redo it later
            if b["work"] is not None:
                b["work"].wait()
                b["work"] = None
            b["pending"] = 0
            b["dirty"] = True
            return
        b["pending"] -= 1
        if b["pending"] == 0:
            self._launch(b)

    def _launch(self, b: Dict):
        grads = [q.grad for q in b["params"]]
        if self._stream is not None:
            for ev in b["events"]:          # every producer stream, not just the one of the last hook
                if ev is not None:
                    self._stream.wait_event(ev)
            self._stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._stream):
                torch._foreach_copy_(b["views"], grads)      # one multi-tensor copy per bucket
                b["work"] = dist.all_reduce(b["flat"], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        else:
            torch._foreach_copy_(b["views"], grads)
            b["work"] = dist.all_reduce(b["flat"], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        b["dirty"] = False

    def _launch_all_streams(self, b: Dict):
        """Late launch from sync_gradients(): backward has returned, so the engine has already joined its streams with the
        caller's; the recorded events are waited for as well."""
        self._launch(b)

    def sync_gradients(self):
        """Wait for every bucket, write the averaged gradients back into `p.grad`, re-arm for the next step."""
        if self.world_size == 1:
            return
        inv = 1.0 / self.world_size
        for b in self.buckets:
            if b["pending"] != 0 or b["dirty"] or b["work"] is None:
                # a parameter received no gradient this step, or gradients were accumulated over several passes: reduce what is there
                for pi, p in enumerate(b["params"]):
                    if p.grad is None:
                        p.grad = torch.zeros_like(b["views"][pi])
                self._launch_all_streams(b)
            b["work"].wait()
            if self._stream is not None:
                torch.cuda.current_stream().wait_stream(self._stream)
            b["flat"].mul_(inv)
            for p, v in zip(b["params"], b["views"]):
                if p.grad is None:
                    p.grad = torch.empty_like(v)
            torch._foreach_copy_([p.grad for p in b["params"]], b["views"])
            b["pending"] = len(b["params"])
            b["work"] = None

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []
