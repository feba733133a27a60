"""One optimisation step of the ScorePerformer model as a reusable object: forward, backward, (all-reduce), clip, AdamW.

Mirrors what `Trainer.run_epoch` + `Optimizer.step` do per batch (experiments/trainer.py:449-452, optimizers.py:151-169:
backward, `clip_grad_norm_(2.0)`, AdamW(lr 2e-4, wd 1e-6), `zero_grad`) with two B200-first changes:

* parameters, gradients, both AdamW moments and a bf16 weight shadow live in flat buffers (every `p.data` / `p.grad` is a
  view), so data parallelism is a single NCCL all-reduce of 46 MB over NVLink per step and the optimiser is ONE kernel:
  clip-by-global-norm + AdamW in one pass over memory (csrc/optim.cu); weight-gradient kernels accumulate
  straight into the gradient buffer (fused.DIRECT_GRAD) and ONE cast kernel per step refreshes the bf16 shadow every
  forward GEMM reads its weight from (fused.SHADOW_ACTIVE);
* forward + backward (the whole kernel sequence, ~1 000 launches) is captured once in a CUDA graph and replayed, because
  at ~15 ms per step the Python dispatch of the eager path (~19 ms) would otherwise be the bottleneck.  Dropout stays
  random across replays through a device-side counter mixed into the kernels' seeds (kernels.RNG_OFFSET).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.distributed as dist
from torch import Tensor

from . import fused, kernels as K


class TrainStep:
    def __init__(self, model: torch.nn.Module, lr: float = 2e-4, weight_decay: float = 1e-6, grad_clip: Optional[float] = 2.0,
                 use_graph: bool = True, process_group=None, betas=(0.9, 0.999), eps: float = 1e-8):
        self.model = model
        self.grad_clip = grad_clip
        self.lr, self.weight_decay, self.betas, self.eps = lr, weight_decay, tuple(betas), eps
        self.use_graph = use_graph
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.params = [p for p in model.parameters() if p.requires_grad]
        dev = self.params[0].device
        if not dev.type == "cuda":
            raise RuntimeError("TrainStep needs a CUDA model: scoreperformer_b200 has no CPU fallback")
        # parameters, gradients and the bf16 weight shadow are three flat buffers with identical layouts (offsets rounded to
        # 8 elements so every bf16 view is 16-byte aligned for TMA)
        offsets, n = [], 0
        for p in self.params:
            offsets.append(n)
            n += (p.numel() + 7) // 8 * 8
        self.flat_param = torch.zeros(n, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.flat_shadow = torch.zeros(n, dtype=torch.bfloat16, device=dev)
        for p, off in zip(self.params, offsets):
            view = self.flat_param[off:off + p.numel()].view_as(p)
            view.copy_(p.data)
            p.data = view
            p.grad = self.flat_grad[off:off + p.numel()].view_as(p)
            p._spb_shadow = self.flat_shadow[off:off + p.numel()].view_as(p)
        self.offsets = offsets
        self.flat_m = torch.zeros(n, dtype=torch.float32, device=dev)        # AdamW exp_avg
        self.flat_v = torch.zeros(n, dtype=torch.float32, device=dev)        # AdamW exp_avg_sq
        self.opt_step = torch.zeros(1, dtype=torch.int64, device=dev)        # device-side step number (CUDA-graph safe)
        if use_graph and getattr(model, "perf_encoder", None) is not None:
            model.perf_encoder.exact_latent_shapes = False      # static segment tables: the step must not sync with the host
        self.rng_offset = torch.zeros(1, dtype=torch.int64, device=dev)
        K.RNG_OFFSET = self.rng_offset
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.static_batch: Optional[Dict[str, Tensor]] = None
        self.static_loss: Optional[Tensor] = None
        self.static_losses: Optional[Dict[str, Tensor]] = None
        self.launches_per_step = 0
        self._warm = 0
        self._copy_stream = None
        self._stage: Optional[Dict[str, Tensor]] = None

    # ------------------------------------------------------------------ pieces
    def _forward_backward(self, batch: Dict[str, Tensor]):
        self.rng_offset.add_(1)
        self.flat_grad.zero_()
        K.cast_bf16(self.flat_param, out=self.flat_shadow)       # ONE cast kernel refreshes every bf16 weight of the step
        fused.DIRECT_GRAD = fused.SHADOW_ACTIVE = True
        try:
            out = self.model(**batch)
            out.loss.backward()
        finally:
            fused.DIRECT_GRAD = fused.SHADOW_ACTIVE = False
        return out

    def _update(self):
        if self.world > 1:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=self.group)
        norm = torch.linalg.vector_norm(self.flat_grad) if self.grad_clip is not None else None
        self.opt_step.add_(1)
        # averaging over ranks, clipping and AdamW: one pass over the flat buffers
        K.adamw_step(self.flat_param, self.flat_grad, self.flat_m, self.flat_v, None, norm, self.opt_step, lr=self.lr,
                     betas=self.betas, eps=self.eps, weight_decay=self.weight_decay, max_norm=self.grad_clip or 0.0,
                     grad_scale=1.0 / self.world)

    # ------------------------------------------------------------------ input pipeline (host batches)
    def prefetch(self, batch: Dict[str, Tensor]) -> None:
        """Start copying the NEXT step's (pinned) host batch to the device on a side stream, so the transfer overlaps the step
        that is currently running.  Consume it with `step_prefetched()`."""
        dev = self.flat_grad.device
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._copied = torch.cuda.Event()
            self._stage_free = torch.cuda.Event()
            self._stage_free.record()
        if self._stage is None or any(self._stage[k].shape != v.shape for k, v in batch.items()):
            self._stage = {k: torch.empty(v.shape, dtype=v.dtype, device=dev) for k, v in batch.items()}
        self._copy_stream.wait_event(self._stage_free)           # the previous step has taken its copy out of the stage
        with torch.cuda.stream(self._copy_stream):
            for k, v in batch.items():
                self._stage[k].copy_(v, non_blocking=True)
            self._copied.record()

    def step_prefetched(self) -> Tensor:
        """One training step on the batch handed to the last `prefetch()` call."""
        assert self._stage is not None, "call prefetch(batch) first"
        cur = torch.cuda.current_stream()
        cur.wait_event(self._copied)
        if not self.use_graph or self._warm < 3 or self.graph is None:
            batch = {k: v.clone() for k, v in self._stage.items()}
            self._stage_free.record()
            return self.step(batch)
        for k, v in self._stage.items():
            self.static_batch[k].copy_(v, non_blocking=True)      # device-to-device, ~10 us
        self._stage_free.record()
        self.graph.replay()
        if self.world > 1:
            self._update()
        self.losses = self.static_losses
        return self.static_loss

    def set_lr(self, lr: float):
        """ExponentialLR etc. (experiments/optimizers.py:121-149): the rate is baked into a captured graph, so re-capture."""
        if lr != self.lr:
            self.lr = lr
            self.graph = None

    def optimizer_state_dict(self) -> dict:
        """`torch.optim.AdamW.state_dict()`-shaped view of the flat moments, so reference checkpoints interoperate."""
        state = {}
        for i, (p, off) in enumerate(zip(self.params, self.offsets)):
            state[i] = {"step": self.opt_step.clone().float().squeeze(0),
                        "exp_avg": self.flat_m[off:off + p.numel()].view_as(p).clone(),
                        "exp_avg_sq": self.flat_v[off:off + p.numel()].view_as(p).clone()}
        group = {"lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": self.weight_decay, "amsgrad": False,
                 "params": list(range(len(self.params)))}
        return {"state": state, "param_groups": [group]}

    def load_optimizer_state_dict(self, sd: dict):
        for i, (p, off) in enumerate(zip(self.params, self.offsets)):
            st = sd["state"].get(i)
            if st is None:
                continue
            self.flat_m[off:off + p.numel()].view_as(p).copy_(st["exp_avg"])
            self.flat_v[off:off + p.numel()].view_as(p).copy_(st["exp_avg_sq"])
            self.opt_step.fill_(int(st["step"]))
        g = sd["param_groups"][0]
        self.lr, self.betas, self.eps, self.weight_decay = g["lr"], tuple(g["betas"]), g["eps"], g["weight_decay"]
        self.graph = None

    def _capture(self, batch: Dict[str, Tensor]):
        self.static_batch = {k: v.clone() for k, v in batch.items()}
        self.graph = torch.cuda.CUDAGraph()
        before = K.LAUNCHES
        with torch.cuda.graph(self.graph):
            out = self._forward_backward(self.static_batch)
            if self.world == 1:
                self._update()
        self.launches_per_step = K.LAUNCHES - before
        self.static_loss = out.loss.detach()
        self.static_losses = {k: v.detach() for k, v in out.losses.items()}

    # ------------------------------------------------------------------ public
    def step(self, batch: Dict[str, Tensor]) -> Tensor:
        """Run one training step on `batch` (device tensors, or pinned host tensors: they are copied in asynchronously).
        Returns the loss tensor (device, detached)."""
        dev = self.flat_grad.device
        if not self.use_graph or self._warm < 3:
            # eager (also used to warm up lazily-initialised state before capture)
            b = {k: v.to(dev, non_blocking=True) for k, v in batch.items()}
            before = K.LAUNCHES
            out = self._forward_backward(b)
            self._update()
            self.launches_per_step = K.LAUNCHES - before
            self._warm += 1
            self.losses = {k: v.detach() for k, v in out.losses.items()}
            return out.loss.detach()
        if self.graph is None:
            torch.cuda.synchronize()
            self._capture({k: v.to(dev) for k, v in batch.items()})
        for k, v in batch.items():
            self.static_batch[k].copy_(v, non_blocking=True)
        self.graph.replay()
        if self.world > 1:
            self._update()
        self.losses = self.static_losses
        return self.static_loss
