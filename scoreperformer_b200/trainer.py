"""What `train.py` needs around the step to run multi-GPU: the pieces of
scoreperformer/experiments/trainer.py that touch the path, re-stated for one process per GPU.

* `build_dataloader` (trainer.py:166-174) with a `DistributedSampler`: every rank draws its own disjoint shard of an epoch's
  permutation (`set_epoch` reseeds it), the reference's `shuffle` / `num_workers` / `pin_memory` / `collate_fn` arguments unchanged;
* the epoch loop (trainer.py:380-470): one `TrainStep` per batch, the evaluator's metrics after every training step
  (trainer.py:462-464), `ExponentialLR` stepped per epoch (experiments/optimizers.py:151-169) through the device-side learning rate;
  the next batch's host-to-device copy is issued before the host waits for the current loss;
* checkpoints in the reference's layout (trainer.py:296-347: `{"experiment": {...}, "model": {"config", "state_dict"},
  "optimizer": ...}`), written by rank 0 only, the optimizer part `torch.optim.AdamW`-shaped (TrainStep.optimizer_state_dict), and
  `load_checkpoint` to resume on any world size.

Callbacks, TensorBoard, the experiment / OmegaConf configuration objects and best-model bookkeeping stay with the reference's Trainer:
they never touch the device."""
import json
import math
import os
from typing import Callable, Dict, Iterable, List, Optional

import torch
import torch.distributed as dist
from torch.utils.data import DataLoader, Dataset
from torch.utils.data.distributed import DistributedSampler

from .train_step import TrainStep


def _rank_world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def make_sampler(dataset: Dataset, world: int, rank: int, shuffle: bool, seed: int, is_train: bool = True) -> Optional[DistributedSampler]:
    """The rank's shard of every epoch's permutation; None for a single process (the DataLoader shuffles itself)."""
    if world <= 1:
        return None
    return DistributedSampler(dataset, num_replicas=world, rank=rank, shuffle=is_train and shuffle, seed=seed, drop_last=is_train)


class DataParallelTrainer:
    def __init__(self, model, train_dataset: Optional[Dataset], collator: Optional[Callable], batch_size: int, lr: float = 2e-4,
                 lr_gamma: float = 1.0, weight_decay: float = 1e-6, grad_clip: float = 2.0, output_dir: Optional[str] = None,
                 evaluator=None, shuffle: bool = True, num_workers: int = 0, pin_memory: bool = True, seed: int = 0,
                 use_graph: bool = True, prepare_inputs: Optional[Callable] = None, model_config: Optional[dict] = None):
        """`collator(samples) -> batch` as in the reference; `prepare_inputs(batch) -> dict of tensors` defaults to
        `model.prepare_inputs` when the collated batch is not a dict already (model.py:335-356)."""
        self.model, self.train_dataset, self.collator = model, train_dataset, collator
        self.batch_size, self.shuffle, self.num_workers, self.pin_memory, self.seed = batch_size, shuffle, num_workers, pin_memory, seed
        self.lr0, self.lr_gamma = lr, lr_gamma
        self.output_dir, self.model_config = output_dir, model_config
        self.prepare_inputs = prepare_inputs
        self.rank, self.world = _rank_world()
        self.step = TrainStep(model, lr=lr, weight_decay=weight_decay, grad_clip=grad_clip, use_graph=use_graph)
        if evaluator is not None:
            self.step.set_evaluator(evaluator)
        self.epoch, self.global_step = 0, 0
        self.sampler: Optional[DistributedSampler] = None
        self.log_history: List[Dict[str, float]] = []

    # ------------------------------------------------------------------ trainer.py:166-174
    def build_dataloader(self, dataset: Dataset, is_train: bool = True) -> DataLoader:
        sampler = make_sampler(dataset, self.world, self.rank, self.shuffle, self.seed, is_train)
        if sampler is not None and is_train:
            self.sampler = sampler
        gen = torch.Generator()
        gen.manual_seed(self.seed)
        return DataLoader(dataset, batch_size=self.batch_size, shuffle=(sampler is None and is_train and self.shuffle), sampler=sampler,
                          num_workers=self.num_workers, collate_fn=self.collator, pin_memory=self.pin_memory, drop_last=is_train,
                          generator=gen if sampler is None else None)

    def _inputs(self, batch) -> Dict[str, torch.Tensor]:
        if isinstance(batch, dict):
            return batch
        fn = self.prepare_inputs or self.model.prepare_inputs
        return fn(batch)

    # ------------------------------------------------------------------ trainer.py:380-470
    def fit(self, epochs: int, max_steps: Optional[int] = None, save_every: Optional[int] = None, log_every: int = 0) -> List[Dict[str, float]]:
        """Train for `epochs` epochs (or until `max_steps` global steps); returns the log history (one entry per step on every
        rank: loss and the evaluator's metrics are device tensors until `log_every` / the end asks for them)."""
        loader = self.build_dataloader(self.train_dataset, is_train=True)
        first_epoch = self.epoch
        pending: List[tuple] = []                 # (global_step, loss tensor, metrics dict of tensors): read back lazily
        done = False
        for epoch in range(first_epoch, epochs):
            self.epoch = epoch
            if self.sampler is not None:
                self.sampler.set_epoch(epoch)
            elif self.shuffle:
                loader.generator.manual_seed(self.seed + epoch)
            self.step.set_lr(self.lr0 * self.lr_gamma ** epoch)              # ExponentialLR, stepped per epoch
            it: Iterable = iter(loader)
            nxt = next(it, None)
            if nxt is not None:
                self.step.prefetch(self._inputs(nxt))
            while nxt is not None:
                loss = self.step.step_prefetched()
                nxt = next(it, None)
                if nxt is not None:
                    self.step.prefetch(self._inputs(nxt))                    # overlaps the step that was just enqueued
                self.global_step += 1
                pending.append((self.global_step, loss, dict(self.step.metrics or {})))
                if log_every and self.global_step % log_every == 0:
                    self._flush(pending)
                if save_every and self.global_step % save_every == 0:
                    self.save_checkpoint()
                if max_steps is not None and self.global_step >= max_steps:
                    done = True
                    break
            if done:
                break
            self.epoch = epoch + 1
        self._flush(pending)
        return self.log_history

    def _flush(self, pending: List[tuple]) -> None:
        for step, loss, metrics in pending:
            entry = {"step": step, "loss": float(loss)}
            entry.update({k: float(v) for k, v in metrics.items()})
            self.log_history.append(entry)
        pending.clear()

    # ------------------------------------------------------------------ trainer.py:296-347
    def checkpoint_path(self) -> str:
        return os.path.join(self.output_dir or ".", f"checkpoint_{self.global_step:d}.pt")

    def save_checkpoint(self, path: Optional[str] = None) -> Optional[str]:
        """Rank 0 writes, everyone waits.  Layout of the reference's `_save_checkpoint`."""
        path = path or self.checkpoint_path()
        if self.rank == 0:
            os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
            state = {"epoch": self.epoch, "global_step": self.global_step, "lr0": self.lr0, "lr_gamma": self.lr_gamma, "seed": self.seed}
            torch.save({"experiment": {"config": None, "trainer": None, "state": json.dumps(state)},
                        "model": {"config": self.model_config, "state_dict": self.model.state_dict()},
                        "optimizer": self.step.optimizer_state_dict()}, path)
        if self.world > 1:
            dist.barrier()
        return path if self.rank == 0 else None

    def load_checkpoint(self, path: str, strict: bool = True) -> None:
        ckpt = torch.load(path, map_location=self.step.flat_param.device, weights_only=False)
        self.model.load_state_dict(ckpt["model"]["state_dict"], strict=strict)
        self.step.mark_weights_changed()
        if ckpt.get("optimizer") is not None:
            self.step.load_optimizer_state_dict(ckpt["optimizer"])
        state = ckpt["experiment"].get("state")
        if state:
            state = json.loads(state)
            self.epoch, self.global_step = int(state.get("epoch", 0)), int(state.get("global_step", 0))
            self.lr0, self.lr_gamma = float(state.get("lr0", self.lr0)), float(state.get("lr_gamma", self.lr_gamma))

    def close(self) -> None:
        """Release the captured graphs (they hold NCCL kernels) -- before dist.destroy_process_group()."""
        self.step.close()


def steps_per_epoch(n_samples: int, batch_size: int, world: int = 1) -> int:
    """Batches a rank sees per epoch with `drop_last` (the DistributedSampler pads nothing away: floor on both levels)."""
    return math.floor(math.floor(n_samples / world) / batch_size)
