"""Packed batches and on-device MixedLM masking.

The reference collators hand the trainer int64 tensors: performance tokens, their MASKed copy, the labels, score tokens, three
segment-id tensors, direction labels and two bool masks -- 466 bytes per note-tuple, 15.3 MB per 64 x 512 step, most of it
redundant: the masked copy and the labels are functions of the performance tokens (data/collators/performance.py:239-255,
`MixedLMPerformanceCollator.mask_sequence`), the masks are functions of the sequence lengths, and no token id needs more than 9
bits.  `pack_batch` keeps the information (65 bytes per tuple, 2.1 MB per step: uint16 tokens, int32 segment ids, uint8
directions, int32 lengths) and `unpack_batch` rebuilds the reference's tensors on the GPU with one kernel (csrc/collate.cu),
bit-exactly (tests/test_kernels_gpu.py::test_unpack_batch_matches_reference_collator against vectors of the reference's own
`mask_sequence`).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional, Sequence, Tuple

import torch
from torch import Tensor


@dataclass(frozen=True)
class PackedBatchSpec:
    """Collator settings of recipes/scoreperformer/base.yaml:62-65 (MixedLMScorePerformanceCollator)."""
    mask_token_id: int = 1
    mask_ignore_token_ids: Tuple[int, ...] = (0, 1, 2, 3)                  # PAD, MASK, SOS, EOS
    mask_ignore_token_dims: Tuple[int, ...] = (0, 1, 2, 4, 6, 7, 8, 9)      # everything except the four performance fields
    label_pad_ignored_dims: bool = True
    label_pad_token_id: int = -100
    pad_token_id: int = 0

    @property
    def ignore_ids(self) -> Tuple[int, ...]:
        return tuple(sorted({*self.mask_ignore_token_ids, self.pad_token_id}))      # performance.py:231

    def bits(self) -> Tuple[int, int]:
        ids = self.ignore_ids
        if any(not 0 <= i < 32 for i in ids) or any(not 0 <= d < 32 for d in self.mask_ignore_token_dims):
            raise ValueError("mask_ignore_token_ids / _dims must lie in [0, 32) for the device-side masking")
        return sum(1 << d for d in set(self.mask_ignore_token_dims)), sum(1 << i for i in ids)


def mixlm_mask_sequence(seq: Tensor, spec: PackedBatchSpec = PackedBatchSpec()) -> Tuple[Tensor, Tensor]:
    """(masked tokens, labels) of `MixedLMPerformanceCollator.mask_sequence` for int64 `seq` [..., F]: the host-side statement of
    what the unpack kernel computes (used by `pack_batch`'s self-check and by the CPU tests)."""
    ids = torch.tensor(spec.ignore_ids, dtype=seq.dtype, device=seq.device)
    no_mask = (seq[..., None] == ids).any(-1)
    dim_ignored = torch.zeros(seq.shape[-1], dtype=torch.bool, device=seq.device)
    dim_ignored[[d for d in spec.mask_ignore_token_dims if d < seq.shape[-1]]] = True
    masked = torch.where(~no_mask & ~dim_ignored, torch.full_like(seq, spec.mask_token_id), seq)
    label_mask = ~no_mask & ~dim_ignored if spec.label_pad_ignored_dims else ~no_mask
    labels = torch.where(label_mask, seq, torch.full_like(seq, spec.label_pad_token_id))
    return masked, labels


def _lengths(mask: Tensor) -> Tensor:
    """Sequence lengths of a right-padded bool mask [B, T]; raises if the mask is not a prefix mask."""
    lengths = mask.sum(dim=1).to(torch.int32)
    if not torch.equal(mask, torch.arange(mask.shape[1], device=mask.device)[None] < lengths[:, None]):
        raise ValueError("packed batches need right-padded sequences (mask == arange(T) < length)")
    return lengths


def pack_batch(batch: Dict[str, Tensor], spec: PackedBatchSpec = PackedBatchSpec(), pin: bool = True, check: bool = False) -> Dict[str, Tensor]:
    """Host side: the dict `ScorePerformer.prepare_inputs` would pass to the model (int64 tensors; `masked_perf` / `labels`
    optional and only verified, never shipped) -> compact (pinned) host tensors.  `check=True` verifies that the batch really is
    what the packed form can express (masking = the collator's function of `perf`, masks = prefix masks)."""
    perf = batch["perf"]
    B, T, Fp = perf.shape
    if int(perf.max()) >= 65536 or int(perf.min()) < 0:
        raise ValueError("token ids must fit uint16")
    out = {"perf": perf.to(torch.int32).to(torch.uint16), "perf_len": _lengths(batch["perf_mask"])}
    if batch.get("score") is not None:
        out["score"] = batch["score"].to(torch.int32).to(torch.uint16)
        out["score_len"] = _lengths(batch["score_mask"]) if batch.get("score_mask") is not None else out["perf_len"].clone()
    if batch.get("bars") is not None:
        out["segs"] = torch.stack([batch["bars"], batch["beats"], batch["onsets"]]).to(torch.int32)
    if batch.get("directions") is not None:
        if int(batch["directions"].max()) > 255 or int(batch["directions"].min()) < 0:
            raise ValueError("direction labels must fit uint8")
        out["dirs"] = batch["directions"].to(torch.uint8)
    if batch.get("deadpan_mask") is not None:
        out["deadpan"] = batch["deadpan_mask"].to(torch.uint8)
    if check and batch.get("labels") is not None:
        masked, labels = mixlm_mask_sequence(perf, spec)
        if not (torch.equal(masked, batch["masked_perf"]) and torch.equal(labels, batch["labels"])):
            raise ValueError("masked_perf / labels are not MixedLMPerformanceCollator.mask_sequence(perf) under this spec")
    out = {k: v.contiguous() for k, v in out.items()}
    return {k: v.pin_memory() for k, v in out.items()} if pin and torch.cuda.is_available() else out


def packed_bytes(packed: Dict[str, Tensor]) -> int:
    return sum(v.numel() * v.element_size() for v in packed.values())


def unpack_batch(packed: Dict[str, Tensor], spec: PackedBatchSpec = PackedBatchSpec(), out: Optional[Dict[str, Tensor]] = None,
                 mixlm: bool = True) -> Dict[str, Tensor]:
    """Device side: compact device tensors -> the model's input dict (int64 tokens, bool masks, MixedLM masked tokens + labels),
    one kernel.  `out` (a dict of preallocated tensors of the right shapes, e.g. the static batch of a captured graph) is written
    in place when given."""
    from .. import kernels as K
    perf = packed["perf"]
    if not perf.is_cuda:
        raise RuntimeError("unpack_batch runs on the GPU: copy the packed tensors to the device first (TrainStep.prefetch_packed)")
    B, T, Fp = perf.shape
    dev = perf.device
    have = {k: packed.get(k) is not None for k in ("score", "segs", "dirs")}

    def buf(name, shape, dtype):
        if out is not None and name in out:
            t = out[name]
            assert t.shape == torch.Size(shape) and t.dtype == dtype and t.is_contiguous(), name
            return t
        return torch.empty(shape, dtype=dtype, device=dev)

    res = {"perf": buf("perf", (B, T, Fp), torch.int64), "perf_mask": buf("perf_mask", (B, T), torch.bool)}
    if mixlm:
        res["masked_perf"] = buf("masked_perf", (B, T, Fp), torch.int64)
        res["labels"] = buf("labels", (B, T, Fp), torch.int64)
    Fs = Fd = 0
    segs_out = None
    if have["score"]:
        Fs = packed["score"].shape[-1]
        res["score"] = buf("score", (B, T, Fs), torch.int64)
        res["score_mask"] = buf("score_mask", (B, T), torch.bool)
    if have["segs"]:
        if out is not None and all(k in out for k in ("bars", "beats", "onsets")) and _adjacent(out["bars"], out["beats"], out["onsets"]):
            segs_out = None            # written straight into the three adjacent tensors below
            res["bars"], res["beats"], res["onsets"] = out["bars"], out["beats"], out["onsets"]
            segs_ptr = out["bars"]
        else:
            segs_out = torch.empty((3, B, T), dtype=torch.int64, device=dev)
            res["bars"], res["beats"], res["onsets"] = segs_out[0], segs_out[1], segs_out[2]
            segs_ptr = segs_out
    if have["dirs"]:
        Fd = packed["dirs"].shape[-1]
        res["directions"] = buf("directions", (B, T, Fd), torch.int64)
    ignore_dims, ignore_ids = spec.bits()
    K.unpack_batch(perf, packed.get("score"), packed.get("segs"), packed.get("dirs"), packed["perf_len"], packed.get("score_len"),
                   res["perf"], res.get("masked_perf"), res.get("labels"), res.get("score"), segs_ptr if have["segs"] else None,
                   res.get("directions"), res["perf_mask"], res.get("score_mask"), B, T, Fp, Fs, Fd, ignore_dims, ignore_ids,
                   spec.mask_token_id, spec.label_pad_token_id, spec.label_pad_ignored_dims)
    if out is not None and have["segs"] and segs_out is not None:
        for k in ("bars", "beats", "onsets"):
            if k in out:
                out[k].copy_(res[k])
                res[k] = out[k]
    if packed.get("deadpan") is not None:
        if out is not None and "deadpan_mask" in out:
            out["deadpan_mask"].copy_(packed["deadpan"])
            res["deadpan_mask"] = out["deadpan_mask"]
        else:
            res["deadpan_mask"] = packed["deadpan"].to(torch.bool)
    return res


def _adjacent(a: Tensor, b: Tensor, c: Tensor) -> bool:
    n = a.numel() * a.element_size()
    return a.is_contiguous() and b.is_contiguous() and c.is_contiguous() and b.data_ptr() == a.data_ptr() + n and c.data_ptr() == b.data_ptr() + n
