"""Synthetic SPMuple batches of the shapes named in BASELINE.json.

The layout is what the reference collators deliver (data/collators/performance.py:42-54,
score_performance.py:55-108): int64 tokens / segments / labels / directions, bool masks
(True = valid), PAD=0 MASK=1 SOS=2 EOS=3 in every field, label ignore index -100.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

PERF_SIZES: Dict[str, int] = {  # SPMupleWindow vocabulary, SURVEY.md Appendix A.1
    "Bar": 260, "Position": 132, "Pitch": 92, "Velocity": 132, "Duration": 133, "Tempo": 125,
    "TimeSig": 26, "PositionShift": 69, "NotesInOnset": 16, "PositionInOnset": 16,
    "RelOnsetDev": 165, "RelPerfDuration": 85,
}
SCORE_KEYS = list(PERF_SIZES)[:10]
PREDICTED_FIELDS = (3, 5, 10, 11)  # Velocity, Tempo, RelOnsetDev, RelPerfDuration (base.yaml:64-65)
DIRECTION_CLASSES: Dict[str, int] = {
    "dynamic/absolute": 10, "dynamic/hairpin": 3, "dynamic/accent": 3, "tempo/absolute": 14,
    "tempo/relative": 11, "articulation/arpeggiate": 2, "articulation/fermata": 2,
    "articulation/staccato": 2, "articulation/tenuto": 2,
}


def make_batch(batch_size: int, seq_len: int, seed: int = 1234, num_tokens: Optional[Dict[str, int]] = None,
               direction_classes: Optional[Dict[str, int]] = None, full_length: bool = False,
               deadpan_last: bool = True) -> Dict[str, torch.Tensor]:
    """Build one synthetic training batch on the CPU with a seeded generator.

    Returns the dict `ScorePerformer.prepare_inputs` would (models/scoreperformer/model.py:343-372).
    """
    num_tokens = num_tokens or PERF_SIZES
    direction_classes = direction_classes or DIRECTION_CLASSES
    g = torch.Generator().manual_seed(seed)
    B, T = batch_size, seq_len
    sizes = list(num_tokens.values())
    F = len(sizes)

    perf = torch.stack([torch.randint(4, v, (B, T), generator=g) for v in sizes], dim=-1)
    new_onset = (torch.rand(B, T, generator=g) < 0.6).long()
    new_onset[:, 0] = 0
    onset = 4 + torch.cumsum(new_onset, dim=1)
    beat = 4 + (onset - 4) // 3
    bar = 4 + (beat - 4) // 4
    perf[..., 0] = bar.clamp(max=sizes[0] - 1)

    lengths = torch.randint((3 * T) // 4, T + 1, (B,), generator=g)
    lengths[0] = T
    if full_length:
        lengths[:] = T
    mask = torch.arange(T)[None, :] < lengths[:, None]

    perf = perf * mask[..., None]
    bars, beats, onsets = bar * mask, beat * mask, onset * mask

    n_score = sum(1 for k in num_tokens if k in SCORE_KEYS)
    score = perf[..., :n_score].clone()

    pred = torch.zeros(F, dtype=torch.bool)
    pred[[i for i in PREDICTED_FIELDS if i < F]] = True
    maskable = (perf > 3) & pred[None, None, :]
    masked_perf = torch.where(maskable, torch.ones_like(perf), perf)
    labels = torch.where(maskable, perf, torch.full_like(perf, -100))

    directions = torch.stack(
        [torch.randint(0, c, (B, T), generator=g) for c in direction_classes.values()], dim=-1
    ) * mask[..., None]

    deadpan_mask = torch.zeros(B, dtype=torch.bool)
    if deadpan_last and B > 1:
        deadpan_mask[-1] = True

    return {
        "perf": perf, "perf_mask": mask, "score": score, "score_mask": mask.clone(),
        "masked_perf": masked_perf, "labels": labels,
        "bars": bars, "beats": beats, "onsets": onsets,
        "directions": directions, "deadpan_mask": deadpan_mask,
    }


class SyntheticTokenizer:
    """Stand-in for the OctupleM tokenizer where only `token_values` is needed (ScorePerformerEvaluator, evaluator.py:30-35):
    a deterministic value per token of every field -- special tokens 0, then an affine ramp whose step depends on the field."""

    def __init__(self, num_tokens: Optional[Dict[str, int]] = None):
        self.num_tokens = dict(num_tokens or PERF_SIZES)

    def token_values(self, normalize: bool = False):
        import numpy as np
        out = {}
        for i, (key, v) in enumerate(self.num_tokens.items()):
            vals = np.zeros(v, dtype=np.float32)
            vals[4:] = (np.arange(v - 4, dtype=np.float32) - 0.25 * v) * (0.5 + 0.125 * i)
            out[key] = vals / np.abs(vals).max() if normalize else vals
        return out


class SyntheticDataset(torch.utils.data.Dataset):
    """`n` synthetic note-tuple sequences of `seq_len` notes, one sample = one row of `make_batch` (the fields of
    ScorePerformer.forward).  Stands in for ScorePerformanceDataset where no corpus is available (benchmarks, trainer tests)."""

    def __init__(self, n: int, seq_len: int, seed: int = 1234, **kwargs):
        self.rows = make_batch(n, seq_len, seed=seed, **kwargs)
        self.n = n

    def __len__(self) -> int:
        return self.n

    def __getitem__(self, i: int) -> Dict[str, torch.Tensor]:
        return {k: v[i] for k, v in self.rows.items()}


def collate_rows(samples) -> Dict[str, torch.Tensor]:
    """Collator of SyntheticDataset: stacks the per-sample fields (all sequences have one length)."""
    return {k: torch.stack([s[k] for s in samples]) for k in samples[0]}
