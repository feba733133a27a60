"""Attention intermediates (reference: scoreperformer/modules/transformer/attend.py:26-33).

The core algorithm of the reference `Attend` (bias materialisation + SDPA) lives in csrc/attention.cu."""
from dataclasses import dataclass
from typing import Optional

from torch import Tensor


@dataclass
class AttentionIntermediates:
    keys: Optional[Tensor] = None
    values: Optional[Tensor] = None
    qk_similarities: Optional[Tensor] = None

    def to_tuple(self):
        return self.keys, self.values, self.qk_similarities
