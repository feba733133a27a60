"""TEST INFRASTRUCTURE ONLY -- deterministic, init-independent weights for parity runs.

Goldens must not depend on the reference's RNG consumption order at construction time, so every
parity run (reference here, oracle + CUDA path on the GPU box) overwrites the model's parameters
with values derived from the *state_dict key name* alone.  Tied tensors (the per-field embedding
modules appear under score_encoder/perf_encoder/perf_decoder/lm_head) are
filled once, under the first key that references the storage.
"""
from __future__ import annotations

import zlib
from typing import Dict

import torch


def _gen(key: str, seed: int) -> torch.Generator:
    return torch.Generator().manual_seed((zlib.crc32(key.encode()) + 7919 * seed) % (2 ** 31))


@torch.no_grad()
def fill_model_(model: torch.nn.Module, seed: int = 0) -> None:
    """Overwrite every floating-point parameter of `model` in place (CPU or CUDA)."""
    seen = {}
    for key, t in model.state_dict(keep_vars=True).items():
        if not t.is_floating_point() or key.endswith("token_values") or key.endswith("class_weights"):
            continue
        ptr = t.data_ptr()
        if ptr in seen:
            continue
        seen[ptr] = key
        g = _gen(key, seed)
        shape = tuple(t.shape)
        r = torch.randn(shape, generator=g, dtype=torch.float32)
        leaf = key.rsplit(".", 1)[-1]
        if leaf == "learned_logslopes":
            val = t.detach().cpu().float() + 0.1 * r
        elif leaf == "index_weight":
            val = 0.5 * r
            val[0] = 0.0                                   # PAD row stays zero (embeddings.py:83-89)
        elif leaf == "bias":
            val = 0.05 * r
            if key.endswith("linear.bias") and ".vae_head." not in key and t.shape[0] % 2 == 0:
                val[: t.shape[0] // 2] += 1.0             # AdaLN: gamma bias 1, beta bias 0 (layers.py:37-39)
        elif leaf == "weight" and t.ndim == 1:
            val = 1.0 + 0.1 * r                            # LayerNorm gains
        elif t.ndim == 2:
            val = r / (t.shape[1] ** 0.5)
        else:
            val = 0.1 * r
        t.data.copy_(val.to(device=t.device, dtype=t.dtype))


def state_dict_checksum(sd: Dict[str, torch.Tensor]) -> float:
    return float(sum(float(v.double().abs().sum()) for v in sd.values() if v.is_floating_point()))
