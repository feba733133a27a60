"""Residual and AdaptiveLayerNorm (reference: scoreperformer/modules/layers.py:13-47)."""
from typing import Optional

import torch
from torch import nn, Tensor

from .. import fused, kernels as K


class Residual(nn.Module):
    def __init__(self, dim: int, scale_residual: bool = False, scale_residual_constant: float = 1.):
        super().__init__()
        self.residual_scale = nn.Parameter(torch.ones(dim)) if scale_residual else None
        self.scale_residual_constant = scale_residual_constant

    def forward(self, x, residual):
        if self.residual_scale is not None:
            residual = residual * self.residual_scale
        if self.scale_residual_constant != 1:
            residual = residual * self.scale_residual_constant
        return x + residual


class LayerNorm(nn.LayerNorm):
    """nn.LayerNorm whose CUDA forward is the spb_layer_norm kernel (same parameters / state_dict keys)."""

    def forward(self, x: Tensor, out_fp32: bool = True) -> Tensor:
        if not x.is_cuda:
            raise RuntimeError("scoreperformer_b200 runs on CUDA only (no CPU fallback)")
        return fused.layer_norm(x, self.weight, self.bias, out_fp32=out_fp32, eps=self.eps)


class AdaptiveLayerNorm(nn.Module):
    """gamma(c) * LN(x) + beta(c), (gamma, beta) = Linear(condition); modules/layers.py:31-47."""

    def __init__(self, dim: int, condition_dim: int, eps: float = 1e-5):
        super().__init__()
        self.dim = dim
        self.eps = eps
        self.linear = nn.Linear(condition_dim, dim * 2)
        self.linear.bias.data[:dim] = 1
        self.linear.bias.data[dim:] = 0

    def forward(self, x: Tensor, condition: Optional[Tensor] = None):
        """Inference-only stand-alone form; training goes through the fused stack (Transformer.forward)."""
        if condition is None:
            raise NotImplementedError("AdaptiveLayerNorm without a condition is not used by any recipe")
        condition = condition.unsqueeze(1) if condition.ndim == 2 else condition
        b, t, d = x.shape
        with torch.no_grad():
            gb = fused.linear(condition.expand(b, t, -1).reshape(b * t, -1), self.linear.weight, self.linear.bias)
            y, _, _ = K.layer_norm_fwd(x.reshape(b * t, d).float().contiguous(), None, None, gb, out_dtype=torch.bfloat16,
                                       eps=self.eps, need_stats=False)
        return y.float().view(b, t, d)
