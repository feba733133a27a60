"""GLU feed-forward block (reference: scoreperformer/modules/transformer/feedforward.py:13-64)."""
from dataclasses import dataclass

import torch.nn as nn

from ... import fused, kernels as K
from ..constructor import Constructor, ModuleConfig


class GLU(nn.Module):
    def __init__(self, dim_in, dim_out, activation):
        super().__init__()
        self.act = activation
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        raise NotImplementedError("GLU is evaluated inside FeedForward.forward (fused)")


@dataclass
class FeedForwardConfig(ModuleConfig):
    dim: int = 512
    mult: int = 4
    glu: bool = False
    swish: bool = False
    post_act_ln: bool = False
    dropout: float = 0.
    no_bias: bool = True


class FeedForward(nn.Module, Constructor):
    def __init__(self, dim: int = 512, mult: int = 4, glu: bool = False, swish: bool = False, post_act_ln: bool = False,
                 dropout: float = 0., no_bias: bool = True):
        super().__init__()
        inner_dim = int(dim * mult)
        activation = nn.SiLU() if swish else nn.GELU()
        project_in = nn.Sequential(nn.Linear(dim, inner_dim, bias=not no_bias), activation) if not glu else GLU(dim, inner_dim, activation)
        self.ff = nn.Sequential(
            project_in,
            nn.LayerNorm(inner_dim) if post_act_ln else nn.Identity(),
            nn.Dropout(dropout),
            nn.Linear(inner_dim, dim, bias=not no_bias),
        )
        self.inner_dim = inner_dim
        self.dropout = dropout
        self.glu, self.swish, self.post_act_ln, self.no_bias = glu, swish, post_act_ln, no_bias

    @property
    def fused_supported(self) -> bool:
        """The sm_100a path covers the recipes' FFN: GLU + SiLU, no post-activation LayerNorm, bias-free output projection."""
        return self.glu and self.swish and not self.post_act_ln and self.no_bias

    def forward(self, x):
        if not self.fused_supported:
            raise NotImplementedError("scoreperformer_b200.FeedForward: only the GLU+SiLU configuration of the recipes is implemented")
        shape = x.shape
        u = fused.linear(x.reshape(-1, shape[-1]), self.ff[0].proj.weight, self.ff[0].proj.bias)
        p = self.dropout if self.training else 0.0
        h = fused.GLUFn.apply(u, p, K.seed_from_torch() if p > 0 else 0)
        return fused.linear(h, self.ff[3].weight, out_fp32=True).view(shape)
