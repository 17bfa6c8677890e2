"""Stub of diffusers.models.attention: FeedForward/GEGLU (exact-erf GELU), AdaLayerNorm placeholder."""
import torch.nn.functional as F
from torch import nn

from .attention_processor import Attention  # noqa: F401  (re-export, as diffusers does)


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        h, g = self.proj(x).chunk(2, dim=-1)
        return h * F.gelu(g)


class FeedForward(nn.Module):
    def __init__(self, dim, dim_out=None, mult=4, dropout=0.0, activation_fn="geglu", final_dropout=False):
        super().__init__()
        assert activation_fn == "geglu"
        inner = dim * mult
        self.net = nn.ModuleList([GEGLU(dim, inner), nn.Dropout(dropout), nn.Linear(inner, dim_out or dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class AdaLayerNorm(nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("AdaLayerNorm is never instantiated on this path")
