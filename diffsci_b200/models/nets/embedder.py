"""Conditioning embedders of the porous-media models (reference nets/embedder.py:199-246): they map the conditioning
dictionary ``y`` to a [B, model_channels] vector that PUNetG / ADM add to the time embedding (punetg.py:400-410,
adm.py:199-211).  They run ONCE per sampling run here (the conditioning is constant along the trajectory; the reference
re-evaluates them at every network evaluation), on tiny [B, M] tensors, as plain torch modules -- the user-supplied
``conditional_embedding`` seam takes ANY torch module, so these are off the per-evaluation hot path by construction;
their output enters the fused network path as one [B, M] vector (and its gradient comes back out of it in training).
State-dict keys match the reference (``gaussian_proj.W``, ``net.{0,2,4}.{weight,bias}``, ``embedders.i.*``)."""
from __future__ import annotations

import math

import torch
from torch import nn


class GaussianFourierProjection(nn.Module):
    """commonlayers.py:161-190 (callable form, for embedders; the networks use layers.FourierParams + dsk_fourier)."""

    def __init__(self, embed_dim: int, scale: float = 30.0):
        super().__init__()
        self.register_buffer("W", torch.randn(embed_dim // 2) * scale)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        xp = 2 * math.pi * x[..., None] * self.W
        return torch.cat([torch.sin(xp), torch.cos(xp)], dim=-1)


class PorosityEmbedder(nn.Module):
    """embedder.py:199-228: y['porosity'] [B, 1] -> Fourier features -> Linear/SiLU/Linear/SiLU/Linear -> [B, dembed]."""

    def __init__(self, dembed: int, scale: float = 30.0):
        super().__init__()
        self.dembed, self.scale = dembed, scale
        self.gaussian_proj = GaussianFourierProjection(dembed, scale)
        self.net = nn.Sequential(nn.Linear(dembed, 4 * dembed), nn.SiLU(), nn.Linear(4 * dembed, 4 * dembed), nn.SiLU(),
                                 nn.Linear(4 * dembed, dembed))

    def forward(self, x):
        return self.net(self.gaussian_proj(x["porosity"].squeeze(-1)))

    def export_description(self):
        return {"dembed": self.dembed, "scale": self.scale}


class CompositeEmbedder(nn.Module):
    """embedder.py:231-246: sum of several embedders' outputs."""

    def __init__(self, embedders):
        super().__init__()
        self.embedders = nn.ModuleList(embedders)

    def forward(self, x):
        return torch.sum(torch.stack([e(x) for e in self.embedders], dim=0), dim=0)

    def export_description(self):
        return {f"embedder_{i}": e.export_description() for i, e in enumerate(self.embedders)}
