"""Parameter holders with the reference's state-dict key names and default initialisers.

These modules own weights only; they have NO forward.  All arithmetic is done by the CUDA kernels
(diffsci_b200.ops); calling one of them is a bug and raises.  Initialisers follow the laws of the
torch layers the reference instantiates (kaiming-uniform(a=sqrt(5)) == U(-1/sqrt(fan_in), +) for
Conv/Linear weight and bias; xavier-uniform in_proj / zero biases for nn.MultiheadAttention).
"""
from __future__ import annotations

import math

import torch
from torch import nn


class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError(f"{type(self).__name__} only stores parameters; compute goes through diffsci_b200.ops")


def _uniform_(t: torch.Tensor, bound: float):
    with torch.no_grad():
        return t.uniform_(-bound, bound)


class ConvParams(_Holder):
    """weight [Cout, Cin, k(,k)(,k)], bias [Cout] -- torch.nn.Conv{2,3}d keys."""
    circular = False

    def __init__(self, cin: int, cout: int, ksize: int, ndim: int, bias: bool = True):
        super().__init__()
        self.cin, self.cout, self.ksize, self.ndim = cin, cout, ksize, ndim
        self.weight = nn.Parameter(torch.empty((cout, cin) + (ksize,) * ndim))
        bound = 1.0 / math.sqrt(cin * ksize ** ndim)
        _uniform_(self.weight, bound)
        self.bias = nn.Parameter(_uniform_(torch.empty(cout), bound)) if bias else None


class CircularConvParams(_Holder):
    """CircularConv2d / CircularConv3d (commonlayers.py:918-1032; all spatial axes periodic): the reference wraps a Conv
    as ``self.conv``, so the state-dict keys are ``<name>.conv.weight`` / ``<name>.conv.bias``."""
    circular = True

    def __init__(self, cin: int, cout: int, ksize: int, ndim: int, bias: bool = True):
        super().__init__()
        self.cin, self.cout, self.ksize, self.ndim = cin, cout, ksize, ndim
        self.conv = ConvParams(cin, cout, ksize, ndim, bias)

    @property
    def weight(self):
        return self.conv.weight

    @property
    def bias(self):
        return self.conv.bias


def make_conv(cin: int, cout: int, ksize: int, ndim: int, bias: bool = True, convolution_type: str = "default"):
    if convolution_type == "circular":
        return CircularConvParams(cin, cout, ksize, ndim, bias)
    return ConvParams(cin, cout, ksize, ndim, bias)


class LinearParams(_Holder):
    def __init__(self, cin: int, cout: int, bias: bool = True):
        super().__init__()
        self.weight = nn.Parameter(_uniform_(torch.empty(cout, cin), 1.0 / math.sqrt(cin)))
        self.bias = nn.Parameter(_uniform_(torch.empty(cout), 1.0 / math.sqrt(cin))) if bias else None


class NormParams(_Holder):
    """weight/bias [C] of GroupNorm / GroupRMSNorm (ones / zeros)."""

    def __init__(self, channels: int, affine: bool = True):
        super().__init__()
        if affine:
            self.weight = nn.Parameter(torch.ones(channels))
            self.bias = nn.Parameter(torch.zeros(channels))
        else:
            self.weight = self.bias = None


class _Act(_Holder):
    """parameter-free placeholder so that Sequential indices match the reference (net.0/.2/.4)."""


class TimeBlockParams(_Holder):
    """ResnetTimeBlock (commonlayers.py:516-550): net = Linear(M,4M) SiLU Linear(4M,4M) SiLU Linear(4M,C)."""

    def __init__(self, embed: int, out: int):
        super().__init__()
        self.net = nn.Sequential(LinearParams(embed, 4 * embed), _Act(), LinearParams(4 * embed, 4 * embed), _Act(),
                                 LinearParams(4 * embed, out))


class MHAParams(_Holder):
    """nn.MultiheadAttention(C, num_heads=1, batch_first=True) parameter layout."""

    def __init__(self, channels: int):
        super().__init__()
        self.in_proj_weight = nn.Parameter(torch.empty(3 * channels, channels))
        nn.init.xavier_uniform_(self.in_proj_weight)
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * channels))
        self.out_proj = LinearParams(channels, channels)
        with torch.no_grad():
            self.out_proj.bias.zero_()


class AttentionParams(_Holder):
    def __init__(self, channels: int):
        super().__init__()
        self.mhattn = MHAParams(channels)


class FourierParams(_Holder):
    """GaussianFourierProjection (commonlayers.py:161-173): buffer W ~ N(0, scale^2), [embed/2]."""

    def __init__(self, embed_dim: int, scale: float = 30.0):
        super().__init__()
        self.register_buffer("W", torch.randn(embed_dim // 2) * scale)


class ConditionDrop(nn.Module):
    """commonlayers.py:1100-1127: during training, replace a sample's conditioning vector by a (learnable) null embedding
    with probability p.  Acts on the [B, M] conditioning vectors, i.e. before they enter the fused network path; it is a
    real (callable) module, unlike the holders above."""

    def __init__(self, p: float, hidden_dim: int, null_is_learnable: bool = True):
        super().__init__()
        self.p = p
        if null_is_learnable:
            self.null_embedding = nn.Parameter(torch.randn(1, hidden_dim))
        else:
            self.register_buffer("null_embedding", torch.zeros(1, hidden_dim))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not self.training or self.p == 0.0:
            return x
        mask_shape = (x.shape[0],) + (1,) * (x.ndim - 1)
        mask = torch.bernoulli(torch.full(mask_shape, 1 - self.p, device=x.device))
        return torch.where(mask == 1, x, self.null_embedding)
