"""MLPUncond -- drop-in for diffsci.models.nets.mlp.MLPUncond (reference nets/mlp.py:4-58).

cat[x, t] -> (Linear, act)* -> Linear, each layer one fused GEMM+bias+activation launch
(dsk_gemm_f32).  State-dict keys ``net.{i}.weight/bias`` match the reference's nn.Sequential.
"""
from __future__ import annotations

from typing import Any, Optional

import torch
from torch import nn

from ... import ops
from ..._lib import require_cuda
from .layers import LinearParams, _Act

_ACT_CODE = {nn.ReLU: 2, nn.SiLU: 1}


class MLPUncond(nn.Module):
    ydim = 0          # MLPCond: width of the conditioning vector concatenated after [x, t]

    def __init__(self, dim: int, hidden_dims=[10], nonlinearity: nn.Module = nn.ReLU(), dropout: float = 0.0):
        super().__init__()
        if type(nonlinearity) not in _ACT_CODE:
            raise NotImplementedError(f"diffsci_b200.MLPUncond: activation {type(nonlinearity).__name__} not built "
                                      "(ReLU and SiLU are fused into the GEMM epilogue)")
        self.dim, self.act = dim, _ACT_CODE[type(nonlinearity)]
        self.dropout = dropout
        layers, d = [], dim + 1 + self.ydim
        for h in hidden_dims:
            layers += [LinearParams(d, h), _Act()]
            if dropout > 0:
                layers.append(_Act())      # Dropout slot keeps the reference's Sequential indices
            d = h
        layers.append(LinearParams(d, dim))
        self.net = nn.Sequential(*layers)
        self._plans: dict[Any, "_MLPPlan"] = {}
        self.precision = "fp32"

    def forward(self, x: torch.Tensor, t: torch.Tensor, y: Optional[torch.Tensor] = None) -> torch.Tensor:
        require_cuda(x, "MLP input")
        if (y is None) != (self.ydim == 0):
            raise TypeError(f"{type(self).__name__}.forward: y must{' not' if self.ydim == 0 else ''} be given")
        if self.training and self.dropout > 0:
            raise NotImplementedError("diffsci_b200 MLP: training-mode dropout not built")
        if y is not None:
            y = y.float().expand(x.shape[0], self.ydim).contiguous()        # one condition broadcasts over the batch
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            from .graph import NetFunction        # training: hand-written backward (graph.build_mlp)
            g = self.train_graph(x.shape[0], x.device)
            if y is not None:
                g.y_in.copy_(y.detach())
            return NetFunction.apply(g, x, t, None, *self.parameters())
        plan = self.plan(x.shape[0], tuple(x.shape[1:]), x.device)
        plan.xin.copy_(x.reshape(plan.xin.shape))
        return plan.forward(plan.xin, t.float().contiguous(), y=y).reshape(x.shape).clone()

    def plan(self, B: int, spatial: tuple, device, precision: Optional[str] = None) -> "_MLPPlan":
        key = (B, str(device))
        sig = tuple(p.data_ptr() for p in self.parameters())
        plan = self._plans.get(key)
        if plan is None or plan.sig != sig:
            if len(self._plans) >= 4:
                self._plans.clear()
            with torch.inference_mode(False), torch.no_grad():   # persistent buffers must be normal tensors
                plan = self._plans[key] = _MLPPlan(self, B, device, sig)
        return plan

    def train_graph(self, B: int, device, *_, **__):
        from .graph import build_mlp
        key = ("train", B, str(device))
        sig = tuple(p.data_ptr() for p in self.parameters())
        g = self._plans.get(key)
        if g is None or g.sig != sig:
            for k in [k for k in self._plans if k[0] == "train"]:
                del self._plans[k]
            with torch.inference_mode(False), torch.no_grad():
                g = self._plans[key] = build_mlp(self, B, device)
        return g

    def _apply(self, fn, *a, **k):
        self._plans = {}
        return super()._apply(fn, *a, **k)


class _MLPPlan:
    act_dtype = torch.float32

    def __init__(self, net: MLPUncond, B: int, device, sig):
        self.net, self.B, self.sig = net, B, sig
        f32 = dict(dtype=torch.float32, device=device)
        self.xin = torch.empty((B, net.dim), **f32)
        self.cat = torch.empty((B, net.dim + 1), **f32)
        self.cat_y = torch.empty((B, net.dim + 1 + net.ydim), **f32) if net.ydim else None
        self.layers = [m for m in net.net if isinstance(m, LinearParams)]
        self.h = [torch.empty((B, m.weight.shape[0]), **f32) for m in self.layers]
        self.F = self.h[-1]

    def prepare(self):
        pass

    def forward(self, xin: torch.Tensor, cnoise: torch.Tensor, y: Optional[torch.Tensor] = None, **_) -> torch.Tensor:
        h = ops.concat_channels(xin.view(self.B, -1), cnoise.view(self.B, 1), out=self.cat)
        if self.cat_y is not None:                      # MLPCond: cat[x, t, y] (nets/mlp.py:118-120)
            h = ops.concat_channels(h, y.view(self.B, -1), out=self.cat_y)
        for i, m in enumerate(self.layers):
            last = i == len(self.layers) - 1
            h = ops.linear(h, m.weight, m.bias, act=0 if last else self.net.act, out=self.h[i])
        return h


class MLPCond(MLPUncond):
    """MLPCond (reference nets/mlp.py:61-121): cat[x, t, y] -> (Linear, act)* -> Linear, y a [B, ydim] conditioning vector.
    Same state-dict keys as the reference (``net.<i>.weight/bias``).  With ``KarrasModule(model, cfg, conditional=True)`` it is
    driven through the module's generic denoiser-network call ``model(x_scaled, c_noise, y)`` (no ``plan`` attribute exposed to
    the graph engine: the conditioning is a per-call argument here, not a time-embedding offset)."""

    engine_native = False     # KarrasModule / the sampler engine call it as a generic denoiser network: model(x, c_noise, y)

    def __init__(self, dim: int, ydim: int, hidden_dims=[10], nonlinearity: nn.Module = nn.ReLU(), dropout: float = 0.0):
        self.ydim = int(ydim)
        super().__init__(dim, hidden_dims, nonlinearity, dropout)
