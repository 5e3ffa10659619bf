"""Precision budget of split-operand tensor-core modes (TEST INFRASTRUCTURE ONLY; CPU emulation, no product code).

Every contraction of the PUNetG oracle (convolutions, linears, QK^T, PV) is replaced by an emulation of what a tcgen05
kernel would compute with operands rounded to a 16/19-bit format and optionally split into hi + lo (+ lo2) parts:
``sum over kept (i, j) part pairs of contract(a_i, b_j)`` with fp32 accumulation; everything else (norms, SiLU, residual
stream, time embedding) stays fp32 exactly as in the fp32 storage mode of the product.  Prints max-rel / L2 of the network
output against the fp64 evaluation of the same oracle and against its plain fp32 evaluation (the reference's arithmetic,
nets/punetg.py:389-416).

    python oracle/split_budget.py [--mc 64] [--size 32] [--dim 3]
"""
from __future__ import annotations

import argparse
import math
import os
import sys
import types

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import nets_oracle as N  # noqa: E402


def rnd(x, fmt):
    if fmt == "bf16":
        return x.bfloat16().float()
    if fmt == "fp16":
        return x.half().float()
    if fmt == "tf32":                                   # round to nearest, 10 explicit mantissa bits
        i = x.contiguous().view(torch.int32)
        return ((i + 0x1000) & ~0x1FFF).view(torch.float32)
    raise ValueError(fmt)


def split(x, fmt, n):
    parts, r = [], x
    for _ in range(n):
        p = rnd(r, fmt)
        parts.append(p)
        r = r - p
    return parts


MODES = {
    # name: (format, parts of A (activations), parts of B (weights), kept (i, j) pairs), MMAs per k-step
    "bf16x1": ("bf16", 1, 1, [(0, 0)]),
    "fp16x1": ("fp16", 1, 1, [(0, 0)]),
    "tf32x1": ("tf32", 1, 1, [(0, 0)]),
    "bf16x2 (act hi+lo)": ("bf16", 2, 1, [(0, 0), (1, 0)]),
    "fp16x2 (act hi+lo)": ("fp16", 2, 1, [(0, 0), (1, 0)]),
    "bf16x3": ("bf16", 2, 2, [(0, 0), (0, 1), (1, 0)]),
    "fp16x3": ("fp16", 2, 2, [(0, 0), (0, 1), (1, 0)]),
    "tf32x3": ("tf32", 2, 2, [(0, 0), (0, 1), (1, 0)]),
    "bf16x6": ("bf16", 3, 3, [(0, 0), (0, 1), (1, 0), (1, 1), (0, 2), (2, 0)]),
}


class Shim(types.SimpleNamespace):
    """Stands in for torch.nn.functional inside nets_oracle: contractions emulated, the rest delegated."""

    def __init__(self, mode, storage=None):
        super().__init__()
        self.fmt, self.na, self.nb, self.pairs = MODES[mode]
        self.storage = storage                          # None: fp32 tensors between kernels; else 16-bit storage format

    def st(self, x):
        return x if self.storage is None else rnd(x, self.storage)

    def silu(self, x):                                  # every norm + SiLU output is a stored tensor
        return self.st(F.silu(x))

    def __getattr__(self, k):
        return getattr(F, k)

    def _contract(self, fn, a, b):
        ap, bp = split(a, self.fmt, self.na), split(b, self.fmt, self.nb)
        out = None
        for i, j in sorted(self.pairs, key=lambda p: -(p[0] + p[1])):     # small terms first
            t = fn(ap[i], bp[j])
            out = t if out is None else out + t
        return out

    def conv2d(self, x, w, b=None, padding=0):
        y = self._contract(lambda a, c: F.conv2d(a, c, None, padding=padding), x, w)
        return self.st(y if b is None else y + b.view(1, -1, 1, 1))

    def conv3d(self, x, w, b=None, padding=0):
        y = self._contract(lambda a, c: F.conv3d(a, c, None, padding=padding), x, w)
        return self.st(y if b is None else y + b.view(1, -1, 1, 1, 1))

    def linear(self, x, w, b=None):
        y = self._contract(lambda a, c: F.linear(a, c), x, w)
        return y if b is None else y + b


def make_attention(shim):
    def mha(x_nc, sd, prefix, residual):
        B, C = x_nc.shape[:2]
        S = x_nc.shape[2:]
        tok = x_nc.reshape(B, C, -1).transpose(1, 2)
        qkv = shim.linear(tok, sd[prefix + "mhattn.in_proj_weight"], sd[prefix + "mhattn.in_proj_bias"])
        q, k, v = qkv.chunk(3, dim=-1)
        s = shim._contract(lambda a, c: a @ c.transpose(1, 2), q, k) / math.sqrt(C)
        a = shim._contract(lambda p, c: p @ c, torch.softmax(s, dim=-1), v)
        o = shim.linear(a, sd[prefix + "mhattn.out_proj.weight"], sd[prefix + "mhattn.out_proj.bias"])
        o = o.transpose(1, 2).reshape(B, C, *S)
        return x_nc + o if residual else o
    return mha


def relmax(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mc", type=int, default=64)
    ap.add_argument("--size", type=int, default=32)
    ap.add_argument("--dim", type=int, default=3)
    ap.add_argument("--modes", default=",".join(MODES))
    a = ap.parse_args()
    import diffsci_b200 as d
    torch.manual_seed(0)
    cfg = d.PUNetGConfig(dimension=a.dim, model_channels=a.mc)
    ocfg = types.SimpleNamespace(**cfg.export_description())
    from diffsci_b200.models.nets.punetg import PUNetG
    sd = N.synth_state_dict([(k, tuple(v.shape)) for k, v in PUNetG(cfg).state_dict().items()], seed=3)
    x = torch.randn(1, 1, *([a.size] * a.dim))
    t = torch.tensor([0.3])
    ref32 = N.punetg_forward(sd, ocfg, x, t)
    ref64 = N.punetg_forward({k: v.double() for k, v in sd.items()}, ocfg, x.double(), t.double())
    print(f"PUNetG dim={a.dim} mc={a.mc} size={a.size}: reference fp32 vs fp64 max-rel {relmax(ref32, ref64):.2e} "
          f"L2 {rel_l2(ref32, ref64):.2e}")
    print(f"{'mode':22s} {'MMAs':>4s}  {'max-rel vs fp64':>15s} {'L2 vs fp64':>11s}  {'max-rel vs fp32':>15s}")
    real_F, real_mha, real_block = N.F, N.mha_self_attention, N.resnet_block_c
    runs = [(m, None) for m in a.modes.split(",")] + [("bf16x1", "bf16"), ("fp16x1", "fp16")]
    for mode, storage in runs:
        shim = Shim(mode, storage)
        N.F, N.mha_self_attention = shim, make_attention(shim)
        N.resnet_block_c = lambda *aa, **kw: shim.st(real_block(*aa, **kw))      # block output (residual add) is stored
        try:
            out = N.punetg_forward(sd, ocfg, x, t)
        finally:
            N.F, N.mha_self_attention, N.resnet_block_c = real_F, real_mha, real_block
        name = mode + (f" + {storage} storage" if storage else "")
        print(f"{name:26s} {len(MODES[mode][3]):4d}  {relmax(out, ref64):15.2e} {rel_l2(out, ref64):11.2e}  "
              f"{relmax(out, ref32):15.2e}", flush=True)


if __name__ == "__main__":
    main()
