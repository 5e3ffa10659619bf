"""Where do two precision modes of a PUNetG plan diverge?  Runs one evaluation in each mode and compares the plan's level
buffers (final contents) -- a coarse bisection over the network.    python tools/diag_precision_layers.py [dim] [mc] [size] [B]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diffsci_b200 as d  # noqa: E402

dim, mc, size, B = (int(a) for a in (sys.argv[1:5] + ["2", "64", "64", "2"][len(sys.argv) - 1:]))
modes = sys.argv[5:] or ["fp32_ffma", "fp32"]
dev = "cuda:0"
torch.manual_seed(0)
net = d.PUNetG(d.PUNetGConfig(dimension=dim, model_channels=mc), precision=modes[0]).to(dev).eval()
x = torch.randn(B, 1, *([size] * dim), device=dev)
t = torch.tensor([0.3, -0.8, 0.1, 1.0][:B], device=dev)
snap = {}
for m in modes:
    net.precision = m
    with torch.no_grad():
        y = net(x, t)
    plan = net.plan(B, tuple(x.shape[2:]), x.device)
    s = {"out": y.double().cpu()}
    for name in ("X", "XU", "Y", "P"):
        for l, b in enumerate(getattr(plan, name)):
            s[f"{name}[{l}]"] = b.double().cpu()
    s["XA"], s["XA2"] = plan.XA.double().cpu(), plan.XA2.double().cpu()
    snap[m] = s
a, b = snap[modes[0]], snap[modes[1]]
for k in a:
    if a[k].shape == b[k].shape:
        e = float((a[k] - b[k]).abs().max() / a[k].abs().max().clamp_min(1e-30))
        l2 = float((a[k] - b[k]).norm() / a[k].norm().clamp_min(1e-30))
        print(f"{k:8s} max-rel {e:.2e}  L2 {l2:.2e}  shape {tuple(a[k].shape)}")
