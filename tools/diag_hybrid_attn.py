import math, sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diffsci_b200 as d
from diffsci_b200.models.nets import graph as G
DEV = "cuda:0"
torch.manual_seed(21)
kw = dict(dimension=2, model_channels=16)
net = d.PUNetG(d.PUNetGConfig(**kw), precision="bf16").to(DEV).train()
x, t = torch.randn(8, 1, 28, 28), torch.randn(8) * 0.5
dF = torch.randn(8, 1, 28, 28)
grads = {}
for hybrid in (True, False, "fp32"):
    G.HYBRID_ATTENTION = hybrid is True
    net.precision = "fp32" if hybrid == "fp32" else "bf16"
    net._plans.clear(); net.zero_grad()
    net(x.to(DEV), t.to(DEV)).backward(dF.to(DEV))
    grads[hybrid] = {k: p.grad.double().cpu().clone() for k, p in net.named_parameters()}
def rl2(a, b): return float((a - b).norm() / b.norm())
for k in grads[True]:
    if "mhattn" in k:
        a, b, c = grads[True][k], grads[False][k], grads["fp32"][k]
        if k.endswith("in_proj_weight") or k.endswith("in_proj_bias"):
            n = a.shape[0] // 3
            for j, nm in enumerate("QKV"):
                sl = slice(j * n, (j + 1) * n)
                print(k, nm, f"hyb-vs-f32core {rl2(a[sl], b[sl]):.3e}  hyb-vs-fp32net {rl2(a[sl], c[sl]):.3e}  f32core-vs-fp32net {rl2(b[sl], c[sl]):.3e}  norm {float(c[sl].norm()):.3e}")
        else:
            print(k, f"hyb-vs-f32core {rl2(a, b):.3e}  hyb-vs-fp32net {rl2(a, c):.3e}  f32core-vs-fp32net {rl2(b, c):.3e} norm {float(c.norm()):.3e}")
