"""Diagnostic: training-gradient error (global L2 vs fp64 autograd of the oracle) of PUNetG-3D mc=64 in bf16 / fp32 mode,
circular vs zero-padded convolutions, same weights and inputs."""
import math, sys, types, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diffsci_b200 as d
from oracle import nets_oracle as N
DEV = "cuda:0"
torch.manual_seed(9)
kw = dict(dimension=3, model_channels=64, channel_expansion=[2], convolution_type="circular")
base = d.PUNetG(d.PUNetGConfig(**kw))
sd = {k: v.detach().double() for k, v in base.state_dict().items()}
sd0 = {k.replace(".conv.weight", ".weight").replace(".conv.bias", ".bias") if not k.startswith(("downsamplers", "upsamplers"))
       else k.replace(".conv.conv.", ".conv."): v for k, v in sd.items()}
x, t = torch.randn(2, 1, 8, 16, 16), torch.tensor([0.3, -0.8])
dF = torch.randn(2, 1, 8, 16, 16)
for ct, s in (("circular", sd), ("default", sd0)):
    cfg = types.SimpleNamespace(**d.PUNetGConfig(**dict(kw, convolution_type=ct)).export_description())
    sdg = {k: v.clone().requires_grad_(True) for k, v in s.items()}
    N.punetg_forward(sdg, cfg, x.double(), t.double()).backward(dF.double())
    for prec in ("fp32", "bf16"):
        net = d.PUNetG(d.PUNetGConfig(**dict(kw, convolution_type=ct)), precision=prec)
        net.load_state_dict({k: v.float() for k, v in s.items()})
        net = net.to(DEV).train()
        F_ = net(x.to(DEV), t.to(DEV))
        F_.backward(dF.to(DEV))
        num = den = 0.0
        worst = (0, "")
        for k, p in net.named_parameters():
            e = float((p.grad.cpu().double() - sdg[k].grad).pow(2).sum()); n = float(sdg[k].grad.pow(2).sum())
            num += e; den += n
            worst = max(worst, (math.sqrt(e / max(n, 1e-300)), k))
        print(f"{ct:9s} {prec}: global L2 {math.sqrt(num / den):.3e}; worst tensor {worst[1]} {worst[0]:.3e}", flush=True)
