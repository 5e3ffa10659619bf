import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import diffsci_b200 as d
dev = "cuda:0"
torch.manual_seed(0)
def run(kw, name, shape=(2, 1, 32, 32, 32)):
    cfg = d.PUNetGConfig(**kw)
    net = d.PUNetG(cfg, precision="fp32").to(dev).eval()
    x = torch.randn(*shape, device=dev); t = torch.tensor([0.3, -0.8], device=dev)
    with torch.no_grad():
        ref = net(x, t).float().cpu()
        net.precision = "fp16s32"
        y = net(x, t).float().cpu()
    err = float((y - ref).abs().max() / ref.abs().max())
    print(f"{name}: fp16s32 vs fp32 mode max rel {err:.3e}", flush=True)
    assert err < 3e-3, err
run(dict(dimension=3, model_channels=64, channel_expansion=[2]), "zero-padded 3-D")
run(dict(dimension=3, model_channels=64, channel_expansion=[2], convolution_type="circular"), "circular 3-D")
run(dict(dimension=3, model_channels=64, channel_expansion=[2], bias=False), "bias=False 3-D")
run(dict(dimension=3, model_channels=64, channel_expansion=[2, 4]), "two levels 3-D", shape=(2, 1, 16, 32, 32))
run(dict(dimension=2, model_channels=64, channel_expansion=[2]), "2-D (fp32 conv1 outputs)", shape=(2, 1, 64, 64))
print("ok")
