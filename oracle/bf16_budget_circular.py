"""bf16 error budget of PERIODIC networks, measured on the REFERENCE ITSELF (build container only).

PUNetGConfig(convolution_type="circular") networks are about twice as sensitive to rounding as their zero-padded twins
(same weights, same input): the reference's own fp32 run is 2x further from fp64, and under torch.autocast(bfloat16) its
L2 error doubles (4.6e-2 vs 2.1e-2 at mc=64).  tests/test_gpu_circular.py states its bf16 tolerances against these
numbers (output: profiles/r1_bf16_budget_circular.txt).  Columns: max-rel, L2-rel of autocast-bf16 vs fp64; max-rel fp32 vs fp64.

    python oracle/bf16_budget_circular.py
"""
import sys, torch
import os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import refload
refload.load_reference()
from diffsci.models.nets.punetg import PUNetG
from diffsci.models.nets.punetg_config import PUNetGConfig
torch.set_num_threads(8)
def relmax(a,b): return float((a.double()-b.double()).abs().max()/b.double().abs().max())
def rel_l2(a,b): return float((a.double()-b.double()).norm()/b.double().norm())
for kw, shape in [(dict(dimension=3, model_channels=64, channel_expansion=[2]), (2,1,8,16,16)),
                  (dict(dimension=2, model_channels=8), (2,1,16,24))]:
    for seed in (9, 10):
        res = {}
        for ct in ("default", "circular"):
            torch.manual_seed(seed)
            net = PUNetG(PUNetGConfig(**kw, convolution_type=ct)).eval()
            torch.manual_seed(100+seed)
            x, t = torch.randn(*shape), torch.tensor([0.3, -0.8])
            with torch.no_grad():
                y32 = net(x, t)
                with torch.autocast("cpu", dtype=torch.bfloat16):
                    yac = net(x, t).float()
                y64 = net.double()(x.double(), t.double())
            res[ct] = (relmax(yac, y64), rel_l2(yac, y64), relmax(y32, y64))
        print(kw.get('model_channels'), seed, {k: tuple(f"{v:.2e}" for v in r) for k, r in res.items()})
