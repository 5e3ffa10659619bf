"""How far is the REFERENCE ITSELF from fp64 truth when it runs in bf16?  (build container only)

The north star asks for denoiser outputs "within 1e-3 relative in bf16".  The reference has no bf16 mode of its own; the
two ways a user would run it in bf16 are torch.autocast(bfloat16) (bf16 operands for conv / linear / matmul, fp32
normalisation) and module.bfloat16() (everything in bf16).  This script measures, on the synthetic-weight networks of the
golden fixtures, max|F - F64| / max|F64| for: reference fp32, reference under autocast-bf16, reference .bfloat16().
The numbers are quoted in DESIGN.md section 2 next to this repo's bf16 mode measured the same way on the B200.

    python oracle/bf16_budget.py
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refload  # noqa: E402
from nets_oracle import synth_state_dict  # noqa: E402

diffsci = refload.load_reference()
from diffsci.models.nets.punetg import PUNetG  # noqa: E402
from diffsci.models.nets.punetg_config import PUNetGConfig  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")


def relmax(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


for name, cfgk, shape in [("punetg2d_mc8", None, None), ("punetg3d_mc8", None, None),
                          ("punetg2d mc=64 32x32", dict(dimension=2, model_channels=64), (2, 1, 32, 32)),
                          ("punetg3d mc=64 16^3", dict(dimension=3, model_channels=64), (1, 1, 16, 16, 16))]:
    if cfgk is None:
        g = torch.load(os.path.join(GOLD, name + ".pt"), weights_only=False)
        cfgk, x, t, seed = g["cfg"], g["x"], g["t"], g["seed"]
    else:
        torch.manual_seed(5)
        x, t, seed = torch.randn(shape), torch.randn(shape[0]) * 0.6 - 0.6, 11
    net = PUNetG(PUNetGConfig(**cfgk)).eval()
    man = [(k, list(v.shape)) for k, v in net.state_dict().items()]
    net.load_state_dict(synth_state_dict(man, seed))
    with torch.no_grad():
        y32 = net(x, t)
        net64 = PUNetG(PUNetGConfig(**cfgk)).double().eval()
        net64.load_state_dict(synth_state_dict(man, seed, torch.float64))
        y64 = net64(x.double(), t.double())
        with torch.autocast("cpu", dtype=torch.bfloat16):
            yac = net(x, t)
        try:
            nb = PUNetG(PUNetGConfig(**cfgk)).eval()
            nb.load_state_dict(synth_state_dict(man, seed))
            nb = nb.bfloat16()
            ybf = nb(x.bfloat16(), t.bfloat16()).float()
            sbf = f"{relmax(ybf, y64):.2e} / {rel_l2(ybf, y64):.2e}"
        except Exception as e:  # noqa: BLE001
            sbf = f"failed ({type(e).__name__})"
    print(f"{name:26s} max-rel / rel-L2 vs fp64:  ref fp32 {relmax(y32, y64):.2e} / {rel_l2(y32, y64):.2e}   "
          f"ref autocast-bf16 {relmax(yac.float(), y64):.2e} / {rel_l2(yac.float(), y64):.2e}   ref .bfloat16() {sbf}", flush=True)
