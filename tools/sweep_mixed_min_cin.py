"""Which layers of the fp16x2m mode need split (hi + lo) activations?  Sweeps punetg.MIXED_MIN_CIN (contractions with fewer
input channels take plain fp16 activations; storage between kernels stays fp32) and reports, per setting, the full-size
denoiser error against the CPU oracle (bench.denoiser_check) and the time of one batched evaluation.
  python tools/sweep_mixed_min_cin.py [--workload c4] [--batch 8] [--min-cin 128,256,100000]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from diffsci_b200.models.nets import punetg as P  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c4")
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--min-cin", default="128,256,100000")
a = ap.parse_args()
dev = torch.device("cuda:0")
module, net, cfg, shape, nsteps, integ, _, flops = bench.build_workload(a.workload, dev, "fp16x2m")
res = []
ref = None
for mc in (int(v) for v in a.min_cin.split(",")):
    P.MIXED_MIN_CIN = mc
    net._plans.clear()
    module._engines.clear()
    chk = bench.denoiser_check(module, net, cfg, shape, dev, ("fp16x2m",))
    plan = net.plan(a.batch, shape[1:], dev)
    plan.xin.copy_(torch.randn(plan.xin.shape, device=dev).to(plan.xin.dtype))
    cn = torch.zeros(a.batch, device=dev)
    for _ in range(3):
        plan.forward(plan.xin, cn)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        plan.forward(plan.xin, cn)
    e1.record()
    torch.cuda.synchronize()
    row = {"workload": a.workload, "mixed_min_cin": mc, "denoiser_max_rel": chk["max_rel"]["fp16x2m"],
           "denoiser_rel_l2": chk["rel_l2"]["fp16x2m"], "ms_per_evaluation": e0.elapsed_time(e1) / 5, "batch": a.batch}
    print(json.dumps(row), flush=True)
    net._plans.clear()
    torch.cuda.empty_cache()
