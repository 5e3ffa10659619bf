"""Full-size denoiser error (vs the CPU oracle, bench.denoiser_check) and time of one batched evaluation for precision modes:
  python tools/check_mode.py [--workload c4] [--batch 8] [--modes fp16s32,fp16x2m]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c4")
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--modes", default="fp16s32,fp16x2m")
a = ap.parse_args()
dev = torch.device("cuda:0")
modes = a.modes.split(",")
module, net, cfg, shape, nsteps, integ, _, flops = bench.build_workload(a.workload, dev, modes[0])
chk = bench.denoiser_check(module, net, cfg, shape, dev, modes)
for m in modes:
    net.precision = m
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
    print(json.dumps({"workload": a.workload, "precision": m, "env_DSK_Y16": os.environ.get("DSK_Y16", "0"),
                      "denoiser_max_rel": chk["max_rel"][m], "denoiser_rel_l2": chk["rel_l2"][m],
                      "ms_per_evaluation": e0.elapsed_time(e1) / 5, "batch": a.batch}), flush=True)
    net._plans.clear()
    torch.cuda.empty_cache()
