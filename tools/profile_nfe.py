"""One eager network evaluation (PUNetG forward) bracketed by cudaProfilerStart/Stop, for ncu launch lists:
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv \
      python tools/profile_nfe.py [--workload c4] [--batch 2] [--precision bf16]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c4")
ap.add_argument("--batch", type=int, default=2)
ap.add_argument("--precision", default="bf16")
ap.add_argument("--reps", type=int, default=1)
a = ap.parse_args()
dev = torch.device("cuda:0")
module, net, cfg, shape, nsteps, integ, _, flops = bench.build_workload(a.workload, dev, a.precision)
plan = net.plan(a.batch, shape[1:], dev)
plan.xin.copy_(torch.randn(plan.xin.shape, device=dev).to(plan.xin.dtype))
cn = torch.zeros(a.batch, device=dev)
for _ in range(2):
    plan.forward(plan.xin, cn)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.profiler.start()
e0.record()
for _ in range(a.reps):
    plan.forward(plan.xin, cn)
e1.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
ms = e0.elapsed_time(e1) / a.reps
print(f"one NFE, B={a.batch}, {a.precision}: {ms:.3f} ms  ({flops * a.batch / ms / 1e9:.1f} TFLOP/s model-level)")
