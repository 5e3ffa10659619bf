"""A few EDM training iterations (EDMTrainer.step) bracketed by cudaProfilerStart/Stop, for ncu launch lists:
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv \
      python tools/profile_train.py [--workload c3train] [--batch 8]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
import diffsci_b200 as d  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c3train")
ap.add_argument("--batch", type=int, default=0)
ap.add_argument("--precision", default="bf16")
ap.add_argument("--reps", type=int, default=1)
a = ap.parse_args()
dev = torch.device("cuda:0")
module, net, cfg, shape, metric, batch, flops = bench.build_train_workload(a.workload, dev, a.precision)
B = a.batch or batch
tr = d.EDMTrainer(module, ema=d.ModelEMA(net, ema_type="traditional", decay=0.999))
x = (torch.randn(B, *shape) * 0.5).to(dev)
for _ in range(2):
    tr.step(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.profiler.start()
e0.record()
for _ in range(a.reps):
    loss = tr.step(x)
e1.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
ms = e0.elapsed_time(e1) / a.reps
print(f"one training iteration {a.workload}, B={B}, {a.precision}: {ms:.3f} ms  ({3 * flops * B / ms / 1e9:.1f} TFLOP/s model-level), "
      f"loss {float(loss):.4f}")
