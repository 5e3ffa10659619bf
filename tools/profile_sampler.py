"""Two eager Heun steps of C4 sampling (fused sampler stages visible as individual launches) bracketed by
cudaProfilerStart/Stop, for ncu captures of the sampler-stage kernel:
  ncu --profile-from-start off --set full --clock-control none -k regex:sampler_stage -c 2 -o out python tools/profile_sampler.py
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c4")
ap.add_argument("--batch", type=int, default=8)
a = ap.parse_args()
dev = torch.device("cuda:0")
module, net, cfg, shape, nsteps, integ, _, flops = bench.build_workload(a.workload, dev, "bf16")
module.use_cuda_graphs = False
wn = torch.randn(a.batch, *shape, device=dev)
module.propagate_white_noise(wn, nsteps=2)
torch.cuda.synchronize()
torch.cuda.profiler.start()
module.propagate_white_noise(wn, nsteps=2)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
