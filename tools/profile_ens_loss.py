"""The fused ensemble-loss kernel (csrc/train.cu: ens_loss_kernel) at a C4-sized problem: CUDA-event timing of
dsk_ensemble_loss_fwd_bwd alone, algorithmic bytes (F + noise read, dF written: 12 B per ensemble element; x: 4 B per
element) over the measured time, with cudaProfilerStart/Stop around two launches for
  ncu --profile-from-start off --set full --clock-control none -k regex:ens_loss -c 2 -o out python tools/profile_ens_loss.py
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from diffsci_b200._lib import lib, check, ptr, stream  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--ensemble", type=int, default=4)
ap.add_argument("--side", type=int, default=64)
ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()
dev = torch.device("cuda:0")
B, E, C, S = a.batch, a.ensemble, 1, a.side ** 3
torch.manual_seed(0)
x = torch.randn(B, C, S, device=dev) * 0.5
noise = torch.randn(B, E, C, S, device=dev)
F = torch.randn(B * E, C, S, device=dev)
dF = torch.empty_like(F)
sigma = torch.exp(torch.randn(B, device=dev) * 1.2 - 1.2)
c_out = sigma * 0.5 / torch.sqrt(sigma ** 2 + 0.25)
c_skip = 0.25 / (sigma ** 2 + 0.25)
s1 = torch.full((B,), 1.0 / (B * E * C * S), device=dev)
s2 = torch.full((B,), 0.5 / (B * max(E * (E - 1) / 2, 1) * C * S), device=dev)
loss = torch.zeros((), device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > the 126 MB L2
out = {}
for kind, name in ((0, "huber"), (2, "CRPS")):
    def launch():
        check(lib.dsk_ensemble_loss_fwd_bwd(ptr(F), ptr(x), ptr(noise), ptr(sigma), ptr(c_out), ptr(c_skip), ptr(s1), ptr(s2),
                                            None, 0, ptr(loss), ptr(dF), B, E, C, S, kind, stream()))
    for _ in range(3):
        launch()
    ts = []
    for _ in range(a.reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        launch()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    nbytes = 12 * B * E * C * S + 4 * B * C * S
    out[name] = {"ms": ms, "algorithmic_bytes": nbytes, "GB_per_s": nbytes / ms / 1e6}
torch.cuda.profiler.start()
for kind in (0, 2):
    check(lib.dsk_ensemble_loss_fwd_bwd(ptr(F), ptr(x), ptr(noise), ptr(sigma), ptr(c_out), ptr(c_skip), ptr(s1), ptr(s2),
                                        None, 0, ptr(loss), ptr(dF), B, E, C, S, kind, stream()))
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(json.dumps({"B": B, "E": E, "S": S, **out}))
