"""The last convolution (C -> 1) alone: python tools/profile_convout.py [--batch 8] [--size 64] [--cin 64]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from diffsci_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--size", type=int, default=64)
ap.add_argument("--cin", type=int, default=64)
ap.add_argument("--cout", type=int, default=1)
ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()
dev = torch.device("cuda:0")
torch.manual_seed(0)
w = torch.randn(a.cout, a.cin, 3, 3, 3, device=dev) * 0.02
pc = ops.PackedConv(w, torch.zeros(a.cout, device=dev), 3, torch.bfloat16)
xs = [torch.randn(a.batch, a.size, a.size, a.size, a.cin, device=dev).bfloat16() for _ in range(3)]
out = torch.empty(a.batch, a.cout, a.size, a.size, a.size, device=dev)
for i in range(3):
    ops.conv(xs[i % 3], pc, out=out, out_nchw=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(a.reps):
    ops.conv(xs[i % 3], pc, out=out, out_nchw=True)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.reps
gb = a.batch * a.size ** 3 * a.cin * 2 / 1e9
print(f"convout {a.cin}->{a.cout} @ {a.size}^3 B={a.batch}: {ms * 1e3:.1f} us  {gb / ms * 1e3:.0f} GB/s of input")
