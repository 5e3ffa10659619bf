"""The attention block alone at the C4 / C5 bottom-level shape (L = 4096 tokens, C = 256):
  python tools/profile_attention.py [--batch 8] [--tokens 4096] [--channels 256]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from diffsci_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--tokens", type=int, default=4096)
ap.add_argument("--channels", type=int, default=256)
ap.add_argument("--reps", type=int, default=10)
a = ap.parse_args()
dev = torch.device("cuda:0")
torch.manual_seed(0)
B, L, C = a.batch, a.tokens, a.channels
bf = dict(dtype=torch.bfloat16, device=dev)
bufs = ops.attention_tc_buffers(B, L, C, dev)
tok = torch.randn(B, L, C, device=dev).bfloat16()
wi = ops.PackedLinear(torch.randn(3 * C, C, device=dev) / C ** 0.5)
wo = ops.PackedLinear(torch.randn(C, C, device=dev) / C ** 0.5)
bi, bo = torch.zeros(3 * C, device=dev), torch.zeros(C, device=dev)
out = torch.empty(B, L, C, **bf)
for _ in range(2):
    ops.self_attention_tc(tok, wi, bi, wo, bo, bufs, out, True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.reps):
    ops.self_attention_tc(tok, wi, bi, wo, bo, bufs, out, True)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.reps
fl = B * (2 * L * C * 4 * C + 4 * L * L * C)
print(f"attention B={B} L={L} C={C}: {ms * 1e3:.1f} us  ({fl / ms / 1e9:.1f} TFLOP/s)")
