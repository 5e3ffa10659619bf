"""Characterise the accumulation error of tcgen05 kind::f16 MMAs (fp32 accumulators in TMEM): C = A B^T with operands that are
EXACT in fp16, against fp64, as a function of the number of chained MMA instructions (K / 16).  Random-sign and positive-only
operands (a rounding mode that truncates shows up as a one-sided error growing linearly with the chain length).
    python tools/probe_tc_accum.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffsci_b200 import ops  # noqa: E402

dev = "cuda:0"
torch.manual_seed(0)
M, N = 256, 128
print(f"{'K':>6s} {'chain':>6s} {'kind':>9s} {'max-rel':>10s} {'mean signed rel':>16s} {'rms rel':>10s}   (fp32 RN dot for comparison: max-rel)")
for K in (64, 256, 1024, 1728, 4096, 16384):
    for kind in ("randsign", "positive"):
        A = torch.randn(M, K)
        B = torch.randn(N, K)
        if kind == "positive":
            A, B = A.abs(), B.abs()
        A, B = A.half(), B.half()
        ref = A.double() @ B.double().t()
        out = torch.empty(M, N, device=dev)
        ops.gemm_bf16_tc(A.to(dev), B.to(dev), out, M=M, N=N, K=K, lda=K, ldb=K, ldc=N)
        got = out.double().cpu()
        scale = ref.abs().max()
        err = (got - ref) / scale
        f32 = (A.float().to(dev) @ B.float().to(dev).t()).double().cpu()       # cuBLAS fp32 (TF32 off by default)
        e32 = ((f32 - ref) / scale).abs().max()
        print(f"{K:6d} {K // 16:6d} {kind:>9s} {err.abs().max():10.2e} {err.mean():16.2e} {err.pow(2).mean().sqrt():10.2e}   {e32:.2e}")
