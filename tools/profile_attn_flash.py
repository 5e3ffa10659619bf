"""dsk_attn_flash alone at the C4 / C5 bottom-level shape (L = 4096 tokens, C = 256) next to the multi-launch GEMM attention
core it replaces (two QK^T passes + P V with the probabilities in HBM): CUDA-event timing, algorithmic FLOPs 4 B L^2 C.
  python tools/profile_attn_flash.py [--batch 8] [--tokens 4096] [--channels 256] [--dtype fp16]
With cudaProfilerStart/Stop around two flash launches for
  ncu --profile-from-start off --set full --clock-control none -k regex:attn_flash -c 2 -o out python tools/profile_attn_flash.py"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from diffsci_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--tokens", type=int, default=4096)
ap.add_argument("--channels", type=int, default=256)
ap.add_argument("--dtype", default="fp16")
ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()
dev = torch.device("cuda:0")
torch.manual_seed(0)
B, L, C = a.batch, a.tokens, a.channels
dt = torch.float16 if a.dtype == "fp16" else torch.bfloat16
qkv = torch.randn(B * L, 3 * C, device=dev).to(dt)
out = torch.empty(B * L, C, dtype=dt, device=dev)
old = ops.attention_tc_buffers(B, L, C, dev, dt)


def flash():
    ops.attn_flash(qkv, out, B, L, C)


def gemm_chain():
    ops.attn_softmax_qk(qkv, old["probs"], old["rowstat"], B, L, C)
    ops.gemm_bf16_tc(old["probs"], qkv, old["ao"], M=L, N=C, K=L, lda=L, ldb=3 * C, ldc=C, batch=B, strideA=L * L, strideB=L * 3 * C,
                     strideC=L * C, b_off=2 * C, transB=True)


def timeit(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


fl = 4.0 * B * L * L * C
res = {"shape": {"batch": B, "tokens": L, "channels": C, "dtype": a.dtype}, "algorithmic_flops": fl}
for name, fn in (("flash", flash), ("gemm_chain", gemm_chain)):
    ms = timeit(fn)
    res[name] = {"us": ms * 1e3, "TFLOP_per_s": fl / ms / 1e9}
err = float((out.float() - old["ao"].float()).abs().max() / old["ao"].float().abs().max())
res["flash_vs_gemm_chain_max_rel"] = err
torch.cuda.profiler.start()
flash()
flash()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(json.dumps(res))
