"""Merged-depth-tap convolution (conv_tc2_kernel<.., T = 2>, DSK_CONV_T2) against torch's fp32 conv3d of the same rounded operands, and its time.
Run once per setting of DSK_CONV_T2 (the library reads it once):  DSK_CONV_T2=0 python tools/check_conv_t2.py ; DSK_CONV_T2=1 python ...
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from diffsci_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
quick = "--quick" in sys.argv


def case(B, S, cin, cout, prec, residual, stats, reps=6, D=None):
    torch.manual_seed(1)
    D = D or S
    w = torch.randn(cout, cin, 3, 3, 3, device=dev) * 0.05
    bias = torch.randn(cout, device=dev)
    h16 = torch.bfloat16 if prec == "bf16" else torch.float16
    odt = torch.float32 if prec == "fp16s32" else h16
    pc = ops.PackedConv(w, bias, 3, h16)
    x = torch.randn(B, D, S, S, cin, device=dev).to(h16)
    cb = torch.randn(B, cout, device=dev)
    res = torch.randn(B, D, S, S, cout, device=dev).to(odt) if residual else None
    out = torch.empty(B, D, S, S, cout, device=dev, dtype=odt)
    st = ops.conv_stats_buffer(B, cout, dev) if stats else None
    ops.conv(x, pc, out=out, stats=st, residual=res, chan_bias=cb)
    ref = F.conv3d(x.float().permute(0, 4, 1, 2, 3), w.to(h16).float(), bias, padding=1) + cb[:, :, None, None, None]
    ref = ref.permute(0, 2, 3, 4, 1)
    if res is not None:
        ref = ref + res.float()
    err = float((out.float() - ref).abs().max() / ref.abs().max())
    tol = 1e-5 if odt == torch.float32 else (8e-3 if h16 == torch.bfloat16 else 1e-3)
    msg = f"{prec} {cin}->{cout} B={B} {D}x{S}x{S} res={int(residual)} stats={int(stats)}: max rel err {err:.2e}"
    if stats:
        s = st.sum(dim=1)                                     # [B, Cout, 2]
        o = out.float().reshape(B, -1, cout)
        e1 = float((s[..., 0] - o.sum(1)).abs().max() / o.sum(1).abs().max())
        e2 = float((s[..., 1] - (o * o).sum(1)).abs().max() / (o * o).sum(1).abs().max())
        msg += f", stats err {e1:.1e} / {e2:.1e}"
        assert e1 < 1e-3 and e2 < 1e-4, msg
    e0, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        ops.conv(x, pc, out=out, stats=st, residual=res, chan_bias=cb)
    e1_.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1_) / reps * 1e3
    fl = 2.0 * B * D * S * S * cin * cout * 27
    print(msg + f";  {us:.1f} us  {fl / us / 1e6:.0f} TFLOP/s", flush=True)
    assert err < tol, msg


print("DSK_CONV_T2 =", os.environ.get("DSK_CONV_T2", "(default)"), flush=True)
case(1, 32, 64, 64, "fp16s32", False, False)
case(2, 32, 64, 64, "fp16s32", True, True)
case(1, 32, 64, 64, "bf16", True, True)
case(2, 32, 128, 64, "fp16", False, True, D=6)
case(1, 32, 64, 192, "fp16s32", True, False, D=5)
if not quick:
    case(2, 64, 64, 64, "bf16", False, False)
    case(8, 64, 64, 64, "bf16", False, False)
    case(8, 64, 64, 64, "bf16", True, True)
    case(8, 64, 64, 64, "fp16", False, False)
    case(8, 64, 64, 64, "fp16", False, True)         # conv1 of a full-resolution block in the fp16s32 mode: fp16 out + statistics
    case(8, 64, 64, 64, "fp16s32", False, True)
    case(8, 64, 64, 64, "fp16s32", False, False)
    case(8, 64, 64, 64, "fp16s32", True, True)
print("ok")
