"""The dominant kernel alone (tcgen05 conv C->C k3 at full resolution) for `ncu --set full`:
  ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 3 -c 2 -o gpurun_out/prof python tools/profile_conv.py
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from diffsci_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=2)
ap.add_argument("--size", type=int, default=64)
ap.add_argument("--cin", type=int, default=64)
ap.add_argument("--cout", type=int, default=64)
ap.add_argument("--reps", type=int, default=6)
ap.add_argument("--stats", action="store_true")
ap.add_argument("--residual", action="store_true")
ap.add_argument("--precision", default="bf16", choices=["bf16", "fp16", "fp16s32", "fp16x2", "fp32"])
ap.add_argument("--quantize", default="", choices=["", "bf16", "zero"],
                help="input values exactly representable in bf16 (few mantissa bits toggle; split lo halves are zero) or all zero: "
                     "separates data-dependent power throttling from pipeline effects")
ap.add_argument("--clocks", action="store_true", help="sample nvidia-smi SM clock / power while the timed loop runs")
a = ap.parse_args()
dev = torch.device("cuda:0")
torch.manual_seed(0)
w = torch.randn(a.cout, a.cin, 3, 3, 3, device=dev) * 0.02
wdt = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp16s32": torch.float16, "fp16x2": torch.float16, "fp32": ops.SPLIT}[a.precision]
adt = {"bf16": torch.bfloat16, "fp16": torch.float16}.get(a.precision, torch.float32)
pc = ops.PackedConv(w, torch.zeros(a.cout, device=dev), 3, wdt)
xs = [torch.randn(a.batch, a.size, a.size, a.size, a.cin, device=dev) for _ in range(3)]
if a.quantize == "bf16":
    xs = [x.bfloat16().float() for x in xs]
    w = w.bfloat16().float()
    pc = ops.PackedConv(w, torch.zeros(a.cout, device=dev), 3, wdt)
elif a.quantize == "zero":
    xs = [torch.zeros_like(x) for x in xs]
xs = [x.to(adt) for x in xs]
if a.precision == "fp16s32":
    xs = [x.half() for x in xs]                   # plain fp16 activations, fp32 output
elif adt == torch.float32:
    xs = [ops.split_f16(x) for x in xs]          # split-fp16 activations (hi | lo), fp32 output
out = torch.empty(a.batch, a.size, a.size, a.size, a.cout, device=dev, dtype=adt)
st = ops.conv_stats_buffer(a.batch, a.cout, dev) if a.stats else None
res = torch.randn_like(out) if a.residual else None
for i in range(3):
    ops.conv(xs[i % 3], pc, out=out, stats=st, residual=res)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
mon = None
if a.clocks:
    import subprocess
    mon = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "50"],
                           stdout=subprocess.PIPE, text=True)
e0.record()
for i in range(a.reps):
    ops.conv(xs[i % 3], pc, out=out, stats=st, residual=res)
e1.record()
torch.cuda.synchronize()
if mon is not None:
    mon.terminate()
    rows = [r.split(",") for r in mon.stdout.read().strip().splitlines() if "," in r]
    if rows:
        clk = sorted(float(r[0]) for r in rows)
        print(f"  clocks under load: median {clk[len(clk) // 2]:.0f} MHz, min {clk[0]:.0f}, max power {max(float(r[1]) for r in rows):.0f} W "
              f"({len(rows)} samples)")
ms = e0.elapsed_time(e1) / a.reps
fl = 2.0 * a.batch * a.size ** 3 * a.cin * a.cout * 27
print(f"conv3d {a.precision} {a.cin}->{a.cout} @ {a.size}^3 B={a.batch} stats={a.stats} residual={a.residual}: {ms * 1e3:.1f} us  {fl / ms / 1e9:.1f} TFLOP/s")
