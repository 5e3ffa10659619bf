"""GPU parity tests of the tensor-core precision modes (north_star tolerance contract: 1e-5 fp32 / 1e-3 16-bit):

* fp16 operands / storage on every tcgen05 kernel (same kernels as bf16, the format is a run-time flag);
* split-fp16 operands (hi + lo, 3 tcgen05 MMAs per k-step, fp32 accumulate): the tensor-core fp32-parity mode -- kernels
  against fp64 ATen on the SAME fp32 inputs, networks against the CPU oracle's fp32 evaluation (the reference's arithmetic,
  nets/punetg.py:389-416, commonlayers.py:809-836).
Stated tolerances per test; oracle/split_budget.py is the CPU emulation these numbers were budgeted with."""
import math
import types

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def relmax(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def to_cl(x):
    if x.ndim == 4:
        x = x.unsqueeze(2)
    return x.permute(0, 2, 3, 4, 1).contiguous().to(DEV)


def from_cl(y, ndim):
    y = y.float().cpu().permute(0, 4, 1, 2, 3)
    return y.squeeze(2) if ndim == 2 else y


@pytest.fixture(scope="module")
def ops():
    from diffsci_b200 import ops as o
    return o


def test_split_f16_roundtrip(ops):
    """dsk_split_f16: hi + 2^-11 lo reproduces an fp32 tensor to 2^-22 relative (lo is stored times 2^11, which keeps it a
    normal fp16 number for every |v| > 2^-14 * 2^11 * 2^-11... i.e. down to fp16's own subnormal threshold on v)."""
    torch.manual_seed(0)
    x = (torch.randn(1000, 64) * 3).to(DEV)
    s = ops.split_f16(x)
    assert s.shape == (1000, 128) and s.dtype == torch.float16
    rec = s[:, :64].double() + s[:, 64:].double() / 2048
    assert torch.equal(s[:, :64], x.half())
    err = ((rec - x.double()).abs() / (2.0 ** -22 * x.double().abs() + 2.0 ** -35)).max()
    assert float(err) <= 1.0, float(err)


SPLIT_CONV_CASES = [
    # ndim, B, Cin, Cout, spatial, up2
    (3, 2, 64, 64, (8, 16, 16), False),      # CTA-pair kernel, fused statistics
    (3, 1, 128, 64, (4, 16, 8), False),      # single-CTA kernel (one w-tile)
    (3, 1, 64, 128, (6, 16, 16), False),     # N tile 128, ragged plane count
    (2, 3, 128, 128, (16, 16), False),
    (3, 1, 128, 64, (4, 8, 8), True),        # sub-pixel UpSampler form
    (2, 2, 64, 64, (8, 16), True),
]


@pytest.mark.parametrize("ndim,B,Cin,Cout,sp,up2", SPLIT_CONV_CASES)
@pytest.mark.parametrize("mode", ["split3", "split2", "fp16"])
def test_conv_tensor_core_formats(ops, ndim, B, Cin, Cout, sp, up2, mode):
    """k = 3 convolution on tcgen05 with split operands (3 MMAs: fp32-class; 2 MMAs: activations split, fp16 weights) and with
    plain fp16 operands, + bias + per-sample channel bias + residual, against fp64 ATen on the same fp32 inputs."""
    torch.manual_seed(3)
    x = torch.randn(B, Cin, *sp)
    w = torch.randn(Cout, Cin, *([3] * ndim)) / math.sqrt(Cin * 3 ** ndim)
    b = torch.randn(Cout) * 0.1
    cb = torch.randn(B, Cout) * 0.3
    xr = F.interpolate(x, scale_factor=2, mode="nearest") if up2 else x
    conv = F.conv2d if ndim == 2 else F.conv3d
    ref = conv(xr.double(), w.double(), b.double(), padding=1)
    res = torch.randn(ref.shape)
    ref = ref + cb.double().view(B, Cout, *([1] * ndim)) + res.double()
    wdt = {"split3": ops.SPLIT, "split2": torch.float16, "fp16": torch.float16}[mode]
    pc = ops.PackedConv(w.to(DEV), b.to(DEV), ndim, wdt, subpixel=up2)
    xcl = to_cl(x)
    if mode == "fp16":
        y = ops.conv(xcl.half(), pc, chan_bias=cb.to(DEV), residual=to_cl(res).half(), up2=up2)
        assert y.dtype == torch.float16
        tol = 3e-3               # fp16 operands (2^-11) and one fp16 rounding of the stored result
    else:
        y = ops.conv(ops.split_f16(xcl), pc, chan_bias=cb.to(DEV), residual=to_cl(res), up2=up2)
        assert y.dtype == torch.float32
        tol = 3e-6 if mode == "split3" else 1.5e-3
    err = relmax(from_cl(y, ndim), ref)
    print(f"{mode} conv {ndim}-D {Cin}->{Cout} {sp} up2={up2}: max-rel {err:.2e}")
    assert err < tol, err


def test_conv_split_fused_statistics(ops):
    """The fused norm statistics of a split convolution (fp32 output): norm_act(conv_stats=...) writing a SPLIT tensor equals
    GroupNorm + SiLU of the fp32 conv output to fp32 accuracy."""
    torch.manual_seed(4)
    B, C, sp = 2, 64, (8, 16, 16)
    x = torch.randn(B, C, *sp)
    w = torch.randn(C, C, 3, 3, 3) / math.sqrt(C * 27)
    g, be = 1 + 0.1 * torch.randn(C), 0.1 * torch.randn(C)
    pc = ops.PackedConv(w.to(DEV), None, 3, ops.SPLIT)
    xs = ops.split_f16(to_cl(x))
    assert ops.conv_stats_supported(xs.shape, xs.dtype, pc, out_dtype=torch.float32)
    st = ops.conv_stats_buffer(B, C, DEV)
    y = ops.conv(xs, pc, stats=st)
    n = torch.empty(y.shape[:-1] + (2 * C,), dtype=torch.float16, device=DEV)
    ops.norm_act(y, g.to(DEV), be.to(DEV), C, 0, True, out=n, conv_stats=st)
    rec = (n[..., :C].double() + n[..., C:].double() / 2048).float()
    ref = F.silu(F.group_norm(F.conv3d(x.double(), w.double(), padding=1), C, g.double(), be.double(), 1e-5))
    err = relmax(from_cl(rec, 3), ref)
    assert err < 5e-6, err


@pytest.mark.parametrize("cout", [1, 3])
def test_convout_split(ops, cout):
    """Last convolution (C -> 1..3) with split operands, NC(D)HW fp32 output (convout_tc.cu)."""
    torch.manual_seed(5)
    x = torch.randn(2, 64, 6, 16, 16)
    w = torch.randn(cout, 64, 3, 3, 3) / math.sqrt(64 * 27)
    b = torch.randn(cout) * 0.1
    ref = F.conv3d(x.double(), w.double(), b.double(), padding=1)
    pc = ops.PackedConv(w.to(DEV), b.to(DEV), 3, ops.SPLIT)
    y = ops.conv(ops.split_f16(to_cl(x)), pc, out_nchw=True)
    err = relmax(y.cpu(), ref)
    assert err < 3e-6, err
    pc16 = ops.PackedConv(w.to(DEV), b.to(DEV), 3, torch.float16)
    y16 = ops.conv(to_cl(x).half(), pc16, out_nchw=True)
    assert relmax(y16.cpu(), ref) < 2e-3


@pytest.mark.parametrize("M,N,K,batch", [(256, 192, 128, 1), (200, 64, 64, 3), (512, 512, 256, 2)])
def test_gemm_split(ops, M, N, K, batch):
    """dsk_gemm_tc with split operands (K-major A and B): fp32-class A B^T + bias + fp32 residual."""
    torch.manual_seed(6)
    A = torch.randn(batch, M, K)
    Bm = torch.randn(batch, N, K) / math.sqrt(K)
    bias = torch.randn(N)
    res = torch.randn(batch, M, N)
    ref = A.double() @ Bm.double().transpose(1, 2) + bias.double() + res.double()
    As, Bs = ops.split_f16(A.to(DEV)), ops.split_f16(Bm.to(DEV))
    out = torch.empty(batch, M, N, device=DEV)
    ops.gemm_split_tc(As, Bs, out, M=M, N=N, K=K, lda=2 * K, ldb=2 * K, ldc=N, a_lo=K, b_lo=K, bias=bias.to(DEV),
                      residual=res.to(DEV), batch=batch, strideA=M * 2 * K, strideB=N * 2 * K, strideC=M * N)
    err = relmax(out.cpu(), ref)
    assert err < 2e-6, err


def test_attention_split_vs_oracle(ops):
    """Single-head self-attention on fp32 tokens through the split tensor-core GEMMs vs the CPU oracle (fp64)."""
    from oracle import nets_oracle as N
    torch.manual_seed(7)
    B, C, sp = 2, 128, (4, 4, 8)
    Lq = sp[0] * sp[1] * sp[2]
    x = torch.randn(B, C, *sp)
    sd = {"a.mhattn.in_proj_weight": torch.randn(3 * C, C) / math.sqrt(C), "a.mhattn.in_proj_bias": torch.randn(3 * C) * 0.1,
          "a.mhattn.out_proj.weight": torch.randn(C, C) / math.sqrt(C), "a.mhattn.out_proj.bias": torch.randn(C) * 0.1}
    for residual in (False, True):
        ref = N.mha_self_attention(x.double(), {k: v.double() for k, v in sd.items()}, "a.", residual)
        tok = to_cl(x).view(B, Lq, C)
        wi = ops.PackedLinear(sd["a.mhattn.in_proj_weight"].to(DEV), ops.SPLIT)
        wo = ops.PackedLinear(sd["a.mhattn.out_proj.weight"].to(DEV), ops.SPLIT)
        out = torch.empty(B, Lq, C, device=DEV)
        ops.self_attention_split(tok, wi, sd["a.mhattn.in_proj_bias"].to(DEV), wo, sd["a.mhattn.out_proj.bias"].to(DEV),
                                 ops.attention_split_buffers(B, Lq, C, DEV), out, residual)
        got = out.cpu().view(B, *sp, C).permute(0, 4, 1, 2, 3)
        err = relmax(got, ref)
        assert err < 5e-6, (residual, err)


def test_attention_fp16_vs_oracle(ops):
    """The 16-bit tensor-core attention (softmax(QK^T) written once) with fp16 operands."""
    from oracle import nets_oracle as N
    torch.manual_seed(8)
    B, C, sp = 2, 128, (4, 4, 8)
    Lq = sp[0] * sp[1] * sp[2]
    x = torch.randn(B, C, *sp)
    sd = {"a.mhattn.in_proj_weight": torch.randn(3 * C, C) / math.sqrt(C), "a.mhattn.in_proj_bias": torch.randn(3 * C) * 0.1,
          "a.mhattn.out_proj.weight": torch.randn(C, C) / math.sqrt(C), "a.mhattn.out_proj.bias": torch.randn(C) * 0.1}
    ref = N.mha_self_attention(x, sd, "a.", True)
    errs = {}
    for dt in (torch.float16, torch.bfloat16):
        tok = to_cl(x).view(B, Lq, C).to(dt)
        wi = ops.PackedLinear(sd["a.mhattn.in_proj_weight"].to(DEV), dt)
        wo = ops.PackedLinear(sd["a.mhattn.out_proj.weight"].to(DEV), dt)
        out = torch.empty(B, Lq, C, device=DEV, dtype=dt)
        ops.self_attention_tc(tok, wi, sd["a.mhattn.in_proj_bias"].to(DEV), wo, sd["a.mhattn.out_proj.bias"].to(DEV),
                              ops.attention_tc_buffers(B, Lq, C, DEV, dt), out, True)
        errs[dt] = relmax(out.float().cpu().view(B, *sp, C).permute(0, 4, 1, 2, 3), ref)
    print(f"attention vs oracle: fp16 {errs[torch.float16]:.2e}, bf16 {errs[torch.bfloat16]:.2e}")
    assert errs[torch.float16] < 3e-3 and errs[torch.bfloat16] < 2e-2
    assert errs[torch.float16] < errs[torch.bfloat16]


@pytest.mark.parametrize("mode", [0, 1])
def test_norm_16bit_formats_and_split_output(ops, mode):
    """norm + SiLU: fp16 in/out (incl. the single-kernel slab path) and fp32 -> split output."""
    from oracle import nets_oracle as N
    torch.manual_seed(9)
    for B, C, sp in [(2, 64, (8, 8, 8)), (160, 32, (7, 7))]:
        x = torch.randn(B, C, *sp) * 2 + 1.0
        g, b = 1 + 0.1 * torch.randn(C), 0.1 * torch.randn(C)
        xcl = x.reshape(B, C, -1).permute(0, 2, 1).contiguous().to(DEV)
        back = lambda y: y.double().cpu().permute(0, 2, 1).reshape(x.shape)  # noqa: E731
        xh = x.half().float()
        refh = F.silu(F.group_norm(xh, C, g, b, 1e-5) if mode == 0 else N.group_rms_norm(xh, C, g, b))
        yh = ops.norm_act(xcl.half(), g.to(DEV), b.to(DEV), C, mode, True)
        assert yh.dtype == torch.float16 and relmax(back(yh), refh) < 1.5e-3
        ref = F.silu(F.group_norm(x.double(), C, g.double(), b.double(), 1e-5) if mode == 0
                     else N.group_rms_norm(x.double(), C, g.double(), b.double()))
        ys = torch.empty(B, xcl.shape[1], 2 * C, dtype=torch.float16, device=DEV)
        ops.norm_act(xcl, g.to(DEV), b.to(DEV), C, mode, True, out=ys)
        rec = ys[..., :C].double() + ys[..., C:].double() / 2048
        assert relmax(back(rec), ref) < 3e-6


NET_CASES = [
    ("punetg3d", dict(dimension=3, model_channels=64), (1, 1, 32, 32, 32)),
    ("punetg2d", dict(dimension=2, model_channels=64), (2, 1, 64, 64)),
    ("punetg2d_mc128", dict(dimension=2, model_channels=128), (2, 1, 28, 28)),
]


@pytest.mark.parametrize("name,kw,shape", NET_CASES)
def test_network_precision_modes_vs_oracle(name, kw, shape):
    """One evaluation of the default-width networks in every precision mode vs the CPU oracle (fp32 = the reference's
    arithmetic, nets/punetg.py:389-416) -- the measured error column of DESIGN.md's 1 / 2 / 3-MMA table.
    Stated tolerances (max-rel of the network output):
      * fp32 modes (tensor-core split-fp16 x3, and CUDA-core FFMA): budgeted against fp64 truth -- a random-weight network
        amplifies fp32 rounding by a factor that depends on the network (the ORACLE's own fp32 result is 1e-6 ... 1e-4 from its
        fp64 result here), so the bound is 3x the oracle's own fp32-vs-fp64 distance + 5e-6;
      * fp16x2 (activations hi + lo, fp16 weights, 2 MMAs): 1e-3 vs the oracle's fp32 -- north_star's 16-bit tolerance;
      * fp16x2m (activations split only in the >= 128-channel layers) and fp16s32 (plain fp16 operands, fp32 storage): 1.5e-3 --
        bench.py selects either only for a network on which it measures <= 8e-4;  fp16 (1 MMA, fp16 storage): 5e-3;  bf16: 4e-2 (measured 1.8e-2 ... 3.0e-2: which way individual bf16 roundings fall depends on the summation order of the epilogues)."""
    import diffsci_b200 as d
    from oracle import nets_oracle as N
    torch.manual_seed(0)
    cfg = d.PUNetGConfig(**kw)
    net = d.PUNetG(cfg, precision="fp32").to(DEV).eval()
    sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    ocfg = types.SimpleNamespace(**cfg.export_description())
    x = torch.randn(*shape)
    t = torch.tensor([0.3, -0.8][:shape[0]])
    ref = N.punetg_forward(sd, ocfg, x, t)
    ref64 = N.punetg_forward({k: v.double() for k, v in sd.items()}, ocfg, x.double(), t.double())
    d0 = relmax(ref, ref64)
    tol = {"fp16x2": 1e-3, "fp16x2m": 1.5e-3, "fp16s32": 1.5e-3, "fp16": 5e-3, "bf16": 4e-2}
    got = {}
    for prec in ("fp32", "fp32_ffma", "fp16x2", "fp16x2m", "fp16s32", "fp16", "bf16"):
        net.precision = prec
        with torch.no_grad():
            y = net(x.to(DEV), t.to(DEV)).cpu()
        got[prec] = (relmax(y, ref), rel_l2(y, ref), relmax(y, ref64))
    print(name, f"oracle fp32 vs fp64 {d0:.2e};", {k: f"{v[0]:.2e}/{v[1]:.2e} (vs fp64 {v[2]:.2e})" for k, v in got.items()})
    for prec in ("fp32", "fp32_ffma"):
        assert got[prec][2] < 3 * d0 + 5e-6, (prec, got[prec], d0)
    for prec, t_ in tol.items():
        assert got[prec][0] < t_, (prec, got[prec])
    assert got["fp16"][0] < got["bf16"][0]


def test_sampling_fp16_graph_engine():
    """Heun sampling through the graph engine in fp16 and in the tensor-core fp32 mode vs the FFMA fp32 mode: same x_T, 4 steps."""
    import diffsci_b200 as d
    torch.manual_seed(0)
    net = d.PUNetG(d.PUNetGConfig(dimension=3, model_channels=64), precision="fp32_ffma").to(DEV).eval()
    mod = d.KarrasModule(net, d.KarrasModuleConfig.from_edm())
    wn = torch.randn(2, 1, 16, 16, 16)
    ref = mod.propagate_white_noise(wn.to(DEV), nsteps=4).cpu()
    net.precision = "fp32"
    a = mod.propagate_white_noise(wn.to(DEV), nsteps=4).cpu()
    net.precision = "fp16"
    b = mod.propagate_white_noise(wn.to(DEV), nsteps=4).cpu()
    print(f"Heun-4 vs FFMA fp32: tensor-core fp32 {relmax(a, ref):.2e}, fp16 {relmax(b, ref):.2e}")
    assert relmax(a, ref) < 1e-4 and relmax(b, ref) < 3e-2
