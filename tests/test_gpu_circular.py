"""GPU parity tests of circular ('periodic') convolution (reference CircularConv2d/3d,
nets/commonlayers.py:918-1032, PUNetGConfig(convolution_type="circular")): the padding kernel, every convolution kernel
path (CUDA-core wrap; tcgen05 single-CTA / CTA-pair / sub-pixel up-conv / N=16 convout tile on the halo-padded copy),
data and weight gradients, and whole networks against goldens recorded from the LIVE reference."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
BF = torch.bfloat16


def relmax(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def to_cl(x, dtype=torch.float32):
    if x.ndim == 4:
        x = x.unsqueeze(2)
    return x.permute(0, 2, 3, 4, 1).contiguous().to(DEV).to(dtype)


def from_cl(y, ndim):
    y = y.float().cpu().permute(0, 4, 1, 2, 3)
    return y.squeeze(2) if ndim == 2 else y


def circ_conv(x, w, b, ndim, up2=False):
    """The reference's forward: F.pad(mode='circular') per axis, then a padding-free conv (commonlayers.py:955-972)."""
    if up2:
        x = F.interpolate(x, scale_factor=2, mode="nearest")
    p = w.shape[-1] // 2
    if p:
        x = F.pad(x, (p, p) * ndim, mode="circular")
    return (F.conv2d if ndim == 2 else F.conv3d)(x, w, b)


@pytest.fixture(scope="module")
def ops():
    from diffsci_b200 import ops as o
    return o


@pytest.mark.parametrize("ndim,shape", [(2, (2, 3, 5, 7)), (2, (3, 64, 6, 4)), (3, (2, 8, 3, 4, 5)), (3, (1, 1, 4, 4, 6))])
@pytest.mark.parametrize("dtype", [torch.float32, BF])
def test_pad_circular(ops, ndim, shape, dtype):
    torch.manual_seed(0)
    x = torch.randn(*shape).to(dtype)
    ref = F.pad(x.float(), (1, 1) * ndim, mode="circular")
    y = ops.pad_circular(to_cl(x.float(), dtype), ndim)
    assert torch.equal(from_cl(y, ndim), ref)


CASES = [
    # ndim, B, Cin, Cout, spatial, k, up2
    (2, 2, 8, 8, (12, 20), 3, False),        # CUDA-core kernel, vector gather
    (2, 1, 3, 5, (9, 11), 3, False),         # ragged channels
    (2, 2, 1, 16, (14, 14), 3, False),       # convin-like (the few-channel kernels hand over to the wrapping kernel)
    (3, 2, 16, 1, (4, 6, 8), 3, False),      # convout-like
    (3, 1, 16, 8, (4, 6, 8), 3, True),       # conv(nearest_up2(x)), wrap in the upsampled grid
    (2, 2, 24, 8, (5, 6), 1, False),         # 1x1: no padding at all
    (3, 1, 64, 64, (4, 16, 8), 3, False),    # tcgen05, one tile pair
    (3, 2, 64, 64, (5, 20, 12), 3, False),   # tcgen05, ragged tiles + wrap
    (3, 1, 128, 64, (4, 16, 16), 3, False),  # two K chunks
    (3, 1, 64, 64, (3, 32, 16), 3, False),   # merged depth taps (conv_tc2 T = 2: 34-row patches) on the padded copy
    (3, 1, 64, 128, (2, 16, 8), 3, False),   # N_TILE = 128
    (2, 3, 64, 64, (16, 8), 3, False),       # 2-D: planes are samples (no depth halo)
    (2, 5, 128, 128, (7, 7), 3, False),      # single-CTA kernel (odd number of w-tiles)
    (3, 2, 128, 64, (4, 8, 16), 3, True),    # sub-pixel up-conv on the padded low-resolution input
    (2, 3, 128, 64, (8, 16), 3, True),
    (3, 1, 64, 1, (6, 10, 12), 3, False),    # convout: in-plane taps as N (convout_tc) on the padded copy
    (2, 2, 128, 3, (9, 20), 3, False),       # convout, 3 output channels, ragged tiles
    (3, 2, 1, 64, (5, 16, 8), 3, False),     # convin on the tensor cores: the im2col-row builders wrap their gather
    (3, 1, 2, 64, (4, 9, 11), 3, False),     # ... ragged tiles
    (2, 3, 3, 128, (20, 12), 3, False),
]


@pytest.mark.parametrize("ndim,B,Cin,Cout,sp,k,up2", CASES)
@pytest.mark.parametrize("dtype", [torch.float32, BF])
def test_circular_conv_forward(ops, ndim, B, Cin, Cout, sp, k, up2, dtype):
    """fp32 mode: 2e-5 (accumulation order).  bf16 mode: one bf16 rounding of the stored output (6e-3 of the range)."""
    import diffsci_b200
    torch.manual_seed(5)
    tc = (dtype == BF and diffsci_b200.TC_CONV_ENABLED and k == 3 and Cin % 64 == 0 and (Cout % 64 == 0 or Cout <= 16))
    q = (lambda t: t.bfloat16().float()) if dtype == BF else (lambda t: t)
    x = q(torch.randn(B, Cin, *sp))
    w = torch.randn(Cout, Cin, *([k] * ndim)) / math.sqrt(Cin * k ** ndim)
    w = q(w) if tc else w
    b = torch.randn(Cout) * 0.1
    ref0 = circ_conv(x, w, b, ndim, up2)
    pc = ops.PackedConv(w.to(DEV), b.to(DEV), ndim, BF if tc else torch.float32, subpixel=bool(up2 and tc), circular=True)
    xc = to_cl(x, dtype)
    y0 = ops.conv(xc, pc, up2=up2)
    tol = 2e-5 if dtype == torch.float32 else 6e-3
    assert relmax(from_cl(y0, ndim), ref0) < tol, relmax(from_cl(y0, ndim), ref0)
    # must differ from the zero-padded convolution (the test would otherwise not see the wrap)
    if k == 3:
        zero = (F.conv2d if ndim == 2 else F.conv3d)(F.interpolate(x, scale_factor=2) if up2 else x, w, b, padding=1)
        assert relmax(zero, ref0) > 1e-2
    if Cout > 4:   # epilogue operands
        cb = torch.randn(B, Cout) * 0.3
        res = q(torch.randn_like(ref0))
        ref = ref0 + res + (0 if up2 else cb.view(B, Cout, *([1] * ndim)))
        y = ops.conv(xc, pc, chan_bias=None if up2 else cb.to(DEV), residual=to_cl(res, dtype), up2=up2)
        assert relmax(from_cl(y, ndim), ref) < tol
        if tc and ops.conv_stats_supported(tuple(xc.shape), xc.dtype, pc, up2=up2):
            st = ops.conv_stats_buffer(B, Cout, DEV)
            st.fill_(float("nan"))
            ys = ops.conv(xc, pc, chan_bias=None if up2 else cb.to(DEV), residual=to_cl(res, dtype), up2=up2, stats=st)
            assert torch.equal(ys, y)
            tot = st.double().sum(1)
            yf = y.double().flatten(1, 3)
            assert relmax(tot[..., 0], yf.sum(1)) < 2e-3 and relmax(tot[..., 1], (yf * yf).sum(1)) < 2e-3
    elif tc:       # convout: fp32 NC(D)HW output straight from the tensor-core tile
        yn = ops.conv(xc, pc, out_nchw=True)
        assert yn.dtype == torch.float32 and relmax(yn.cpu(), ref0) < 2e-5


@pytest.mark.parametrize("ndim,B,Cin,Cout,sp,k,up2", CASES)
@pytest.mark.parametrize("dtype", [torch.float32, BF])
def test_circular_conv_backward(ops, ndim, B, Cin, Cout, sp, k, up2, dtype):
    """dW (dsk_conv_wgrad with d->circular) and dX (circular conv of dY with the flipped weights) vs autograd of the
    reference formulation."""
    import diffsci_b200
    torch.manual_seed(6)
    q = (lambda t: t.bfloat16().float()) if dtype == BF else (lambda t: t)
    x = q(torch.randn(B, Cin, *sp)).requires_grad_(True)
    w = (torch.randn(Cout, Cin, *([k] * ndim)) / math.sqrt(Cin * k ** ndim)).requires_grad_(True)
    y = circ_conv(x, w, None, ndim, up2)
    dy = q(torch.randn_like(y))
    y.backward(dy)
    tc = dtype == BF and diffsci_b200.TC_CONV_ENABLED and k == 3 and Cin % 64 == 0 and Cout % 64 == 0
    wdt = BF if tc else torch.float32
    xc, dyc = to_cl(x.detach(), dtype), to_cl(dy, dtype)
    wd = w.detach().to(DEV)
    Bq, D, H, W, _ = dyc.shape
    gw = torch.full_like(wd, 7.0)
    if up2 and tc:      # as the training graph does it: materialised upsample, then the tensor-core weight gradient
        u = ops.upsample2x(xc, ndim)
        desc = ops.conv_desc(Bq, D, H, W, Cin, Cout, k, ndim, False, wdt, dtype, dtype, True)
        ws = torch.empty(max(ops.conv_wgrad_ws_bytes(desc), 1), dtype=torch.uint8, device=DEV)
        ops.conv_wgrad(desc, u, dyc, gw, ws)
    else:
        desc = ops.conv_desc(Bq, D, H, W, Cin, Cout, k, ndim, up2, wdt, dtype, dtype, True)
        ws = torch.empty(max(ops.conv_wgrad_ws_bytes(desc), 1), dtype=torch.uint8, device=DEV)
        ops.conv_wgrad(desc, xc, dyc, gw, ws)
    assert relmax(gw.cpu(), w.grad) < (2e-4 if tc else 2e-5), ("wgrad", relmax(gw.cpu(), w.grad))
    pd = ops.PackedConv(wd, None, ndim, BF if tc else torch.float32, dgrad=True, circular=True)
    du = ops.conv(dyc, pd)
    if up2:
        dx = torch.empty_like(xc)
        ops.upsample2x_bwd(du, dx, ndim)
    else:
        dx = du
    tol_x = 2e-5 if dtype == torch.float32 else 1.2e-2
    assert relmax(from_cl(dx, ndim), x.grad) < tol_x, ("dgrad", relmax(from_cl(dx, ndim), x.grad))


# ------------------------------------------------------------------------------------------------ networks
def build(g, precision="fp32"):
    import diffsci_b200 as d
    from oracle.nets_oracle import synth_state_dict
    net = (d.ADM(d.ADMConfig(**g["cfg"]), precision=precision) if g.get("kind") == "adm" else
           d.PUNetG(d.PUNetGConfig(**g["cfg"]), precision=precision))
    assert list(net.state_dict().keys()) == [k for k, _ in g["manifest"]]      # <name>.conv.weight keys, reference order
    net.load_state_dict(synth_state_dict(g["manifest"], g["seed"]))
    return net.to(DEV).eval()


@pytest.mark.parametrize("name", ["circ_punetg2d", "circ_punetg3d", "circ_adm2d"])
def test_circular_network_vs_live_reference(golden, name):
    import diffsci_b200 as d
    from oracle import karras_oracle as K
    from tests.test_oracle_vs_golden import oracle_net
    g = golden(name)
    net = build(g)
    with torch.no_grad():
        y = net(g["x"].to(DEV), g["t"].to(DEV)).cpu()
    e_ref = relmax(g["y"], g["y64"])
    assert relmax(y, g["y64"]) < max(3 * e_ref, 2e-5) and relmax(y, g["y"]) < max(4 * e_ref, 2e-5)
    with torch.no_grad():
        yb = build(g, "bf16")(g["x"].to(DEV), g["t"].to(DEV)).cpu()
    # bf16 storage of activations between layers.  Periodic nets are ~2x more rounding-sensitive than zero-padded ones: the
    # REFERENCE ITSELF under torch.autocast(bfloat16) is 1.8e-2 .. 2.8e-2 (L2) from fp64 on mc=8 circular nets
    # (oracle/bf16_budget_circular.py -> profiles/r1_bf16_budget_circular.txt).  Stated tolerance: L2 4e-2, max-rel 6e-2.
    assert relmax(yb, g["y64"]) < 6e-2 and rel_l2(yb, g["y64"]) < 4e-2, (relmax(yb, g["y64"]), rel_l2(yb, g["y64"]))
    # Heun sampling through the graph engine, budgeted against fp64 truth
    mod = d.KarrasModule(net, d.KarrasModuleConfig.from_edm())
    n, wn = g["nsteps"], g["white_noise"]
    truth = K.sample_from_white_noise(oracle_net(g, torch.float64), wn.double(), n, "heun", record_history=True)
    out = mod.propagate_white_noise(wn.to(DEV), nsteps=n, record_history=True).cpu()
    assert relmax(out, truth) <= 3.0 * relmax(g["heun_hist"], truth) + 5e-5
    # loss + gradients of the live reference
    net.train()
    mod.train()
    net.zero_grad()
    mod._injected_loss_noise = g["loss_noise"]
    L = mod.loss_fn(g["loss_x"].to(DEV), g["loss_sigma"].to(DEV))
    L.backward()
    assert abs(float(L.detach()) - float(g["loss_huber"])) < 5e-5 * abs(float(g["loss_huber"]))
    params = dict(net.named_parameters())
    for k, ref in g["loss_huber_grads"].items():
        e = relmax(params[k].grad.cpu(), ref)
        assert e < 3e-4, (k, e)


def test_circular_network_tensor_core_path():
    """PUNetG-3D mc=64 with circular convolutions in bf16 (every conv on tcgen05 through the halo-padded copy, fused norm
    statistics, sub-pixel up-convs) vs the fp64 oracle; forward and training gradients."""
    import types
    import diffsci_b200 as d
    from oracle import nets_oracle as N
    torch.manual_seed(9)
    kw = dict(dimension=3, model_channels=64, channel_expansion=[2], convolution_type="circular")
    net = d.PUNetG(d.PUNetGConfig(**kw), precision="bf16").to(DEV).eval()
    sd = {k: v.detach().cpu().double() for k, v in net.state_dict().items()}
    cfg = types.SimpleNamespace(**d.PUNetGConfig(**kw).export_description())
    x, t = torch.randn(2, 1, 8, 16, 16), torch.tensor([0.3, -0.8])
    ref = N.punetg_forward(sd, cfg, x.double(), t.double())
    with torch.no_grad():
        y = net(x.to(DEV), t.to(DEV)).cpu()
    cfg0 = types.SimpleNamespace(**dict(vars(cfg), convolution_type="default"))
    sd0 = {k.replace(".conv.weight", ".weight").replace(".conv.bias", ".bias") if not k.startswith(("downsamplers", "upsamplers"))
           else k.replace(".conv.conv.", ".conv."): v for k, v in sd.items()}
    ref0 = N.punetg_forward(sd0, cfg0, x.double(), t.double())
    assert relmax(ref0, ref) > 5e-2                                                    # the wrap matters on this input
    # the same weights with zero padding (the round-1 path) set the bf16 error scale of this network
    net0 = d.PUNetG(d.PUNetGConfig(**dict(kw, convolution_type="default")), precision="bf16").to(DEV).eval()
    net0.load_state_dict({k: v.float() for k, v in sd0.items()})
    with torch.no_grad():
        y0 = net0(x.to(DEV), t.to(DEV)).cpu()
    e, e0 = rel_l2(y, ref), rel_l2(y0, ref0)
    print(f"circular bf16 L2 {e:.3e} max {relmax(y, ref):.3e}; zero-padded bf16 L2 {e0:.3e} max {relmax(y0, ref0):.3e}")
    # the reference's own autocast-bf16 run of this network: L2 4.6e-2 circular vs 2.1e-2 zero-padded (same weights; periodic
    # nets are ~2x more rounding-sensitive, profiles/r1_bf16_budget_circular.txt).  Stated tolerance: L2 6e-2, max-rel 8e-2.
    assert e < 6e-2 and e0 < 3e-2 and relmax(y, ref) < 8e-2, (e, e0, relmax(y, ref))
    # training gradients: global L2 against fp64 autograd of the oracle.  fp32 mode pins the logic (wrap in forward, dgrad
    # and wgrad); in bf16 the periodic net is ~2x more rounding-sensitive than its zero-padded twin (fp32: 3.9e-5 vs 1.1e-5,
    # bf16: 1.2e-1 vs 5.9e-2 measured, tests/diag_circular_grads.py), so the bf16 gate is relative to the twin.
    dF = torch.randn(2, 1, 8, 16, 16)
    err = {}
    for ct, s_, c_ in (("circular", sd, cfg), ("default", sd0, cfg0)):
        sdg = {k: v.clone().requires_grad_(True) for k, v in s_.items()}
        N.punetg_forward(sdg, c_, x.double(), t.double()).backward(dF.double())
        for prec in ("fp32", "bf16"):
            n_ = d.PUNetG(d.PUNetGConfig(**dict(kw, convolution_type=ct)), precision=prec)
            n_.load_state_dict({k: v.float() for k, v in s_.items()})
            n_ = n_.to(DEV).train()
            n_(x.to(DEV), t.to(DEV)).backward(dF.to(DEV))
            num = den = 0.0
            for k, p in n_.named_parameters():
                num += float((p.grad.cpu().double() - sdg[k].grad).pow(2).sum())
                den += float(sdg[k].grad.pow(2).sum())
            err[ct, prec] = math.sqrt(num / den)
    print("gradient global-L2 errors:", {k: f"{v:.2e}" for k, v in err.items()})
    assert err["circular", "fp32"] < 2e-4 and err["default", "fp32"] < 1e-4
    assert err["default", "bf16"] < 8e-2 and err["circular", "bf16"] < 3 * err["default", "bf16"]


@pytest.mark.parametrize("ndim,shape", [(3, (2, 6, 10, 12, 64)), (2, (3, 1, 9, 16, 128)), (3, (1, 1, 4, 8, 64))])
def test_norm_apply_padded_and_prepadded_conv(ops, ndim, shape):
    """The norm's apply pass writing the halo-padded conv input directly (dsk_norm_apply_padded) == apply + dsk_pad_circular,
    bit for bit; a circular tcgen05 convolution on that pre-padded input (d->circular = 2) == the same convolution padding
    internally, bit for bit (including the fused statistics)."""
    torch.manual_seed(8)
    B, D, H, W, C = shape
    if ndim == 2:
        D = 1
    x = torch.randn(B, D, H, W, C, device=DEV).bfloat16()
    g, b_ = torch.randn(C, device=DEV), torch.randn(C, device=DEV)
    for mode in (0, 1):
        ws = ops.norm_ws(B, D * H * W, C, DEV)
        ref = ops.pad_circular(ops.norm_act(x, g, b_, C, mode, True, ws=ws), ndim)
        ws2 = ops.norm_ws(B, D * H * W, C, DEV)
        ops.norm_act(x, g, b_, C, mode, True, ws=ws2, table_only=True)
        out = torch.full_like(ref, float("nan"))
        ops.norm_apply_padded(x, ws2, out, ndim)
        assert torch.equal(out, ref)
    w = torch.randn(C, C, *([3] * ndim), device=DEV) / math.sqrt(C * 3 ** ndim)
    pc = ops.PackedConv(w, torch.randn(C, device=DEV), ndim, BF, circular=True)
    res = torch.randn(B, D, H, W, C, device=DEV).bfloat16()
    y_int = ops.conv(x, pc, residual=res)
    y_pre = ops.conv(ops.pad_circular(x, ndim), pc, residual=res, prepadded=True)
    assert torch.equal(y_int, y_pre)
    if ops.conv_stats_supported(tuple(x.shape), x.dtype, pc):
        s1, s2 = ops.conv_stats_buffer(B, C, DEV), ops.conv_stats_buffer(B, C, DEV)
        ops.conv(x, pc, residual=res, stats=s1)
        ops.conv(ops.pad_circular(x, ndim), pc, residual=res, stats=s2, prepadded=True)
        assert torch.equal(s1, s2)


def test_circular_adm_tensor_core_path():
    """ADM mc=64 with circular block convolutions in bf16 (tcgen05 through the halo-padded copy, sub-pixel up-convs, the plan's
    grown-on-demand pad workspace under graph capture) vs the fp64 oracle, next to its zero-padded twin."""
    import diffsci_b200 as d
    from oracle import nets_oracle as N
    from tests.test_oracle_vs_golden import cfg_for
    kw = dict(input_channels=3, output_channels=3, model_channels=64, channel_expansion=[2], number_resnet_attn_block=2)
    errs = {}
    for ct in ("circular", "default"):
        kwc = dict(kw, convolution_type=ct)
        net = d.ADM(d.ADMConfig(**kwc), precision="bf16")
        man = [(k, list(v.shape)) for k, v in net.state_dict().items()]
        sd = N.synth_state_dict(man, 777)
        net.load_state_dict(sd)
        net = net.to(DEV).eval()
        torch.manual_seed(3)
        x, t = torch.randn(2, 3, 32, 32), torch.tensor([0.3, -1.0])
        cfg = cfg_for("adm", kwc)
        ref = N.adm_forward({k: v.double() for k, v in sd.items()}, cfg, x.double(), t.double())
        with torch.no_grad():
            y = net(x.to(DEV), t.to(DEV)).cpu()
        errs[ct] = rel_l2(y, ref)
        mod = d.KarrasModule(net, d.KarrasModuleConfig.from_edm())
        s = mod.sample(2, [3, 32, 32], nsteps=3)                       # graph engine: capture + replay
        assert torch.isfinite(s).all()
    print(f"ADM bf16 L2 vs fp64: circular {errs['circular']:.3e}, zero-padded {errs['default']:.3e}")
    assert errs["default"] < 3e-2 and errs["circular"] < max(3 * errs["default"], 3e-2)
