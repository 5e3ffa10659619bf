"""GPU parity tests of the backward (training) path: every K2 entry point of the C ABI against ATen autograd on the CPU
(the library the reference's loss.backward() dispatches to), then whole-network parameter gradients of the native
PUNetG / ADM against autograd through the CPU oracle and against gradients recorded from the live reference."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
BF = torch.bfloat16


def relmax(a, b):
    return float((a.double().cpu() - b.double().cpu()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def rel_l2(a, b):
    return float((a.double().cpu() - b.double().cpu()).norm() / b.double().norm().clamp_min(1e-30))


def to_cl(x, dtype=torch.float32):
    if x.ndim == 4:
        x = x.unsqueeze(2)
    return x.permute(0, 2, 3, 4, 1).contiguous().to(DEV).to(dtype)


def from_cl(y, ndim):
    y = y.float().cpu().permute(0, 4, 1, 2, 3)
    return y.squeeze(2) if ndim == 2 else y


@pytest.fixture(scope="module")
def ops():
    from diffsci_b200 import ops as o
    return o


# ------------------------------------------------------------------------------------------------ GEMM variants
@pytest.mark.parametrize("transA,transB", [(False, False), (False, True), (True, False), (True, True)])
def test_gemm_ex(ops, transA, transB):
    torch.manual_seed(0)
    batch, M, N, K = 3, 70, 45, 133
    A = torch.randn(batch, K, M) if transA else torch.randn(batch, M, K)
    Bm = torch.randn(batch, N, K) if transB else torch.randn(batch, K, N)
    C0 = torch.randn(batch, M, N)
    bias = torch.randn(N)
    opA = A.transpose(1, 2) if transA else A
    opB = Bm.transpose(1, 2) if transB else Bm
    ref = 0.7 * (opA.double() @ opB.double()) + bias.double() + 0.5 * C0.double()
    out = C0.clone().to(DEV)
    ops.gemm_ex(A.to(DEV), Bm.to(DEV), out, M=M, N=N, K=K, lda=A.shape[2], ldb=Bm.shape[2], ldc=N, bias=bias.to(DEV),
                transA=transA, transB=transB, alpha=0.7, beta=0.5, batch=batch, strideA=A[0].numel(), strideB=Bm[0].numel(),
                strideC=M * N)
    assert relmax(out, ref) < 2e-6


# ------------------------------------------------------------------------------------------------ convolution backward
CONV_CASES = [
    # ndim, B, Cin, Cout, spatial, k, up2
    (2, 2, 8, 8, (12, 20), 3, False),
    (2, 2, 1, 16, (14, 14), 3, False),      # convin-like
    (2, 2, 16, 1, (7, 7), 3, False),        # convout-like, odd size
    (2, 1, 3, 5, (9, 11), 3, False),        # ragged channels
    (3, 2, 8, 16, (6, 8, 10), 3, False),
    (3, 1, 16, 8, (4, 6, 8), 3, True),      # conv(nearest_up2(x))
    (2, 2, 8, 12, (8, 8), 3, True),
    (2, 2, 24, 8, (5, 6), 1, False),        # 1x1 (ADM residual conv)
    (2, 2, 24, 8, (5, 6), 1, True),
    (3, 1, 64, 64, (8, 8, 8), 3, False),
    (2, 3, 64, 128, (16, 24), 3, False),
    (2, 5, 128, 64, (7, 7), 3, False),      # tiles mostly out of bounds (MNIST bottom level)
    (3, 2, 64, 128, (6, 10, 12), 3, False),
    (3, 1, 128, 128, (16, 16, 16), 3, False),
    (2, 2, 1, 64, (14, 14), 3, False),      # few-channel weight-gradient kernel (wide side 32 / 64 / 128)
    (3, 1, 1, 64, (6, 8, 10), 3, False),
    (2, 2, 64, 1, (7, 9), 3, False),
    (3, 2, 128, 3, (4, 6, 8), 3, False),
    (2, 2, 3, 128, (9, 11), 3, False),
    (3, 1, 32, 2, (5, 6, 7), 3, False),
]


@pytest.mark.parametrize("ndim,B,Cin,Cout,sp,k,up2", CONV_CASES)
@pytest.mark.parametrize("dtype", [torch.float32, BF])
def test_conv_backward(ops, ndim, B, Cin, Cout, sp, k, up2, dtype):
    """dX via dsk_conv_fwd on dgrad-packed weights (+ upsample backward), dW via dsk_conv_wgrad, db via dsk_channel_sum."""
    torch.manual_seed(2)
    x = torch.randn(B, Cin, *sp).to(dtype).float().requires_grad_(True)
    w = (torch.randn(Cout, Cin, *([k] * ndim)) / math.sqrt(Cin * k ** ndim)).requires_grad_(True)
    b = torch.zeros(Cout, requires_grad=True)
    xr = F.interpolate(x, scale_factor=2, mode="nearest") if up2 else x
    y = (F.conv2d if ndim == 2 else F.conv3d)(xr, w, b, padding=k // 2)
    dy = torch.randn_like(y).to(dtype).float()
    wq = w if dtype == torch.float32 else w       # weights stay fp32 on the CUDA-core path
    y.backward(dy)

    # eligibility as the training graph decides it
    import diffsci_b200
    tc = dtype == BF and diffsci_b200.TC_CONV_ENABLED and k == 3 and Cin % 64 == 0 and Cout % 64 == 0
    wdt = BF if tc else torch.float32
    xc, dyc = to_cl(x.detach(), dtype), to_cl(dy, dtype)
    wd = wq.detach().to(DEV)
    Bq, D, H, W, _ = dyc.shape
    desc = ops.conv_desc(Bq, D, H, W, Cin, Cout, k, ndim, up2, wdt, dtype, dtype)
    ws = torch.empty(max(ops.conv_wgrad_ws_bytes(desc), 1), dtype=torch.uint8, device=DEV)
    gw = torch.full_like(wd, 7.0)
    ops.conv_wgrad(desc, xc, dyc, gw, ws)
    tol_w = 2e-4 if tc else 2e-5     # bf16 products are exact in fp32; only the summation order differs (fp32 in TMEM / registers)
    assert relmax(gw, w.grad) < tol_w, ("wgrad", relmax(gw, w.grad))
    gw2 = gw.clone()
    ops.conv_wgrad(desc, xc, dyc, gw2, ws, accumulate=True)
    assert relmax(gw2, 2 * w.grad) < tol_w

    gb = torch.empty(Cout, device=DEV)
    bws = torch.empty(ops.bwd_ws_bytes(Bq, D * H * W, Cout), dtype=torch.uint8, device=DEV)
    ops.channel_sum(dyc, gb, bws, False)
    assert relmax(gb, b.grad) < 1e-5
    gbs = torch.empty(Bq, Cout, device=DEV)
    ops.channel_sum(dyc, gbs, bws, True)
    assert relmax(gbs, dy.flatten(2).sum(-1)) < 1e-5

    tcd = dtype == BF and diffsci_b200.TC_CONV_ENABLED and k == 3 and Cin % 64 == 0 and Cout % 64 == 0
    pd = ops.PackedConv(wd, None, ndim, BF if tcd else torch.float32, dgrad=True)
    du = ops.conv(dyc, pd)
    if up2:
        dx = torch.empty_like(xc)
        ops.upsample2x_bwd(du, dx, ndim)
    else:
        dx = du
    tol_x = 2e-5 if dtype == torch.float32 else 1.2e-2     # bf16: one rounding of the stored result (+ bf16 weights on tcgen05)
    assert relmax(from_cl(dx, ndim), x.grad) < tol_x, ("dgrad", relmax(from_cl(dx, ndim), x.grad))
    if not up2:   # accumulation through the residual operand
        prev = torch.randn_like(x.grad).to(dtype).float()
        acc = to_cl(prev, dtype)
        ops.conv(dyc, pd, out=acc, residual=acc)
        assert relmax(from_cl(acc, ndim), x.grad + prev) < 2 * tol_x


# ------------------------------------------------------------------------------------------------ norm backward
def _ref_norm(x, G, w, b, mode, silu, fsc, fsh):
    from oracle.nets_oracle import group_rms_norm
    y = F.group_norm(x, G, w, b, 1e-5) if mode == 0 else group_rms_norm(x, G, w, b)
    if fsc is not None:
        shp = fsc.shape + (1,) * (x.ndim - 2)
        y = y * fsc.view(shp) + fsh.view(shp)
    return F.silu(y) if silu else y


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("group_all", [False, True])           # G == C (PUNetG) / G == 1 (ADM)
@pytest.mark.parametrize("film,silu,affine", [(False, True, True), (True, True, True), (False, False, True), (False, True, False)])
@pytest.mark.parametrize("dtype", [torch.float32, BF])
def test_norm_act_backward(ops, mode, group_all, film, silu, affine, dtype):
    torch.manual_seed(3)
    B, C, sp = 3, 16, (6, 10)
    G = 1 if group_all else C
    x = (torch.randn(B, C, *sp) * 1.5 + 0.3).to(dtype).float().requires_grad_(True)
    w = (1 + 0.2 * torch.randn(C)).requires_grad_(True) if affine else None
    b = (0.1 * torch.randn(C)).requires_grad_(True) if affine else None
    fsc = (1 + 0.3 * torch.randn(B, C)).requires_grad_(True) if film else None
    fsh = (0.2 * torch.randn(B, C)).requires_grad_(True) if film else None
    y = _ref_norm(x, G, w, b, mode, silu, fsc, fsh)
    dy = torch.randn_like(y).to(dtype).float()
    y.backward(dy)
    dres = torch.randn_like(x).to(dtype).float()

    xc, dyc, rc = to_cl(x.detach(), dtype), to_cl(dy, dtype), to_cl(dres, dtype)
    S = sp[0] * sp[1]
    fws = torch.empty(int(ops.lib.dsk_norm_ws_bytes(B, S, C)), dtype=torch.uint8, device=DEV)
    wd = w.detach().to(DEV) if affine else None
    bd = b.detach().to(DEV) if affine else None
    fs = fsc.detach().to(DEV) if film else None
    fh = fsh.detach().to(DEV) if film else None
    yc = ops.norm_act(xc, wd, bd, G, mode, silu, film_scale=fs, film_shift=fh, ws=fws)
    tol = 3e-5 if dtype == torch.float32 else 1.5e-2
    assert relmax(from_cl(yc, 2), y.detach()) < tol
    ws = torch.empty(ops.bwd_ws_bytes(B, S, C), dtype=torch.uint8, device=DEV)
    dg = torch.empty(C, device=DEV) if affine else None
    db = torch.empty(C, device=DEV) if affine else None
    dfs = torch.empty(B, C, device=DEV) if film else None
    dfh = torch.empty(B, C, device=DEV) if film else None
    dx = torch.empty_like(xc)
    ops.norm_act_bwd(xc, dyc, dx, wd, bd, G, mode, silu, fws, ws, dgamma=dg, dbeta=db, film_scale=fs, dfilm_scale=dfs, dfilm_shift=dfh)
    ptol = 5e-5 if dtype == torch.float32 else 5e-5     # parameter sums are fp32/fp64 on exact inputs in both modes
    assert relmax(from_cl(dx, 2), x.grad) < tol, ("dx", relmax(from_cl(dx, 2), x.grad))
    if affine:
        assert relmax(dg, w.grad) < ptol and relmax(db, b.grad) < ptol
    if film:
        assert relmax(dfs, fsc.grad) < ptol and relmax(dfh, fsh.grad) < ptol
    # accumulation operand, in place
    ops.norm_act_bwd(xc, dyc, rc, wd, bd, G, mode, silu, fws, ws, dgamma=dg, dbeta=db, dres=rc, film_scale=fs, dfilm_scale=dfs,
                     dfilm_shift=dfh)
    assert relmax(from_cl(rc, 2), x.grad + dres) < 2 * tol


# ------------------------------------------------------------------------------------------------ pooling / upsampling / small ops
@pytest.mark.parametrize("ndim,sp", [(2, (8, 12)), (2, (7, 9)), (3, (4, 6, 8))])
@pytest.mark.parametrize("is_max", [True, False])
@pytest.mark.parametrize("dtype", [torch.float32, BF])
def test_pool_backward(ops, ndim, sp, is_max, dtype):
    torch.manual_seed(4)
    B, C = 2, 8
    x = torch.randn(B, C, *sp).to(dtype).float()
    if is_max:   # ties (likely in bf16): the first maximum in scan order takes the gradient, as ATen
        x[:, :, ::2] = x[:, :, ::2].round()
    x.requires_grad_(True)
    pool = {(2, True): F.max_pool2d, (2, False): F.avg_pool2d, (3, True): F.max_pool3d, (3, False): F.avg_pool3d}[(ndim, is_max)]
    y = pool(x, 2)
    dy = torch.randn_like(y).to(dtype).float()
    y.backward(dy)
    xc, dyc = to_cl(x.detach(), dtype), to_cl(dy, dtype)
    dx = torch.full_like(xc, 5.0)
    ops.pool2x_bwd(xc, dyc, dx, ndim, is_max)
    tol = 1e-6 if dtype == torch.float32 else 5e-3
    assert relmax(from_cl(dx, ndim), x.grad) < tol
    prev = to_cl(torch.randn_like(x.grad), dtype)
    want = x.grad + from_cl(prev, ndim)
    ops.pool2x_bwd(xc, dyc, prev, ndim, is_max, dres=prev)
    assert relmax(from_cl(prev, ndim), want) < 2e-2 if dtype == BF else relmax(from_cl(prev, ndim), want) < 1e-6


@pytest.mark.parametrize("ndim,sp", [(2, (5, 6)), (3, (3, 4, 5))])
def test_upsample_backward_and_small_ops(ops, ndim, sp):
    torch.manual_seed(5)
    B, C = 2, 8
    x = torch.randn(B, C, *sp, requires_grad=True)
    y = F.interpolate(x, scale_factor=2, mode="nearest")
    dy = torch.randn_like(y)
    y.backward(dy)
    dx = torch.empty_like(to_cl(x.detach()))
    ops.upsample2x_bwd(to_cl(dy), dx, ndim)
    assert relmax(from_cl(dx, ndim), x.grad) < 1e-6
    # add_ex over dtype mixes, split_channels, colsum, silu, softmax backward
    a, b = torch.randn(1000), torch.randn(1000)
    out = torch.empty(1000, device=DEV, dtype=BF)
    ops.add_ex(a.to(DEV), b.to(DEV).to(BF), out)
    assert relmax(out, a + b.to(BF).float()) < 8e-3
    out32 = torch.empty(1000, device=DEV)
    ops.add_ex(a.to(DEV).to(BF), None, out32)
    assert torch.equal(out32.cpu(), a.to(BF).float())
    dyc = torch.randn(7, 5, 12, device=DEV)
    da, db = torch.empty(7, 5, 8, device=DEV), torch.empty(7, 5, 4, device=DEV)
    ra = torch.randn(7, 5, 8, device=DEV)
    ops.split_channels(dyc, da, db, ra=ra)
    assert torch.equal(da, dyc[..., :8] + ra) and torch.equal(db, dyc[..., 8:])
    m = torch.randn(300, 70)
    cs = torch.empty(70, device=DEV)
    ops.colsum(m.to(DEV), cs)
    assert relmax(cs, m.double().sum(0)) < 1e-6
    z = (torch.randn(500) * 3).requires_grad_(True)
    g = torch.randn(500)
    F.silu(z).backward(g)
    sa, sb = torch.empty(500, device=DEV), torch.empty(500, device=DEV)
    ops.silu_fwd(z.detach().to(DEV), sa)
    ops.silu_bwd(z.detach().to(DEV), g.to(DEV), sb)
    assert relmax(sa, F.silu(z.detach())) < 1e-6 and relmax(sb, z.grad) < 2e-6
    s = torch.randn(6, 40, requires_grad=True)
    p = torch.softmax(s, -1)
    gp = torch.randn(6, 40)
    p.backward(gp)
    dP = gp.clone().to(DEV)
    ops.softmax_bwd_rows(p.detach().to(DEV), dP, 6, 40)
    assert relmax(dP, s.grad) < 2e-6


# ------------------------------------------------------------------------------------------------ whole networks
def _oracle_grads(g, x, t, dF, dtype=torch.float64):
    """autograd through the CPU oracle (fp64): d<F, dF>/dtheta for every parameter."""
    from oracle import nets_oracle as N
    from tests.test_gpu_nets import build_net  # noqa: F401  (same synthetic weights)
    sd = {k: v.to(dtype) for k, v in N.synth_state_dict(g["manifest"], g["seed"]).items()}
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items() if not k.endswith(".W")}
    full = dict(sd, **leaves)

    class Cfg:
        pass
    import diffsci_b200 as d
    cfg = (d.PUNetGConfig if g["kind"] == "punetg" else d.ADMConfig)(**g["cfg"])
    fwd = N.punetg_forward if g["kind"] == "punetg" else N.adm_forward
    Fo = fwd(full, cfg, x.to(dtype), t.to(dtype))
    (Fo * dF.to(dtype)).sum().backward()
    return Fo.detach(), {k: v.grad for k, v in leaves.items()}


@pytest.mark.parametrize("name", ["punetg2d_mc8", "punetg3d_mc8", "punetg2d_multi", "adm2d_mc8", "adm2d_add"])
def test_net_backward_fp32(golden, name):
    from tests.test_gpu_nets import build_net
    g = golden(name)
    net = build_net(g).train()
    x, t = g["x"], g["t"]
    torch.manual_seed(11)
    dF = torch.randn_like(g["y"])
    F64, ref = _oracle_grads(g, x, t, dF)
    out = net(x.to(DEV), t.to(DEV))
    assert out.requires_grad
    assert relmax(out.detach(), F64) < 5e-5
    out.backward(dF.to(DEV))
    worst = ("", 0.0)
    for k, p in net.named_parameters():
        assert p.grad is not None, k
        e = relmax(p.grad, ref[k])
        if e > worst[1]:
            worst = (k, e)
    print(f"{name}: worst parameter-gradient max-rel error vs fp64 autograd = {worst[1]:.2e} ({worst[0]})")
    assert worst[1] < 2e-4, worst
    # a second step reuses the graph (no stale state) and accumulates into .grad like autograd
    out2 = net(x.to(DEV), t.to(DEV))
    out2.backward(dF.to(DEV))
    k0, p0 = next(iter(net.named_parameters()))
    assert relmax(p0.grad, 2 * ref[k0]) < 2e-4


@pytest.mark.parametrize("name", ["punetg2d_mc8", "punetg3d_mc8", "adm2d_mc8"])
def test_net_backward_bf16(golden, name):
    from tests.test_gpu_nets import build_net
    g = golden(name)
    net = build_net(g, "bf16").train()
    x, t = g["x"], g["t"]
    torch.manual_seed(11)
    dF = torch.randn_like(g["y"])
    _, ref = _oracle_grads(g, x, t, dF)
    net(x.to(DEV), t.to(DEV)).backward(dF.to(DEV))
    num = sum(float((p.grad.double().cpu() - ref[k]).pow(2).sum()) for k, p in net.named_parameters())
    den = sum(float(ref[k].pow(2).sum()) for k, _ in net.named_parameters())
    err = math.sqrt(num / den)
    print(f"{name} bf16: global parameter-gradient L2 error vs fp64 autograd = {err:.2e}")
    # bf16 storage of every activation AND activation gradient through ~40 layers; stated tolerance 6e-2 of the global norm
    assert err < 6e-2


def test_training_loss_grads_vs_live_reference(golden):
    """loss_fn -> backward through the native PUNetG: gradients recorded from the LIVE reference (oracle/make_goldens.py)."""
    import diffsci_b200 as d
    from tests.test_gpu_nets import build_net
    g = golden("sampler_punetg2d")
    net = build_net(golden("punetg2d_mc8")).train()
    for metric in ("huber", "mse"):
        mod = d.KarrasModule(net, d.KarrasModuleConfig.from_edm(loss_metric=metric)).train()
        for use_mask in (False, True):
            net.zero_grad()
            mod._injected_loss_noise = g["loss_noise"]
            L = mod.loss_fn(g["loss_x"].to(DEV), g["loss_sigma"].to(DEV), None, g["loss_mask"].to(DEV) if use_mask else None)
            L.backward()
            key = f"loss_{metric}{'_mask' if use_mask else ''}"
            assert abs(float(L) - float(g[key])) < 5e-5 * abs(float(g[key]))
            grads = dict(net.named_parameters())
            for k, ref in g[key + "_grads"].items():
                e = relmax(grads[k].grad, ref)
                assert e < 3e-4, (key, k, e)


# ------------------------------------------------------------------------------------------------ fused training step
@pytest.mark.parametrize("name,metric", [("punetg2d_mc8", "huber"), ("adm2d_mc8", "mse")])
def test_trainer_steps_match_oracle(golden, name, metric):
    """EDMTrainer.step (noising, forward, fused loss, hand-written backward, fused AdamW + EMA) against the CPU oracle:
    autograd through the oracle network + the oracle's AdamW / EMA restatements, 3 iterations on identical sigma / noise."""
    import diffsci_b200 as d
    from oracle import karras_oracle as K, nets_oracle as N
    from tests.test_gpu_nets import build_net
    g = golden(name)
    net = build_net(g).train()
    mod = d.KarrasModule(net, d.KarrasModuleConfig.from_edm(loss_metric=metric)).train()
    ema = d.ModelEMA(net, ema_type="traditional", decay=0.9)
    tr = d.EDMTrainer(mod, lr=2e-3, ema=ema)
    cfg = (d.PUNetGConfig if g["kind"] == "punetg" else d.ADMConfig)(**g["cfg"])
    fwd = N.punetg_forward if g["kind"] == "punetg" else N.adm_forward
    dt = torch.float64
    sd = {k: v.to(dt) for k, v in N.synth_state_dict(g["manifest"], g["seed"]).items()}
    names = [k for k, _ in net.named_parameters()]
    P = {k: sd[k].clone() for k in names}
    Mo = {k: torch.zeros_like(v) for k, v in P.items()}
    Vo = {k: torch.zeros_like(v) for k, v in P.items()}
    Sh = {k: v.clone() for k, v in P.items()}
    torch.manual_seed(21)
    shape = g["x"].shape
    for it in range(1, 4):
        x = torch.randn(shape) * 0.5
        sigma = torch.exp(torch.randn(shape[0]) * 1.2 - 1.2)
        noise = torch.randn(shape)
        loss = tr.step(x.to(DEV), sigma=sigma.to(DEV), noise=noise.to(DEV))
        graph = net.train_graph(shape[0], tuple(shape[2:]), torch.device(DEV))
        g_gpu = {k: gv.detach().double().cpu() for k, gv in zip(names, graph.grads())}
        # (1) loss and gradients at the parameters the step started from
        leaves = {k: v.clone().requires_grad_(True) for k, v in P.items()}
        full = dict(sd, **leaves)
        L = K.edm_loss(lambda xx, tt: fwd(full, cfg, xx, tt), x.to(dt), sigma.to(dt), noise.to(dt), loss_metric=metric)
        L.backward()
        assert abs(float(loss) - float(L)) < 1e-4 * abs(float(L)), (it, float(loss), float(L))
        worst_g = max(relmax(g_gpu[k], leaves[k].grad) for k in names)
        assert worst_g < 5e-4, (it, worst_g)
        # (2) the fused AdamW + EMA arithmetic.  Adam divides by sqrt(v): for elements whose gradient is within rounding
        # of zero the update is +-lr whatever the sign noise says, so the optimizer restatement is fed the gradients the
        # device produced (checked above) -- what is compared here is the update rule, bit-close.
        for k in names:
            P[k], Mo[k], Vo[k] = K.adamw_step(P[k], g_gpu[k], Mo[k], Vo[k], it, lr=2e-3)
            Sh[k] = K.ema_update(Sh[k], P[k], 0.9)
        worst = max(relmax(p.detach(), P[k]) for k, p in net.named_parameters())
        worst_s = max(relmax(ema.selected_profile()["params"][k], Sh[k]) for k in names)
        assert worst < 2e-6 and worst_s < 2e-6, (it, worst, worst_s)
    assert ema.num_updates == 3 and tr.nstep == 3
    # inference after training sees the updated weights (packed copies were invalidated)
    net.eval()
    with torch.no_grad():
        y = net(g["x"].to(DEV), g["t"].to(DEV)).cpu()
    yo = fwd(dict(sd, **P), cfg, g["x"].to(dt), g["t"].to(dt))
    assert relmax(y, yo) < 2e-3


def test_trainer_equals_autograd_seam(golden):
    """The fused trainer and KarrasModule.loss_fn + backward (autograd seam) produce the same gradients (bf16 mode,
    tcgen05 kernels where the shapes allow: PUNetG-2D mc=64)."""
    import diffsci_b200 as d
    torch.manual_seed(5)
    cfgk = dict(dimension=2, model_channels=64)
    net = d.PUNetG(d.PUNetGConfig(**cfgk), precision="bf16").to(DEV).train()
    mod = d.KarrasModule(net, d.KarrasModuleConfig.from_edm()).train()
    x = (torch.randn(4, 1, 32, 32) * 0.5).to(DEV)
    sigma = torch.exp(torch.randn(4) * 1.2 - 1.2).to(DEV)
    noise = torch.randn(4, 1, 32, 32).to(DEV)
    mod._injected_loss_noise = noise
    L = mod.loss_fn(x, sigma)
    L.backward()
    ref = [p.grad.clone() for p in net.parameters()]
    tr = d.EDMTrainer(mod, lr=0.0, weight_decay=0.0)
    L2 = tr.step(x, sigma=sigma, noise=noise)
    graph = net.train_graph(4, (32, 32), x.device)
    assert abs(float(L2) - float(L)) < 1e-6 * abs(float(L))
    for a, b in zip(graph.grads(), ref):
        assert torch.equal(a, b)
    # and against fp32 mode: the bf16 step is a faithful (bf16-accurate) gradient
    net32 = d.PUNetG(d.PUNetGConfig(**cfgk), precision="fp32").to(DEV).train()
    net32.load_state_dict(net.state_dict())
    mod32 = d.KarrasModule(net32, d.KarrasModuleConfig.from_edm()).train()
    mod32._injected_loss_noise = noise
    mod32.loss_fn(x, sigma).backward()
    num = sum(float((a - p.grad).double().pow(2).sum()) for a, p in zip(ref, net32.parameters()))
    den = sum(float(p.grad.double().pow(2).sum()) for p in net32.parameters())
    err = math.sqrt(num / den)
    print(f"PUNetG-2D mc=64 bf16-vs-fp32 global gradient L2 error: {err:.2e}")
    assert err < 6e-2


# ------------------------------------------------------------------------------------------------ grouped time-MLP layer
@pytest.mark.parametrize("act,shared,B", [(1, False, 5), (0, False, 37), (0, True, 6), (1, True, 3), (1, False, 70), (0, True, 33),
                                          (1, True, 64)])
def test_grouped_linear_bwd(ops, act, shared, B):
    """dsk_grouped_linear (+ pre-activation output) and dsk_grouped_linear_bwd against autograd of F.linear / F.silu."""
    torch.manual_seed(8)
    dims = [(64, 24), (40, 24), (600, 24)] if shared else [(64, 16), (24, 48), (8, 700)]     # (N, K); K > 512 sweeps twice
    x0 = torch.randn(B, dims[0][1], dtype=torch.float64)
    xs = [x0.clone().requires_grad_(True)] if shared else [torch.randn(B, k, dtype=torch.float64, requires_grad=True) for _, k in dims]
    ws = [(torch.randn(n, k, dtype=torch.float64) / math.sqrt(k)).requires_grad_(True) for n, k in dims]
    bs = [torch.randn(n, dtype=torch.float64, requires_grad=True) for n, _ in dims]
    dys = [torch.randn(B, n, dtype=torch.float64) for n, _ in dims]
    tot = 0
    for i, (w, b, dy) in enumerate(zip(ws, bs, dys)):
        z = F.linear(xs[0] if shared else xs[i], w, b)
        tot = tot + ((F.silu(z) if act else z) * dy).sum()
    tot.backward()
    f = lambda t: t.detach().float().to(DEV).contiguous()  # noqa: E731
    xd = [f(xs[0])] * len(dims) if shared else [f(x) for x in xs]
    wd, bd, dyd = [f(w) for w in ws], [f(b) for b in bs], [f(d) for d in dys]
    yd = [torch.empty(B, n, device=DEV) for n, _ in dims]
    zd = [torch.empty(B, n, device=DEV) for n, _ in dims] if act else None
    layer = ops.GroupedLinear(xd, wd, bd, yd, act, zd)
    layer.run()
    for i, (w, b) in enumerate(zip(ws, bs)):
        z = F.linear(xs[0] if shared else xs[i], w, b).detach()
        assert relmax(yd[i], F.silu(z) if act else z) < 1e-5
        if act:
            assert relmax(zd[i], z) < 1e-5
    dzd = [torch.empty_like(d) for d in dyd] if act else dyd
    dwd, dbd = [torch.empty_like(w) for w in wd], [torch.empty_like(b) for b in bd]
    dxd = ([torch.empty_like(xd[0])] + [None] * (len(dims) - 1)) if shared else [torch.empty_like(x) for x in xd]
    layer.backward_tables(dyd, dzd, dwd, dbd, dxd, shared_dx=shared)()
    for i in range(len(dims)):
        assert relmax(dwd[i], ws[i].grad) < 2e-5, i
        assert relmax(dbd[i], bs[i].grad) < 2e-5, i
        if not shared:
            assert relmax(dxd[i], xs[i].grad) < 2e-5, i
    if shared:
        assert relmax(dxd[0], xs[0].grad) < 2e-5


# ------------------------------------------------------------------------------------------------ tensor-core attention
@pytest.mark.parametrize("residual", [False, True])
@pytest.mark.parametrize("B,H,W", [(3, 8, 16), (8, 7, 7)])
def test_attention_tc_forward_backward(residual, B, H, W):
    """TrainGraph.attention on the tcgen05 path (bf16 operands, MN-major A^T B products) against fp64 autograd of
    nn.MultiheadAttention(C, 1 head) on the same bf16-rounded input and weights.  (8, 7, 7): 49 tokens -- off the
    tensor-core score path, the hybrid form (projections + their gradients on tcgen05, the L x L part in fp32)."""
    from diffsci_b200.models.nets.graph import TrainGraph, Var
    torch.manual_seed(31)
    C = 64
    L = H * W
    mha = torch.nn.MultiheadAttention(C, 1, batch_first=True)
    with torch.no_grad():
        for p in mha.parameters():
            p.copy_((torch.randn_like(p) * (0.1 if p.ndim == 1 else C ** -0.5)).bfloat16().float())
    holder = torch.nn.Module()
    holder.mhattn = mha
    holder = holder.to(DEV)
    g = TrainGraph(holder, B, torch.device(DEV), "bf16", 2)
    x = Var(g.empty((B, 1, H, W, C)))
    y = g.attention(x, holder.mhattn, residual)
    g.finalize(y)
    kinds = {getattr(f, "__qualname__", "").split(".")[1] for f in g.fwd}
    assert kinds == ({"_attention_tc"} if L % 8 == 0 else {"_attention_hybrid"}), kinds
    xv = torch.randn(B, L, C).bfloat16()
    dyv = torch.randn(B, L, C).bfloat16()
    x.t.copy_(xv.view(B, 1, H, W, C))
    g.run_forward()
    y.g.copy_(dyv.view(B, 1, H, W, C))
    g.run_backward()
    ref = torch.nn.MultiheadAttention(C, 1, batch_first=True).double()
    ref.load_state_dict({k: v.double().cpu() for k, v in mha.state_dict().items()})
    xr = xv.double().requires_grad_(True)
    out = ref(xr, xr, xr, need_weights=False)[0]
    if residual:
        out = out + xr
    out.backward(dyv.double())
    # bf16 storage of Q|K|V, P, the attention output and every intermediate gradient: stated tolerance 2e-2 of the range
    assert relmax(y.t.float().view(B, L, C), out.detach()) < 2e-2
    assert relmax(x.g.float().view(B, L, C), xr.grad) < 2e-2
    grads = dict(zip((n for n, _ in holder.named_parameters()), g.grads()))
    for n, p in ref.named_parameters():
        e = relmax(grads["mhattn." + n], p.grad)
        assert e < 2e-2, (n, e)


def test_hybrid_attention_training_path():
    """The hybrid attention inside a whole bf16 training graph (PUNetG-2D mc=16 on 28x28: 49 tokens x 64 channels at the
    bottom level): global parameter-gradient error vs fp64 autograd of the oracle no worse than the same graph with the
    all-fp32 CUDA-core attention (A/B switch graph.HYBRID_ATTENTION).  The path itself is pinned in isolation by
    test_attention_tc_forward_backward[8-7-7]; this narrow default-init net sits at 6-7e-2 on either path (bf16 storage)."""
    import types
    import diffsci_b200 as d
    from diffsci_b200.models.nets import graph as G
    from oracle import nets_oracle as N
    torch.manual_seed(21)
    kw = dict(dimension=2, model_channels=16)                 # bottom level: 64 channels, 28 -> 7
    net = d.PUNetG(d.PUNetGConfig(**kw), precision="bf16").to(DEV).train()
    sd = {k: v.detach().cpu().double() for k, v in net.state_dict().items()}
    cfg = types.SimpleNamespace(**d.PUNetGConfig(**kw).export_description())
    x, t = torch.randn(8, 1, 28, 28), torch.randn(8) * 0.5
    dF = torch.randn(8, 1, 28, 28)
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items() if not k.endswith(".W")}
    N.punetg_forward(dict(sd, **leaves), cfg, x.double(), t.double()).backward(dF.double())
    den = sum(float(v.grad.pow(2).sum()) for v in leaves.values())
    err = {}
    for hybrid in (True, False):
        G.HYBRID_ATTENTION = hybrid
        try:
            net._plans.clear()
            net.zero_grad()
            g = net.train_graph(8, (28, 28), DEV)
            assert any("_attention_hybrid" in getattr(f, "__qualname__", "") for f in g.fwd) == hybrid
            net(x.to(DEV), t.to(DEV)).backward(dF.to(DEV))
            err[hybrid] = math.sqrt(sum(float((p.grad.double().cpu() - leaves[k].grad).pow(2).sum())
                                        for k, p in net.named_parameters()) / den)
        finally:
            G.HYBRID_ATTENTION = True
    print(f"global gradient L2 error vs fp64: hybrid {err[True]:.3e}, fp32-core attention {err[False]:.3e}")
    assert err[True] < max(1.25 * err[False], 6e-2) and err[True] < 1e-1, err


# ------------------------------------------------------------------------------------------------ training-mode dropout
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_dropout_kernel(dtype):
    """dsk_dropout: y in {0, x / (1-p)}, keep fraction 1-p, the mask depends on (seed, stream, index) only -- so the backward
    launch regenerates the forward's mask --, streams and seeds are independent, dres is added after masking."""
    from diffsci_b200 import ops
    torch.manual_seed(0)
    n, p = 1_000_003, 0.3
    x = (torch.randn(n, device=DEV) + 3.0).to(dtype)
    y = ops.dropout(x, p, 1234, 5)
    keep = y != 0
    assert abs(float(keep.float().mean()) - (1 - p)) < 3e-3
    assert relmax(y[keep].float(), (x[keep].float() / (1 - p)).to(dtype).float()) < 1e-6
    assert torch.equal(ops.dropout(x, p, 1234, 5), y)
    other_stream, other_seed = ops.dropout(x, p, 1234, 6) != 0, ops.dropout(x, p, 1235, 5) != 0
    for o in (other_stream, other_seed):        # independent masks agree on (1-p)^2 + p^2 of the positions
        assert abs(float((o == keep).float().mean()) - ((1 - p) ** 2 + p ** 2)) < 5e-3
    dy, dres = torch.randn(n, device=DEV).to(dtype), torch.randn(n, device=DEV).to(dtype)
    dx = ops.dropout(dy, p, 1234, 5, dres=dres)
    ref = torch.where(keep, dy.float() / (1 - p), torch.zeros_like(dy.float())) + dres.float()
    assert relmax(dx.float(), ref.to(dtype).float()) < (1e-6 if dtype == torch.float32 else 8e-3)
    assert torch.equal(ops.dropout(x, 0.0, 1, 1), x)
    with pytest.raises(RuntimeError):
        ops.dropout(x, 1.0, 1, 1)


@pytest.mark.parametrize("name", ["punetg2d_mc8", "punetg3d_mc8", "adm2d_mc8"])
def test_net_backward_with_dropout(golden, name):
    """dropout > 0 on the training path: the masks the device drew (regenerated per site from the pinned seed) are fed to the
    fp64 oracle; forward output and every parameter gradient agree, i.e. the backward applied the same masks."""
    import diffsci_b200 as d
    from diffsci_b200 import ops
    from oracle import nets_oracle as N
    g = golden(name)
    p = 0.25
    cfg_kw = dict(g["cfg"], dropout=p)
    net = (d.PUNetG(d.PUNetGConfig(**cfg_kw)) if g["kind"] == "punetg" else d.ADM(d.ADMConfig(**cfg_kw)))
    net.load_state_dict(N.synth_state_dict(g["manifest"], g["seed"]))
    net = net.to(DEV).train()
    x, t = g["x"], g["t"]
    graph = net.train_graph(x.shape[0], tuple(x.shape[2:]), torch.device(DEV))
    assert len(graph.dropout_sites) == 14
    graph.fixed_dropout_seed = 987654321
    out = net(x.to(DEV), t.to(DEV))
    torch.manual_seed(11)
    dF = torch.randn_like(g["y"])
    out.backward(dF.to(DEV))
    masks = {}
    for site, sid, shape in graph.dropout_sites:
        m = ops.dropout(torch.ones(shape, device=DEV), p, graph.fixed_dropout_seed, sid)      # [B, D, H, W, C], pre-scaled
        m = m.permute(0, 4, 1, 2, 3)
        masks[site] = (m.squeeze(2) if len(x.shape) == 4 else m).cpu().double()
    keep_frac = sum(float((m != 0).sum()) for m in masks.values()) / sum(m.numel() for m in masks.values())
    assert abs(keep_frac - (1 - p)) < 0.02
    sd = {k: v.double() for k, v in N.synth_state_dict(g["manifest"], g["seed"]).items()}
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items() if not k.endswith(".W")}
    cfg = (d.PUNetGConfig if g["kind"] == "punetg" else d.ADMConfig)(**cfg_kw)
    fwd = N.punetg_forward if g["kind"] == "punetg" else N.adm_forward
    Fo = fwd(dict(sd, **leaves), cfg, x.double(), t.double(), dropout_masks=masks)
    (Fo * dF.double()).sum().backward()
    assert relmax(out.detach(), Fo.detach()) < 5e-5
    worst = max((relmax(pp.grad, leaves[k].grad), k) for k, pp in net.named_parameters())
    assert worst[0] < 2e-4, worst
    # a fresh seed per forward when nothing is pinned; evaluation ignores dropout; no-grad training-mode forward fails loudly
    graph.fixed_dropout_seed = None
    a, b = net(x.to(DEV), t.to(DEV)).detach(), net(x.to(DEV), t.to(DEV)).detach()
    assert not torch.equal(a, b)
    with torch.no_grad(), pytest.raises(NotImplementedError):
        net(x.to(DEV), t.to(DEV))
    net.eval()
    with torch.no_grad():
        e = net(x.to(DEV), t.to(DEV))
    assert relmax(e, g["y"]) < 5e-5
