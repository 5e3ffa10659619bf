"""GPU parity tests, kernel level: every C-ABI compute entry point against the CPU oracle / ATen fp32
reference on seeded inputs.  Tolerances are stated per test (fp32 accumulation-order noise unless noted)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def relmax(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def to_cl(x):
    """NC(D)HW cpu -> channels-last [B, D, H, W, C] on the GPU."""
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


CONV_CASES = [
    # ndim, B, Cin, Cout, spatial, k, up2
    (2, 2, 8, 8, (12, 20), 3, False),
    (2, 1, 1, 16, (28, 28), 3, False),      # convin-like (Cin = 1)
    (2, 2, 16, 1, (7, 7), 3, False),        # convout-like (Cout = 1), odd spatial size
    (2, 1, 3, 5, (9, 11), 3, False),        # ragged channels
    (3, 2, 8, 16, (6, 8, 10), 3, False),
    (3, 1, 16, 8, (4, 6, 8), 3, True),      # fused nearest x2 upsample
    (2, 2, 8, 12, (8, 8), 3, True),
    (2, 2, 24, 8, (5, 6), 1, False),        # 1x1 conv (ADM residual path)
    (3, 1, 64, 64, (8, 8, 8), 3, False),    # C4-like channel count
]


@pytest.mark.parametrize("ndim,B,Cin,Cout,sp,k,up2", CONV_CASES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_conv_fwd(ops, ndim, B, Cin, Cout, sp, k, up2, dtype):
    torch.manual_seed(1)
    x = torch.randn(B, Cin, *sp)
    w = torch.randn(Cout, Cin, *([k] * ndim)) / math.sqrt(Cin * k ** ndim)
    b = torch.randn(Cout) * 0.1
    xs = x.to(dtype).float()
    xr = F.interpolate(xs, scale_factor=2, mode="nearest") if up2 else xs
    ref = (F.conv2d if ndim == 2 else F.conv3d)(xr, w, b, padding=k // 2)
    cb = torch.randn(B, Cout) * 0.3
    res = torch.randn_like(ref).to(dtype).float()
    ref = ref + cb.view(B, Cout, *([1] * ndim)) + res
    pc = ops.PackedConv(w.to(DEV), b.to(DEV), ndim, torch.float32)
    y = ops.conv(to_cl(x).to(dtype), pc, chan_bias=cb.to(DEV), residual=to_cl(res).to(dtype), up2=up2)
    tol = 2e-5 if dtype == torch.float32 else 8e-3   # bf16: one rounding of the stored output
    assert relmax(from_cl(y, ndim), ref) < tol
    # NC(D)HW fp32 output variant (convout)
    y2 = ops.conv(to_cl(x).to(dtype), pc, up2=up2, out_nchw=True)
    ref2 = (F.conv2d if ndim == 2 else F.conv3d)(xr, w, b, padding=k // 2)
    assert relmax(y2.cpu(), ref2) < 2e-5


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("B,C,G,sp", [(2, 8, 8, (5, 7)), (2, 16, 1, (6, 6)), (1, 64, 64, (4, 8, 8)), (3, 32, 4, (9,)),
                                       (2, 256, 256, (16, 16))])
def test_norm_act(ops, mode, B, C, G, sp):
    from oracle import nets_oracle as N
    torch.manual_seed(2)
    x = torch.randn(B, C, *sp) * 2 + 3.0          # non-zero mean: exercises the variance cancellation
    g, b = 1 + 0.1 * torch.randn(C), 0.1 * torch.randn(C)
    fs, fh = torch.randn(B, C), torch.randn(B, C)
    ref = F.group_norm(x, G, g, b, 1e-5) if mode == 0 else N.group_rms_norm(x, G, g, b)
    xcl = x.reshape(B, C, -1).permute(0, 2, 1).contiguous().to(DEV)
    y = ops.norm_act(xcl, g.to(DEV), b.to(DEV), G, mode, True)
    assert relmax(y.cpu().permute(0, 2, 1).reshape(x.shape), F.silu(ref)) < 1e-5
    y = ops.norm_act(xcl, g.to(DEV), b.to(DEV), G, mode, True, film_scale=fs.to(DEV), film_shift=fh.to(DEV))
    shp = (B, C) + (1,) * len(sp)
    assert relmax(y.cpu().permute(0, 2, 1).reshape(x.shape), F.silu(ref * fs.view(shp) + fh.view(shp))) < 1e-5
    yb = ops.norm_act(xcl.bfloat16(), g.to(DEV), b.to(DEV), G, mode, False)
    xb = x.bfloat16().float()
    refb = F.group_norm(xb, G, g, b, 1e-5) if mode == 0 else N.group_rms_norm(xb, G, g, b)
    assert relmax(yb.float().cpu().permute(0, 2, 1).reshape(x.shape), refb) < 8e-3


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("B,C,sp", [(64, 128, (28, 28)), (160, 32, (7, 7)), (40, 256, (14, 14)), (5, 1024, (3, 5))])
def test_norm_act_small_samples_one_kernel(ops, mode, B, C, sp):
    """Per-channel bf16 norms on many small samples (MNIST-size tensors): the single-kernel slab path (norm_slab_kernel: one
    read, statistics + fold + apply in shared memory) vs ATen on the same bf16 input, with and without FiLM / SiLU, the
    table-only form, and the (mean, rstd) it leaves for the backward."""
    from oracle import nets_oracle as N
    torch.manual_seed(5)
    x = (torch.randn(B, C, *sp) * 2 + 1.5).bfloat16()
    g, b = 1 + 0.1 * torch.randn(C), 0.1 * torch.randn(C)
    fs, fh = torch.randn(B, C), torch.randn(B, C)
    xf = x.float()
    ref = F.group_norm(xf, C, g, b, 1e-5) if mode == 0 else N.group_rms_norm(xf, C, g, b)
    S = sp[0] * sp[1]
    xcl = xf.reshape(B, C, S).permute(0, 2, 1).contiguous().to(DEV).bfloat16()
    back = lambda y: y.float().cpu().permute(0, 2, 1).reshape(xf.shape)  # noqa: E731
    ws = ops.norm_ws(B, S, C, DEV)
    y = ops.norm_act(xcl, g.to(DEV), b.to(DEV), C, mode, True, ws=ws)
    assert relmax(back(y), F.silu(ref)) < 8e-3
    # statistics left for the backward: (mean, rstd) per (b, c) behind the (scale, shift) table
    st = ws.view(torch.float32)[2 * B * C: 4 * B * C].view(B, C, 2).cpu()
    mean = xf.flatten(2).mean(-1)
    var = xf.flatten(2).var(-1, unbiased=False) if mode == 0 else (xf.flatten(2) ** 2).mean(-1)
    assert relmax(st[..., 1], 1 / torch.sqrt(var + 1e-5)) < 1e-5
    if mode == 0:
        assert relmax(st[..., 0], mean) < 1e-5
    shp = (B, C, 1, 1)
    y = ops.norm_act(xcl, g.to(DEV), b.to(DEV), C, mode, False, film_scale=fs.to(DEV), film_shift=fh.to(DEV))
    assert relmax(back(y), ref * fs.view(shp) + fh.view(shp)) < 8e-3
    y = ops.norm_act(xcl, None, None, C, mode, True)
    plain = F.group_norm(xf, C, None, None, 1e-5) if mode == 0 else N.group_rms_norm(xf, C, torch.ones(C), torch.zeros(C))
    assert relmax(back(y), F.silu(plain)) < 8e-3


@pytest.mark.parametrize("ndim,sp", [(2, (8, 10)), (2, (7, 9)), (3, (4, 6, 8))])
@pytest.mark.parametrize("is_max", [True, False])
def test_pool_add_layout(ops, ndim, sp, is_max):
    torch.manual_seed(3)
    x = torch.randn(2, 8, *sp)
    fn = {(2, True): F.max_pool2d, (2, False): F.avg_pool2d, (3, True): F.max_pool3d, (3, False): F.avg_pool3d}[ndim, is_max]
    y = ops.pool2x(to_cl(x), ndim, is_max)
    assert relmax(from_cl(y, ndim), fn(x, 2)) < 1e-6
    a = ops.add(to_cl(x), to_cl(2 * x))
    assert relmax(from_cl(a, ndim), 3 * x) < 1e-6
    cl = ops.nchw_to_cl(x.to(DEV), torch.float32, ndim)
    assert torch.equal(cl.cpu(), to_cl(x).cpu())
    assert torch.equal(ops.cl_to_nchw(cl, ndim).cpu(), x)
    cc = ops.concat_channels(to_cl(x), to_cl(x[:, :3]))
    assert torch.equal(from_cl(cc, ndim), torch.cat([x, x[:, :3]], 1))


def test_fourier_and_grouped_linear(ops):
    from oracle import nets_oracle as N
    torch.manual_seed(4)
    t = torch.tensor([-3.4, 0.0, 0.35, 2.19])        # c_noise range of EDM: 0.5*log(0.002..80)
    W = torch.randn(32) * 30.0
    out = ops.fourier(t.to(DEV), W.to(DEV))
    assert float((out.cpu() - N.fourier(t, W)).abs().max()) < 2e-4   # |arg| ~ 1e3 rad: fp32 argument rounding
    B = 5
    xs = [torch.randn(B, k) for k in (16, 16, 64)]
    ws = [torch.randn(n, k) / math.sqrt(k) for n, k in ((64, 16), (24, 16), (8, 64))]
    bs = [torch.randn(w.shape[0]) for w in ws]
    ys = [torch.empty(B, w.shape[0], device=DEV) for w in ws]
    for act, fn in ((0, lambda v: v), (1, F.silu), (2, F.relu)):
        g = ops.GroupedLinear([x.to(DEV) for x in xs], [w.to(DEV) for w in ws], [b.to(DEV) for b in bs], ys, act)
        g.run()
        for x, w, b, y in zip(xs, ws, bs, ys):
            assert relmax(y.cpu(), fn(F.linear(x, w, b))) < 1e-5


@pytest.mark.parametrize("M,N,K,batch,transB", [(200, 70, 33, 1, True), (128, 64, 256, 3, True), (50, 130, 75, 2, False),
                                                 (1000, 128, 3, 1, True)])
def test_gemm_and_softmax(ops, M, N, K, batch, transB):
    torch.manual_seed(5)
    A = torch.randn(batch, M, K)
    Bm = torch.randn(batch, N, K) if transB else torch.randn(batch, K, N)
    bias = torch.randn(N)
    ref = 0.7 * (A @ (Bm.transpose(1, 2) if transB else Bm)) + bias
    out = torch.empty(batch, M, N, device=DEV)
    ops.gemm(A.to(DEV), Bm.to(DEV), out, M=M, N=N, K=K, lda=K, ldb=K if transB else N, ldc=N, bias=bias.to(DEV),
             transB=transB, alpha=0.7, batch=batch, strideA=M * K, strideB=N * K, strideC=M * N)
    assert relmax(out.cpu(), ref) < 1e-5
    sm = ops.softmax_rows(out.clone(), batch * M, N)
    assert relmax(sm.cpu(), torch.softmax(ref, -1)) < 1e-5


def test_attention(ops):
    from oracle import nets_oracle as N
    torch.manual_seed(6)
    B, C, sp = 2, 32, (6, 7)
    x = torch.randn(B, C, *sp)
    sd = {"a.mhattn.in_proj_weight": torch.randn(3 * C, C) / math.sqrt(C), "a.mhattn.in_proj_bias": torch.randn(3 * C) * 0.1,
          "a.mhattn.out_proj.weight": torch.randn(C, C) / math.sqrt(C), "a.mhattn.out_proj.bias": torch.randn(C) * 0.1}
    Lq = sp[0] * sp[1]
    for residual in (False, True):
        ref = N.mha_self_attention(x, sd, "a.", residual)
        f32 = dict(dtype=torch.float32, device=DEV)
        bufs = dict(qkv=torch.empty(B * Lq, 3 * C, **f32), scores=torch.empty(B, Lq, Lq, **f32),
                    ao=torch.empty(B * Lq, C, **f32), out=torch.empty(B, Lq, C, **f32))
        tok = x.reshape(B, C, Lq).permute(0, 2, 1).contiguous().to(DEV)
        o = ops.self_attention_f32(tok, *[sd[k].to(DEV) for k in sd], bufs, residual)
        assert relmax(o.cpu().permute(0, 2, 1).reshape(x.shape), ref) < 1e-5


def test_lincomb_and_philox(ops):
    torch.manual_seed(7)
    x, r1, r2, z = (torch.randn(1000) for _ in range(4))
    o = ops.lincomb(x.to(DEV), 1.0, r1.to(DEV), -0.3, r2.to(DEV), 0.7, z.to(DEV), 2.0)
    assert relmax(o.cpu(), x - 0.3 * r1 + 0.7 * r2 + 2.0 * z) < 1e-6
    n = ops.philox_normal((1 << 20,), 1234, 3, DEV).cpu()
    assert abs(float(n.mean())) < 5e-3 and abs(float(n.std()) - 1) < 5e-3
    assert abs(float((n ** 3).mean())) < 2e-2 and abs(float((n ** 4).mean()) - 3) < 5e-2
    assert torch.equal(n, ops.philox_normal((1 << 20,), 1234, 3, DEV).cpu())          # counter-based: reproducible
    assert not torch.equal(n, ops.philox_normal((1 << 20,), 1234, 4, DEV).cpu())      # new stream id -> new draws
    # Kolmogorov-Smirnov distance against the normal CDF
    s, _ = torch.sort(n.double())
    cdf = 0.5 * (1 + torch.erf(s / math.sqrt(2)))
    ks = float((cdf - torch.arange(1, len(s) + 1).double() / len(s)).abs().max())
    assert ks < 2e-3


def test_loss_ema_adamw():
    from diffsci_b200._lib import lib, check, ptr, stream
    from oracle import karras_oracle as K
    torch.manual_seed(8)
    B, shape = 3, (2, 5, 6)
    x, noise = torch.randn(B, *shape) * 0.5, torch.randn(B, *shape)
    Fv = torch.randn(B, *shape, requires_grad=True)
    sigma = torch.exp(torch.randn(B) * 1.2 - 1.2)
    mask = (torch.rand(B, *shape) > 0.6).float()
    for kind, name in ((0, "huber"), (1, "mse")):
        for m in (None, mask):
            ref = K.edm_loss(lambda xs, cn: Fv, x, sigma, noise, name, m)
            (gref,) = torch.autograd.grad(ref, Fv)
            loss = torch.zeros((), device=DEV)
            dF = torch.empty(B, *shape, device=DEV)
            keep = [Fv.detach().to(DEV), x.to(DEV), noise.to(DEV), sigma.to(DEV), None if m is None else m.to(DEV)]
            check(lib.dsk_edm_loss_fwd_bwd(*[ptr(t) for t in keep], ptr(loss), ptr(dF), B, shape[0],
                                           shape[1] * shape[2], 0.5, kind, stream()))
            assert abs(float(loss) - float(ref.detach())) < 2e-5 * abs(float(ref.detach()))
            assert relmax(dF.cpu(), gref) < 2e-5
    # EMA + AdamW multi-tensor kernels
    ps = [torch.randn(n) for n in (7, 1000, 33)]
    gs = [torch.randn_like(p) for p in ps]
    sh = [torch.randn_like(p) for p in ps]
    dps, dgs, dsh = [[t.to(DEV) for t in ts] for ts in (ps, gs, sh)]
    mk = lambda ts: torch.tensor([t.data_ptr() for t in ts], dtype=torch.int64, device=DEV)  # noqa: E731
    numel = torch.tensor([p.numel() for p in ps], dtype=torch.int64, device=DEV)
    t_sh, t_p = mk(dsh), mk(dps)        # keep the pointer tables alive until the launch is enqueued
    check(lib.dsk_ema_update(ptr(t_sh), ptr(t_p), ptr(numel), 3, 1000, 0.9, stream()))
    for s_, p_, d_ in zip(sh, ps, dsh):
        assert relmax(d_.cpu(), K.ema_update(s_, p_, 0.9)) < 1e-6
    ms, vs = [torch.zeros_like(p) for p in ps], [torch.zeros_like(p) for p in ps]
    dms, dvs = [t.to(DEV) for t in ms], [t.to(DEV) for t in vs]
    shadow = [t.clone() for t in dsh]
    tabs = [mk(ts) for ts in (dps, dgs, dms, dvs, shadow)]
    for step in (1, 2, 3):
        check(lib.dsk_adamw_ema_step(*[ptr(t) for t in tabs], ptr(numel), 3,
                                     1000, 1e-3, 0.9, 0.999, 1e-8, 1e-4, step, 0.99, 1.0, stream()))
        for i in range(3):
            ps[i], ms[i], vs[i] = K.adamw_step(ps[i], gs[i], ms[i], vs[i], step)
    for i in range(3):
        assert relmax(dps[i].cpu(), ps[i]) < 1e-5
        assert relmax(dms[i].cpu(), ms[i]) < 1e-5 and relmax(dvs[i].cpu(), vs[i]) < 5e-5   # v: (1-b2)*g*g association
    ref = [torch.optim.AdamW([torch.nn.Parameter(p.clone())], lr=1e-3, weight_decay=1e-4) for p in
           (torch.randn(5),)]  # sanity: oracle adamw == torch.optim.AdamW
    p0 = ref[0].param_groups[0]["params"][0]
    g0 = torch.randn(5)
    start = p0.detach().clone()
    p0.grad = g0.clone()
    ref[0].step()
    pe, _, _ = K.adamw_step(start, g0, torch.zeros(5), torch.zeros(5), 1)
    assert relmax(pe, p0.detach()) < 1e-6


TC_CASES = [
    # ndim, B, Cin, Cout, spatial
    (3, 1, 64, 64, (4, 16, 8)),          # exactly one tile pair
    (3, 2, 64, 64, (5, 20, 12)),         # ragged tiles in d, h and w (TMA zero fill + masked stores)
    (3, 1, 128, 64, (4, 16, 16)),        # two K chunks
    (3, 1, 64, 128, (2, 16, 8)),         # N_TILE = 128
    (3, 1, 128, 256, (4, 8, 8)),         # two N tiles of 128
    (3, 1, 64, 192, (2, 16, 8)),         # N_TILE = 64 x 3
    (2, 3, 64, 64, (16, 8)),             # 2-D: the batch is the plane axis
    (2, 5, 128, 128, (28, 28)),          # MNIST-like, odd number of planes
    (3, 1, 64, 64, (16, 32, 32)),        # 64 tiles
    (3, 2, 64, 64, (32, 32, 32)),        # > 148 tiles: persistent loop, accumulator double buffering
    (2, 4, 128, 128, (7, 7)),            # odd number of w-tiles, even number of plane groups: CTA pairs along the plane axis
    (2, 8, 64, 64, (14, 7)),
    (3, 2, 64, 128, (4, 16, 24)),        # 3-D, 3 w-tiles: pairs along depth
    (3, 1, 128, 64, (5, 32, 16)),        # merged depth taps (T = 2: rows in whole 32-row tiles, 64-channel output tiles): odd plane
                                         # count, two K chunks
    (3, 2, 64, 192, (3, 64, 16)),        # merged depth taps: three N tiles of 64, two 32-row tiles per plane
]


@pytest.mark.parametrize("ndim,B,Cin,Cout,sp", TC_CASES)
def test_conv_tcgen05(ops, ndim, B, Cin, Cout, sp):
    """tcgen05 implicit-GEMM conv (bf16 operands, fp32 TMEM accumulation) vs ATen fp32 conv on the same
    bf16-rounded operands.  Tolerance: one bf16 rounding of the stored output (2^-8 relative to the value,
    checked as 6e-3 of the output range) -- accumulation itself is fp32."""
    torch.manual_seed(11)
    x = torch.randn(B, Cin, *sp).bfloat16().float()
    w = (torch.randn(Cout, Cin, *([3] * ndim)) / math.sqrt(Cin * 3 ** ndim)).bfloat16().float()
    b = torch.randn(Cout) * 0.1
    cb = torch.randn(B, Cout) * 0.3
    ref0 = (F.conv2d if ndim == 2 else F.conv3d)(x, w, b, padding=1)
    res = torch.randn_like(ref0).bfloat16().float()
    ref = ref0 + cb.view(B, Cout, *([1] * ndim)) + res
    pc = ops.PackedConv(w.to(DEV), b.to(DEV), ndim, torch.bfloat16)
    y = ops.conv(to_cl(x).bfloat16(), pc, chan_bias=cb.to(DEV), residual=to_cl(res).bfloat16())
    err = relmax(from_cl(y, ndim), ref)
    assert err < 6e-3, err
    y0 = ops.conv(to_cl(x).bfloat16(), pc)          # no epilogue operands
    assert relmax(from_cl(y0, ndim), ref0) < 6e-3
    # against the FFMA kernel on the same operands the only difference is summation order (+ output rounding)
    pc32 = ops.PackedConv(w.to(DEV), b.to(DEV), ndim, torch.float32)
    y32 = ops.conv(to_cl(x).bfloat16(), pc32)
    assert relmax(y0.float(), y32.float()) < 6e-3


def test_upsample2x(ops):
    torch.manual_seed(12)
    for ndim, sp in ((2, (5, 6)), (3, (3, 4, 5))):
        x = torch.randn(2, 16, *sp).bfloat16()
        y = ops.upsample2x(to_cl(x.float()).bfloat16(), ndim)
        assert torch.equal(from_cl(y, ndim), F.interpolate(x.float(), scale_factor=2, mode="nearest"))


@pytest.mark.parametrize("M,N,K,batch", [(128, 64, 64, 1), (300, 200, 136, 2), (64, 512, 256, 1), (1000, 96, 64, 3)])
def test_gemm_bf16_tcgen05(ops, M, N, K, batch):
    """tcgen05 batched GEMM vs fp32 matmul on the same bf16-rounded operands (fp32 accumulation both sides)."""
    torch.manual_seed(13)
    A = torch.randn(batch, M, K).bfloat16()
    Bm = torch.randn(batch, N, K).bfloat16()
    bias = torch.randn(N)
    res = torch.randn(batch, M, N).bfloat16()
    ref = 0.5 * (A.float() @ Bm.float().transpose(1, 2)) + bias + res.float()
    ldc = (N + 7) // 8 * 8
    out = torch.zeros(batch, M, ldc, device=DEV)
    resp = torch.zeros(batch, M, ldc, dtype=torch.bfloat16)
    resp[..., :N] = res
    ops.gemm_bf16_tc(A.to(DEV), Bm.to(DEV), out, M=M, N=N, K=K, lda=K, ldb=K, ldc=ldc, bias=bias.to(DEV),
                     residual=resp.to(DEV), alpha=0.5, batch=batch, strideA=M * K, strideB=N * K, strideC=M * ldc)
    assert relmax(out.cpu()[..., :N], ref) < 1e-5
    # shared A (stride 0), per-row bias, bf16 output: the V^T projection pattern
    rb = torch.randn(M)
    ref2 = A[0].float() @ Bm.float().transpose(1, 2) + rb.view(1, M, 1)
    out2 = torch.zeros(batch, M, ldc, dtype=torch.bfloat16, device=DEV)
    ops.gemm_bf16_tc(A[0].contiguous().to(DEV), Bm.to(DEV), out2, M=M, N=N, K=K, lda=K, ldb=K, ldc=ldc, bias=rb.to(DEV),
                     bias_rows=True, batch=batch, strideA=0, strideB=N * K, strideC=M * ldc)
    assert relmax(out2.float().cpu()[..., :N], ref2) < 6e-3


@pytest.mark.parametrize("transA,transB", [(True, False), (False, True), (True, True)])
@pytest.mark.parametrize("M,N,K,batch", [(128, 64, 64, 1), (200, 136, 72, 2), (64, 256, 264, 1), (256, 192, 1000, 3)])
def test_gemm_bf16_tcgen05_mn_major(ops, M, N, K, batch, transA, transB):
    """MN-major operands: A stored [K, M] and / or B stored [K, N] -- the A^T B products of the backward passes."""
    torch.manual_seed(15)
    A = torch.randn(batch, K, M).bfloat16() if transA else torch.randn(batch, M, K).bfloat16()
    Bm = torch.randn(batch, K, N).bfloat16() if transB else torch.randn(batch, N, K).bfloat16()
    opA = A.float().transpose(1, 2) if transA else A.float()
    opB = Bm.float() if transB else Bm.float().transpose(1, 2)
    ref = 0.25 * (opA @ opB)
    out = torch.zeros(batch, M, N, device=DEV)
    ops.gemm_bf16_tc(A.to(DEV), Bm.to(DEV), out, M=M, N=N, K=K, lda=A.shape[2], ldb=Bm.shape[2], ldc=N, alpha=0.25, batch=batch,
                     strideA=A[0].numel(), strideB=Bm[0].numel(), strideC=M * N, transA=transA, transB=transB)
    assert relmax(out.cpu(), ref) < 1e-5


def test_softmax_bwd_rows_bf16(ops):
    torch.manual_seed(16)
    rows, cols = 37, 1024
    P = torch.softmax(torch.randn(rows, cols) * 2, -1).bfloat16()
    dP = torch.randn(rows, cols)
    ref = P.float() * (dP - (dP * P.float()).sum(-1, keepdim=True))
    dS = torch.empty(rows, cols, dtype=torch.bfloat16, device=DEV)
    from diffsci_b200._lib import lib, check, ptr, stream
    Pd, dPd = P.to(DEV), dP.to(DEV)
    check(lib.dsk_softmax_bwd_rows_bf16(ptr(Pd), ptr(dPd), ptr(dS), rows, cols, stream()))
    assert relmax(dS.float().cpu(), ref) < 6e-3


@pytest.mark.parametrize("B,L,C", [(2, 128, 64), (3, 200, 128), (1, 1024, 256), (2, 520, 64)])
def test_attn_softmax_qk_fused(ops, B, L, C):
    """dsk_attn_softmax_qk (two QK^T passes, statistics + exp in the GEMM epilogues) vs softmax of the fp32 product of the same
    bf16 operands; ragged L (not a multiple of the 256-column tile), several column tiles per row."""
    torch.manual_seed(17)
    qkv = (torch.randn(B * L, 3 * C) * 1.5).bfloat16()
    q, k = qkv.float().view(B, L, 3 * C)[..., :C], qkv.float().view(B, L, 3 * C)[..., C:2 * C]
    ref = torch.softmax(q @ k.transpose(1, 2) * C ** -0.5, -1)
    bufs = ops.attention_tc_buffers(B, L, C, DEV)
    bufs["probs"].fill_(float("nan"))
    P = ops.attn_softmax_qk(qkv.to(DEV), bufs["probs"], bufs["rowstat"], B, L, C)
    assert relmax(P.float().cpu(), ref) < 6e-3                       # one bf16 rounding of the stored probability
    assert float((P.float().sum(-1) - 1).abs().max()) < 2e-2


def _attn_ref(qkv, B, L, C):
    t = qkv.float().view(B, L, 3 * C)
    q, k, v = t[..., :C], t[..., C:2 * C], t[..., 2 * C:]
    return torch.softmax(q @ k.transpose(1, 2) * C ** -0.5, -1) @ v


@pytest.mark.parametrize("B,L,C,dtype", [(2, 256, 256, torch.float16), (1, 1024, 256, torch.bfloat16), (3, 200, 128, torch.float16),
                                         (2, 136, 256, torch.float16), (1, 128, 128, torch.bfloat16), (2, 640, 128, torch.float16),
                                         (1, 4096, 256, torch.float16)])
def test_attn_flash(ops, B, L, C, dtype):
    """dsk_attn_flash (scores / probabilities / output accumulator in TMEM, P as the TMEM A operand) vs softmax(QK^T/sqrt(C)) V in
    fp32 on the same 16-bit operands: ragged L (key masking in the last tile, partial query tiles), one and many key tiles,
    both channel counts, 16-bit / fp32 / split outputs.  Tolerance: one 16-bit rounding of the probabilities (and of the
    stored output for the 16-bit mode) relative to the output range."""
    torch.manual_seed(31)
    qkv = (torch.randn(B * L, 3 * C) * 1.5).to(dtype)
    ref = _attn_ref(qkv, B, L, C)
    tol = 1.5e-3 if dtype == torch.float16 else 1e-2
    q = qkv.to(DEV)
    o32 = torch.full((B * L, C), float("nan"), device=DEV)
    ops.attn_flash(q, o32, B, L, C)
    assert relmax(o32.cpu().view(B, L, C), ref) < tol, relmax(o32.cpu().view(B, L, C), ref)
    o16 = torch.full((B * L, C), float("nan"), dtype=dtype, device=DEV)
    ops.attn_flash(q, o16, B, L, C)
    assert relmax(o16.float().cpu().view(B, L, C), ref) < tol + (6e-4 if dtype == torch.float16 else 5e-3)
    assert torch.equal(o16, o32.to(dtype))                           # the same accumulator, rounded once
    if dtype == torch.float16:
        os_ = torch.full((B * L, 2 * C), float("nan"), dtype=dtype, device=DEV)
        ops.attn_flash(q, os_, B, L, C)
        assert torch.equal(os_[:, :C], o16)
        rec = os_[:, :C].float() + os_[:, C:].float() / 2048.0
        assert float((rec - o32).abs().max()) <= 2.0 ** -21 * float(o32.abs().max())


def test_attn_flash_moving_maximum(ops):
    """Scores that grow along the key axis by far more than the lazy-maximum threshold (2^8): every row's reference maximum
    has to move, several times, and the accumulated output has to be rescaled in TMEM each time."""
    torch.manual_seed(32)
    B, L, C = 2, 1024, 256
    t = torch.randn(B, L, 3 * C)
    t[..., :C] *= 2.0
    t[..., C:2 * C] *= (0.25 + 6.0 * torch.arange(L).view(1, L, 1) / L)          # key norms grow 25-fold along the sequence
    qkv = t.view(B * L, 3 * C).half()
    ref = _attn_ref(qkv, B, L, C)
    s = (qkv.float().view(B, L, 3 * C)[..., :C] @ qkv.float().view(B, L, 3 * C)[..., C:2 * C].transpose(1, 2)) * C ** -0.5
    first, last = s[..., :128].amax(-1), s.amax(-1)
    assert float(((last - first) * 1.4427 > 8).float().mean()) > 0.9               # the test does exercise the rescale path
    out = torch.full((B * L, C), float("nan"), device=DEV)
    ops.attn_flash(qkv.to(DEV), out, B, L, C)
    assert relmax(out.cpu().view(B, L, C), ref) < 1.5e-3, relmax(out.cpu().view(B, L, C), ref)


@pytest.mark.parametrize("split", [False, True])
def test_self_attention_flash_block(ops, split):
    """The whole attention block on the flash core (projection GEMMs + dsk_attn_flash) vs the oracle's MultiheadAttention
    restatement: fp16 tokens (16-bit modes) and fp32 tokens with split projections (fp16x2 / fp16x2m modes)."""
    from oracle import nets_oracle as N
    torch.manual_seed(33)
    B, C, sp = 2, 256, (16, 24)
    Lq = sp[0] * sp[1]
    x = torch.randn(B, C, *sp)
    if not split:
        x = x.half().float()
    sd = {"a.mhattn.in_proj_weight": torch.randn(3 * C, C) / math.sqrt(C), "a.mhattn.in_proj_bias": torch.randn(3 * C) * 0.1,
          "a.mhattn.out_proj.weight": torch.randn(C, C) / math.sqrt(C), "a.mhattn.out_proj.bias": torch.randn(C) * 0.1}
    if not split:
        sd = {k: (v.half().float() if k.endswith("weight") else v) for k, v in sd.items()}
    fmt = ops.SPLIT if split else torch.float16
    tok = x.reshape(B, C, Lq).permute(0, 2, 1).contiguous().to(DEV)
    tok = tok if split else tok.half()
    wi = ops.PackedLinear(sd["a.mhattn.in_proj_weight"].to(DEV), fmt)
    wo = ops.PackedLinear(sd["a.mhattn.out_proj.weight"].to(DEV), fmt)
    bufs = ops.attention_flash_buffers(B, Lq, C, DEV, torch.float16, split=split)
    fn = ops.self_attention_flash_split if split else ops.self_attention_flash
    for residual in (False, True):
        ref = N.mha_self_attention(x, sd, "a.", residual)
        out = torch.empty(B, Lq, C, dtype=tok.dtype, device=DEV)
        fn(tok, wi, sd["a.mhattn.in_proj_bias"].to(DEV), wo, sd["a.mhattn.out_proj.bias"].to(DEV), bufs, out, residual)
        got = out.float().cpu().permute(0, 2, 1).reshape(x.shape)
        # split: Q|K|V and P rounded once to fp16, everything else fp32-class; 16-bit: + fp16 storage of the block's output
        assert relmax(got, ref) < (1e-3 if split else 3e-3), relmax(got, ref)


def test_attention_tcgen05(ops):
    from oracle import nets_oracle as N
    torch.manual_seed(14)
    B, C, sp = 2, 64, (8, 16)
    Lq = sp[0] * sp[1]
    x = torch.randn(B, C, *sp).bfloat16().float()
    sd = {"a.mhattn.in_proj_weight": (torch.randn(3 * C, C) / math.sqrt(C)).bfloat16().float(),
          "a.mhattn.in_proj_bias": torch.randn(3 * C) * 0.1,
          "a.mhattn.out_proj.weight": (torch.randn(C, C) / math.sqrt(C)).bfloat16().float(),
          "a.mhattn.out_proj.bias": torch.randn(C) * 0.1}
    bf = dict(dtype=torch.bfloat16, device=DEV)
    bufs = ops.attention_tc_buffers(B, Lq, C, DEV)
    tok = x.reshape(B, C, Lq).permute(0, 2, 1).contiguous().bfloat16().to(DEV)
    wi = ops.PackedLinear(sd["a.mhattn.in_proj_weight"].to(DEV))
    wo = ops.PackedLinear(sd["a.mhattn.out_proj.weight"].to(DEV))
    for residual in (False, True):
        ref = N.mha_self_attention(x, sd, "a.", residual)
        out = torch.empty(B, Lq, C, **bf)
        ops.self_attention_tc(tok, wi, sd["a.mhattn.in_proj_bias"].to(DEV), wo, sd["a.mhattn.out_proj.bias"].to(DEV), bufs, out,
                              residual)
        got = out.float().cpu().permute(0, 2, 1).reshape(x.shape)
        # bf16 storage of Q|K|V, P and the attention output: stated tolerance 2e-2 of the output range
        assert relmax(got, ref) < 2e-2, relmax(got, ref)


@pytest.mark.parametrize("ndim,B,Cin,Cout,sp", [(3, 1, 64, 64, (2, 16, 8)), (3, 2, 128, 64, (3, 10, 12)), (2, 3, 64, 128, (16, 8)),
                                                 (2, 2, 128, 64, (7, 9)), (3, 1, 256, 128, (4, 8, 8)),
                                                 (2, 4, 128, 64, (7, 7)), (3, 2, 64, 64, (4, 8, 8))])   # last two: pairs along the plane axis
def test_upconv_subpixel_tcgen05(ops, ndim, B, Cin, Cout, sp):
    """conv3(nearest_up2(x)) in sub-pixel (phase-decomposed, pre-summed taps) form on tcgen05 vs ATen on the same bf16 inputs.
    The pre-summed weights are rounded to bf16 after summation, so the tolerance is 2 bf16 roundings of the range."""
    torch.manual_seed(21)
    x = torch.randn(B, Cin, *sp).bfloat16().float()
    w = (torch.randn(Cout, Cin, *([3] * ndim)) / math.sqrt(Cin * 3 ** ndim)).bfloat16().float()
    b = torch.randn(Cout) * 0.1
    ref = (F.conv2d if ndim == 2 else F.conv3d)(F.interpolate(x, scale_factor=2, mode="nearest"), w, b, padding=1)
    res = torch.randn_like(ref).bfloat16().float()
    pc = ops.PackedConv(w.to(DEV), b.to(DEV), ndim, torch.bfloat16, subpixel=True)
    y = ops.conv(to_cl(x).bfloat16(), pc, residual=to_cl(res).bfloat16(), up2=True)
    assert y.shape[1:4] == tuple((1,) + tuple(2 * s for s in sp)) if ndim == 2 else tuple(2 * s for s in sp)
    assert relmax(from_cl(y, ndim), ref + res) < 1e-2


@pytest.mark.parametrize("ndim,B,Cin,Cout,sp", [(3, 2, 64, 1, (4, 16, 8)), (3, 1, 64, 3, (3, 10, 12)), (2, 3, 128, 2, (20, 12)),
                                                (3, 1, 64, 1, (32, 32, 32)), (2, 5, 64, 1, (28, 28)), (3, 2, 128, 1, (5, 7, 9)),
                                                (2, 2, 64, 4, (16, 16))])
def test_convout_tcgen05(ops, ndim, B, Cin, Cout, sp):
    """convout (few output channels): the taps-as-N tensor-core kernel (convout_tc.cu; Cout <= 3) and the N = 16 tile it
    falls back to (Cout = 4); fp32 NC(D)HW and bf16 channels-last outputs."""
    torch.manual_seed(22)
    x = torch.randn(B, Cin, *sp).bfloat16().float()
    w = (torch.randn(Cout, Cin, *([3] * ndim)) / math.sqrt(Cin * 3 ** ndim)).bfloat16().float()
    b = torch.randn(Cout) * 0.1
    ref = (F.conv2d if ndim == 2 else F.conv3d)(x, w, b, padding=1)
    pc = ops.PackedConv(w.to(DEV), b.to(DEV), ndim, torch.bfloat16)
    y = ops.conv(to_cl(x).bfloat16(), pc, out_nchw=True)
    assert y.dtype == torch.float32 and relmax(y.cpu(), ref) < 2e-5       # fp32 accumulation, fp32 store
    y2 = ops.conv(to_cl(x).bfloat16(), pc)
    assert relmax(from_cl(y2, ndim), ref) < 6e-3


@pytest.mark.parametrize("B,Cin,Cout,sp,up2", [(2, 64, 64, (4, 16, 16), False), (1, 128, 128, (6, 32, 16), False),
                                               (3, 64, 128, (2, 16, 32), False), (2, 128, 64, (4, 8, 16), True),
                                               (1, 64, 64, (64, 64, 64), False), (5, 64, 64, (32, 48), False),
                                               (3, 64, 64, (5, 32, 16), False), (2, 128, 64, (6, 32, 32), False),   # merged depth taps
                                               (2, 128, 256, (16, 16), False), (3, 128, 64, (8, 16), True)])
def test_conv_fused_norm_statistics(ops, B, Cin, Cout, sp, up2):
    """dsk_conv_fwd_stats: the per-(sample, channel) sum / sum of squares left by the cta_group::2 conv epilogue equal the
    statistics of the stored output, and dsk_norm_act_prestat reproduces dsk_norm_act on it (3-D and 2-D)."""
    torch.manual_seed(41)
    nd = len(sp)
    full = sp if nd == 3 else (1,) + tuple(sp)
    x = torch.randn(B, *full, Cin, device=DEV).bfloat16()
    w = torch.randn(Cout, Cin, *([3] * nd), device=DEV) / math.sqrt(3 ** nd * Cin)
    pc = ops.PackedConv(w, torch.randn(Cout, device=DEV), nd, torch.bfloat16, subpixel=up2)
    if nd == 2 and not ops.conv_stats_supported(tuple(x.shape), x.dtype, pc, up2=up2):
        pytest.skip("fused statistics are enabled for 3-D convolutions only (set DSK_CONV_STATS_2D=1 to test the 2-D epilogue)")
    assert ops.conv_stats_supported(tuple(x.shape), x.dtype, pc, up2=up2)
    cb = torch.randn(B, Cout, device=DEV)
    osp = tuple(2 * s for s in sp) if up2 else tuple(sp)
    osp = osp if nd == 3 else (1,) + osp
    res = torch.randn(B, *osp, Cout, device=DEV).bfloat16()
    st = ops.conv_stats_buffer(B, Cout, DEV)
    st.fill_(float("nan"))                                   # every slot must be written (or zero-filled) by the launch
    y = ops.conv(x, pc, chan_bias=None if up2 else cb, residual=res, up2=up2, stats=st)
    y0 = ops.conv(x, pc, chan_bias=None if up2 else cb, residual=res, up2=up2)
    assert torch.equal(y, y0)
    tot = st.double().sum(1)                                 # [B, Cout, 2]
    yf = y.double().flatten(1, 3)
    # statistics are taken on the fp32 values before the bf16 rounding of the store
    assert relmax(tot[..., 0], yf.sum(1)) < 2e-3
    assert relmax(tot[..., 1], (yf * yf).sum(1)) < 2e-3
    g, b_ = torch.randn(Cout, device=DEV), torch.randn(Cout, device=DEV)
    for mode in (0, 1):
        ref = ops.norm_act(y, g, b_, Cout, mode, True)
        got = ops.norm_act(y, g, b_, Cout, mode, True, conv_stats=st)
        assert relmax(got.float(), ref.float()) < 1e-2
    # fp32-storage modes: fp16 operands, fp32 output and fp32 residual -- residual add and statistics happen in the STORE mapping
    # of the epilogue (coalesced residual reads, reduce-scatter over the lanes of a chunk); the sums are of the stored fp32 values
    xh = x.half()
    pch = ops.PackedConv(w, pc.bias, nd, torch.float16, subpixel=up2)
    res32 = torch.randn(B, *osp, Cout, device=DEV)
    out32 = torch.empty(B, *osp, Cout, device=DEV)
    st.fill_(float("nan"))
    y32 = ops.conv(xh, pch, out=out32, chan_bias=None if up2 else cb, residual=res32, up2=up2, stats=st)
    y32_0 = ops.conv(xh, pch, out=torch.empty_like(out32), chan_bias=None if up2 else cb, residual=res32, up2=up2)
    assert torch.equal(y32, y32_0)
    tot = st.double().sum(1)
    yf = y32.double().flatten(1, 3)
    assert relmax(tot[..., 0], yf.sum(1)) < 1e-5
    assert relmax(tot[..., 1], (yf * yf).sum(1)) < 1e-5
    ref32 = ops.norm_act(y32, g, b_, Cout, 0, True, out=torch.empty(y32.shape, dtype=torch.float16, device=DEV))
    got32 = ops.norm_act(y32, g, b_, Cout, 0, True, out=torch.empty(y32.shape, dtype=torch.float16, device=DEV), conv_stats=st)
    assert relmax(got32.float(), ref32.float()) < 2e-3


@pytest.mark.parametrize("ndim,B,Cin,Cout,sp", [(3, 2, 1, 64, (5, 16, 8)), (3, 1, 2, 64, (4, 9, 11)), (2, 3, 3, 128, (20, 12)),
                                                (2, 4, 1, 128, (28, 28)), (3, 1, 1, 64, (32, 32, 32)), (2, 2, 4, 256, (7, 9))])
def test_convin_tcgen05(ops, ndim, B, Cin, Cout, sp):
    """convin (Cin <= 4) as an im2col-row tcgen05 GEMM (convin_tc.cu) vs ATen on the same bf16-rounded operands."""
    torch.manual_seed(23)
    x = torch.randn(B, Cin, *sp).bfloat16().float()
    w = (torch.randn(Cout, Cin, *([3] * ndim)) / math.sqrt(Cin * 3 ** ndim))
    b = torch.randn(Cout) * 0.1
    ref = (F.conv2d if ndim == 2 else F.conv3d)(x, w.bfloat16().float(), b, padding=1)
    pc = ops.PackedConv(w.to(DEV), b.to(DEV), ndim, torch.float32)        # the few-channel path keeps fp32 packed weights
    y = ops.conv(to_cl(x).bfloat16(), pc)
    assert y.dtype == torch.bfloat16 and relmax(from_cl(y, ndim), ref) < 6e-3


@pytest.mark.parametrize("ndim,B,Cin,Cout,sp", [(3, 2, 1, 64, (5, 16, 8)), (3, 1, 2, 64, (4, 9, 11)), (2, 3, 3, 128, (20, 12)),
                                                (2, 2, 1, 64, (28, 28)), (2, 2, 4, 256, (7, 9))])
def test_convin_fp32_storage_paths(ops, ndim, B, Cin, Cout, sp):
    """The first layer of the fp32-storage modes: (1) fp32 in / fp32 out on the CUDA cores (pixel-major few-input-channel kernel)
    vs ATen fp32; (2) the same tensors with operand16 = fp16: im2col on the tensor cores with operands rounded to fp16 while
    gathering (convin_tc.cu, fp32 output through the staged epilogue) vs ATen on fp16-rounded operands."""
    torch.manual_seed(24)
    x = torch.randn(B, Cin, *sp)
    w = torch.randn(Cout, Cin, *([3] * ndim)) / math.sqrt(Cin * 3 ** ndim)
    b = torch.randn(Cout) * 0.1
    conv = F.conv2d if ndim == 2 else F.conv3d
    pc = ops.PackedConv(w.to(DEV), b.to(DEV), ndim, torch.float32)
    y = ops.conv(to_cl(x), pc)
    assert y.dtype == torch.float32 and relmax(from_cl(y, ndim), conv(x, w, b, padding=1)) < 2e-5
    y16 = ops.conv(to_cl(x), pc, operand16=torch.float16)
    ref16 = conv(x.half().float(), w.half().float(), b, padding=1)
    assert y16.dtype == torch.float32 and relmax(from_cl(y16, ndim), ref16) < 2e-5
    # (3) operand16 = SPLIT: [x_hi | x_lo | x_hi] . [w_hi | w_hi | w_lo] in one im2col row where 3 * taps * Cin <= 128 (else the
    # CUDA-core kernel answers): fp32-class against the fp32 reference
    ys = ops.conv(to_cl(x), pc, operand16=ops.SPLIT)
    assert ys.dtype == torch.float32 and relmax(from_cl(ys, ndim), conv(x, w, b, padding=1)) < 2e-5
    # 16-bit in / out through the same kernel (two epilogue warp sets, staged stores)
    yh = ops.conv(to_cl(x).half(), pc)
    assert yh.dtype == torch.float16 and relmax(from_cl(yh, ndim), ref16) < 1e-3


@pytest.mark.parametrize("ndim,B,Cin,Cout,sp,up2", [(3, 2, 64, 64, (4, 16, 16), False), (3, 1, 128, 128, (4, 16, 16), False),
                                                    (2, 3, 64, 128, (16, 16), False), (3, 1, 128, 64, (2, 8, 8), True),
                                                    (2, 2, 64, 64, (7, 9), False)])
def test_conv_tc_fp32_residual_16bit_out(ops, ndim, B, Cin, Cout, sp, up2):
    """dsk_conv_desc.res_dtype = DSK_RES_F32: fp16 operands, fp32 residual added in the epilogue, fp16 OUTPUT (the operand copy a
    block of an fp32-storage mode writes for the single convolution that reads it) -- CTA-pair kernel and the single-CTA
    kernel (odd tile counts), with the per-warp staged bias + chan_bias."""
    torch.manual_seed(25)
    x = torch.randn(B, Cin, *sp).half().float()
    w = (torch.randn(Cout, Cin, *([3] * ndim)) / math.sqrt(Cin * 3 ** ndim)).half().float()
    b = torch.randn(Cout) * 0.1
    cb = torch.randn(B, Cout) * 0.3
    conv = F.conv2d if ndim == 2 else F.conv3d
    xr = F.interpolate(x, scale_factor=2, mode="nearest") if up2 else x
    ref = conv(xr, w, b, padding=1) + cb.view(B, Cout, *([1] * ndim))
    res = torch.randn_like(ref)                                            # fp32 residual, not representable in fp16
    pc = ops.PackedConv(w.to(DEV), b.to(DEV), ndim, torch.float16, subpixel=up2)
    out = torch.empty(to_cl(ref).shape, dtype=torch.float16, device=DEV)
    y = ops.conv(to_cl(x).half(), pc, out=out, chan_bias=cb.to(DEV), residual=to_cl(res), up2=up2)
    tol = 1.2e-3 if not up2 else 3e-3                                      # one fp16 rounding of the output (sub-pixel: + summed taps)
    assert y.dtype == torch.float16 and relmax(from_cl(y, ndim), ref + res) < tol
    y32 = ops.conv(to_cl(x).half(), pc, out=torch.empty_like(out, dtype=torch.float32), chan_bias=cb.to(DEV), residual=to_cl(res),
                   up2=up2)
    assert relmax(from_cl(y32, ndim), ref + res) < (2e-5 if not up2 else 2e-3)
    assert torch.equal(y, y32.half())                                      # the same accumulator + residual, rounded once


@pytest.mark.parametrize("ndim,sp", [(3, (4, 6, 8)), (2, (10, 12))])
@pytest.mark.parametrize("is_max", [True, False])
def test_pool2x_fp32_to_operand_copy(ops, ndim, sp, is_max):
    """dsk_pool2x_f32: float4 pooling of an fp32 tensor into fp32, or rounded once into the fp16 / bf16 operand copy."""
    torch.manual_seed(26)
    x = torch.randn(2, 64, *sp)
    pool = (F.max_pool3d if ndim == 3 else F.max_pool2d) if is_max else (F.avg_pool3d if ndim == 3 else F.avg_pool2d)
    ref = pool(x, 2)
    xc = to_cl(x)
    y = ops.pool2x(xc, ndim, is_max)
    assert y.dtype == torch.float32 and relmax(from_cl(y, ndim), ref) < 1e-6
    for dt in (torch.float16, torch.bfloat16):
        o = torch.empty(y.shape, dtype=dt, device=DEV)
        ops.pool2x(xc, ndim, is_max, out=o)
        assert torch.equal(o, y.to(dt))
