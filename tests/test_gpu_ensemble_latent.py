"""GPU parity tests of SURVEY 8(f)-4: the ensemble losses of EnsembleKarrasModule (karrasmodule_new.py:963-1149 with the
ensemble-aware Huber / MSE / CRPS metrics, custom_losses.py:536-690, 765-865) and the latent-diffusion wrapper of
KarrasModule (karrasmodule.py:583-587, 893-896, 1192-1234) -- against goldens recorded from the LIVE reference
(oracle/make_goldens.py --only ensemble) and against the CPU oracle, through the C ABI."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
KINDS = {"huber": 0, "mse": 1, "CRPS": 2}


def relmax(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("metric", ["huber", "mse", "CRPS"])
@pytest.mark.parametrize("B,E,C,spatial,mask_c", [
    (3, 1, 1, (8, 8), 0), (2, 2, 3, (5, 7), 1), (4, 3, 2, (16, 16), 2), (2, 5, 1, (4, 4, 4), 1), (3, 8, 2, (12, 12), 0),
    (2, 9, 1, (8, 8), 1), (1, 16, 2, (6, 6), 2), (5, 4, 1, (64, 64), 0), (2, 7, 4, (3, 3), 4)])
def test_ensemble_loss_kernel_vs_oracle(metric, B, E, C, spatial, mask_c):
    """dsk_ensemble_loss_fwd_bwd (one launch: D, metric, reduction, dL/dF) vs oracle.ensemble_metric + autograd in fp64, for
    both access widths (C*S % 4), every ensemble-size template up to 16, broadcast (1-channel) and per-channel masks."""
    from diffsci_b200.models.karras.karrasmodule_new import _ensemble_scales, _EnsembleLossFn
    from oracle import karras_oracle as K
    torch.manual_seed(B * 100 + E)
    x = torch.randn(B, C, *spatial) * 0.5
    noise = torch.randn(B, E, C, *spatial)
    sigma = torch.exp(torch.randn(B) * 1.2 - 1.2)
    F = torch.randn(B * E, C, *spatial)
    mask = None
    if mask_c:
        mask = (torch.rand(B, mask_c, *spatial) > 0.5).float()
        mask[-1] = 1.0
    _, c_out, c_skip, _ = K.edm_precond(sigma.double())
    S = x[0, 0].numel()
    lam_mean = K.edm_loss_weight(sigma).mean()

    # fp64 truth: the metric of the reference applied to D = c_out F + c_skip x_noised, times mean(lambda)
    F64 = F.double().requires_grad_(True)
    xn = (x.double().unsqueeze(1) + K.bcast(sigma.double(), x).unsqueeze(1) * noise.double())
    D = K.bcast(c_out, x).unsqueeze(1) * F64.view(B, E, C, *spatial) + K.bcast(c_skip, x).unsqueeze(1) * xn
    L64 = K.edm_loss_weight(sigma.double()).mean() * K.ensemble_metric(D, x.double(), metric,
                                                                        None if mask is None else mask.double())
    L64.backward()

    Fd = F.to(DEV).requires_grad_(True)
    md = None if mask is None else mask.to(DEV)
    s1, s2, mk = _ensemble_scales(KINDS[metric], lam_mean.to(DEV), B, E, C, S, md)
    L = _EnsembleLossFn.apply(Fd, x.to(DEV), noise.to(DEV).contiguous(), sigma.to(DEV), c_out.float().to(DEV),
                              c_skip.float().to(DEV), s1, s2, mk, KINDS[metric], E)
    L.backward()
    assert abs(float(L) - float(L64)) <= 2e-5 * abs(float(L64)) + 1e-7, (float(L), float(L64))
    # sign() / clamp() gradients flip at |r| ~ ulp: compare in the max norm relative to the largest entry
    assert relmax(Fd.grad.cpu(), F64.grad) < 2e-5


def test_ensemble_noise_add_bit_exact():
    from diffsci_b200._lib import lib, check, ptr, stream
    torch.manual_seed(3)
    B, E, shape = 3, 5, (2, 7, 9)
    x, nz, sg = torch.randn(B, *shape).to(DEV), torch.randn(B, E, *shape).to(DEV), torch.rand(B).to(DEV) + 0.1
    out = torch.empty(B * E, *shape, device=DEV)
    check(lib.dsk_ensemble_noise_add(ptr(x), ptr(nz), ptr(sg), ptr(out), B, E, x[0].numel(), stream()))
    ref = torch.addcmul(x.unsqueeze(1), sg.view(B, 1, 1, 1, 1), nz).reshape(B * E, *shape)    # one fused multiply-add, like the kernel
    assert torch.equal(out, ref) or relmax(out, ref) < 2e-7


@pytest.mark.parametrize("metric", ["huber", "mse", "CRPS"])
def test_ensemble_module_vs_live_reference(golden, metric):
    """EnsembleKarrasModule.loss_fn(n_ensemble=3 | 1, mask | None) -> backward through the native PUNetG: losses and
    parameter gradients recorded from the live reference."""
    import diffsci_b200 as d
    from tests.test_gpu_nets import build_net
    g = golden("ensemble_punetg2d")
    net = build_net(golden(g["net"])).train()
    cfg = d.EnsembleKarrasModuleConfig.from_edm(loss_metric=metric)
    cfg.ensemble_size_train = g["E"]
    mod = d.EnsembleKarrasModule(net, cfg).train()
    for tag in ("E3", "E3_mask", "E1", "E1_mask"):
        ref = g["cases"][f"{metric}_{tag}"]
        single = tag.startswith("E1")
        net.zero_grad()
        mod._injected_ensemble_noise = g["noise1"] if single else g["noise"]
        L = mod.loss_fn(g["x"].to(DEV), g["sigma"].to(DEV), None, g["mask"].to(DEV) if tag.endswith("mask") else None,
                        n_ensemble=1 if single else g["E"])
        L.backward()
        assert abs(float(L) - float(ref["loss"])) < 5e-5 * abs(float(ref["loss"])), (metric, tag, float(L), float(ref["loss"]))
        grads = dict(net.named_parameters())
        for k, gr in ref["grads"].items():
            e = relmax(grads[k].grad, gr)
            assert e < 5e-4, (metric, tag, k, e)


def test_ensemble_module_plain_sizes_fall_back_to_per_sample_weighting(golden):
    """All ensemble sizes at 1: the metric is elementwise and lambda weights each sample (= KarrasModule.loss_fn);
    "CRPS" is then not a recognised name (karrasmodule_new.py:844-846, 877-893)."""
    import diffsci_b200 as d
    from tests.test_gpu_nets import build_net
    g = golden("sampler_punetg2d")
    net = build_net(golden("punetg2d_mc8")).train()
    mod = d.EnsembleKarrasModule(net, d.EnsembleKarrasModuleConfig.from_edm(loss_metric="huber")).train()
    mod._injected_loss_noise = g["loss_noise"]
    L = mod.loss_fn(g["loss_x"].to(DEV), g["loss_sigma"].to(DEV), None, None)
    assert abs(float(L) - float(g["loss_huber"])) < 5e-5 * abs(float(g["loss_huber"]))
    with pytest.raises(ValueError):
        d.EnsembleKarrasModule(net, d.EnsembleKarrasModuleConfig.from_edm(loss_metric="CRPS"))
    cfg = d.EnsembleKarrasModuleConfig.from_edm(loss_metric="CRPS")
    cfg.ensemble_size_train = 32
    big = d.EnsembleKarrasModule(net, cfg).train()
    with pytest.raises(NotImplementedError):
        big.loss_fn(g["loss_x"].to(DEV), g["loss_sigma"].to(DEV), None, None, n_ensemble=32)


def test_ensemble_training_step_bf16_and_ema(golden):
    """training_step with config.ensemble_size_train on the bf16 tensor-core path + the EMA hook: the loss agrees with the fp32
    path within the bf16 budget, every parameter receives a gradient, on_before_zero_grad moves the shadow weights."""
    import diffsci_b200 as d
    from tests.test_gpu_nets import build_net
    g = golden("ensemble_punetg2d")
    losses = {}
    for precision in ("fp32", "bf16"):
        net = build_net(golden(g["net"]), precision).train()
        cfg = d.EnsembleKarrasModuleConfig.from_edm(loss_metric="CRPS", ema_enabled=True, ema_decay=0.5)
        cfg.ensemble_size_train = g["E"]
        mod = d.EnsembleKarrasModule(net, cfg).train()
        mod._injected_ensemble_noise = g["noise"]
        torch.manual_seed(5)
        L = mod.training_step(g["x"].to(DEV), 0)
        L.backward()
        assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in net.parameters())
        losses[precision] = float(L)
        before = {k: v.clone() for k, v in mod.ema_tracker.profiles[0]["params"].items()}
        with torch.no_grad():
            for p in net.parameters():
                p.add_(0.01)
        mod.on_before_zero_grad(None)
        after = mod.ema_tracker.profiles[0]["params"]
        assert mod.ema_tracker.num_updates == 1 and any(not torch.equal(before[k], after[k]) for k in before)
    assert losses["fp32"] == losses["fp32"] and abs(losses["bf16"] - losses["fp32"]) < 3e-2 * abs(losses["fp32"])


# ----------------------------------------------------------------------------------------------- latent-diffusion wrapper
def test_latent_wrapper_vs_live_reference(golden):
    import diffsci_b200 as d
    from tests.test_gpu_nets import build_net
    from tests.test_oracle_vs_golden import toy_autoencoder, oracle_net
    from oracle import karras_oracle as K
    g = golden("latent_punetg2d")
    net = build_net(golden(g["net"])).train()
    ae = toy_autoencoder().to(DEV)
    mod = d.KarrasModule(net, d.KarrasModuleConfig.from_edm(), autoencoder=ae).train()
    assert mod.latent_model and not any(p.requires_grad for p in ae.parameters())
    mod._injected_loss_noise = g["noise"]
    L = mod.loss_fn(g["x"].to(DEV), g["sigma"].to(DEV))
    L.backward()
    assert abs(float(L) - float(g["loss"])) < 5e-5 * abs(float(g["loss"]))
    grads = dict(net.named_parameters())
    for k, gr in g["grads"].items():
        assert relmax(grads[k].grad, gr) < 3e-4, k
    assert all(p.grad is None for p in ae.parameters())
    mod.eval(), net.eval()
    wn = g["white_noise"]
    net64 = oracle_net(golden(g["net"]), torch.float64)
    lat64 = K.sample_from_white_noise(net64, wn.double(), 4, "heun")
    budget = 2.0 * relmax(g["sample_latent"].double(), lat64) + 2e-5
    lat = mod.propagate_white_noise(wn.to(DEV), nsteps=4, return_in_latent_space=True).cpu()
    assert relmax(lat, g["sample_latent"]) <= budget
    dec = mod.propagate_white_noise(wn.to(DEV), nsteps=4).cpu()
    assert dec.shape == g["sample_decoded"].shape == (2, 1, 32, 32)
    assert relmax(dec, g["sample_decoded"]) <= budget
    hist = mod.propagate_white_noise(wn.to(DEV), nsteps=4, record_history=True).cpu()
    assert hist.shape == g["sample_hist_decoded"].shape and relmax(hist, g["sample_hist_decoded"]) <= budget
    torch.manual_seed(77)
    api = mod.sample(2, [1, 16, 16], nsteps=4, is_latent_shape=True).cpu()
    torch.manual_seed(77)                       # its own x_T, hence its own fp64 budget (chained random-weight evaluations)
    wn2 = torch.randn(2, 1, 16, 16)
    with torch.no_grad():
        truth = toy_autoencoder().double().decode(K.sample_from_white_noise(net64, wn2.double(), 4, "heun"))
    assert relmax(api, g["sample_api"]) <= 2.0 * relmax(g["sample_api"], truth) + 2e-5
    # a data-space shape: the latent shape comes from the encoder
    out = mod.sample(3, [1, 32, 32], nsteps=3)
    assert out.shape == (3, 1, 32, 32) and torch.isfinite(out).all()
    assert mod.sample(3, [1, 32, 32], nsteps=3, return_in_latent_space=True).shape == (3, 1, 16, 16)
    with pytest.raises(ValueError):       # encode_y needs a conditional autoencoder whose encode(x, y) returns (z, y')
        d.KarrasModule(net, d.KarrasModuleConfig.from_edm(), autoencoder=ae, encode_y=True)


def test_ema_weights_for_sampling(golden):
    """karrasmodule_new.py:1400-1405, 2208-2227: in eval mode sample() runs on the EMA weights (apply_to, sample, restore) --
    with the engine cached from an earlier run on the live weights."""
    import diffsci_b200 as d
    from tests.test_gpu_nets import build_net
    g = golden("punetg2d_mc8")
    net = build_net(g, "bf16")
    cfg = d.EnsembleKarrasModuleConfig.from_edm(ema_enabled=True, ema_decay=0.0)
    mod = d.EnsembleKarrasModule(net, cfg)
    shape = list(g["x"].shape[1:])

    def draw(use_ema=None):
        torch.manual_seed(21)
        return mod.sample(2, shape, nsteps=3, use_ema=use_ema).cpu()
    mod.eval()
    live0 = draw(use_ema=False)
    assert torch.equal(draw(), live0)                         # shadow == live weights right after construction
    mod.train()
    with torch.no_grad():
        for k, p in net.named_parameters():
            if k.endswith("weight") and p.ndim >= 4:
                p.mul_(0.5)
    mod.on_before_zero_grad(None)                             # decay 0: the shadow becomes the updated weights ...
    with torch.no_grad():
        for k, p in net.named_parameters():
            if k.endswith("weight") and p.ndim >= 4:
                p.mul_(2.0)                                   # ... and the live weights go back
    before = {k: v.clone() for k, v in net.state_dict().items()}
    mod.eval()
    ema = draw()
    assert not torch.equal(ema, live0) and torch.equal(draw(use_ema=False), live0)
    assert all(torch.equal(v, before[k]) for k, v in net.state_dict().items())      # restored
    mod.train()
    assert torch.equal(draw(), live0)                         # training mode: live weights


@pytest.mark.parametrize("metric", ["huber", "mse"])
def test_dynamic_loss_weight_vs_live_reference(golden, metric):
    """KarrasModuleConfig.from_edm(dynamic_loss_weight=n): loss = mean(lambda exp(-u) l + u), u = DynamicLossWeight(c_noise);
    per-sample loss sums from dsk_precond_loss_rows; gradients of the network AND of the uncertainty head vs the live reference."""
    import diffsci_b200 as d
    from tests.test_gpu_nets import build_net
    g = golden("dynweight_punetg2d")
    net = build_net(golden(g["net"])).train()
    mod = d.KarrasModule(net, d.KarrasModuleConfig.from_edm(dynamic_loss_weight=g["nhidden"], loss_metric=metric)).train()
    assert list(mod.dynamic_loss_weight.state_dict().keys()) == list(g["dlw_state"].keys())
    mod.dynamic_loss_weight.load_state_dict(g["dlw_state"])
    mod.dynamic_loss_weight.to(DEV)
    for tag in ("", "_mask"):
        ref = g[metric + tag]
        net.zero_grad()
        mod.dynamic_loss_weight.zero_grad()
        mod._injected_loss_noise = g["noise"]
        L = mod.loss_fn(g["x"].to(DEV), g["sigma"].to(DEV), None, g["mask"].to(DEV) if tag else None)
        L.backward()
        assert abs(float(L) - float(ref["loss"])) < 5e-5 * abs(float(ref["loss"])), (tag, float(L), float(ref["loss"]))
        grads = dict(net.named_parameters())
        for k, gr in ref["grads"].items():
            assert relmax(grads[k].grad, gr) < 3e-4, (tag, k)
        for k, p in mod.dynamic_loss_weight.named_parameters():
            assert relmax(p.grad, ref["dlw_grads"][k]) < 5e-5, (tag, k, relmax(p.grad, ref["dlw_grads"][k]))
