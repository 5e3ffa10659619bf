"""Pin the CPU oracle (oracle/*.py) against golden vectors produced by the LIVE reference
(oracle/make_goldens.py).  Runs without a GPU."""
import types

import pytest
import torch

from oracle import karras_oracle as K
from oracle import nets_oracle as N

TOL32 = 2e-5   # oracle fp32 vs reference fp32: same ATen ops, op-order rounding only


def relmax(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def cfg_for(kind, kw):
    if kind == "punetg":
        d = dict(input_channels=1, output_channels=1, dimension=2, model_channels=64, channel_expansion=[2, 4],
                 number_resnet_downward_block=2, number_resnet_upward_block=2, number_resnet_attn_block=2,
                 number_resnet_before_attn_block=2, number_resnet_after_attn_block=2, transition_scale_factor=2,
                 first_resblock_norm="GroupLN", second_resblock_norm="GroupRMS", affine_norm=True,
                 attn_residual=False, bias=True)
    else:
        d = dict(input_channels=1, output_channels=1, dimension=2, model_channels=64, time_embed_dim=64,
                 output_embed_dim=256, channel_expansion=[2, 4], number_resnet_downward_block=2,
                 number_resnet_upward_block=2, number_resnet_attn_block=2, number_resnet_before_attn_block=2,
                 number_resnet_after_attn_block=2, kernel_size=3, transition_scale_factor=2,
                 first_resblock_norm="GroupLN", second_resblock_norm="GroupRMS", num_groups=1,
                 skip_integration_type="concat", attn_residual=True, decoder_type=1)
    d.update(kw)
    c = types.SimpleNamespace(**d)
    if kind == "adm":
        c.middle_block_attn_config = ([False] * c.number_resnet_before_attn_block +
                                      [True] * (c.number_resnet_attn_block - 1) + [False] +
                                      [False] * c.number_resnet_after_attn_block)
    return c


def oracle_net(g, dtype=torch.float32):
    sd = N.synth_state_dict(g["manifest"], g["seed"], dtype)
    if g["kind"] == "punetg":
        cfg = cfg_for("punetg", g["cfg"])
        return lambda x, t: N.punetg_forward(sd, cfg, x, t)
    if g["kind"] == "adm":
        cfg = cfg_for("adm", g["cfg"])
        return lambda x, t: N.adm_forward(sd, cfg, x, t)
    return lambda x, t: N.mlp_uncond_forward(sd, x, t, g["cfg"]["act"])


def test_schedule_and_scalars(golden):
    g = golden("numerics")
    for n, ref in g["steps"].items():
        assert torch.equal(K.edm_steps(n), ref), n
    c_in, c_out, c_skip, c_noise = K.edm_precond(g["sigma"])
    for a, b in ((c_in, "c_in"), (c_out, "c_out"), (c_skip, "c_skip"), (c_noise, "c_noise")):
        assert torch.equal(a, g[b]), b
    assert torch.equal(K.edm_loss_weight(g["sigma"]), g["loss_weight"])
    assert torch.equal(K.edm_sigma_from_normal(g["sigma_xi"]), g["sigma_from_xi"])


def test_ema_betas(golden):
    g = golden("numerics")
    for s, v in g["ema_power_exp"].items():
        assert K.power_function_exp_from_std(s) == v
    for (s, n), v in g["ema_power_beta"].items():
        assert K.ema_beta("power", n, std=s) == v
    for n, v in g["ema_trad_beta"].items():
        assert K.ema_beta("traditional", n, halflife_steps=100.0, rampup_ratio=0.5) == v
    # known answers of reference tests/test_karras_ema.py:23-52
    assert torch.allclose(K.ema_update(torch.zeros(2), torch.full((2,), 2.0), 0.5), torch.ones(2))
    assert K.ema_beta("power", 1) == 0.0
    e = golden("ema")
    shadow = None
    for step in e["traj"]:
        if shadow is None:
            # ModelEMA.reset clones the params at construction; first shadow = lerp(init, p1, 1-beta)
            # => recover init from the recorded first update.
            shadow = {k: (step["shadow"][k] - (1 - step["beta"]) * step["params"][k]) / step["beta"]
                      for k in step["shadow"]}
        for k in shadow:
            shadow[k] = K.ema_update(shadow[k], step["params"][k], step["beta"])
            assert torch.allclose(shadow[k], step["shadow"][k], atol=1e-6)


@pytest.mark.parametrize("name", ["punetg2d_mc8", "punetg3d_mc8", "punetg2d_multi", "adm2d_mc8", "adm2d_add",
                                  "mlp_silu", "nobias_punetg2d", "nobias_punetg3d"])
def test_net_forward(golden, name):
    g = golden(name)
    y = oracle_net(g)(g["x"], g["t"])
    assert y.shape == g["y"].shape
    assert relmax(y, g["y"]) < TOL32, relmax(y, g["y"])
    y64 = oracle_net(g, torch.float64)(g["x"].double(), g["t"].double())
    assert relmax(y64, g["y64"]) < 1e-12


@pytest.mark.parametrize("name,netname", [("sampler_mlp", "mlp_silu"), ("sampler_punetg2d", "punetg2d_mc8")])
def test_denoiser_samplers_loss(golden, name, netname):
    g = golden(name)
    net = oracle_net(golden(netname))
    n = g["nsteps"]
    assert relmax(K.denoiser(net, g["den_x"], g["den_sigma"]), g["den_D"]) < TOL32
    assert relmax(K.score(net, g["den_x"], g["den_sigma"]), g["den_score"]) < TOL32
    wn = g["white_noise"]
    # N chained evaluations of a random-weight network amplify fp32 rounding chaotically (the reference's own
    # fp32 run sits ~1e-3..1e-2 from an fp64 run of the same chain).  Budget against fp64 truth: the oracle may
    # differ from the reference by at most what the reference itself differs from fp64, plus 2e-5.
    net64 = oracle_net(golden(netname), torch.float64)
    nz64 = [z.double() for z in g["noises"]]

    def check(ref, integrator, **kw):
        o32 = K.sample_from_white_noise(net, wn, n, integrator, **kw)
        kw64 = dict(kw)
        if "noises" in kw64:
            kw64["noises"] = nz64
        o64 = K.sample_from_white_noise(net64, wn.double(), n, integrator, **kw64)
        budget = 2.0 * relmax(ref.double(), o64) + 2e-5
        assert relmax(o32, ref) <= budget, (integrator, relmax(o32, ref), budget)
        return o32

    h = check(g["heun_hist"], "heun", record_history=True)
    assert torch.equal(h[0], wn * 80.0)
    check(g["euler"], "euler")
    check(g["em"], "euler-maruyama", noises=g["noises"])
    check(g["em_interval"], "euler-maruyama", noises=g["noises"], langevin_const=0.5, langevin_interval=(0.1, 10.0))
    check(g["karras"], "karras", noises=g["noises"])
    check(g["karras_custom"], "karras", noises=g["noises"], s_churn=10, s_tmin=0.01, s_tmax=1.0, s_noise=1.0)
    for metric in ("huber", "mse"):
        for use_mask in (False, True):
            L = K.edm_loss(net, g["loss_x"], g["loss_sigma"], g["loss_noise"], metric,
                           g["loss_mask"] if use_mask else None)
            ref = g[f"loss_{metric}{'_mask' if use_mask else ''}"]
            assert abs(float(L) - float(ref)) <= 2e-5 * abs(float(ref)), (metric, use_mask)


def test_sampler_edge_nsteps2(golden):
    g = golden("sampler_edge")
    net = oracle_net(golden("mlp_silu"))
    assert relmax(K.sample_from_white_noise(net, g["white_noise"], 2, "heun"), g["heun2"]) < 2e-5
    # nsteps=1 is outside the reference's domain: create_steps(2) divides by n-2 = 0
    assert not torch.isfinite(K.edm_steps(2)).all()


# ----------------------------------------------------------------------------- conditional path
def cond_oracle_nets(g, dtype=torch.float32, batch=True):
    """-> (net_cond, net_uncond or None) of a tests/golden/cond_*.pt fixture.  batch: per-sample conditions (y_batch)
    or the single condition sample() broadcasts (y_one, unsqueezed as karrasmodule.py:914-915 does)."""
    sd = N.synth_state_dict(g["manifest"], g["seed"], dtype)
    y = g["y_batch"] if batch else {k: v.unsqueeze(0) for k, v in g["y_one"].items()}
    y = {k: v.to(dtype) for k, v in y.items()}
    ye = N.porosity_embedder(sd, "conditional_embedding.", y["porosity"])
    if g.get("kind", "punetg") == "adm":
        acfg = cfg_for("adm", g["cfg"])
        return (lambda x, t: N.adm_forward(sd, acfg, x, t, ye)), (lambda x, t: N.adm_forward(sd, acfg, x, t))
    cfg = cfg_for("punetg", g["cfg"])
    if "cond" in y:
        return (lambda x, t: N.punetg_cond_forward(sd, cfg, x, t, [y["cond"]], ye)), None
    return (lambda x, t: N.punetg_forward(sd, cfg, x, t, ye)), (lambda x, t: N.punetg_forward(sd, cfg, x, t))


@pytest.mark.parametrize("name", ["cond_punetg2d_embed", "cond_punetg3d_chan", "cond_adm2d_embed"])
def test_conditional_path(golden, name):
    g = golden(name)
    nc, nu = cond_oracle_nets(g)
    assert relmax(nc(g["x"], g["t"]), g["net_y"]) < TOL32
    nc64, _ = cond_oracle_nets(g, torch.float64)
    assert relmax(nc64(g["x"].double(), g["t"].double()), g["net_y64"]) < 1e-12
    for key in [k for k in g if k.startswith("den_D_g")]:
        gd = float(key[len("den_D_g"):])
        net = K.guided_net(nc, nu, gd)
        assert relmax(K.denoiser(net, g["den_x"], g["den_sigma"]), g[key]) < TOL32, key
    assert relmax(K.score(nc, g["den_x"], g["den_sigma"]), g["den_score_g1.0"]) < TOL32
    # sampling with ONE condition broadcast over the batch; budget against fp64 truth as above
    n, wn = g["nsteps"], g["white_noise"]
    oc, ou = cond_oracle_nets(g, batch=False)
    oc64, ou64 = cond_oracle_nets(g, torch.float64, batch=False)
    for key in [k for k in g if k.startswith("heun_hist_g")]:
        gd = float(key[len("heun_hist_g"):])
        o32 = K.sample_from_white_noise(K.guided_net(oc, ou, gd), wn, n, "heun", record_history=True)
        o64 = K.sample_from_white_noise(K.guided_net(oc64, ou64, gd), wn.double(), n, "heun", record_history=True)
        assert relmax(o32, g[key]) <= 2.0 * relmax(g[key].double(), o64) + 2e-5, key
    o32 = K.sample_from_white_noise(oc, wn, n, "euler-maruyama", noises=g["noises"])
    o64 = K.sample_from_white_noise(oc64, wn.double(), n, "euler-maruyama", noises=[z.double() for z in g["noises"]])
    assert relmax(o32, g["em_g1.0"]) <= 2.0 * relmax(g["em_g1.0"].double(), o64) + 2e-5
    L = K.edm_loss(nc, g["loss_x"], g["loss_sigma"], g["loss_noise"], "huber")
    assert abs(float(L) - float(g["loss_huber"])) <= 2e-5 * abs(float(g["loss_huber"]))


# ----------------------------------------------------------------------------- circular convolution
@pytest.mark.parametrize("name", ["circ_punetg2d", "circ_punetg3d", "circ_adm2d"])
def test_circular_punetg(golden, name):
    g = golden(name)
    net = oracle_net(g)
    assert relmax(net(g["x"], g["t"]), g["y"]) < TOL32
    net64 = oracle_net(g, torch.float64)
    assert relmax(net64(g["x"].double(), g["t"].double()), g["y64"]) < 1e-12
    n, wn = g["nsteps"], g["white_noise"]
    o32 = K.sample_from_white_noise(net, wn, n, "heun", record_history=True)
    o64 = K.sample_from_white_noise(net64, wn.double(), n, "heun", record_history=True)
    assert relmax(o32, g["heun_hist"]) <= 2.0 * relmax(g["heun_hist"].double(), o64) + 2e-5
    L = K.edm_loss(net, g["loss_x"], g["loss_sigma"], g["loss_noise"], "huber")
    assert abs(float(L) - float(g["loss_huber"])) <= 2e-5 * abs(float(g["loss_huber"]))


# ----------------------------------------------------------------------------- VP / VE / SR3
def precond_case(g, dtype=torch.float32):
    """-> (net with the fixture's scaled last layer, scheduler tag, preconditioner kind)."""
    base = torch.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", g["net"] + ".pt"),
                      weights_only=False)
    sd = N.synth_state_dict(base["manifest"], base["seed"], dtype)
    last = "convout." if base["kind"] == "punetg" else [k for k, _ in base["manifest"] if k.endswith(".weight")][-1][:-6]
    for k in (last + "weight", last + "bias"):
        sd[k] = sd[k] * g["out_scale"]
    if base["kind"] == "punetg":
        cfg = cfg_for("punetg", base["cfg"])
        net = lambda x, t: N.punetg_forward(sd, cfg, x, t)  # noqa: E731
    else:
        net = lambda x, t: N.mlp_uncond_forward(sd, x, t, base["cfg"]["act"])  # noqa: E731
    tag = "edm" if g["tag"] == "sr3" else g["tag"]
    return net, tag, g["tag"]


@pytest.mark.parametrize("name", ["precond_vp_mlp", "precond_ve_mlp", "precond_sr3_mlp", "precond_vp_punetg2d",
                                  "precond_ve_punetg2d"])
def test_vp_ve_sr3(golden, name):
    g = golden(name)
    net, tag, kind = precond_case(g)
    net64, _, _ = precond_case(g, torch.float64)
    fns = K.SchedFns(tag)
    n = g["nsteps"]
    assert torch.allclose(K.generic_steps(tag, n + 1), g["steps"], rtol=1e-6, atol=0)
    assert relmax(K.generic_precond(kind, g["den_sigma"], fns)[3], g["den_cnoise"]) < 1e-6
    assert relmax(K.generic_denoiser(net, g["den_x"], g["den_sigma"], kind, fns), g["den_D"]) < TOL32
    assert relmax(K.generic_score(net, g["den_x"], g["den_sigma"], kind, fns), g["den_score"]) < 5 * TOL32
    x0 = g["white_noise"] * g["maximum_scale"]
    for key, integ, nz in (("heun_hist", "heun", None), ("euler", "euler", None), ("em", "euler-maruyama", g["noises"])):
        o32 = K.generic_propagate(net, x0, n, tag, kind, integ, record_history=key == "heun_hist", noises=nz)
        o64 = K.generic_propagate(net64, x0.double(), n, tag, kind, integ, record_history=key == "heun_hist",
                                  noises=None if nz is None else [z.double() for z in nz])
        assert relmax(o32, g[key]) <= 2.0 * relmax(g[key].double(), o64) + 2e-5, (key, relmax(o32, g[key]))
    L = K.generic_loss(net, g["loss_x"], g["loss_sigma"], g["loss_noise"], kind, fns)
    assert abs(float(L) - float(g["loss_huber"])) <= 2e-5 * abs(float(g["loss_huber"]))


# ----------------------------------------------------------------------------- ensemble losses, latent wrapper
def _leaf_net(g_net, dtype=torch.float32):
    sd = N.synth_state_dict(g_net["manifest"], g_net["seed"], dtype)
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and k != "time_projection.W"}
    full = dict(sd, **leaves)
    cfg = cfg_for("punetg", g_net["cfg"])
    return (lambda x, t: N.punetg_forward(full, cfg, x, t)), leaves


@pytest.mark.parametrize("metric", ["huber", "mse", "CRPS"])
def test_ensemble_losses(golden, metric):
    """oracle.ensemble_loss vs EnsembleKarrasModule.loss_fn of the LIVE reference (loss and parameter gradients)."""
    g = golden("ensemble_punetg2d")
    for tag in ("E3", "E3_mask", "E1", "E1_mask"):
        ref = g["cases"][f"{metric}_{tag}"]
        net, leaves = _leaf_net(golden(g["net"]))
        single = tag.startswith("E1")
        L = K.ensemble_loss(net, g["x"], g["sigma"], g["noise1"] if single else g["noise"], metric,
                            g["mask"] if tag.endswith("mask") else None, single=single)
        assert abs(float(L.detach()) - float(ref["loss"])) <= 5e-6 * abs(float(ref["loss"])), (metric, tag)
        L.backward()
        for k, gr in ref["grads"].items():
            assert relmax(leaves[k].grad, gr) < 2e-4, (metric, tag, k, relmax(leaves[k].grad, gr))


def toy_autoencoder():
    """The fixed encode / decode pair of oracle/make_goldens.py: ToyAutoencoder (same numbers)."""
    class Toy(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.enc, self.dec = torch.nn.Conv2d(1, 1, 1), torch.nn.Conv2d(1, 1, 1)
            with torch.no_grad():
                self.enc.weight.fill_(0.8), self.enc.bias.fill_(0.05), self.dec.weight.fill_(0.7), self.dec.bias.fill_(0.02)

        def encode(self, x):
            return self.enc(torch.nn.functional.avg_pool2d(x, 2))

        def decode(self, z):
            return torch.nn.functional.interpolate(self.dec(z), scale_factor=2, mode="nearest")
    return Toy()


def test_latent_wrapper(golden):
    """karrasmodule.py:583-587, 893-896, 1192-1234: loss on encode(x), decode after sampling."""
    g = golden("latent_punetg2d")
    ae = toy_autoencoder()
    with torch.no_grad():
        z = ae.encode(g["x"])
    assert torch.allclose(z, g["encoded"], rtol=1e-6, atol=1e-7)
    net, leaves = _leaf_net(golden(g["net"]))
    L = K.edm_loss(net, z, g["sigma"], g["noise"], "huber")
    assert abs(float(L.detach()) - float(g["loss"])) <= 5e-6 * abs(float(g["loss"]))
    L.backward()
    for k, gr in g["grads"].items():
        assert relmax(leaves[k].grad, gr) < 2e-4, k
    net32, net64 = oracle_net(golden(g["net"])), oracle_net(golden(g["net"]), torch.float64)
    with torch.no_grad():
        lat = K.sample_from_white_noise(net32, g["white_noise"], 4, "heun")
        lat64 = K.sample_from_white_noise(net64, g["white_noise"].double(), 4, "heun")
        budget = 2.0 * relmax(g["sample_latent"].double(), lat64) + 2e-5
        assert relmax(lat, g["sample_latent"]) <= budget
        assert torch.allclose(ae.decode(g["sample_latent"]), g["sample_decoded"], rtol=1e-6, atol=1e-6)
        assert torch.allclose(g["sample_hist_decoded"][-1], g["sample_decoded"], rtol=1e-6, atol=1e-6)
        assert tuple(g["sample_hist_decoded"].shape) == (5, 2, 1, 32, 32)


@pytest.mark.parametrize("metric", ["huber", "mse"])
def test_dynamic_loss_weight(golden, metric):
    """has_dynamic_loss_weight (karrasmodule.py:594-602, 1243-1278): loss, network gradients and the gradients of the
    uncertainty head vs the LIVE reference."""
    g = golden("dynweight_punetg2d")
    for tag in ("", "_mask"):
        ref = g[metric + tag]
        net, leaves = _leaf_net(golden(g["net"]))
        st = {k: v.clone().requires_grad_(k.startswith("linear")) for k, v in g["dlw_state"].items()}
        L = K.edm_loss(net, g["x"], g["sigma"], g["noise"], metric, g["mask"] if tag else None, dynamic_state=st)
        assert abs(float(L.detach()) - float(ref["loss"])) <= 5e-6 * abs(float(ref["loss"]))
        L.backward()
        for k, gr in ref["grads"].items():
            assert relmax(leaves[k].grad, gr) < 2e-4, (metric, tag, k)
        for k, gr in ref["dlw_grads"].items():
            assert relmax(st[k].grad, gr) < 1e-5, (metric, tag, k)


@pytest.mark.parametrize("name", ["dropout_punetg2d", "dropout_adm2d"])
def test_training_dropout(golden, name):
    """Training-mode dropout (commonlayers.py:829-831; adm.py:323-329): the oracle with the recorded keep masks vs the LIVE
    reference run with the same masks injected -- forward output and parameter gradients."""
    g = golden(name)
    base = golden(g["net"])
    sd = N.synth_state_dict(base["manifest"], base["seed"])
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and not k.endswith(".W")}
    cfg = cfg_for(base["kind"], dict(base["cfg"], dropout=g["p"]))
    masks = {k: v.float() / (1.0 - g["p"]) for k, v in g["keep"].items()}
    fwd = N.punetg_forward if base["kind"] == "punetg" else N.adm_forward
    y = fwd(dict(sd, **leaves), cfg, base["x"], base["t"], dropout_masks=masks)
    assert relmax(y.detach(), g["y"]) < TOL32
    (y * g["dF"]).sum().backward()
    for k, gr in g["grads"].items():
        assert relmax(leaves[k].grad, gr) < 2e-4, k
    frac = sum(float(v.float().sum()) for v in g["keep"].values()) / sum(v.numel() for v in g["keep"].values())
    assert abs(frac - (1 - g["p"])) < 0.02


@pytest.mark.parametrize("name", ["nobias_punetg2d", "nobias_punetg3d"])
def test_bias_false_punetg(golden, name):
    """PUNetGConfig(bias=False) (punetg.py:188-216, 389-394): denoiser, Heun / Euler-Maruyama sampling and loss + gradients of
    the LIVE reference with bias-free convolutions and the appended ones channel."""
    g = golden(name)
    net = oracle_net(g)
    assert relmax(K.denoiser(net, g["den_x"], g["den_sigma"]), g["den_D"]) < TOL32
    net64 = oracle_net(g, torch.float64)
    wn, n = g["white_noise"], g["nsteps"]
    for key, integ, kw in (("heun_hist", "heun", dict(record_history=True)), ("em", "euler-maruyama", dict(noises=g["noises"]))):
        o32 = K.sample_from_white_noise(net, wn, n, integ, **kw)
        kw64 = {k: ([z.double() for z in v] if k == "noises" else v) for k, v in kw.items()}
        o64 = K.sample_from_white_noise(net64, wn.double(), n, integ, **kw64)
        assert relmax(o32, g[key]) <= 2.0 * relmax(g[key].double(), o64) + 2e-5, key
    netg, leaves = _leaf_net(g)
    L = K.edm_loss(netg, g["loss_x"], g["loss_sigma"], g["loss_noise"], "huber")
    assert abs(float(L.detach()) - float(g["loss_huber"])) <= 5e-6 * abs(float(g["loss_huber"]))
    L.backward()
    for k, gr in g["loss_huber_grads"].items():
        assert relmax(leaves[k].grad, gr) < 2e-4, k
