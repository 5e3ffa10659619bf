"""GPU parity tests of the conditional path: vector conditioning through a conditional embedding,
PUNetGCond's channel conditioning, classifier-free guidance as one 2B-sample evaluation -- against golden vectors
recorded from the LIVE reference (oracle/make_goldens.py --only cond) and against the CPU oracle, through the C ABI."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def relmax(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def build(g, precision="fp32"):
    import diffsci_b200 as d
    from diffsci_b200.models.nets.embedder import PorosityEmbedder
    from oracle.nets_oracle import synth_state_dict
    if g.get("kind", "punetg") == "adm":
        cfg = d.ADMConfig(**g["cfg"])
        net = d.ADM(cfg, conditional_embedding=PorosityEmbedder(cfg.output_embed_dim), precision=precision)
        assert list(net.state_dict().keys()) == [k for k, _ in g["manifest"]]
        net.load_state_dict(synth_state_dict(g["manifest"], g["seed"]))
        net = net.to(DEV).eval()
        return net, d.KarrasModule(net, d.KarrasModuleConfig.from_edm(), conditional=True)
    cfg = d.PUNetGConfig(**g["cfg"])
    emb = PorosityEmbedder(cfg.model_channels)
    if "cond" in g["y_batch"]:
        net = d.PUNetGCond(cfg, conditional_embedding=emb, channel_conditional_items=["cond"], precision=precision)
    else:
        net = d.PUNetG(cfg, conditional_embedding=emb, precision=precision)
    assert list(net.state_dict().keys()) == [k for k, _ in g["manifest"]]      # same keys, same order as the reference
    net.load_state_dict(synth_state_dict(g["manifest"], g["seed"]))
    net = net.to(DEV).eval()
    mod = d.KarrasModule(net, d.KarrasModuleConfig.from_edm(), conditional=True)
    return net, mod


def dev(y):
    return {k: v.to(DEV) for k, v in y.items()}


@pytest.mark.parametrize("name", ["cond_punetg2d_embed", "cond_punetg3d_chan", "cond_adm2d_embed"])
def test_conditional_forward_and_denoiser(golden, name):
    g = golden(name)
    net, mod = build(g)
    with torch.no_grad():
        y = net(g["x"].to(DEV), g["t"].to(DEV), dev(g["y_batch"])).cpu()
    e_ref = relmax(g["net_y"], g["net_y64"])
    assert relmax(y, g["net_y64"]) < max(3 * e_ref, 2e-5)
    assert relmax(y, g["net_y"]) < max(4 * e_ref, 2e-5)
    for key in [k for k in g if k.startswith("den_D_g")]:
        gd = float(key[len("den_D_g"):])
        with torch.no_grad():
            D, _ = mod.get_denoiser(g["den_x"].to(DEV), g["den_sigma"].to(DEV), dev(g["y_batch"]), guidance=gd)
        assert relmax(D.cpu(), g[key]) < 5e-5, (key, relmax(D.cpu(), g[key]))
    with torch.no_grad():
        sc = mod.get_score(g["den_x"].to(DEV), g["den_sigma"].to(DEV), dev(g["y_batch"]))
    assert relmax(sc.cpu(), g["den_score_g1.0"]) < 5e-5
    # conditional=False ignores y (karrasmodule.py:703: the conditional call needs self.conditional)
    import diffsci_b200 as d
    if "cond" not in g["y_batch"]:
        mod_u = d.KarrasModule(net, d.KarrasModuleConfig.from_edm(), conditional=False)
        with torch.no_grad():
            D, _ = mod_u.get_denoiser(g["den_x"].to(DEV), g["den_sigma"].to(DEV), dev(g["y_batch"]))
        assert relmax(D.cpu(), g["den_D_g0.0"]) < 5e-5
    else:
        with pytest.raises(TypeError):              # CFG on a channel-conditioned net: no unconditional evaluation exists
            mod.get_denoiser(g["den_x"].to(DEV), g["den_sigma"].to(DEV), dev(g["y_batch"]), guidance=2.0)


@pytest.mark.parametrize("name", ["cond_punetg2d_embed", "cond_punetg3d_chan", "cond_adm2d_embed"])
def test_conditional_sampling(golden, name):
    """Heun / Euler-Maruyama with y through the graph engine (conditioning written once per run; CFG = one 2B-sample
    evaluation, mixed inside the fused stage) vs histories of the live reference, budgeted against fp64 truth."""
    import diffsci_b200 as d
    from oracle import karras_oracle as K
    from tests.test_oracle_vs_golden import cond_oracle_nets
    g = golden(name)
    net, mod = build(g)
    n, wn = g["nsteps"], g["white_noise"]
    oc64, ou64 = cond_oracle_nets(g, torch.float64, batch=False)
    for key in [k for k in g if k.startswith("heun_hist_g")]:
        gd = float(key[len("heun_hist_g"):])
        truth = K.sample_from_white_noise(K.guided_net(oc64, ou64, gd), wn.double(), n, "heun", record_history=True)
        budget = 3.0 * relmax(g[key], truth) + 5e-5
        for graphs in (True, False):
            mod.use_cuda_graphs = graphs
            out = mod.propagate_white_noise(wn.to(DEV), dev(g["y_one"]), gd, n, record_history=True).cpu()
            assert relmax(out, truth) <= budget, (key, graphs, relmax(out, truth), budget)
        assert mod.last_nfe == 2 * n - 1
    truth = K.sample_from_white_noise(oc64, wn.double(), n, "euler-maruyama", noises=[z.double() for z in g["noises"]])
    integ = d.name_to_integrator("euler-maruyama")
    integ.reset_noise(injected=g["noises"])
    mod.use_cuda_graphs = True
    out = mod.propagate_white_noise(wn.to(DEV), dev(g["y_one"]), 1.0, n, integrator=integ).cpu()
    assert relmax(out, truth) <= 3.0 * relmax(g["em_g1.0"], truth) + 5e-5
    # a new condition on the same engine (static buffers refilled, no re-capture) changes the result
    y2 = {k: v + 0.3 for k, v in g["y_one"].items()}
    out2 = mod.propagate_white_noise(wn.to(DEV), dev(y2), 1.0, n).cpu()
    out1 = mod.propagate_white_noise(wn.to(DEV), dev(g["y_one"]), 1.0, n).cpu()
    assert relmax(out1, g["heun_hist_g1.0"][-1]) < 1e-2 and relmax(out2, out1) > 1e-4
    # sample(): public entry point with y
    torch.manual_seed(3)
    s = mod.sample(3, list(wn.shape[1:]), y=g["y_one"], nsteps=3)
    assert s.shape == (3,) + tuple(wn.shape[1:]) and torch.isfinite(s).all()


@pytest.mark.parametrize("name", ["cond_punetg2d_embed", "cond_punetg3d_chan", "cond_adm2d_embed"])
def test_conditional_sampling_bf16(golden, name):
    from oracle import karras_oracle as K
    from tests.test_oracle_vs_golden import cond_oracle_nets
    g = golden(name)
    net, mod = build(g, "bf16")
    oc64, ou64 = cond_oracle_nets(g, torch.float64)
    with torch.no_grad():
        y = net(g["x"].to(DEV), g["t"].to(DEV), dev(g["y_batch"])).cpu()
    assert relmax(y, g["net_y64"]) < 3e-2
    for gd in ((1.0, 2.5) if ou64 is not None else (1.0,)):
        truth = K.denoiser(K.guided_net(oc64, ou64, gd), g["den_x"].double(), g["den_sigma"].double())
        with torch.no_grad():
            D, _ = mod.get_denoiser(g["den_x"].to(DEV), g["den_sigma"].to(DEV), dev(g["y_batch"]), guidance=gd)
        assert relmax(D.cpu(), truth) < 3e-2, gd


@pytest.mark.parametrize("name", ["cond_punetg2d_embed", "cond_punetg3d_chan", "cond_adm2d_embed"])
def test_conditional_loss_and_gradients(golden, name):
    """loss_fn(x, sigma, y) -> backward: the native backward returns d(loss)/d(ye), torch autograd carries it into the
    user's embedder; every recorded gradient of the live reference (network AND embedder) is matched."""
    g = golden(name)
    net, mod = build(g)
    net.train()
    mod.train()
    net.zero_grad()
    mod._injected_loss_noise = g["loss_noise"]
    L = mod.loss_fn(g["loss_x"].to(DEV), g["loss_sigma"].to(DEV), dev(g["y_batch"]))
    L.backward()
    assert abs(float(L) - float(g["loss_huber"])) < 5e-5 * abs(float(g["loss_huber"]))
    params = dict(net.named_parameters())
    seen_emb = False
    for k, ref in g["loss_huber_grads"].items():
        assert params[k].grad is not None, k
        e = relmax(params[k].grad.cpu(), ref)
        assert e < 3e-4, (k, e)
        seen_emb |= k.startswith("conditional_embedding.")
    assert seen_emb
