"""GPU parity tests of the VP / VE / SR3 configurations (reference preconditioners.py:56-136,
noisesamplers.py:44-110, schedulers.py:393-448 and the non-constant-scaling branch of Scheduler.rhs :275-293):
KarrasModuleConfig.from_vp / from_ve / a custom SR3 config on the native networks, through the Integrator.step seam,
against goldens recorded from the LIVE reference (oracle/make_goldens.py --only precond)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
NAMES = ["precond_vp_mlp", "precond_ve_mlp", "precond_sr3_mlp", "precond_vp_punetg2d", "precond_ve_punetg2d"]


def relmax(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def build(golden, g):
    import diffsci_b200 as d
    from tests.test_gpu_nets import build_net
    net = build_net(golden(g["net"]))
    with torch.no_grad():    # the fixtures scale the last layer (see oracle/make_goldens.py: precond_goldens)
        last = net.convout if hasattr(net, "convout") else [m for m in net.modules() if hasattr(m, "weight") and m.weight.ndim == 2][-1]
        last.weight.mul_(g["out_scale"])
        last.bias.mul_(g["out_scale"])
    if g["tag"] == "vp":
        cfg = d.KarrasModuleConfig.from_vp()
    elif g["tag"] == "ve":
        cfg = d.KarrasModuleConfig.from_ve()
    else:
        from diffsci_b200.models.karras import preconditioners as P, noisesamplers as NS, schedulers as S
        cfg = d.KarrasModuleConfig(preconditioner=P.SR3Preconditioner(), noisesampler=NS.EDMNoiseSampler(),
                                   noisescheduler=S.EDMScheduler())
    return net, d.KarrasModule(net, cfg)


@pytest.mark.parametrize("name", NAMES)
def test_denoiser_sampling_loss(golden, name):
    import diffsci_b200 as d
    from oracle import karras_oracle as K
    from tests.test_oracle_vs_golden import precond_case
    g = golden(name)
    net, mod = build(golden, g)
    n = g["nsteps"]
    sch = mod.config.noisescheduler
    assert torch.allclose(sch.create_steps(n + 1).cpu(), g["steps"], rtol=1e-6, atol=0)
    assert abs(float(sch.maximum_scale) - g["maximum_scale"]) <= 1e-6 * g["maximum_scale"]
    assert relmax(mod.config.noisesampler.loss_weighting(g["den_sigma"]), g["loss_weight"]) < 1e-6
    D, cn = mod.get_denoiser(g["den_x"].to(DEV), g["den_sigma"].to(DEV))
    assert relmax(D.cpu(), g["den_D"]) < 5e-5 and relmax(cn.cpu(), g["den_cnoise"]) < 1e-6
    assert relmax(mod.get_score(g["den_x"].to(DEV), g["den_sigma"].to(DEV)).cpu(), g["den_score"]) < 2e-4
    # sampling: budget against the fp64 oracle (chained evaluations), as for the EDM samplers
    net64, tag, kind = precond_case(g, torch.float64)
    wn = g["white_noise"]
    x0 = wn.double() * g["maximum_scale"]
    for key, integ, nz in (("heun_hist", "heun", None), ("euler", "euler", None), ("em", "euler-maruyama", g["noises"])):
        truth = K.generic_propagate(net64, x0, n, tag, kind, integ, record_history=key == "heun_hist",
                                    noises=None if nz is None else [z.double() for z in nz])
        integrator = d.name_to_integrator(integ)
        integrator.reset_noise(injected=nz)
        out = mod.propagate_white_noise(wn.to(DEV), nsteps=n, integrator=integrator, record_history=key == "heun_hist").cpu()
        e, budget = relmax(out, truth), 3.0 * relmax(g[key], truth) + 5e-5
        assert e <= budget, (key, e, budget)
    # loss (general-preconditioner form of the fused loss kernel) and gradients of the live reference
    mod._injected_loss_noise = g["loss_noise"]
    if not hasattr(net, "train_graph"):       # MLPUncond has no native backward (the toy config is sampling-only): loss value
        with torch.no_grad():
            L = mod.loss_fn(g["loss_x"].to(DEV), g["loss_sigma"].to(DEV))
        assert abs(float(L) - float(g["loss_huber"])) < 5e-5 * abs(float(g["loss_huber"]))
        return
    net.train()
    mod.train()
    net.zero_grad()
    L = mod.loss_fn(g["loss_x"].to(DEV), g["loss_sigma"].to(DEV))
    L.backward()
    assert abs(float(L.detach()) - float(g["loss_huber"])) < 5e-5 * abs(float(g["loss_huber"]))
    params = dict(net.named_parameters())
    for k, ref in g["loss_huber_grads"].items():
        e = relmax(params[k].grad.cpu(), ref)
        assert e < 3e-4, (k, e)


@pytest.mark.parametrize("name", NAMES)
def test_general_engine_equals_the_step_seam(golden, name):
    """VP / VE / SR3 on the captured-graph loop (csrc/sampler_general.cu, the default route) vs the same runs on the
    Integrator.step seam (DSK_GENERAL_ENGINE=0) and vs the fp64 oracle budget."""
    import os
    import diffsci_b200 as d
    from oracle import karras_oracle as K
    from tests.test_oracle_vs_golden import precond_case
    g = golden(name)
    net, mod = build(golden, g)
    n, wn = g["nsteps"], g["white_noise"]
    net64, tag, kind = precond_case(g, torch.float64)
    x0 = wn.double() * g["maximum_scale"]
    for key, integ, nz in (("heun_hist", "heun", None), ("euler", "euler", None), ("em", "euler-maruyama", g["noises"])):
        truth = K.generic_propagate(net64, x0, n, tag, kind, integ, record_history=key == "heun_hist",
                                    noises=None if nz is None else [z.double() for z in nz])
        outs = {}
        try:
            for route in ("1", "0"):
                os.environ["DSK_GENERAL_ENGINE"] = route
                integrator = d.name_to_integrator(integ)
                integrator.reset_noise(injected=nz)
                outs[route] = mod.propagate_white_noise(wn.to(DEV), nsteps=n, integrator=integrator,
                                                        record_history=key == "heun_hist").cpu()
        finally:
            os.environ.pop("DSK_GENERAL_ENGINE", None)
        assert any(k[0] == "general" for k in mod._engines)
        budget = 3.0 * relmax(g[key], truth) + 5e-5
        assert relmax(outs["1"], truth) <= budget, (key, relmax(outs["1"], truth), budget)
        assert relmax(outs["1"], outs["0"]) <= budget
