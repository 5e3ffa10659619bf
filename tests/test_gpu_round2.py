"""GPU tests of the round-2 boundary additions: MLPCond, configurable Huber delta, EDMTrainer for non-EDM configurations."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def relmax(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def test_mlpcond_forward_and_gradients():
    """MLPCond (nets/mlp.py:61-121): cat[x, t, y] -> MLP; inference plan and the hand-written backward vs torch autograd on the
    same weights."""
    import diffsci_b200 as d
    torch.manual_seed(0)
    net = d.MLPCond(2, 3, [32, 32], torch.nn.SiLU()).to(DEV)
    ref = torch.nn.Sequential(torch.nn.Linear(6, 32), torch.nn.SiLU(), torch.nn.Linear(32, 32), torch.nn.SiLU(),
                              torch.nn.Linear(32, 2)).double()
    ref.load_state_dict({k[len("net."):]: v.double().cpu() for k, v in net.state_dict().items()})
    x, t, y = torch.randn(64, 2), torch.randn(64), torch.randn(64, 3)
    want = ref(torch.cat([x, t[:, None], y], -1).double())
    with torch.no_grad():
        got = net.eval()(x.to(DEV), t.to(DEV), y.to(DEV))
    assert relmax(got.cpu(), want) < 1e-5
    net.train()
    out = net(x.to(DEV), t.to(DEV), y.to(DEV))
    w = torch.randn(64, 2)
    (out * w.to(DEV)).sum().backward()
    (want * w.double()).sum().backward()
    for (k, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
        assert relmax(p.grad.cpu(), q.grad) < 1e-4, k
    with pytest.raises(TypeError):
        net(x.to(DEV), t.to(DEV))


@pytest.mark.parametrize("delta", [0.25, 2.0])
def test_huber_delta_loss_and_gradient(delta):
    """loss_metric = {"huber": {"delta": d}} (karrasmodule.py:558-562): fused loss value and dL/dF vs torch.nn.HuberLoss(delta)
    around the oracle's denoiser arithmetic."""
    import diffsci_b200 as d
    torch.manual_seed(1)
    net = d.MLPUncond(2, [32], torch.nn.SiLU()).to(DEV)
    mod = d.KarrasModule(net, d.KarrasModuleConfig.from_edm(loss_metric={"huber": {"delta": delta}})).to(DEV)
    x, sigma, noise = torch.randn(128, 2) * 3, torch.exp(torch.randn(128)), torch.randn(128, 2)
    mod._injected_loss_noise = noise
    loss = mod.loss_fn(x.to(DEV), sigma.to(DEV))
    loss.backward()
    ref = torch.nn.Sequential(torch.nn.Linear(3, 32), torch.nn.SiLU(), torch.nn.Linear(32, 2)).double()
    ref.load_state_dict({k[len("net."):]: v.double().cpu() for k, v in net.state_dict().items()})
    s = sigma.double()[:, None]
    sd = 0.5
    xn = x.double() + s * noise.double()
    c_in, c_out, c_skip = 1 / torch.sqrt(s ** 2 + sd ** 2), s * sd / torch.sqrt(s ** 2 + sd ** 2), sd ** 2 / (s ** 2 + sd ** 2)
    F = ref(torch.cat([c_in * xn, 0.5 * torch.log(s)], -1))
    D = c_out * F + c_skip * xn
    lam = (s ** 2 + sd ** 2) / (s * sd) ** 2
    want = (lam * torch.nn.HuberLoss(reduction="none", delta=delta)(D, x.double())).mean()
    want.backward()
    assert abs(float(loss) - float(want)) < 2e-5 * abs(float(want))
    for (k, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
        assert relmax(p.grad.cpu(), q.grad) < 2e-4, k


def test_trainer_general_preconditioner_and_delta():
    """EDMTrainer with a VP configuration and a Huber delta: the fused iteration's loss equals KarrasModule.loss_fn on the same
    sigma / noise (per-sample coefficient vectors path of the loss kernel), and parameters move."""
    import diffsci_b200 as d
    torch.manual_seed(2)
    net = d.PUNetG(d.PUNetGConfig(model_channels=8), precision="fp32").to(DEV)
    mod = d.KarrasModule(net, d.KarrasModuleConfig.from_vp(loss_metric={"huber": {"delta": 0.5}})).to(DEV)
    x = torch.randn(4, 1, 16, 16, device=DEV)
    sigma = mod.config.noisesampler.sample(4).to(DEV)
    noise = torch.randn(4, 1, 16, 16, device=DEV)
    mod._injected_loss_noise = noise
    want = float(mod.loss_fn(x, sigma).detach())
    before = [p.detach().clone() for p in net.parameters()]
    tr = d.EDMTrainer(mod, lr=1e-3)
    got = float(tr.step(x, sigma=sigma, noise=noise))
    assert abs(got - want) < 1e-4 * abs(want), (got, want)
    assert any(not torch.equal(a, b) for a, b in zip(before, net.parameters()))


def test_inpaint_repaint_partial_run_on_the_graph_engine(monkeypatch):
    """SURVEY 8f-1 on the captured-graph engine: partial stretches start at a schedule row, the inpainting blend
    is fused into the step-completing stages, repaint chains engine stretches -- each equal to the Integrator.step seam
    (DSK_PARTIAL_ENGINE=0; the seam itself is pinned against the live reference in test_gpu_nets.py) and replayed from graphs."""
    import diffsci_b200 as d
    torch.manual_seed(3)
    net = d.PUNetG(d.PUNetGConfig(model_channels=8), precision="fp32").to(DEV).eval()
    mod = d.KarrasModule(net, d.KarrasModuleConfig.from_edm())
    sch = mod.config.noisescheduler
    n = 20                                     # Scheduler.repaint's default rsteps = 10 must divide it
    x0 = torch.randn(2, 1, 16, 16, device=DEV)
    mask = (torch.rand(1, 16, 16, device=DEV) > 0.5).float()
    sch.stochastic_integrator.reset_noise(seed=11)
    hist = mod.propagate_toward_noise(x0, nsteps=n, record_history=True, stochastic_integration=True)
    start = torch.randn(2, 1, 16, 16, device=DEV) * 80

    def both(fn):
        sch.integrator.reset_noise(seed=7)
        monkeypatch.setenv("DSK_PARTIAL_ENGINE", "0")
        seam = fn()
        sch.integrator.reset_noise(seed=7)
        monkeypatch.delenv("DSK_PARTIAL_ENGINE")
        eng = next(iter(mod._engines.values()), None)
        before = 0 if eng is None else eng.graph_launches_per_run()
        out = fn()
        eng = next(iter(mod._engines.values()))
        assert eng.use_graphs and eng.graph_launches_per_run() > 0, before
        return seam, out

    for name, fn in [("partial", lambda: mod.propagate_partial_toward_sample(start, 2, 6, nsteps=n, record_history=True)),
                     ("partial-to-end", lambda: mod.propagate_partial_toward_sample(start, 15, None, nsteps=n)),
                     ("inpaint", lambda: mod.propagate_inpaint_toward_sample(start, hist, mask, record_history=True)),
                     ("repaint", lambda: mod.propagate_repaint_toward_sample(start, hist, mask))]:
        seam, out = both(fn)
        assert out.shape == seam.shape, name
        assert relmax(out, seam) < (5e-3 if name == "repaint" else 2e-4), (name, relmax(out, seam))   # repaint: 110 chained evaluations
