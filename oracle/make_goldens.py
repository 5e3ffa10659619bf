"""Generate tests/golden/*.pt by running the LIVE reference (/root/reference) on CPU.

Build-container only (the reference cannot travel to the GPU box).  Run:

    python oracle/make_goldens.py

Every fixture stores: the config kwargs, the (key, shape) manifest of the reference
module's state_dict, the synthetic-weight seed (weights are regenerated with
``nets_oracle.synth_state_dict`` -- they are NOT stored), the seeded inputs and the
reference outputs in fp32 plus an fp64 run (``module.double()``) for tolerance budgeting.
"""
from __future__ import annotations

import os
import re
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refload  # noqa: E402
from nets_oracle import synth_state_dict  # noqa: E402

GRAD_KEYS = re.compile(r"^(net\.|convin|convout|downward_blocks\.0\.0\.|before_block\.0\.conv1|"
                       r"attn_block\.0\.|upsamplers\.1\.|downsamplers\.0\.)")
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def manifest_of(module):
    return [(k, list(v.shape)) for k, v in module.state_dict().items()]


def load_synth(module, seed):
    man = manifest_of(module)
    module.load_state_dict(synth_state_dict(man, seed))
    return man


class _Noise:
    """Replace torch.randn_like with a replay of pre-drawn tensors (integrators.py:68,105;
    karrasmodule.py:591 draw on the device generator -- injected for parity)."""

    def __init__(self, tensors):
        self.tensors = list(tensors)
        self.i = 0

    def __enter__(self):
        self.orig = torch.randn_like

        def fake(x, *a, **k):
            t = self.tensors[self.i].to(x)
            self.i += 1
            assert t.shape == x.shape
            return t
        torch.randn_like = fake
        return self

    def __exit__(self, *a):
        torch.randn_like = self.orig


def inpaint_goldens():
    """Scheduler.inpaint / repaint / propagate_partial / propagate_forward of the LIVE reference on the
    punetg2d_mc8 network (step noise injected)  ->  tests/golden/inpaint_punetg2d.pt.   python oracle/make_goldens.py --only inpaint"""
    os.makedirs(OUT, exist_ok=True)
    refload.load_reference()
    import diffsci.models as M
    from diffsci.models.nets.punetg import PUNetG
    from diffsci.models.nets.punetg_config import PUNetGConfig
    torch.set_num_threads(8)
    kw = dict(dimension=2, model_channels=8)
    net = PUNetG(PUNetGConfig(**kw))
    load_synth(net, 101)                                   # the weights of punetg2d_mc8
    net.eval()
    cfg = M.KarrasModuleConfig.from_edm()
    mod = M.KarrasModule(net, cfg)
    mod.eval()
    torch.manual_seed(303)
    shape, nsteps = (2, 1, 28, 28), 4
    x_orig = torch.randn(*shape) * 0.5
    mask = (torch.rand(*shape[1:]) > 0.5).float()
    start = torch.randn(*shape) * 80.0
    fwd_noise = [torch.randn(*shape) for _ in range(nsteps)]
    re_noise = [torch.randn(*shape) for _ in range(8)]
    out = dict(nsteps=nsteps, x_orig=x_orig, mask=mask, start=start, fwd_noise=fwd_noise, re_noise=re_noise)
    with torch.no_grad():
        with _Noise(fwd_noise):
            hist = mod.propagate_toward_noise(x_orig, nsteps=nsteps, record_history=True, stochastic_integration=True)
        out["fwd_hist"] = hist
        out["fwd_ode"] = mod.propagate_toward_noise(x_orig, nsteps=nsteps)
        out["inpaint_hist"] = mod.propagate_inpaint_toward_sample(start, hist, mask, record_history=True)
        sch = cfg.noisescheduler

        def rhs(x, sigma):
            return mod.get_score(x, sigma, None)
        with _Noise(re_noise):
            out["repaint_hist"] = sch.repaint(start, hist, mask, rhs, nsteps, rsteps=2, nresamples=2, record_history=True)
        out["partial"] = mod.propagate_partial_toward_sample(start, 1, 3, nsteps=nsteps, record_history=True)
    torch.save(out, os.path.join(OUT, "inpaint_punetg2d.pt"))
    print("inpaint_punetg2d", {k: (tuple(v.shape), float(v.abs().max())) for k, v in out.items() if torch.is_tensor(v)})


def cond_goldens():
    """conditional path of the LIVE reference -- PUNetG + PorosityEmbedder (vector conditioning, classifier-
    free guidance) and PUNetGCond (channel conditioning + embedder): network forward, get_denoiser at guidance 0 / 1 /
    2.5, Heun / Euler-Maruyama sampling with y, conditional loss + gradients (incl. the embedder's)
    ->  tests/golden/cond_*.pt.   python oracle/make_goldens.py --only cond"""
    os.makedirs(OUT, exist_ok=True)
    refload.load_reference()
    import diffsci.models as M
    from diffsci.models.nets.punetg import PUNetG, PUNetGCond
    from diffsci.models.nets.punetg_config import PUNetGConfig
    from diffsci.models.nets.embedder import PorosityEmbedder
    torch.set_num_threads(8)
    gk = re.compile(r"^(conditional_embedding\.|convin|convout|downward_blocks\.0\.0\.|before_block\.0\.conv1|"
                    r"attn_block\.0\.|upsamplers\.1\.|time_embedding\.|input_layer|output_layer|"
                    r"encoder\.layers\.0\.input_blocks\.0\.|middle_block\.middle_blocks\.2\.)")

    def case(name, net, kw, shape, seed, chan_shape, cfg_ok, kind="punetg"):
        man = load_synth(net, seed)
        net.eval()
        mod = M.KarrasModule(net, M.KarrasModuleConfig.from_edm(), conditional=True)
        mod.eval()
        torch.manual_seed(seed + 1000)
        B, nsteps = shape[0], 4
        x = torch.randn(*shape)
        t = torch.randn(B) * 0.7
        yb = {"porosity": torch.rand(B, 1)}
        y1 = {"porosity": torch.rand(1)}               # one condition for the whole sample() batch (unsqueezed inside)
        if chan_shape is not None:
            yb["cond"] = torch.randn(B, *chan_shape)
            y1["cond"] = torch.randn(*chan_shape)
        out = dict(kind=kind, cfg=kw, manifest=man, seed=seed, x=x, t=t, y_batch=yb, y_one=y1, nsteps=nsteps)
        with torch.no_grad():
            out["net_y"] = net(x, t, dict(yb))
            out["net_y64"] = net.double()(x.double(), t.double(), {k: v.double() for k, v in yb.items()})
            net.float()
            sg = torch.exp(torch.randn(B) * 1.2 - 1.2)
            xin = torch.randn(*shape) * (1 + sg.view(-1, *([1] * (len(shape) - 1))))
            out.update(den_x=xin, den_sigma=sg)
            for g in ((1.0, 0.0, 2.5) if cfg_ok else (1.0,)):
                out[f"den_D_g{g}"] = mod.get_denoiser(xin, sg, dict(yb), guidance=g)[0]
            out["den_score_g1.0"] = mod.get_score(xin, sg, dict(yb))
            wn = torch.randn(*shape)
            out["white_noise"] = wn
            out["heun_hist_g1.0"] = mod.propagate_white_noise(wn, dict(y1), 1.0, nsteps, record_history=True)
            if cfg_ok:
                out["heun_hist_g2.5"] = mod.propagate_white_noise(wn, dict(y1), 2.5, nsteps, record_history=True)
            noises = [torch.randn(*shape) for _ in range(nsteps)]
            out["noises"] = noises
            with _Noise(noises):
                out["em_g1.0"] = mod.propagate_white_noise(wn, dict(y1), 1.0, nsteps, integrator="euler-maruyama")
        x0 = torch.randn(*shape) * 0.5
        ln = torch.randn(*shape)
        out.update(loss_x=x0, loss_noise=ln, loss_sigma=sg)
        net.train()                                     # dropout probabilities are 0: train == eval numerically
        net.zero_grad()
        with _Noise([ln]):
            L = mod.loss_fn(x0, sg, dict(yb), None)
        L.backward()
        out["loss_huber"] = L.detach()
        out["loss_huber_grads"] = {k: p.grad.clone() for k, p in net.named_parameters() if gk.search(k)}
        torch.save(out, os.path.join(OUT, name + ".pt"))
        print(name, {k: (tuple(v.shape), float(v.abs().max())) for k, v in out.items() if torch.is_tensor(v)})

    if "--adm" in sys.argv:      # added after the PUNetG fixtures were committed; regenerates only the ADM one
        from diffsci.models.nets.adm import ADM, ADMConfig
        kw = dict(input_channels=3, output_channels=3, model_channels=8, time_embed_dim=16, output_embed_dim=32)
        return case("cond_adm2d_embed", ADM(ADMConfig(**kw), conditional_embedding=PorosityEmbedder(32)), kw,
                    (2, 3, 16, 16), 113, None, True, kind="adm")
    kw = dict(dimension=2, model_channels=8)
    case("cond_punetg2d_embed", PUNetG(PUNetGConfig(**kw), conditional_embedding=PorosityEmbedder(8)), kw,
         (2, 1, 16, 16), 111, None, True)
    kw = dict(dimension=3, model_channels=8, input_channels=2, channel_expansion=[2])
    case("cond_punetg3d_chan", PUNetGCond(PUNetGConfig(**kw), conditional_embedding=PorosityEmbedder(8),
                                          channel_conditional_items=["cond"]), kw,
         (2, 1, 8, 8, 8), 112, (1, 8, 8, 8), False)


def circular_goldens():
    """PUNetGConfig(convolution_type='circular') of the LIVE reference (periodic porous media): network
    forward (fp32 + fp64), Heun history, loss + gradients  ->  tests/golden/circ_*.pt.   python oracle/make_goldens.py --only circular"""
    os.makedirs(OUT, exist_ok=True)
    refload.load_reference()
    import diffsci.models as M
    from diffsci.models.nets.punetg import PUNetG
    from diffsci.models.nets.punetg_config import PUNetGConfig
    torch.set_num_threads(8)
    gk = re.compile(r"^(convin|convout|downward_blocks\.0\.0\.|before_block\.0\.conv1|attn_block\.0\.|upsamplers\.0\.|"
                    r"downsamplers\.0\.)")

    def case(name, kw, shape, seed):
        net = PUNetG(PUNetGConfig(**kw))
        man = load_synth(net, seed)
        net.eval()
        mod = M.KarrasModule(net, M.KarrasModuleConfig.from_edm())
        mod.eval()
        torch.manual_seed(seed + 1000)
        B, nsteps = shape[0], 3
        x, t = torch.randn(*shape), torch.randn(B) * 0.7
        out = dict(kind="punetg", cfg=kw, manifest=man, seed=seed, x=x, t=t, nsteps=nsteps)
        with torch.no_grad():
            out["y"] = net(x, t)
            out["y64"] = net.double()(x.double(), t.double())
            net.float()
            wn = torch.randn(*shape)
            out["white_noise"] = wn
            out["heun_hist"] = mod.propagate_white_noise(wn, nsteps=nsteps, record_history=True)
        sg = torch.exp(torch.randn(B) * 1.2 - 1.2)
        x0, ln = torch.randn(*shape) * 0.5, torch.randn(*shape)
        out.update(loss_x=x0, loss_noise=ln, loss_sigma=sg)
        net.train()
        net.zero_grad()
        with _Noise([ln]):
            L = mod.loss_fn(x0, sg, None, None)
        L.backward()
        out["loss_huber"] = L.detach()
        out["loss_huber_grads"] = {k: p.grad.clone() for k, p in net.named_parameters() if gk.search(k)}
        torch.save(out, os.path.join(OUT, name + ".pt"))
        print(name, tuple(out["y"].shape), "fp32-vs-fp64", float((out["y"] - out["y64"]).abs().max() / out["y64"].abs().max()),
              "loss", float(L))

    if "--adm" in sys.argv:      # ADMConfig(convolution_type="circular"): block convs periodic, input / output layers zero-padded
        from diffsci.models.nets.adm import ADM, ADMConfig
        kw = dict(input_channels=3, output_channels=3, model_channels=8, time_embed_dim=16, output_embed_dim=32,
                  convolution_type="circular")
        gk_adm = re.compile(r"^(input_layer|output_layer|encoder\.layers\.0\.input_blocks\.[01]\.|middle_block\.middle_blocks\.2\.|"
                            r"decoder\.layers\.1\.input_blocks\.1\.)")
        net = ADM(ADMConfig(**kw))
        man = load_synth(net, 123)
        net.eval()
        mod = M.KarrasModule(net, M.KarrasModuleConfig.from_edm())
        mod.eval()
        torch.manual_seed(1123)
        shape, nsteps = (2, 3, 16, 16), 3
        x, t = torch.randn(*shape), torch.randn(2) * 0.7
        out = dict(kind="adm", cfg=kw, manifest=man, seed=123, x=x, t=t, nsteps=nsteps)
        with torch.no_grad():
            out["y"] = net(x, t)
            out["y64"] = net.double()(x.double(), t.double())
            net.float()
            wn = torch.randn(*shape)
            out["white_noise"] = wn
            out["heun_hist"] = mod.propagate_white_noise(wn, nsteps=nsteps, record_history=True)
        sg = torch.exp(torch.randn(2) * 1.2 - 1.2)
        x0, ln = torch.randn(*shape) * 0.5, torch.randn(*shape)
        out.update(loss_x=x0, loss_noise=ln, loss_sigma=sg)
        net.train()
        net.zero_grad()
        with _Noise([ln]):
            L = mod.loss_fn(x0, sg, None, None)
        L.backward()
        out["loss_huber"] = L.detach()
        out["loss_huber_grads"] = {k: p.grad.clone() for k, p in net.named_parameters() if gk_adm.search(k)}
        torch.save(out, os.path.join(OUT, "circ_adm2d.pt"))
        print("circ_adm2d fp32-vs-fp64", float((out["y"] - out["y64"]).abs().max() / out["y64"].abs().max()), "loss", float(L.detach()),
              len(out["loss_huber_grads"]), "gradient tensors")
        return
    case("circ_punetg2d", dict(dimension=2, model_channels=8, convolution_type="circular"), (2, 1, 16, 24), 121)
    case("circ_punetg3d", dict(dimension=3, model_channels=8, channel_expansion=[2], convolution_type="circular"),
         (2, 1, 8, 8, 8), 122)


def precond_goldens():
    """VP / VE / SR3 preconditioners, noise samplers and schedulers of the LIVE reference
    (preconditioners.py:56-136, noisesamplers.py:44-110, schedulers.py:393-448, rhs with non-constant s(t) :275-293):
    get_denoiser / get_score, Heun + Euler + Euler-Maruyama sampling, loss + gradients
    ->  tests/golden/precond_*.pt.   python oracle/make_goldens.py --only precond"""
    os.makedirs(OUT, exist_ok=True)
    refload.load_reference()
    import diffsci.models as M
    from diffsci.models.nets.punetg import PUNetG
    from diffsci.models.nets.punetg_config import PUNetGConfig
    from diffsci.models.nets.mlp import MLPUncond
    from diffsci.models.karras import preconditioners as P, noisesamplers as NS, schedulers as S
    torch.set_num_threads(8)

    def config(tag):
        if tag == "vp":
            return M.KarrasModuleConfig.from_vp()
        if tag == "ve":
            return M.KarrasModuleConfig.from_ve()
        return M.KarrasModuleConfig(preconditioner=P.SR3Preconditioner(), noisesampler=NS.EDMNoiseSampler(),
                                    noisescheduler=S.EDMScheduler())

    def case(name, tag, model, man_seed, netname, shape, seed, nsteps=5):
        cfg = config(tag)
        mod = M.KarrasModule(model, cfg)
        mod.eval()
        torch.manual_seed(seed)
        B = shape[0]
        wn = torch.randn(*shape)
        out = dict(tag=tag, net=netname, out_scale=0.02, nsteps=nsteps, white_noise=wn, steps=cfg.noisescheduler.create_steps(nsteps + 1),
                   maximum_scale=float(cfg.noisescheduler.maximum_scale))
        with torch.no_grad():
            torch.manual_seed(seed + 1)
            sg = cfg.noisesampler.sample([B])
            out["sampled_sigma"] = sg
            xin = torch.randn(*shape) * (1 + sg.view(-1, *([1] * (len(shape) - 1))))
            D, cn = mod.get_denoiser(xin, sg)
            out.update(den_x=xin, den_sigma=sg, den_D=D, den_cnoise=cn, den_score=mod.get_score(xin, sg),
                       loss_weight=cfg.noisesampler.loss_weighting(sg))
            out["heun_hist"] = mod.propagate_white_noise(wn, nsteps=nsteps, record_history=True)
            out["euler"] = mod.propagate_white_noise(wn, nsteps=nsteps, integrator="euler")
            noises = [torch.randn(*shape) for _ in range(nsteps)]
            out["noises"] = noises
            with _Noise(noises):
                out["em"] = mod.propagate_white_noise(wn, nsteps=nsteps, integrator="euler-maruyama")
        x0, ln = torch.randn(*shape) * 0.5, torch.randn(*shape)
        out.update(loss_x=x0, loss_noise=ln, loss_sigma=sg)
        model.zero_grad()
        with _Noise([ln]):
            L = mod.loss_fn(x0, sg, None, None)
        L.backward()
        out["loss_huber"] = L.detach()
        out["loss_huber_grads"] = {k: p.grad.clone() for k, p in model.named_parameters() if GRAD_KEYS.search(k)}
        torch.save(out, os.path.join(OUT, name + ".pt"))
        print(name, "heun final absmax", float(out["heun_hist"][-1].abs().max()), "loss", float(L.detach()))

    # synthetic-weight networks are not denoisers: with the VP / VE schedules their trajectories blow up (1e4 .. 1e10) and a
    # comparison would only measure chaos.  The fixtures therefore scale the LAST layer by OUT_SCALE (recorded; the tests
    # apply the same factor), which keeps F small and the trajectories O(sigma_max).
    OUT_SCALE = 0.02
    mlp = MLPUncond(2, [16, 16], nonlinearity=torch.nn.SiLU())
    load_synth(mlp, 106)                                   # the weights of mlp_silu
    p2d = PUNetG(PUNetGConfig(dimension=2, model_channels=8))
    load_synth(p2d, 101)                                   # the weights of punetg2d_mc8
    with torch.no_grad():
        last = [m for m in mlp.modules() if isinstance(m, torch.nn.Linear)][-1]
        for t_ in (last.weight, last.bias, p2d.convout.weight, p2d.convout.bias):
            t_.mul_(OUT_SCALE)
    for tag in ("vp", "ve", "sr3"):
        case(f"precond_{tag}_mlp", tag, mlp, 106, "mlp_silu", (16, 2), 301)
    for tag in ("vp", "ve"):
        case(f"precond_{tag}_punetg2d", tag, p2d, 101, "punetg2d_mc8", (2, 1, 16, 16), 302, nsteps=3)



class _Randn:
    """Replace torch.randn for ONE expected shape with a pre-drawn tensor (EnsembleKarrasModule.loss_fn draws its [B, E, ...]
    noise with torch.randn(shape, device=..., dtype=...), karrasmodule_new.py:1024-1025)."""

    def __init__(self, tensor):
        self.tensor, self.used = tensor, 0

    def __enter__(self):
        self.orig = torch.randn

        def fake(*a, **k):
            shape = tuple(a[0]) if len(a) == 1 and isinstance(a[0], (tuple, list, torch.Size)) else tuple(a)
            if shape == tuple(self.tensor.shape):
                self.used += 1
                return self.tensor.clone()
            return self.orig(*a, **k)
        torch.randn = fake
        return self

    def __exit__(self, *a):
        torch.randn = self.orig


class ToyAutoencoder(torch.nn.Module):
    """A fixed, parameter-carrying encode / decode pair for the latent-diffusion fixtures (NOT an inverse pair -- the wrapper
    never assumes one): encode = 2x2 average pooling then a 1x1 conv, decode = a 1x1 conv then nearest 2x upsampling.
    tests/test_gpu_ensemble_latent.py rebuilds it from the same numbers."""

    def __init__(self):
        super().__init__()
        self.enc = torch.nn.Conv2d(1, 1, 1)
        self.dec = torch.nn.Conv2d(1, 1, 1)
        with torch.no_grad():
            self.enc.weight.fill_(0.8)
            self.enc.bias.fill_(0.05)
            self.dec.weight.fill_(0.7)
            self.dec.bias.fill_(0.02)

    def encode(self, x):
        return self.enc(torch.nn.functional.avg_pool2d(x, 2))

    def decode(self, z):
        return torch.nn.functional.interpolate(self.dec(z), scale_factor=2, mode="nearest")


def ensemble_goldens():
    """SURVEY 8(f)-4 from the LIVE reference: EnsembleKarrasModule.loss_fn (karrasmodule_new.py:963-1149) with the
    ensemble-aware Huber / MSE / CRPS metrics (custom_losses.py:536-690, 765-865), with and without a mask, for E = 3 and
    for the n_ensemble = 1 path of an ensemble-configured module; and the latent-diffusion wrapper of KarrasModule
    (karrasmodule.py:583-587, 1192-1234): loss in the latent space and decode after sampling.
    ->  tests/golden/ensemble_punetg2d.pt, latent_punetg2d.pt.   python oracle/make_goldens.py --only ensemble"""
    os.makedirs(OUT, exist_ok=True)
    refload.load_reference()
    import diffsci.models as M
    from diffsci.models.karras.karrasmodule_new import EnsembleKarrasModule, EnsembleKarrasModuleConfig
    from diffsci.models.nets.punetg import PUNetG
    from diffsci.models.nets.punetg_config import PUNetGConfig
    torch.set_num_threads(8)
    B, E, shape = 3, 3, (1, 16, 16)
    net = PUNetG(PUNetGConfig(dimension=2, model_channels=8))
    load_synth(net, 101)                                   # the weights of punetg2d_mc8
    torch.manual_seed(401)
    x = torch.randn(B, *shape) * 0.5
    sigma = torch.tensor([0.07, 0.6, 2.5])
    noise = torch.randn(B, E, *shape)
    noise1 = torch.randn(B, 1, *shape)
    mask = (torch.rand(B, 1, *shape[1:]) > 0.6).float()
    mask[2] = 1.0                                           # a fully masked sample: the clamp(min=1) branches
    out = dict(net="punetg2d_mc8", B=B, E=E, x=x, sigma=sigma, noise=noise, noise1=noise1, mask=mask, cases={})
    for metric in ("huber", "mse", "CRPS"):
        cfg = EnsembleKarrasModuleConfig.from_edm(loss_metric=metric)
        cfg.ensemble_size_train = E
        mod = EnsembleKarrasModule(net, cfg)
        mod.train()
        for tag, nz, n_ens, mk in (("E3", noise, E, None), ("E3_mask", noise, E, mask), ("E1", noise1, 1, None),
                                   ("E1_mask", noise1, 1, mask)):
            net.zero_grad()
            if n_ens > 1:
                with _Randn(nz) as ctx:
                    L = mod.loss_fn(x, sigma, None, mk, n_ensemble=n_ens)
                assert ctx.used == 1
            else:
                with _Noise([nz[:, 0]]):
                    L = mod.loss_fn(x, sigma, None, mk, n_ensemble=1)
            L.backward()
            out["cases"][f"{metric}_{tag}"] = dict(loss=L.detach().clone(),
                                                   grads={k: p.grad.clone() for k, p in net.named_parameters() if GRAD_KEYS.search(k)})
            print(metric, tag, float(L.detach()))
    torch.save(out, os.path.join(OUT, "ensemble_punetg2d.pt"))

    # ---- latent-diffusion wrapper (plain KarrasModule with an autoencoder)
    ae = ToyAutoencoder()
    mod = M.KarrasModule(net, M.KarrasModuleConfig.from_edm(), autoencoder=ae)
    mod.train()
    torch.manual_seed(402)
    xd = torch.randn(2, 1, 32, 32) * 0.5                    # data space 32 x 32 -> latent 16 x 16
    lat = dict(net="punetg2d_mc8")
    sg = torch.tensor([0.2, 1.7])
    ln = torch.randn(2, 1, 16, 16)
    net.zero_grad()
    with _Noise([ln]):
        L = mod.loss_fn(xd, sg, None, None)
    L.backward()
    lat.update(x=xd, sigma=sg, noise=ln, loss=L.detach().clone(), encoded=ae.encode(xd).detach(),
               grads={k: p.grad.clone() for k, p in net.named_parameters() if GRAD_KEYS.search(k)})
    assert all(p.grad is None for p in ae.parameters()) and not any(p.requires_grad for p in ae.parameters())
    mod.eval()
    wn = torch.randn(2, 1, 16, 16)
    with torch.no_grad():
        lat["white_noise"] = wn
        lat["sample_latent"] = mod.propagate_white_noise(wn, nsteps=4, return_in_latent_space=True)
        lat["sample_decoded"] = mod.propagate_white_noise(wn, nsteps=4)
        lat["sample_hist_decoded"] = mod.propagate_white_noise(wn, nsteps=4, record_history=True)
        torch.manual_seed(77)
        lat["sample_api"] = mod.sample(2, [1, 16, 16], nsteps=4, is_latent_shape=True)
    print("latent loss", float(L.detach()), "decoded absmax", float(lat["sample_decoded"].abs().max()),
          tuple(lat["sample_hist_decoded"].shape))
    torch.save(lat, os.path.join(OUT, "latent_punetg2d.pt"))

    # ---- learned uncertainty weighting (has_dynamic_loss_weight; karrasmodule.py:594-602, 1243-1278)
    dyn = dict(net="punetg2d_mc8", nhidden=8)
    mod = M.KarrasModule(net, M.KarrasModuleConfig.from_edm(dynamic_loss_weight=8))
    mod.train()
    torch.manual_seed(403)
    dlw = mod.dynamic_loss_weight
    with torch.no_grad():
        dlw.fourier_weights.copy_(torch.randn(8))
        dlw.fourier_bias.copy_(torch.rand(8))
        dlw.linear.weight.copy_(torch.randn(1, 8) * 0.3)
        dlw.linear.bias.fill_(0.1)
    dyn["dlw_state"] = {k: v.clone() for k, v in dlw.state_dict().items()}
    xq, sq, nq = torch.randn(3, 1, 16, 16) * 0.5, torch.tensor([0.05, 0.7, 3.0]), torch.randn(3, 1, 16, 16)
    mq = (torch.rand(3, 1, 16, 16) > 0.5).float()
    dyn.update(x=xq, sigma=sq, noise=nq, mask=mq)
    for metric in ("huber", "mse"):
        mod = M.KarrasModule(net, M.KarrasModuleConfig.from_edm(dynamic_loss_weight=8, loss_metric=metric))
        mod.dynamic_loss_weight.load_state_dict(dyn["dlw_state"])
        mod.train()
        for tag, mk in (("", None), ("_mask", mq)):
            net.zero_grad()
            mod.dynamic_loss_weight.zero_grad()
            with _Noise([nq]):
                L = mod.loss_fn(xq, sq, None, mk)
            L.backward()
            dyn[f"{metric}{tag}"] = dict(loss=L.detach().clone(),
                                         grads={k: p.grad.clone() for k, p in net.named_parameters() if GRAD_KEYS.search(k)},
                                         dlw_grads={k: p.grad.clone() for k, p in mod.dynamic_loss_weight.named_parameters()})
            print("dynamic", metric, tag, float(L.detach()))
    torch.save(dyn, os.path.join(OUT, "dynweight_punetg2d.pt"))


def dropout_goldens():
    """Training-mode dropout (commonlayers.py:792, 829-831; adm.py:279, 323-329) from the LIVE reference with INJECTED masks:
    torch.nn.Dropout.forward is replaced by x * keep / (1 - p) with a recorded keep mask per block, so the restatement
    (nets_oracle: dropout_masks) can be pinned on forward output and parameter gradients.
    ->  tests/golden/dropout_{punetg2d,adm2d}.pt.   python oracle/make_goldens.py --only dropout"""
    os.makedirs(OUT, exist_ok=True)
    refload.load_reference()
    from diffsci.models.nets.punetg import PUNetG
    from diffsci.models.nets.punetg_config import PUNetGConfig
    from diffsci.models.nets.adm import ADM, ADMConfig
    torch.set_num_threads(8)
    P = 0.25

    def case(name, net, base_name, seed):
        base = torch.load(os.path.join(OUT, base_name + ".pt"), weights_only=False)
        load_synth(net, base["seed"])
        net.train()
        x, t = base["x"], base["t"]
        names = {id(m): n for n, m in net.named_modules()}
        gen_ = torch.Generator().manual_seed(seed)
        keep = {}
        orig = torch.nn.Dropout.forward

        def fake(self, inp):
            blk = names[id(self)].rsplit(".", 1)[0]
            if self.p == 0.0 or "cond" in names[id(self)]:
                return inp
            keep[blk] = (torch.rand(inp.shape, generator=gen_) >= self.p)
            return inp * keep[blk].to(inp) / (1.0 - self.p)
        torch.nn.Dropout.forward = fake
        try:
            net.zero_grad()
            y = net(x, t)
            torch.manual_seed(seed + 1)
            dF = torch.randn_like(y)
            (y * dF).sum().backward()
        finally:
            torch.nn.Dropout.forward = orig
        out = dict(net=base_name, p=P, keep={k: v.to(torch.uint8) for k, v in keep.items()}, dF=dF, y=y.detach().clone(),
                   grads={k: p.grad.clone() for k, p in net.named_parameters() if GRAD_KEYS.search(k) or "conv2" in k and ".0." in k})
        torch.save(out, os.path.join(OUT, name + ".pt"))
        print(name, len(keep), "sites; |y| max", float(y.detach().abs().max()), "grads", len(out["grads"]))

    case("dropout_punetg2d", PUNetG(PUNetGConfig(dimension=2, model_channels=8, dropout=P)), "punetg2d_mc8", 501)
    base = torch.load(os.path.join(OUT, "adm2d_mc8.pt"), weights_only=False)
    case("dropout_adm2d", ADM(ADMConfig(**dict(base["cfg"], dropout=P))), "adm2d_mc8", 502)


def nobias_goldens():
    """PUNetGConfig(bias=False) (punetg.py:188-216, 389-394): no conv bias, a constant ones channel appended to the network
    input.  Forward (fp32 + fp64), denoiser, Heun / Euler-Maruyama sampling, loss + gradients from the LIVE reference, 2-D
    and 3-D  ->  tests/golden/nobias_punetg{2,3}d.pt.   python oracle/make_goldens.py --only nobias"""
    os.makedirs(OUT, exist_ok=True)
    refload.load_reference()
    import diffsci.models as M
    from diffsci.models.nets.punetg import PUNetG
    from diffsci.models.nets.punetg_config import PUNetGConfig
    torch.set_num_threads(8)
    for name, kw, shape, seed in (("nobias_punetg2d", dict(dimension=2, model_channels=8, bias=False), (2, 1, 16, 16), 111),
                                  ("nobias_punetg3d", dict(dimension=3, model_channels=8, bias=False, input_channels=2,
                                                           output_channels=2), (2, 2, 8, 8, 8), 112)):
        net = PUNetG(PUNetGConfig(**kw))
        man = load_synth(net, seed)
        assert not any(k.endswith("conv1.bias") or k.startswith("convin.bias") for k, _ in man)
        net.eval()
        torch.manual_seed(seed)
        x, t = torch.randn(*shape), torch.tensor([-1.1, 0.7])
        with torch.no_grad():
            y32 = net(x, t)
            y64 = net.double()(x.double(), t.double())
            net.float()
        out = {"kind": "punetg", "cfg": kw, "manifest": man, "seed": seed, "x": x, "t": t, "y": y32, "y64": y64, "nsteps": 4}
        mod = M.KarrasModule(net, M.KarrasModuleConfig.from_edm())
        mod.eval()
        wn = torch.randn(*shape)
        sg = torch.tensor([0.3, 4.0])
        xin = torch.randn(*shape) * (1 + sg.view(-1, *([1] * (len(shape) - 1))))
        with torch.no_grad():
            D, _ = mod.get_denoiser(xin, sg)
            out.update(white_noise=wn, den_x=xin, den_sigma=sg, den_D=D,
                       heun_hist=mod.propagate_white_noise(wn, nsteps=4, record_history=True))
            noises = [torch.randn(*shape) for _ in range(4)]
            out["noises"] = noises
            with _Noise(noises):
                out["em"] = mod.propagate_white_noise(wn, nsteps=4, integrator="euler-maruyama")
        mod.train()
        x0, ln = torch.randn(*shape) * 0.5, torch.randn(*shape)
        net.zero_grad()
        with _Noise([ln]):
            L = mod.loss_fn(x0, sg, None, None)
        L.backward()
        out.update(loss_x=x0, loss_noise=ln, loss_sigma=sg, loss_huber=L.detach().clone(),
                   loss_huber_grads={k: p.grad.clone() for k, p in net.named_parameters() if GRAD_KEYS.search(k)})
        torch.save(out, os.path.join(OUT, name + ".pt"))
        print(name, "fp32-vs-fp64", float((y32 - y64).abs().max() / y64.abs().max()), "loss", float(L.detach()),
              "convin", [s_ for k, s_ in man if k == "convin.weight"])


def fullsteps_goldens(which=("c1", "c2", "c5", "c4")):
    """BASELINE.json's configurations at their REAL step counts through the LIVE reference's own sampling loop
    (KarrasModule.propagate_white_noise, karrasmodule.py:867-931; Scheduler.propagate, schedulers.py:48-89) on CPU, fp32 and --
    for the tolerance budget -- fp64:
      c1  MLPUncond(2, [128, 128, 128], SiLU), Heun 18 steps, B = 64
      c2  PUNetG 2-D mc = 128, 1 x 28 x 28, Heun 40 steps, B = 2
      c4  PUNetG 3-D mc = 64, 1 x 64^3, Heun 64 steps, B = 1                 (138 TFLOP per precision: ~15 / 30 min here)
      c5  PUNetG 2-D mc = 64, 1 x 256 x 256, Euler-Maruyama 256 steps, B = 1, step noise from a seeded CPU generator
    Weights are synth_state_dict(manifest, seed) (not stored), x_T = randn under torch.manual_seed (not stored either: the test
    regenerates it with the same CPU generator).   python oracle/make_goldens.py --only fullsteps [c1 c2 c5 c4]"""
    os.makedirs(OUT, exist_ok=True)
    refload.load_reference()
    import diffsci.models as M
    from diffsci.models.nets.mlp import MLPUncond
    from diffsci.models.nets.punetg import PUNetG
    from diffsci.models.nets.punetg_config import PUNetGConfig
    torch.set_num_threads(os.cpu_count())
    cases = {
        "c1": dict(net="mlp", kw=dict(dim=2, hidden_dims=[128, 128, 128]), shape=(64, 2), nsteps=18, integ="heun", wseed=501,
                   xseed=601),
        "c2": dict(net="punetg", kw=dict(dimension=2, model_channels=128), shape=(2, 1, 28, 28), nsteps=40, integ="heun",
                   wseed=502, xseed=602),
        "c5": dict(net="punetg", kw=dict(dimension=2), shape=(1, 1, 256, 256), nsteps=256, integ="euler-maruyama", wseed=505,
                   xseed=605, nseed=705),
        "c4": dict(net="punetg", kw=dict(dimension=3), shape=(1, 1, 64, 64, 64), nsteps=64, integ="heun", wseed=504, xseed=604),
    }
    import time
    for name in which:
        c = cases[name]
        if c["net"] == "mlp":
            net = MLPUncond(c["kw"]["dim"], c["kw"]["hidden_dims"], nonlinearity=torch.nn.SiLU())
        else:
            net = PUNetG(PUNetGConfig(**c["kw"]))
        man = load_synth(net, c["wseed"])
        net.eval()
        mod = M.KarrasModule(net, M.KarrasModuleConfig.from_edm())
        mod.eval()
        torch.manual_seed(c["xseed"])
        wn = torch.randn(*c["shape"])
        noises = None
        if "nseed" in c:
            g = torch.Generator().manual_seed(c["nseed"])
            noises = [torch.randn(c["shape"], generator=g) for _ in range(c["nsteps"])]
        out = {}
        for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
            m = mod.double() if dt == torch.float64 else mod
            t0 = time.time()
            with torch.no_grad():
                if noises is None:
                    y = m.propagate_white_noise(wn.to(dt), nsteps=c["nsteps"], integrator=c["integ"])
                else:
                    with _Noise(noises):
                        y = m.propagate_white_noise(wn.to(dt), nsteps=c["nsteps"], integrator=c["integ"])
            out[tag] = y.detach()
            print(f"fullsteps {name} {tag}: {time.time() - t0:.1f} s, |y|max {float(y.abs().max()):.4f}", flush=True)
            torch.save(dict(kw=c["kw"], net=c["net"], manifest=man, wseed=c["wseed"], xseed=c["xseed"], nseed=c.get("nseed"),
                            shape=c["shape"], nsteps=c["nsteps"], integrator=c["integ"],
                            out_f32=out.get("f32"), out_f64=out.get("f64")), os.path.join(OUT, f"fullsteps_{name}.pt"))
        d = (out["f32"].double() - out["f64"]).abs()
        print(f"fullsteps {name}: reference fp32 vs fp64 max-abs {float(d.max()):.3e} rms {float(d.pow(2).mean().sqrt()):.3e} "
              f"(field rms {float(out['f64'].pow(2).mean().sqrt()):.3e})", flush=True)


def main():
    if "--only" in sys.argv and sys.argv[sys.argv.index("--only") + 1] == "fullsteps":
        rest = [a for a in sys.argv[sys.argv.index("--only") + 2:] if not a.startswith("-")]
        return fullsteps_goldens(tuple(rest) if rest else ("c1", "c2", "c5", "c4"))
    if "--only" in sys.argv and sys.argv[sys.argv.index("--only") + 1] == "nobias":
        return nobias_goldens()
    if "--only" in sys.argv and sys.argv[sys.argv.index("--only") + 1] == "dropout":
        return dropout_goldens()
    if "--only" in sys.argv and sys.argv[sys.argv.index("--only") + 1] == "ensemble":
        return ensemble_goldens()
    if "--only" in sys.argv and sys.argv[sys.argv.index("--only") + 1] == "precond":
        return precond_goldens()
    if "--only" in sys.argv and sys.argv[sys.argv.index("--only") + 1] == "circular":
        return circular_goldens()
    if "--only" in sys.argv and sys.argv[sys.argv.index("--only") + 1] == "inpaint":
        return inpaint_goldens()
    if "--only" in sys.argv and sys.argv[sys.argv.index("--only") + 1] == "cond":
        return cond_goldens()
    os.makedirs(OUT, exist_ok=True)
    refload.load_reference()
    import diffsci.models as M
    from diffsci.models.nets.punetg import PUNetG
    from diffsci.models.nets.punetg_config import PUNetGConfig
    from diffsci.models.nets.adm import ADM, ADMConfig
    from diffsci.models.nets.mlp import MLPUncond
    from diffsci.models.karras.ema import ModelEMA, _power_function_beta, _power_function_exp_from_std
    from diffsci.models.karras import integrators as I

    torch.set_num_threads(8)

    # ------------------------------------------------------------------ numerics
    sch = M.EDMScheduler()
    pre = M.EDMPreconditioner()
    ns = M.EDMNoiseSampler()
    sig = torch.tensor([0.002, 0.0137, 0.25, 0.5, 1.0, 3.7, 21.0, 80.0])
    g = {
        "steps": {n: sch.create_steps(n) for n in (3, 4, 11, 19, 41, 65, 257)},
        "sigma": sig,
        "c_in": pre.input_scaling(sig), "c_out": pre.output_scaling(sig),
        "c_skip": pre.skip_scaling(sig), "c_noise": pre.noise_conditioner(sig),
        "loss_weight": ns.loss_weighting(sig),
    }
    torch.manual_seed(7)
    xi = torch.randn(16)
    g["sigma_xi"] = xi
    g["sigma_from_xi"] = torch.exp(xi * ns.prior_std + ns.prior_mean)
    g["ema_power_exp"] = {s: _power_function_exp_from_std(s) for s in (0.05, 0.1)}
    g["ema_power_beta"] = {(s, n): _power_function_beta(s, n) for s in (0.05, 0.1) for n in (1, 2, 10, 1000)}
    lin = torch.nn.Linear(2, 1, bias=False)
    ema = ModelEMA(lin, ema_type="traditional", halflife_steps=100.0, rampup_ratio=0.5)
    g["ema_trad_beta"] = {n: ema._traditional_beta(n) for n in (1, 2, 50, 1000)}
    torch.save(g, os.path.join(OUT, "numerics.pt"))

    # ------------------------------------------------------------------ nets
    def net_case(name, module, cfg_kwargs, x, t, seed, kind):
        man = load_synth(module, seed)
        module.eval()
        with torch.no_grad():
            y32 = module(x, t)
            y64 = module.double()(x.double(), t.double())
            module.float()
        torch.save({"kind": kind, "cfg": cfg_kwargs, "manifest": man, "seed": seed,
                    "x": x, "t": t, "y": y32, "y64": y64}, os.path.join(OUT, name + ".pt"))
        print(name, tuple(y32.shape), float(y32.abs().max()),
              "fp32-vs-fp64 maxrel", float((y32 - y64).abs().max() / y64.abs().max()))
        return module

    torch.manual_seed(11)
    kw = dict(dimension=2, model_channels=8)
    p2d = net_case("punetg2d_mc8", PUNetG(PUNetGConfig(**kw)), kw,
                   torch.randn(2, 1, 28, 28), torch.tensor([-1.3, 0.9]), 101, "punetg")
    kw = dict(dimension=3, model_channels=8)
    net_case("punetg3d_mc8", PUNetG(PUNetGConfig(**kw)), kw,
             torch.randn(2, 1, 8, 8, 8), torch.tensor([0.4, -2.0]), 102, "punetg")
    kw = dict(dimension=2, model_channels=8, input_channels=2, output_channels=3, channel_expansion=[2],
              number_resnet_attn_block=3, attn_residual=True)
    net_case("punetg2d_multi", PUNetG(PUNetGConfig(**kw)), kw,
             torch.randn(1, 2, 16, 24), torch.tensor([0.1]), 103, "punetg")
    kw = dict(input_channels=3, output_channels=3, model_channels=8, time_embed_dim=16, output_embed_dim=32)
    net_case("adm2d_mc8", ADM(ADMConfig(**kw)), kw,
             torch.randn(2, 3, 16, 16), torch.tensor([-0.7, 1.1]), 104, "adm")
    kw = dict(input_channels=1, output_channels=1, model_channels=8, time_embed_dim=16, output_embed_dim=32,
              skip_integration_type="add", channel_expansion=[2])
    net_case("adm2d_add", ADM(ADMConfig(**kw)), kw,
             torch.randn(1, 1, 12, 20), torch.tensor([0.3]), 105, "adm")
    mlp = MLPUncond(2, [16, 16], nonlinearity=torch.nn.SiLU())
    kw = dict(dim=2, hidden_dims=[16, 16], act="silu")
    mlp = net_case("mlp_silu", mlp, kw, torch.randn(32, 2), torch.randn(32), 106, "mlp")

    # ------------------------------------------------------------------ denoiser / samplers / loss
    def sampler_case(name, model, shape, nsteps, seed):
        cfg = M.KarrasModuleConfig.from_edm()
        mod = M.KarrasModule(model, cfg)
        mod.eval()
        torch.manual_seed(seed)
        B = shape[0]
        wn = torch.randn(*shape)
        out = {"nsteps": nsteps, "white_noise": wn}
        with torch.no_grad():
            sg = torch.exp(torch.randn(B) * 1.2 - 1.2)
            xin = torch.randn(*shape) * (1 + sg.view(-1, *([1] * (len(shape) - 1))))
            D, cn = mod.get_denoiser(xin, sg)
            out.update(den_x=xin, den_sigma=sg, den_D=D, den_cnoise=cn, den_score=mod.get_score(xin, sg))
            out["heun_hist"] = mod.propagate_white_noise(wn, nsteps=nsteps, record_history=True)
            out["euler"] = mod.propagate_white_noise(wn, nsteps=nsteps, integrator="euler")
            noises = [torch.randn(*shape) for _ in range(nsteps)]
            out["noises"] = noises
            with _Noise(noises):
                out["em"] = mod.propagate_white_noise(wn, nsteps=nsteps, integrator="euler-maruyama")
            cfg.noisescheduler.langevin_const = 0.5
            cfg.noisescheduler.langevin_interval = (0.1, 10.0)
            with _Noise(noises):
                out["em_interval"] = mod.propagate_white_noise(wn, nsteps=nsteps, integrator="euler-maruyama")
            cfg.noisescheduler.langevin_const = 1.0
            cfg.noisescheduler.langevin_interval = None
            with _Noise(noises):
                out["karras"] = mod.propagate_white_noise(wn, nsteps=nsteps, integrator="karras")
            with _Noise(noises):
                out["karras_custom"] = mod.propagate_white_noise(
                    wn, nsteps=nsteps, integrator=I.KarrasIntegrator(s_schurn=10, s_tmin=0.01, s_tmax=1.0, s_noise=1.0))
        # loss + grads (training path)
        x0 = torch.randn(*shape) * 0.5
        ln = torch.randn(*shape)
        mask = (torch.rand(*shape) > 0.7).float()
        out.update(loss_x=x0, loss_noise=ln, loss_sigma=sg, loss_mask=mask)
        for metric in ("huber", "mse"):
            cfg2 = M.KarrasModuleConfig.from_edm(loss_metric=metric)
            mod2 = M.KarrasModule(model, cfg2)
            for use_mask in (False, True):
                model.zero_grad()
                with _Noise([ln]):
                    L = mod2.loss_fn(x0, sg, None, mask if use_mask else None)
                L.backward()
                key = f"loss_{metric}{'_mask' if use_mask else ''}"
                out[key] = L.detach()
                # a representative subset keeps the fixture small: first/last layers, one deep
                # conv, norm affine, time MLP, attention projections
                out[key + "_grads"] = {k: p.grad.clone() for k, p in model.named_parameters()
                                       if GRAD_KEYS.search(k)}
        torch.save(out, os.path.join(OUT, name + ".pt"))
        print(name, "heun final absmax", float(out["heun_hist"][-1].abs().max()))

    sampler_case("sampler_mlp", mlp, (32, 2), 6, 201)
    sampler_case("sampler_punetg2d", p2d, (2, 1, 28, 28), 4, 202)

    # nsteps edge: nsteps=2 is the smallest valid schedule
    cfg = M.KarrasModuleConfig.from_edm()
    mod = M.KarrasModule(mlp, cfg)
    torch.manual_seed(5)
    wn = torch.randn(8, 2)
    with torch.no_grad():
        torch.save({"white_noise": wn, "heun2": mod.propagate_white_noise(wn, nsteps=2)},
                   os.path.join(OUT, "sampler_edge.pt"))

    # ------------------------------------------------------------------ EMA known answers (tests/test_karras_ema.py:23-52)
    lin = torch.nn.Linear(3, 2)
    torch.manual_seed(3)
    ema = ModelEMA(lin, ema_type="traditional", decay=0.9)
    traj = []
    for i in range(3):
        with torch.no_grad():
            for p in lin.parameters():
                p.add_(torch.randn_like(p))
        ema.update(lin)
        traj.append({"params": {k: v.detach().clone() for k, v in lin.named_parameters()},
                     "shadow": {k: v.clone() for k, v in ema.selected_profile()["params"].items()},
                     "beta": ema.last_beta})
    torch.save({"init": None, "traj": traj}, os.path.join(OUT, "ema.pt"))
    print("done ->", OUT)


if __name__ == "__main__":
    main()
