"""CPU oracle for the Karras/EDM numerics of DiffSci (TEST INFRASTRUCTURE ONLY).

This file is a plain-PyTorch (CPU, fp32 or fp64) restatement of the reference algorithm
for the sampler / loss / EMA part of the hot path.  It is *not* product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker / the CPU arm.

The arithmetic of this path lives in PyTorch ATen (third-party; torch 2.11.0+cu128 in this
image, the reference pins no version in requirements.txt:10).  The oracle therefore calls
the same ATen CPU primitives in the same order as the reference call sites quoted on each
function.  Parity is PINNED: ``tests/golden/*.pt`` were produced by running the live
reference (``oracle/make_goldens.py``, importing /root/reference) and
``tests/test_oracle_vs_golden.py`` checks every function here against them.

All functions take/return CPU tensors; ``net`` arguments are callables
``net(x_scaled, c_noise) -> F`` (see ``oracle/nets_oracle.py``).
"""
from __future__ import annotations

import math
from typing import Callable, Optional, Sequence

import numpy as np
import torch

Net = Callable[[torch.Tensor, torch.Tensor], torch.Tensor]


def bcast(v: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """diffsci/torchutils.py:4-40 -- append singleton dims so v[B] broadcasts against x[B,...]."""
    return v.reshape(v.shape + (1,) * (x.ndim - v.ndim)).to(x)


# ----------------------------------------------------------------------------- schedule
def edm_steps(n: int, sigma_min=0.002, sigma_max=80.0, rho=7.0, dtype=torch.float32) -> torch.Tensor:
    """EDMScheduler.create_steps (karras/schedulers.py:377-385): n-1 rho-spaced sigmas then 0.

    The reference evaluates this with 0-dim fp32 buffers; the op order below is the same.
    """
    rho_t = torch.tensor(rho, dtype=dtype)
    smax = torch.tensor(sigma_max, dtype=dtype)
    smin = torch.tensor(sigma_min, dtype=dtype)
    s = torch.arange(n - 1).to(rho_t) / (n - 2)
    start = smax ** (1 / rho_t)
    end = smin ** (1 / rho_t)
    steps = (start + s * (end - start)) ** rho_t
    return torch.cat([steps, torch.zeros([1]).to(steps)])


# ----------------------------------------------------------------------------- preconditioner
def edm_precond(sigma: torch.Tensor, sigma_data: float = 0.5):
    """EDMPreconditioner (karras/preconditioners.py:30-53) -> (c_in, c_out, c_skip, c_noise)."""
    sd = torch.tensor(sigma_data, dtype=sigma.dtype)
    c_skip = sd ** 2 / (sigma ** 2 + sd ** 2)
    c_out = sigma * sd / torch.sqrt(sigma ** 2 + sd ** 2)
    c_in = 1 / torch.sqrt(sigma ** 2 + sd ** 2)
    c_noise = 0.5 * torch.log(sigma)
    return c_in, c_out, c_skip, c_noise


def denoiser(net: Net, x: torch.Tensor, sigma: torch.Tensor, sigma_data: float = 0.5) -> torch.Tensor:
    """KarrasModule.get_denoiser (karras/karrasmodule.py:690-719), unconditional branch."""
    c_in, c_out, c_skip, c_noise = edm_precond(sigma, sigma_data)
    F = net(bcast(c_in, x) * x, c_noise)
    return bcast(c_out, x) * F + bcast(c_skip, x) * x


def guided_net(net_cond: Net, net_uncond: Optional[Net], guidance: float = 1.0, conditional: bool = True) -> Net:
    """The network dispatch of KarrasModule.get_denoiser (karrasmodule.py:703-716): conditional call iff
    ``conditional and guidance != 0``; classifier-free mix (1-g) F_u + g F_c iff additionally guidance != 1."""
    if not (conditional and guidance != 0.0):
        return net_uncond
    if guidance == 1.0:
        return net_cond
    return lambda x, t: (1 - guidance) * net_uncond(x, t) + guidance * net_cond(x, t)


def score(net: Net, x, sigma, sigma_data=0.5):
    """KarrasModule.get_score (karrasmodule.py:721-733)."""
    return (denoiser(net, x, sigma, sigma_data) - x) / (bcast(sigma, x) ** 2)


def edm_rhs(net: Net, x, ti, sigma_data=0.5, stochastic=False, langevin_const=1.0,
            langevin_interval=None):
    """Scheduler.rhs, EDM branch (karras/schedulers.py:247-274; s=1, sigma=t, sigma'=1)."""
    t = ti * torch.ones(x.shape[0]).to(x)
    t_ = bcast(t, x)
    sc = score(net, x, t, sigma_data)
    res = -(t_ * (1 + 0 * t_)) * sc
    if stochastic:
        res = res + (-(langevin_factor(t_, langevin_const, langevin_interval) * sc))
    return res


def langevin_factor(t, langevin_const=1.0, langevin_interval=None):
    """Scheduler.langevin_factor (schedulers.py:219-240) for the EDM functions."""
    std = (1 + 0 * t) ** 2 * (1 + 0 * t) * (1 * t)
    if langevin_interval is not None:
        t0 = t.reshape(-1)[0]
        if not (t0 > langevin_interval[0] and t0 < langevin_interval[1]):
            return 0 * t
    return langevin_const * std + 0 * t


def noise_injection(t, langevin_const=1.0, langevin_interval=None):
    """Scheduler.noise_injection (schedulers.py:242-245)."""
    return torch.sqrt(2 * langevin_factor(t, langevin_const, langevin_interval))


# ----------------------------------------------------------------------------- integrators
def step_euler(net, x, t, dt, **kw):
    """EulerIntegrator.step (karras/integrators.py:29-35)."""
    return x + dt * edm_rhs(net, x, t, **kw)


def step_heun(net, x, t, dt, **kw):
    """HeunIntegrator.step (integrators.py:38-54)."""
    r1 = edm_rhs(net, x, t, **kw)
    if (t + dt) > 0:
        r2 = edm_rhs(net, x + dt * r1, t + dt, **kw)
    elif (t + dt) == 0:
        r2 = r1
    else:
        raise ValueError("t+dt < 0 is not supported")
    return x + 0.5 * (r1 + r2) * dt


def step_euler_maruyama(net, x, t, dt, noise, sigma_data=0.5, langevin_const=1.0,
                        langevin_interval=None):
    """EulerMaruyamaIntegrator.step (integrators.py:57-69) with the N(0,1) draw injected."""
    r = edm_rhs(net, x, t, sigma_data=sigma_data, stochastic=True,
                langevin_const=langevin_const, langevin_interval=langevin_interval)
    return x + r * dt + noise_injection(t, langevin_const, langevin_interval) * noise * torch.sqrt(torch.abs(dt))


def step_karras(net, x, t, dt, noise, nsteps, s_churn=40.0, s_tmin=0.05, s_tmax=50.0,
                s_noise=1.003, sigma_data=0.5):
    """KarrasIntegrator.step (integrators.py:72-113), EDM functions, N(0,1) draw injected."""
    back = min(s_churn / nsteps, np.sqrt(2) - 1)
    if s_tmin is not None and not (s_tmin <= t <= s_tmax):
        back = 0
    sigma = 1 * t
    sigma_n = sigma + back * sigma
    t_n = 1 * sigma_n
    scale, scale_n = 1 + 0 * t, 1 + 0 * t_n
    std = scale_n * torch.sqrt(sigma_n ** 2 - sigma ** 2)
    x_n = (scale_n / scale) * x + std * s_noise * noise
    r1 = edm_rhs(net, x_n, t_n, sigma_data=sigma_data)
    dt_n = (t + dt) - t_n
    x = x_n + dt_n * r1
    if (t + dt) > 0:
        r2 = edm_rhs(net, x, t + dt, sigma_data=sigma_data)
        x = x_n + 0.5 * (r1 + r2) * dt_n
    return x


def propagate_backward(net: Net, x: torch.Tensor, nsteps: int, integrator: str = "heun",
                       record_history: bool = False, noises: Optional[Sequence[torch.Tensor]] = None,
                       sigma_data: float = 0.5, langevin_const: float = 1.0,
                       langevin_interval=None, **integrator_kw):
    """Scheduler.propagate(backward=True) (schedulers.py:48-89) for an EDMScheduler."""
    t = edm_steps(nsteps + 1).to(x)
    dt = torch.diff(t)
    hist = [x] if record_history else None
    for i in range(nsteps):
        if integrator == "euler":
            x = step_euler(net, x, t[i], dt[i], sigma_data=sigma_data)
        elif integrator == "heun":
            x = step_heun(net, x, t[i], dt[i], sigma_data=sigma_data)
        elif integrator == "euler-maruyama":
            x = step_euler_maruyama(net, x, t[i], dt[i], noises[i], sigma_data=sigma_data,
                                    langevin_const=langevin_const, langevin_interval=langevin_interval)
        elif integrator == "karras":
            x = step_karras(net, x, t[i], dt[i], noises[i], nsteps, sigma_data=sigma_data, **integrator_kw)
        else:
            raise ValueError(f"Unknown integrator: {integrator}")
        if record_history:
            hist.append(x)
    return torch.stack(hist, 0) if record_history else x


def sample_from_white_noise(net, white_noise, nsteps, integrator="heun", sigma_max=80.0, **kw):
    """KarrasModule.propagate_white_noise (karrasmodule.py:867-905) without latent decode."""
    return propagate_backward(net, white_noise * sigma_max, nsteps, integrator, **kw)


# ----------------------------------------------------------------------------- training
def edm_sigma_from_normal(xi: torch.Tensor, prior_mean=-1.2, prior_std=1.2):
    """EDMNoiseSampler.sample (karras/noisesamplers.py:35-41) given the N(0,1) draw."""
    return torch.exp(xi * torch.tensor(prior_std, dtype=xi.dtype) + torch.tensor(prior_mean, dtype=xi.dtype))


def edm_loss_weight(sigma, sigma_data=0.5):
    """EDMNoiseSampler.loss_weighting (noisesamplers.py:30-33)."""
    sd = torch.tensor(sigma_data, dtype=sigma.dtype)
    return (sigma ** 2 + sd ** 2) / ((sigma * sd) ** 2)


def dynamic_loss_weight(state, c_noise):
    """DynamicLossWeight.forward (karrasmodule.py:1270-1278): cos features of c_noise, one linear layer -> u [B]."""
    h = torch.cos(c_noise.unsqueeze(1) * state["fourier_weights"] + state["fourier_bias"])
    return (h @ state["linear.weight"].t() + state["linear.bias"]).squeeze(1)


def edm_loss(net, x, sigma, noise, loss_metric="huber", mask=None, sigma_data=0.5, dynamic_state=None):
    """KarrasModule.loss_fn (karrasmodule.py:569-650), single-loss branch, noise injected.  dynamic_state: the state dict of
    a DynamicLossWeight -> weight / exp(u), bias + u with u = u(c_noise) (:594-602)."""
    bs = bcast(sigma, x)
    x_noised = x + bs * noise
    D = denoiser(net, x_noised, sigma, sigma_data)
    w = edm_loss_weight(bs, sigma_data)
    bias = torch.zeros_like(w)
    if dynamic_state is not None:
        u = bcast(dynamic_loss_weight(dynamic_state, edm_precond(sigma, sigma_data)[3]), x)
        w = w / torch.exp(u)
        bias = bias + u
    if loss_metric == "huber":
        l = torch.nn.functional.huber_loss(D, x, reduction="none", delta=1.0)
    elif loss_metric == "mse":
        l = torch.nn.functional.mse_loss(D, x, reduction="none")
    else:
        raise ValueError(loss_metric)
    if mask is not None:
        l = l * (1 - mask.expand_as(l))
    return (w * l + bias).mean()



# ----------------------------------------------------------------------------- ensemble losses
def ensemble_metric(D, x, metric: str, mask=None):
    """The ensemble-aware metrics (custom_losses.py:536-690 Huber / MSE, :765-865 CRPS) as scalars.
    D: [B, E, *shape] -- or [B, *shape], the 4-D branch the same objects take when n_ensemble <= 1 --, x: [B, *shape],
    mask: None | [B, 1 | C, *spatial] (1 = ignore)."""
    single = D.ndim == x.ndim
    if metric == "CRPS":
        D5 = D.unsqueeze(1) if single else D
        B, E = D5.shape[:2]
        flat = D5.reshape(B, E, -1)
        to_target = (flat - x.reshape(B, 1, -1)).abs().mean(2).mean(1)                         # :817-819
        pair = torch.zeros(B, dtype=D.dtype)
        if E > 1:                                                                             # :823-849
            pm = (flat.unsqueeze(2) - flat.unsqueeze(1)).abs().mean(3)                         # [B, E, E]
            iu = torch.triu(torch.ones(E, E), diagonal=1).bool()
            pair = pm[:, iu].sum(1) / max(E * (E - 1) / 2, 1)
        crps = to_target - 0.5 * pair
        if mask is not None:                                                                  # :851-857: a per-sample rescale only
            mexp = mask.expand(x.shape) if mask.shape[1] == 1 else mask
            valid = (~mexp.bool()).reshape(B, -1).sum(1).to(D.dtype).clamp(min=1)
            crps = crps * (valid / flat.shape[2])
        return crps.mean()
    el = (lambda a, b: torch.nn.functional.huber_loss(a, b, reduction="none", delta=1.0)) if metric == "huber" else \
        (lambda a, b: torch.nn.functional.mse_loss(a, b, reduction="none"))

    def masked_mean_4d(pred):                                                                 # :553-560, :610-621
        l = el(pred, x)
        if mask is None:
            return l.mean()
        return (l * (1 - mask)).sum() / (1 - mask).sum().clamp(min=1)
    if single:
        return masked_mean_4d(D)
    B, E = D.shape[:2]
    if metric == "mse":                                                                       # :544-552: member by member
        return torch.stack([masked_mean_4d(D[:, e]) for e in range(E)]).mean()
    l = el(D, x.unsqueeze(1).expand_as(D))                                                    # Huber, :623-690
    dims = tuple(range(1, l.ndim))
    if mask is None:
        return l.mean(dim=dims).mean()
    m5 = mask.unsqueeze(2) if mask.shape[1] == 1 else mask.unsqueeze(1)     # [B,1,1,*sp] | [B,1,C,*sp]
    per_b = (l * (1 - m5)).sum(dim=dims) / (1 - m5).sum(dim=dims).clamp(min=1)                 # the count is NOT expanded over E / C
    return per_b.mean()


def ensemble_loss(net, x, sigma, noise, metric="huber", mask=None, sigma_data=0.5, single=False):
    """EnsembleKarrasModule.loss_fn (karrasmodule_new.py:963-1149) for n_ensemble = noise.shape[1] > 1, and -- single=True,
    noise [B, 1, ...] -- old_loss_fn (:1151-1235) on an ensemble-configured module: one denoiser call on the B*E rows, the
    scalar metric, times the batch MEAN of lambda(sigma)."""
    B, E = noise.shape[:2]
    xn = (x.unsqueeze(1) + bcast(sigma, x).unsqueeze(1) * noise).reshape(B * E, *x.shape[1:])
    D = denoiser(net, xn, sigma.repeat_interleave(E), sigma_data).reshape(B, E, *x.shape[1:])
    if single:
        assert E == 1
        D = D[:, 0]
    return edm_loss_weight(bcast(sigma, x), sigma_data).mean() * ensemble_metric(D, x, metric, mask)

# ----------------------------------------------------------------------------- EMA
def power_function_exp_from_std(std: float) -> float:
    """karras/ema.py:9-15."""
    target = float(std) ** -2
    roots = np.roots([1.0, 7.0, 16.0 - target, 12.0 - target])
    return float(np.max(roots.real))


def ema_beta(kind: str, next_update: int, decay=0.999, halflife_steps=None, rampup_ratio=None, std=0.05) -> float:
    """ModelEMA._traditional_beta / _power_function_beta (ema.py:18-23, 111-121)."""
    if kind == "power":
        if next_update <= 1:
            return 0.0
        return float((1.0 - 1.0 / next_update) ** (power_function_exp_from_std(std) + 1.0))
    if halflife_steps is None:
        return decay
    hl = float(halflife_steps)
    if rampup_ratio is not None:
        hl = min(hl, max(float(next_update), 1.0) * float(rampup_ratio))
    return float(0.5 ** (1.0 / max(hl, 1e-8)))


def ema_update(shadow: torch.Tensor, param: torch.Tensor, beta: float) -> torch.Tensor:
    """ModelEMA.update inner op (ema.py:147): shadow.lerp_(param, 1-beta)."""
    return torch.lerp(shadow, param, 1.0 - beta)


def adamw_step(p, g, m, v, step, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8, wd=1e-4):
    """torch.optim.AdamW single-tensor update (reference default optimizer, karrasmodule.py:497-500)."""
    p = p * (1 - lr * wd)
    m = torch.lerp(m, g, 1 - b1)
    v = v * b2 + (1 - b2) * g * g
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)) + eps
    p = p - (lr / bc1) * (m / denom)
    return p, m, v


# ----------------------------------------------------------------------------- VP / VE / SR3
class SchedFns:
    """schedulingfunctions.py:39-170 restated for tag in {"edm", "vp", "ve"}: s(t), s'(t), sigma(t), sigma'(t),
    sigma^{-1}, and the flags Scheduler.rhs branches on."""

    def __init__(self, tag: str, beta_data: float = 19.9, beta_min: float = 0.1):
        self.tag, self.bd, self.bm = tag, beta_data, beta_min
        self.constant_scaling = tag in ("edm", "ve")            # constant_scaling_fn
        self.has_pf_score_multiplier = tag == "ve"              # VP's flag is False upstream ("TODO: Set to true")

    def _e(self, t):
        return 0.5 * self.bd * t ** 2 + self.bm * t

    def scaling(self, t):
        return torch.exp(-self._e(t) / 2) if self.tag == "vp" else 1 + 0 * t

    def scaling_deriv(self, t):
        if self.tag == "vp":
            return -(self.bd * t + self.bm) / 2 * torch.exp(-self._e(t) / 2)
        return 0 * t

    def noise(self, t):
        if self.tag == "vp":
            return torch.sqrt(torch.exp(self._e(t)) - 1)
        return torch.sqrt(t) if self.tag == "ve" else 1 * t

    def noise_deriv(self, t):
        if self.tag == "vp":
            ex = torch.exp(self._e(t))
            return ((self.bd * t + self.bm) * ex) / (2 * torch.sqrt(ex - 1))
        return 0.5 / torch.sqrt(t) if self.tag == "ve" else 1 + 0 * t

    def inverse_noise(self, s):
        if self.tag == "vp":
            y = torch.log(s ** 2 + 1)
            return (-self.bm + torch.sqrt(self.bm ** 2 + 2 * self.bd * y)) / self.bd
        return s ** 2 if self.tag == "ve" else 1 * s

    def pf_score_multiplier(self, t):
        assert self.tag == "ve"
        return 0.5 + 0 * t


def generic_steps(tag: str, n: int, epsilon_min=1e-3, sigma_min=0.02, sigma_max=100.0) -> torch.Tensor:
    """VPScheduler.create_steps / VEScheduler.create_steps (schedulers.py:411-414, 437-440); EDM: edm_steps."""
    if tag == "vp":
        eps = torch.tensor(epsilon_min)
        return 1 + (torch.arange(n).to(eps) / (n - 1)) * (eps - 1)
    if tag == "ve":
        smin, smax = torch.tensor(float(sigma_min)), torch.tensor(float(sigma_max))
        return smax ** 2 * (smin ** 2 / smax ** 2) ** (torch.arange(n).to(smin) / (n - 1))
    return edm_steps(n)


def generic_precond(kind: str, sigma: torch.Tensor, fns: Optional[SchedFns] = None, M: int = 1000, sigma_data: float = 0.5):
    """(c_in, c_out, c_skip, c_noise) of VPPreconditioner / VEPreconditioner / SR3Preconditioner / EDMPreconditioner
    (preconditioners.py:30-136)."""
    if kind == "vp":
        return 1 / torch.sqrt(sigma ** 2 + 1.0), -sigma, 1 + 0.0 * sigma, (M - 1) * fns.inverse_noise(sigma)
    if kind == "ve":
        return 1 + 0.0 * sigma, sigma, 1 + 0.0 * sigma, torch.log(0.5 * sigma)
    c_in, c_out, c_skip, c_noise = edm_precond(sigma, sigma_data)
    if kind == "sr3":
        sd = torch.as_tensor(sigma_data, dtype=sigma.dtype)
        c_skip = sd ** 2 / (2 * (sigma ** 2 + sd ** 2))
        c_out = sigma * sd / (2 * torch.sqrt(sigma ** 2 + sd ** 2))
    return c_in, c_out, c_skip, c_noise


def generic_denoiser(net: Net, x, sigma, kind: str, fns: Optional[SchedFns] = None):
    """KarrasModule.get_denoiser (karrasmodule.py:690-719) for any preconditioner."""
    c_in, c_out, c_skip, c_noise = generic_precond(kind, sigma, fns)
    return bcast(c_out, x) * net(bcast(c_in, x) * x, c_noise) + bcast(c_skip, x) * x


def generic_score(net: Net, x, sigma, kind: str, fns: Optional[SchedFns] = None):
    return (generic_denoiser(net, x, sigma, kind, fns) - x) / (bcast(sigma, x) ** 2)


def generic_rhs(net: Net, x, ti, kind: str, fns: SchedFns, stochastic=False, langevin_const=1.0):
    """Scheduler.rhs (schedulers.py:247-294), both branches (constant and non-constant s(t)), backward direction."""
    t = ti * torch.ones(x.shape[0]).to(x)
    t_ = bcast(t, x)
    sigma = fns.noise(t)
    lang = langevin_const * (fns.scaling(t_) ** 2 * fns.noise_deriv(t_) * fns.noise(t_)) + 0 * t_     # langevin_factor :219-241
    if fns.constant_scaling:
        mult = fns.pf_score_multiplier(t_) if fns.has_pf_score_multiplier else bcast(sigma, x) * bcast(fns.noise_deriv(t), x)
        sc = generic_score(net, x, sigma, kind, fns)
        res = -mult * sc
        if stochastic:
            res = res + (-(lang * sc))
        return res
    s, sd = fns.scaling(t_), fns.scaling_deriv(t_)
    mult = s * (fns.noise_deriv(t_) * fns.noise(t_))
    sc = generic_score(net, x / s, sigma, kind, fns)
    res = (sd / s) * x - mult * sc
    if stochastic:
        res = res + (-(lang * 1 / s * sc))
    return res


def generic_propagate(net: Net, x, nsteps: int, tag: str, kind: Optional[str] = None, integrator: str = "heun",
                      record_history: bool = False, noises=None, langevin_const: float = 1.0):
    """Scheduler.propagate(backward=True) (schedulers.py:48-89) with the Euler / Heun / Euler-Maruyama steps
    (integrators.py:29-69) for the VP, VE or EDM scheduler; `kind` names the preconditioner (default: the tag's own)."""
    kind = kind or tag
    fns = SchedFns(tag)
    t = generic_steps(tag, nsteps + 1).to(x)
    dt = torch.diff(t)
    hist = [x] if record_history else None
    for i in range(nsteps):
        ti, dti = t[i], dt[i]
        if integrator == "euler":
            x = x + dti * generic_rhs(net, x, ti, kind, fns)
        elif integrator == "heun":
            r1 = generic_rhs(net, x, ti, kind, fns)
            r2 = generic_rhs(net, x + dti * r1, ti + dti, kind, fns) if (ti + dti) > 0 else r1
            x = x + 0.5 * (r1 + r2) * dti
        elif integrator == "euler-maruyama":
            strength = torch.sqrt(2 * (langevin_const * (fns.scaling(ti) ** 2 * fns.noise_deriv(ti) * fns.noise(ti)) + 0 * ti))
            x = (x + generic_rhs(net, x, ti, kind, fns, stochastic=True, langevin_const=langevin_const) * dti +
                 (strength * noises[i] * torch.sqrt(torch.abs(dti))))
        else:
            raise ValueError(integrator)
        if record_history:
            hist.append(x)
    return torch.stack(hist, 0) if record_history else x


def generic_loss(net: Net, x, sigma, noise, kind: str, fns: Optional[SchedFns] = None, sigma_data=0.5):
    """KarrasModule.loss_fn (karrasmodule.py:569-650), Huber, for any preconditioner; weights of VPNoiseSampler /
    VENoiseSampler (1/sigma^2, noisesamplers.py:59-60, 84-85) or EDMNoiseSampler (SR3 fixtures)."""
    bs = bcast(sigma, x)
    D = generic_denoiser(net, x + bs * noise, sigma, kind, fns)
    w = 1 / (bs ** 2) if kind in ("vp", "ve") else edm_loss_weight(bs, sigma_data)
    l = torch.nn.functional.huber_loss(D, x, reduction="none", delta=1.0)
    return (w * l + torch.zeros_like(w)).mean()
