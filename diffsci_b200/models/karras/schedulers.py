"""Noise schedules and the integration driver (reference karras/schedulers.py:27-448).

``Scheduler`` keeps the reference's public surface (``propagate*``, ``rhs``, ``create_steps``,
``set_temporary_integrator``, the mutable knobs ``maximum_scale`` / ``langevin_const`` /
``langevin_interval``).  Two execution routes share it:

* ``propagate(x, score_fn, ...)`` -- the duck-typed seam for a foreign score function: a host loop
  over ``integrator.step`` whose arithmetic runs in ``dsk_lincomb`` launches;
* ``step_table(nsteps, integrator)`` -- the whole ``(t_i, dt_i, t_hat_i, langevin, noise, ...)``
  schedule precomputed on the host in the reference's fp32 operation order.  The fused CUDA-graph
  engine (``engine.SamplerEngine``) reads it from device memory, which removes every host sync the
  reference has inside the loop (integrators.py:45,49,96; schedulers.py:233,254).
"""
from __future__ import annotations

import functools
from typing import Callable, Optional

import torch
from torch import Tensor

from ... import ops
from ..._lib import TAB_COLS, TAB_T, TAB_DT, TAB_THAT, TAB_LANG, TAB_NOISE, TAB_SQDT, TAB_TNEXT, TAB_CHURN
from . import integrators, schedulingfunctions
from .integrators import _f32

ScoreFunction = Callable[[Tensor, Tensor], Tensor]


class Scheduler(torch.nn.Module):
    def __init__(self, scheduler_fns: schedulingfunctions.SchedulingFunctions, integrator: integrators.Integrator,
                 maximum_scale: float, stochastic_integrator: Optional[integrators.Integrator] = None):
        super().__init__()
        self.scheduler_fns = scheduler_fns
        self._integrator = integrator
        self.maximum_scale = maximum_scale
        if stochastic_integrator is None:
            stochastic_integrator = integrators.EulerMaruyamaIntegrator()
        else:
            assert stochastic_integrator.stochastic is True
        self.stochastic_integrator = stochastic_integrator
        self._temporary_integrator = None
        self.langevin_const = 1.0
        self.langevin_interval = None

    # ------------------------------------------------------------------ integrator selection
    @property
    def integrator(self) -> integrators.Integrator:
        return self._temporary_integrator if self._temporary_integrator is not None else self._integrator

    def set_temporary_integrator(self, integrator):
        if type(integrator) is str:
            integrator = integrators.name_to_integrator(integrator)
        self._temporary_integrator = integrator

    def unset_temporary_integrator(self):
        self._temporary_integrator = None

    # ------------------------------------------------------------------ schedule
    def create_steps(self, n: int) -> Tensor:
        raise NotImplementedError

    @property
    def fused_supported(self) -> bool:
        """True when csrc/sampler.cu implements this scheduler's RHS in closed form (EDM: s=1, sigma=t)."""
        f = self.scheduler_fns
        return bool(f.constant_scaling_fn and f.identity_noise_fn and not f.has_pf_score_multiplier)

    def step_table(self, nsteps: int, integrator: Optional[integrators.Integrator] = None) -> Tensor:
        """fp32 CPU tensor [nsteps+1, DSK_TAB_COLS] (include/diffsci_b200.h); last row is padding."""
        if nsteps < 2:
            raise ValueError("nsteps must be >= 2: create_steps(nsteps+1) divides by nsteps-1 "
                             "(reference schedulers.py:378 yields NaN for nsteps=1)")
        integ = integrator if integrator is not None else self.integrator
        t = self.create_steps(nsteps + 1).float().cpu()
        dt = torch.diff(t)
        tab = torch.zeros((nsteps + 1, TAB_COLS), dtype=torch.float32)
        tab[:nsteps, TAB_T] = t[:nsteps]
        tab[:nsteps, TAB_DT] = dt
        tab[:nsteps, TAB_THAT] = t[:nsteps]
        prog = integ.fused_program
        for i in range(nsteps):
            if prog == "euler-maruyama":
                tab[i, TAB_LANG] = self.langevin_factor(t[i])
                tab[i, TAB_NOISE] = self.noise_injection(t[i])
                tab[i, TAB_SQDT] = torch.sqrt(torch.abs(dt[i]))
            elif prog == "karras":
                that, coef = integ.churn(t[i], nsteps)
                tab[i, TAB_THAT] = that
                tab[i, TAB_CHURN] = coef
        tab[:nsteps - 1, TAB_TNEXT] = tab[1:nsteps, TAB_THAT]
        return tab


    # ------------------------------------------------------------------ general (table-driven) engine, SURVEY 8f-3
    GTAB_COLS = 12
    G_DT, G_P1, G_Q1, G_P2, G_Q2, G_XS1, G_CN1, G_XS2, G_CN2, G_HAS2, G_NZ = range(11)
    GENERAL_PROGRAMS = ("euler", "heun", "euler-maruyama")

    def rhs_coefficients(self, t, preconditioner, stochastic: bool = False):
        """(P, Q, c_in / s, c_noise) at time t such that  rhs(x, t) = P x + Q F(c_in/s x, c_noise)  for the backward
        integration -- Scheduler.rhs (schedulers.py:247-294) with score = (D - z) / sigma^2, D = c_out F + c_skip z,
        z = x / s, folded into two scalars.  Evaluated with the scheduler's and the preconditioner's own objects."""
        f = self.scheduler_fns
        t0 = _f32(t)
        sigma0, dsigma0 = f.noise_fn(t0), f.noise_fn_deriv(t0)
        if f.constant_scaling_fn:
            s, ds = torch.ones_like(t0), torch.zeros_like(t0)
            mult = f.pf_score_multiplier(t0) if f.has_pf_score_multiplier else sigma0 * dsigma0
        else:
            s, ds = f.scaling_fn(t0), f.scaling_fn_deriv(t0)
            mult = f.pf_score_multiplier(t0) if f.has_pf_score_multiplier else s * (dsigma0 * sigma0)
        bm = -mult.double()
        if stochastic:
            bm = bm - self.langevin_factor(t0).double() / s.double()
        sg = sigma0.reshape(1).float()
        c_in, c_out, c_skip, c_noise = (v.reshape(-1)[0].double() for v in (
            preconditioner.input_scaling(sg), preconditioner.output_scaling(sg), preconditioner.skip_scaling(sg),
            preconditioner.noise_conditioner(sg)))
        s, ds, sig2 = s.double(), ds.double(), sigma0.double() ** 2
        return ds / s + bm * (c_skip - 1.0) / (sig2 * s), bm * c_out / sig2, c_in / s, c_noise

    def general_step_table(self, nsteps: int, preconditioner, integrator: Optional[integrators.Integrator] = None) -> Tensor:
        """fp32 CPU tensor [nsteps + 1, GTAB_COLS] for dsk_sampler_stage_general (include/diffsci_b200.h: dsk_gtab_col); the
        last row is zero padding.  Rows follow Scheduler.propagate(backward=True) (schedulers.py:48-89): dt = diff(steps);
        a Heun step whose end time is 0 has no second evaluation (integrators.py:45-53)."""
        if nsteps < 2:
            raise ValueError("nsteps must be >= 2")
        integ = integrator if integrator is not None else self.integrator
        prog = integ.fused_program
        if prog not in self.GENERAL_PROGRAMS:
            raise NotImplementedError(f"general step table: integrator program {prog!r}")
        t = self.create_steps(nsteps + 1).float().cpu()
        dt = torch.diff(t)
        stoch = bool(integ.stochastic)
        tab = torch.zeros((nsteps + 1, self.GTAB_COLS), dtype=torch.float64)
        for i in range(nsteps):
            P1, Q1, xs1, cn1 = self.rhs_coefficients(t[i], preconditioner, stoch)
            tab[i, self.G_DT] = dt[i].double()
            tab[i, self.G_P1], tab[i, self.G_Q1], tab[i, self.G_XS1], tab[i, self.G_CN1] = P1, Q1, xs1, cn1
            t2 = t[i] + dt[i]                      # fp32 sum, as integrators.py:45 evaluates it
            if prog == "heun":
                if float(t2) > 0:
                    P2, Q2, xs2, cn2 = self.rhs_coefficients(t2, preconditioner, False)
                    tab[i, self.G_P2], tab[i, self.G_Q2], tab[i, self.G_XS2], tab[i, self.G_CN2] = P2, Q2, xs2, cn2
                    tab[i, self.G_HAS2] = 1.0
                elif float(t2) < 0:
                    raise ValueError("t+dt < 0 in Heun integrator")
            if prog == "euler-maruyama":
                tab[i, self.G_NZ] = (self.noise_injection(t[i]).double() * torch.sqrt(torch.abs(dt[i])).double())
        return tab.float()

    # ------------------------------------------------------------------ generic seam
    def propagate(self, x: Tensor, score_fn: ScoreFunction, nsteps: int = 100, record_history: bool = False,
                  backward: bool = True, stochastic: bool = False) -> Tensor:
        integrator = self.integrator if not stochastic else self.stochastic_integrator
        return self._run(x, score_fn, nsteps, 0, nsteps, record_history, backward, integrator, full=True)

    def propagate_partial(self, x, score_fn, nsteps: int = 100, initial_step: int = 0, final_step: int = 100,
                          record_history: bool = False, backward: bool = True, stochastic: bool = False):
        if not backward:
            raise NotImplementedError
        integrator = self.integrator if not stochastic else self.stochastic_integrator
        return self._run(x, score_fn, nsteps, initial_step, final_step, record_history, True, integrator, full=False)

    def _run(self, x, score_fn, nsteps, first, last, record_history, backward, integrator, full):
        t = self.create_steps(nsteps + 1).float().cpu()
        skip = 0
        if not backward:
            t, skip = t.flip(0), 1
        dt = torch.diff(t)
        x = x.float().contiguous()
        rhs = functools.partial(self.rhs, score_fn=score_fn, backward=backward, stochastic=integrator.stochastic)
        if hasattr(integrator, "begin_run"):
            integrator.begin_run()
        step = integrator.step
        if integrator.need_fns:
            step = functools.partial(step, scheduler_fns=self.scheduler_fns, nsteps=nsteps)
        nrec = (nsteps + 1) if full else (last - first + 1)
        if record_history:
            history = torch.zeros((nrec,) + tuple(x.shape), dtype=x.dtype, device=x.device)
            history[0 + (skip if full else 0)] = x
        lo, hi = (skip, nsteps) if full else (first, last)
        for k, i in enumerate(range(lo, hi)):
            x = step(x, t[i], dt[i], rhs, noise_strength=self.noise_injection)
            if record_history:
                history[k + 1 + (skip if full else 0)] = x
        return history if record_history else x

    def inpaint(self, x: Tensor, y: Tensor, mask: Tensor, score_fn: ScoreFunction, nsteps: int = 100,
                record_history: bool = False) -> Tensor:
        """Reverse integration with the known region re-imposed after every step from the forward history `y`
        [nsteps+1, B, *shape] (schedulers.py:91-122): x <- step(x); x <- x (1 - mask) + y[-i-2] mask."""
        t = self.create_steps(nsteps + 1).float().cpu()
        dt = torch.diff(t)
        x = x.float().contiguous()
        if record_history:
            history = torch.zeros((nsteps + 1,) + tuple(x.shape), dtype=x.dtype, device=x.device)
            history[0] = x
        rhs = functools.partial(self.rhs, score_fn=score_fn, backward=True)
        if hasattr(self.integrator, "begin_run"):
            self.integrator.begin_run()
        x = ops.mask_blend(x, y[-1], mask)
        for i in range(nsteps):
            x = self.integrator.step(x, t[i], dt[i], rhs, self.noise_injection)
            x = ops.mask_blend(x, y[-i - 2], mask)
            if record_history:
                history[i + 1] = x
        return history if record_history else x

    def repaint(self, x: Tensor, y: Tensor, mask: Tensor, score_fn: ScoreFunction, nsteps: int = 100, rsteps: int = 10,
                nresamples: int = 10, record_history: bool = False, _partial=None) -> Tensor:
        """RePaint resampling (schedulers.py:124-175): every `rsteps` steps, `nresamples` times: re-impose the known
        region, jump back in noise level (renoise) and integrate the same stretch again."""
        if not (nsteps % rsteps) == 0:
            raise ValueError("rsteps should divide nsteps")
        t = self.create_steps(nsteps + 1).float().cpu()
        x = x.float().contiguous()
        # _partial (KarrasModule): integrates a stretch of the schedule on the captured-graph engine instead of the step seam
        partial = self.propagate_partial if _partial is None else _partial
        if _partial is not None and hasattr(self.integrator, "begin_run"):
            self.integrator.begin_run()              # renoise draws from the integrator's stream
        if record_history:
            history = torch.zeros((int(nresamples * (nsteps / rsteps - 1)) + 2,) + tuple(x.shape), dtype=x.dtype,
                                  device=x.device)
            history[0] = x
        x = ops.mask_blend(x, y[-1], mask)
        step, fstep = 0, rsteps
        x = partial(x, score_fn, nsteps, step, fstep)
        step, fstep = fstep, fstep + rsteps
        level = 0
        while fstep <= nsteps:
            x = partial(x, score_fn, nsteps, step, fstep)
            for i in range(nresamples):
                x = ops.mask_blend(x, y[-fstep - 1], mask)
                if record_history:
                    history[level + i + 1] = x
                x = self.renoise(x, t[fstep], t[step])
                x = partial(x, score_fn, nsteps, step, fstep)
            step, fstep = fstep, fstep + rsteps
            level = level + nresamples
        if not step == nsteps:
            raise ValueError("Wrong counting")
        if record_history:
            history[level + 1] = x
            return history
        return x

    def propagate_backward(self, x, score_fn, nsteps: int = 100, record_history: bool = False,
                           stochastic: bool = False):
        return self.propagate(x, score_fn, nsteps, record_history, backward=True, stochastic=stochastic)

    def propagate_forward(self, x, score_fn, nsteps: int = 100, record_history: bool = False,
                          stochastic: bool = False):
        return self.propagate(x, score_fn, nsteps, record_history, backward=False, stochastic=stochastic)

    def langevin_factor(self, t: Tensor, type: str = "const") -> Tensor:
        """langevin_const * s^2 sigma' sigma, gated by langevin_interval (schedulers.py:219-240)."""
        f = self.scheduler_fns
        standard = f.scaling_fn(t) ** 2 * f.noise_fn_deriv(t) * f.noise_fn(t)
        if type != "const":
            raise NotImplementedError
        if self.langevin_interval is not None:
            t0 = t.reshape(-1)[0] if t.ndim > 0 else t
            if not (t0 > self.langevin_interval[0] and t0 < self.langevin_interval[1]):
                return 0 * t
        return self.langevin_const * standard + 0 * t

    def noise_injection(self, t: Tensor) -> Tensor:
        return torch.sqrt(2 * self.langevin_factor(t))

    def rhs(self, x: Tensor, ti: Tensor, score_fn: ScoreFunction, backward: bool = True,
            stochastic: bool = False) -> Tensor:
        """Probability-flow / reverse-SDE drift for a foreign score function (schedulers.py:247-294)."""
        f = self.scheduler_fns
        t0 = _f32(ti)
        sigma0, dsigma0 = f.noise_fn(t0), f.noise_fn_deriv(t0)
        sigma = (sigma0 * torch.ones(x.shape[0])).to(x)
        sgn = 1.0 if backward else -1.0
        if f.constant_scaling_fn:
            mult = f.pf_score_multiplier(t0) if f.has_pf_score_multiplier else sigma0 * dsigma0
            score = score_fn(x, sigma)
            if stochastic:
                return ops.lincomb(None, 0.0, score, -float(mult), score, -sgn * float(self.langevin_factor(t0)))
            return ops.lincomb(None, 0.0, score, -float(mult))
        s, ds = f.scaling_fn(t0), f.scaling_fn_deriv(t0)
        mult = f.pf_score_multiplier(t0) if f.has_pf_score_multiplier else s * (dsigma0 * sigma0)
        score = score_fn(ops.lincomb(x, float(1 / s)), sigma)
        if stochastic:
            return ops.lincomb(x, float(ds / s), score, -float(mult), score,
                               -sgn * float(self.langevin_factor(t0) * 1 / s))
        return ops.lincomb(x, float(ds / s), score, -float(mult))

    def renoise(self, x: Tensor, t, t_noise) -> Tensor:
        """x -> (s_n/s) x + s_n sqrt(sigma_n^2 - sigma^2) xi  (schedulers.py:177-187)."""
        f = self.scheduler_fns
        t, tn = _f32(t), _f32(t_noise)
        std = f.scaling_fn(tn) * torch.sqrt(f.noise_fn(tn) ** 2 - f.noise_fn(t) ** 2)
        z = self.integrator._randn_like(x)
        return ops.lincomb(x, float(f.scaling_fn(tn) / f.scaling_fn(t)), None, 0.0, None, 0.0, z, float(std))

    def apply_noise(self, x: Tensor, nsteps: int = 100, step: int = 0) -> Tensor:
        if step > nsteps:
            raise ValueError(f"Step larger than num of steps:{step}>{nsteps}")
        f = self.scheduler_fns
        ts = self.create_steps(nsteps + 1).float().cpu()[step]
        scale, sigma = f.scaling_fn(ts), f.noise_fn(ts)
        z = self.integrator._randn_like(x)
        return ops.lincomb(x, float(scale), None, 0.0, None, 0.0, z, float(scale * sigma))


class EDMScheduler(Scheduler):
    def __init__(self, sigma_min: float = 0.002, sigma_max: float = 80.0, expoent_steps: float = 7.0,
                 scheduler_fns="EDM"):
        if type(scheduler_fns) is str:
            scheduler_fns = schedulingfunctions.name_to_scheduling_functions(scheduler_fns)
        super().__init__(scheduler_fns, integrators.HeunIntegrator(), sigma_max)
        self.register_buffer("sigma_min", torch.tensor(sigma_min))
        self.register_buffer("sigma_max", torch.tensor(sigma_max))
        self.register_buffer("expoent_steps", torch.tensor(expoent_steps))

    def create_steps(self, n: int) -> Tensor:
        """t_i = (smax^(1/rho) + i/(n-2) (smin^(1/rho) - smax^(1/rho)))^rho, i < n-1; t_{n-1} = 0, in fp32
        with the reference's operation order (schedulers.py:377-385) so the table is bit-identical."""
        rho = self.expoent_steps.detach().float().cpu()
        smax, smin = self.sigma_max.detach().float().cpu(), self.sigma_min.detach().float().cpu()
        s = torch.arange(n - 1).to(rho) / (n - 2)
        start, end = smax ** (1 / rho), smin ** (1 / rho)
        steps = (start + s * (end - start)) ** rho
        if not self.scheduler_fns.identity_noise_fn:
            steps = self.scheduler_fns.inverse_noise_fn(steps)
        return torch.cat([steps, torch.zeros([1]).to(steps)])

    def step_from_time(self, t: Tensor, n: int):
        e = 1 / self.expoent_steps
        step = (n - 1) * (t ** e - self.sigma_max ** e) / (self.sigma_min ** e - self.sigma_max ** e)
        return torch.round(step).int()


class VPScheduler(Scheduler):
    def __init__(self, epsilon_min: float = 0.001, scheduler_fns="VP", *args, **kwargs):
        if type(scheduler_fns) is str:
            scheduler_fns = schedulingfunctions.name_to_scheduling_functions(scheduler_fns, *args, **kwargs)
        one = torch.ones([1])
        sigma_max = (scheduler_fns.noise_fn(one) * scheduler_fns.scaling_fn(one)).item()
        super().__init__(scheduler_fns, integrators.HeunIntegrator(), sigma_max)
        self.register_buffer("epsilon_min", torch.tensor(epsilon_min))

    def create_steps(self, n: int) -> Tensor:
        eps = self.epsilon_min.detach().float().cpu()
        s = torch.arange(n).to(eps) / (n - 1)
        return 1 + s * (eps - 1)

    def step_from_time(self, t: Tensor, n: int):
        """Index of the schedule step at time t: the inverse of create_steps (schedulers.py:417-419)."""
        step = (n - 1) * (t - 1) / (self.epsilon_min - 1)
        return torch.round(step).int()


class VEScheduler(Scheduler):
    def __init__(self, sigma_min: float = 0.02, sigma_max: float = 100, scheduler_fns="VE", *args, **kwargs):
        if type(scheduler_fns) is str:
            scheduler_fns = schedulingfunctions.name_to_scheduling_functions(scheduler_fns, *args, **kwargs)
        super().__init__(scheduler_fns, integrators.HeunIntegrator(), sigma_max)
        self.register_buffer("sigma_min", torch.tensor(float(sigma_min)))
        self.register_buffer("sigma_max", torch.tensor(float(sigma_max)))

    def create_steps(self, n: int) -> Tensor:
        smin, smax = self.sigma_min.detach().float().cpu(), self.sigma_max.detach().float().cpu()
        s = torch.arange(n).to(smin) / (n - 1)
        return smax ** 2 * (smin ** 2 / smax ** 2) ** s

    def step_from_time(self, t: Tensor, n: int):
        """Index of the schedule step at time t: the inverse of create_steps (schedulers.py:446-448)."""
        step = (n - 1) * (torch.log(t) - torch.log(self.sigma_max ** 2)) / (torch.log(self.sigma_min ** 2) -
                                                                          torch.log(self.sigma_max ** 2))
        return torch.round(step).int()
