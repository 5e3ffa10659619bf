"""KarrasModule / KarrasModuleConfig -- drop-in for diffsci.models.karras.karrasmodule
(reference karras/karrasmodule.py:30-1253), restricted to the Karras/EDM hot path of SURVEY.md section 8:
``get_denoiser``, ``get_score``, ``loss_fn``, ``training_step``, ``sample``, ``propagate_white_noise``,
``propagate_toward_sample`` and -- section 8(f) -- ``inpaint`` / ``repaint`` / ``propagate_partial_toward_sample`` /
``propagate_toward_noise`` / ``interpolate_images``, conditional models (``conditional=True``: ``y``, classifier-free
``guidance``) and the VP / VE configurations keep their signatures and semantics.

Where the reference launches ~30 elementwise ATen kernels per network evaluation plus host syncs,
this module drives the fused CUDA stages (csrc/sampler.cu, csrc/train.cu) and, for native networks,
replays one captured CUDA graph per integrator step (engine.SamplerEngine).
Section 8(f)-4: the latent-diffusion wrapper (``autoencoder=``: a frozen user torch module whose ``encode`` / ``decode``
bracket the loss and the sampler, karrasmodule.py:1192-1234) and, in ``karrasmodule_new.py``, the ensemble losses.
Autoregressive, dynamic-loss-weight and multi-space-loss recipes are out of
scope and raise NotImplementedError instead of silently doing something else.
"""
from __future__ import annotations

import os
from typing import Any, Dict, Optional, Union

import torch
from torch import Tensor

from ... import ops
from ..._lib import lib, check, ptr, stream, dt_code, require_cuda
from ...torchutils import broadcast_from_below, dict_to, dict_unsqueeze
from ...utils import get_minibatch_sizes
from . import engine as _engine
from . import integrators, noisesamplers, preconditioners, schedulers

try:  # Lightning is optional: the hot path never needs it, Trainer users get a real LightningModule
    import lightning as _lightning
    _Base = _lightning.LightningModule
except Exception:  # pragma: no cover - lightning is absent in the build image
    class _Base(torch.nn.Module):
        def log(self, *a, **k):
            return None

        @property
        def device(self):
            for t in list(self.parameters()) + list(self.buffers()):
                return t.device
            return torch.device("cpu")


class KarrasModuleConfig:
    def __init__(self, preconditioner: preconditioners.KarrasPreconditioner, noisesampler: noisesamplers.NoiseSampler,
                 noisescheduler: schedulers.Scheduler, loss_metric: Union[str, Dict[str, Any]] = "huber",
                 tag: str = "custom", has_edm_batch_norm: bool = False, dynamic_loss_weight: Optional[int] = None,
                 extra_args: Optional[dict] = None, autoregressive_loss_steps: int = 1,
                 autoregressive_loss_diffusion_steps: int = 100, autoregressive_loss_guidance: float = 1.0,
                 autoregressive_loss_weights=None, autoregressive_loss_maximum_batch_size=None,
                 autoregressive_loss_integrator=None, spatial_shape: tuple = None, focus_radius: float = None):
        self.preconditioner, self.noisesampler, self.noisescheduler = preconditioner, noisesampler, noisescheduler
        self.loss_metric, self.tag = loss_metric, tag
        self.has_edm_batch_norm = has_edm_batch_norm
        self.dynamic_loss_weight = dynamic_loss_weight
        self.autoregressive_loss_steps = autoregressive_loss_steps
        self.autoregressive_loss_diffusion_steps = autoregressive_loss_diffusion_steps
        self.autoregressive_loss_guidance = autoregressive_loss_guidance
        self.autoregressive_loss_weights = autoregressive_loss_weights
        self.autoregressive_loss_maximum_batch_size = autoregressive_loss_maximum_batch_size
        self.autoregressive_loss_integrator = autoregressive_loss_integrator
        self.spatial_shape, self.focus_radius = spatial_shape, focus_radius
        self.extra_args = dict() if extra_args is None else extra_args

    @classmethod
    def _build(cls, tag, pre, smp, sch, loss_metric, own_args, common):
        extra = dict(own_args, loss_metric=loss_metric, **common)
        return cls(preconditioner=pre, noisesampler=smp, noisescheduler=sch, loss_metric=loss_metric, tag=tag,
                   extra_args=extra, **common)

    @classmethod
    def from_edm(cls, sigma_data: float = 0.5, prior_mean: float = -1.2, prior_std: float = 1.2,
                 has_edm_batch_norm: bool = False, dynamic_loss_weight: Optional[int] = None,
                 loss_metric: Union[str, Dict[str, Any]] = "huber", autoregressive_loss_steps: int = 1,
                 autoregressive_loss_diffusion_steps: int = 100, autoregressive_loss_guidance: float = 1.0,
                 autoregressive_loss_weights=None, autoregressive_loss_maximum_batch_size=None,
                 autoregressive_loss_integrator=None, spatial_shape: tuple = None, focus_radius: float = None):
        common = dict(autoregressive_loss_steps=autoregressive_loss_steps,
                      autoregressive_loss_diffusion_steps=autoregressive_loss_diffusion_steps,
                      autoregressive_loss_guidance=autoregressive_loss_guidance,
                      autoregressive_loss_weights=autoregressive_loss_weights,
                      autoregressive_loss_maximum_batch_size=autoregressive_loss_maximum_batch_size,
                      autoregressive_loss_integrator=autoregressive_loss_integrator,
                      spatial_shape=spatial_shape, focus_radius=focus_radius)
        cfg = cls._build("edm", preconditioners.EDMPreconditioner(sigma_data=sigma_data),
                         noisesamplers.EDMNoiseSampler(sigma_data=sigma_data, prior_mean=prior_mean, prior_std=prior_std),
                         schedulers.EDMScheduler(), loss_metric,
                         dict(sigma_data=sigma_data, prior_mean=prior_mean, prior_std=prior_std,
                              has_edm_batch_norm=has_edm_batch_norm, dynamic_loss_weight=dynamic_loss_weight), common)
        cfg.has_edm_batch_norm, cfg.dynamic_loss_weight = has_edm_batch_norm, dynamic_loss_weight
        return cfg

    @classmethod
    def from_vp(cls, beta_data: float = 19.9, beta_min: float = 0.1, epsilon_min: float = 1e-3,
                epsilon_sampler: float = 1e-5, M: int = 1000, loss_metric="huber", **common):
        sch = schedulers.VPScheduler(epsilon_min=epsilon_min, beta_data=beta_data, beta_min=beta_min)
        return cls._build("vp", preconditioners.VPPreconditioner(scheduler=sch, M=M),
                          noisesamplers.VPNoiseSampler(noise_scheduler=sch, epsilon=epsilon_sampler), sch, loss_metric,
                          dict(beta_data=beta_data, beta_min=beta_min, epsilon_min=epsilon_min,
                               epsilon_sampler=epsilon_sampler, M=M), common)

    @classmethod
    def from_ve(cls, sigma_min: float = 0.02, sigma_max: float = 100, loss_metric="huber", **common):
        return cls._build("ve", preconditioners.VEPreconditioner(),
                          noisesamplers.VENoiseSampler(sigma_min=sigma_min, sigma_max=sigma_max),
                          schedulers.VEScheduler(sigma_min=sigma_min, sigma_max=sigma_max), loss_metric,
                          dict(sigma_min=sigma_min, sigma_max=sigma_max), common)

    @classmethod
    def conditionalSR3(cls, sigma_min: float = 0.02, sigma_max: float = 100, loss_metric="huber", **common):
        """karrasmodule.py:291-341, as the reference builds it: EDMScheduler(sigma_min, sigma_max) + SR3Preconditioner +
        ``EDMNoiseSampler(sigma_min=..., sigma_max=...)``.  The reference's EDMNoiseSampler takes (sigma_data, prior_mean,
        prior_std) (noisesamplers.py:20-28), so this factory raises TypeError there -- and here, for the same call: the drop-in
        keeps the reference's error behaviour rather than inventing a sampler the reference never ran."""
        sch = schedulers.EDMScheduler(sigma_min=sigma_min, sigma_max=sigma_max)
        return cls._build("conditionalSR3", preconditioners.SR3Preconditioner(),
                          noisesamplers.EDMNoiseSampler(sigma_min=sigma_min, sigma_max=sigma_max), sch, loss_metric,
                          dict(sigma_min=sigma_min, sigma_max=sigma_max), common)

    def export_description(self) -> dict[str, Any]:
        return dict(tag=self.tag, extra_args=self.extra_args)

    @classmethod
    def load_from_description_with_tag(cls, description: dict[str, Any]):
        tag, extra = description["tag"], description["extra_args"]
        if tag == "custom":
            raise ValueError("Cannot load from a custom tag")
        ctor = {"edm": cls.from_edm, "vp": cls.from_vp, "ve": cls.from_ve, "conditionalSR3": cls.conditionalSR3}.get(tag)
        if ctor is None:
            raise ValueError(f"Unknown tag: {tag}")
        return ctor(**extra)

    @property
    def has_dynamic_loss_weight(self) -> bool:
        return self.dynamic_loss_weight is not None

    def update_loss_metric(self, loss_config):
        self.loss_metric = loss_config
        if "loss_metric" in self.extra_args:
            self.extra_args["loss_metric"] = loss_config


class _EDMLossFn(torch.autograd.Function):
    """loss = mean(lambda(sigma) * l(c_out F + c_skip (x + sigma n), x) * (1 - mask)); backward = dL/dF from
    the same fused launch (csrc/train.cu: edm_loss_kernel)."""

    @staticmethod
    def forward(ctx, F, x, noise, sigma, mask, sigma_data, kind, coeffs=None, delta=1.0):
        """coeffs: None (EDM preconditioner + EDM weighting evaluated in the kernel) or fp32 [B] vectors (c_out, c_skip,
        lambda) from any preconditioner / noise sampler (dsk_precond_loss_fwd_bwd)."""
        B = x.shape[0]
        Cc = x.shape[1] if x.ndim > 1 else 1
        S = x.numel() // (B * Cc)
        loss = torch.zeros((), dtype=torch.float32, device=x.device)
        dF = torch.empty_like(x, dtype=torch.float32)
        if coeffs is None:
            check(lib.dsk_edm_loss_fwd_bwd(ptr(F), ptr(x), ptr(noise), ptr(sigma), ptr(mask), ptr(loss), ptr(dF), B, Cc, S,
                                           float(sigma_data), int(kind), stream()))
        else:
            c_out, c_skip, lam = coeffs
            check(lib.dsk_precond_loss_fwd_bwd_huber(ptr(F), ptr(x), ptr(noise), ptr(sigma), ptr(c_out), ptr(c_skip), ptr(lam),
                                                     ptr(mask), ptr(loss), ptr(dF), B, Cc, S, int(kind), float(delta), stream()))
        ctx.save_for_backward(dF)
        return loss

    @staticmethod
    def backward(ctx, g):
        (dF,) = ctx.saved_tensors
        return dF * g, None, None, None, None, None, None, None, None


class _RowLossFn(torch.autograd.Function):
    """sum_b m[b] * loss_b[b] with loss_b from dsk_precond_loss_rows (per-sample partial sums of the fused loss): gradients
    flow to F (m[b] * dF[b], dF from the same launch) and to the per-sample factor m (loss_b)."""

    @staticmethod
    def forward(ctx, F, m, x, noise, sigma, mask, coeffs, kind, delta=1.0):
        B = x.shape[0]
        Cc = x.shape[1] if x.ndim > 1 else 1
        S = x.numel() // (B * Cc)
        loss_b = torch.zeros((B,), dtype=torch.float32, device=x.device)
        dF = torch.empty_like(x, dtype=torch.float32)
        c_out, c_skip, lam = coeffs
        check(lib.dsk_precond_loss_rows_huber(ptr(F), ptr(x), ptr(noise), ptr(sigma), ptr(c_out), ptr(c_skip), ptr(lam), ptr(mask),
                                              ptr(loss_b), ptr(dF), B, Cc, S, int(kind), float(delta), stream()))
        ctx.save_for_backward(dF, loss_b, m)
        return (m * loss_b).sum()

    @staticmethod
    def backward(ctx, g):
        dF, loss_b, m = ctx.saved_tensors
        return dF * (g * m).view(-1, *([1] * (dF.ndim - 1))), g * loss_b, None, None, None, None, None, None, None


class DynamicLossWeight(torch.nn.Module):
    """Learned log-uncertainty u(c_noise) of the loss weighting (karrasmodule.py:1256-1278): a fixed random cosine feature
    map of the noise conditioner followed by one linear layer; the loss becomes mean(lambda * exp(-u) * l + u).  A [B]-sized
    torch module, like the preconditioner scalars."""

    def __init__(self, nhidden: int, scale: float = 1.0):
        super().__init__()
        self.nhidden = nhidden
        self.register_buffer("fourier_weights", torch.randn(nhidden) * scale)
        self.register_buffer("fourier_bias", torch.rand(nhidden) * scale)
        self.linear = torch.nn.Linear(nhidden, 1)

    def forward(self, x):
        h = torch.cos(x.unsqueeze(1) * self.fourier_weights + self.fourier_bias)
        return self.linear(h).squeeze(1)


def _rowwise_axpy(a_vec: Tensor, z: Tensor, b_vec: Tensor, x: Tensor) -> Tensor:
    """out[b, ...] = a[b]*z[b, ...] + b_vec[b]*x[b, ...] in one launch (dsk_precond_denoise on flat rows)."""
    B = x.shape[0]
    n = x.numel() // B
    out = torch.empty_like(x)
    check(lib.dsk_precond_denoise(ptr(z), ptr(x), ptr(a_vec), ptr(b_vec), None, ptr(out), None, B, 1, n, 0, stream()))
    return out


class KarrasModule(_Base):
    huber_delta = 1.0        # torch.nn.HuberLoss(delta) of loss_metric = {"huber": {"delta": d}} (set_loss_metric)

    def __init__(self, model: torch.nn.Module, config: KarrasModuleConfig, conditional: bool = False,
                 masked: bool = False, autoencoder: Optional[torch.nn.Module] = None,
                 autoencoder_conditional: bool = False, encode_y: bool = False, decode_original_y: bool = False):
        super().__init__()
        if (encode_y or decode_original_y) and not (autoencoder is not None and autoencoder_conditional):
            raise ValueError("encode_y / decode_original_y need a conditional autoencoder (its encode(x, y) returns (z, y'))")
        if autoencoder_conditional and autoencoder is None:
            raise ValueError("autoencoder_conditional=True needs an autoencoder")
        if config.has_edm_batch_norm:
            raise NotImplementedError("has_edm_batch_norm=True: karras/edmbatchnorm.py is empty in the reference "
                                      "(karrasmodule.py:1241 crashes there too)")
        self.model, self.config = model, config
        self.conditional, self.masked = conditional, masked
        self.autoencoder = autoencoder
        if self.autoencoder is not None:
            self.freeze_autoencoder()
        self.autoencoder_conditional = bool(autoencoder_conditional)
        # encode_y: the conditional autoencoder also re-encodes the condition, encode(x, y) -> (z, y') (karrasmodule.py:1201-1212);
        # decode_original_y: sampling decodes with the ORIGINAL y instead of the encoded one (:843-860)
        self.encode_y, self.decode_original_y = bool(encode_y), bool(decode_original_y)
        self.norm = 1.0
        self.set_optimizer_and_scheduler()
        self.set_loss_metric()
        self.edm_batch_norm = None
        self.start_dynamic_loss_weight()
        self._engines: dict[Any, _engine.SamplerEngine] = {}
        self.use_cuda_graphs = True
        self.last_nfe = 0

    def start_dynamic_loss_weight(self):
        """karrasmodule.py:1243-1249 (created after the default optimizer, as in the reference: pass your own optimizer
        over module.parameters() to train it)."""
        self.dynamic_loss_weight = (DynamicLossWeight(self.config.dynamic_loss_weight)
                                    if self.config.has_dynamic_loss_weight else None)

    def freeze_autoencoder(self):
        """karrasmodule.py:457-460: the autoencoder is a fixed pre-/post-processor, never trained here."""
        for param in self.autoencoder.parameters():
            param.requires_grad = False

    def export_description(self) -> dict[str, Any]:
        return dict(config_description=self.config.export_description(), conditional=self.conditional, masked=self.masked,
                    autoencoder=self.autoencoder is not None, autoencoder_conditional=self.autoencoder_conditional,
                    encode_y=self.encode_y)

    # ------------------------------------------------------------------ optimiser / loss config
    def set_optimizer_and_scheduler(self, optimizer=None, scheduler=None, scheduler_interval="step"):
        """Defaults as the reference (karrasmodule.py:476-508): AdamW(1e-3, (0.9, 0.999), wd 1e-4), identity LR."""
        params = [p for p in self.parameters()]
        if optimizer is not None:
            self.optimizer = optimizer
        elif params:
            self.optimizer = torch.optim.AdamW(params, lr=1e-3, betas=(0.9, 0.999), weight_decay=1e-4)
        else:
            self.optimizer = None
        if scheduler is not None:
            self.lr_scheduler = scheduler
        elif self.optimizer is not None:
            self.lr_scheduler = torch.optim.lr_scheduler.LambdaLR(self.optimizer, lr_lambda=lambda step: 1.0 + 0 * step)
        else:
            self.lr_scheduler = None
        self.lr_scheduler_interval = scheduler_interval

    def configure_optimizers(self):
        if self.lr_scheduler is not None:
            return [self.optimizer], [{"scheduler": self.lr_scheduler, "interval": self.lr_scheduler_interval}]
        return self.optimizer

    def set_loss_metric(self):
        lm = self.config.loss_metric
        if isinstance(lm, dict) and len(lm) == 1 and "losses" not in lm:
            name, params = next(iter(lm.items()))
            self.huber_delta = float((params or {}).get("delta", 1.0)) if name == "huber" else 1.0
            lm = name
        if lm not in ("huber", "mse"):
            raise NotImplementedError(f"diffsci_b200.KarrasModule: loss_metric={self.config.loss_metric!r} is out of "
                                      "scope (SURVEY.md section 2 #10); 'huber' (default) and 'mse' are fused")
        self.loss_kind = 0 if lm == "huber" else 1
        self.loss_metric = lm
        if self.loss_kind != 0 or not (isinstance(self.config.loss_metric, dict) and "huber" in self.config.loss_metric):
            self.huber_delta = 1.0

    # ------------------------------------------------------------------ denoiser
    @property
    def latent_model(self) -> bool:
        return self.autoencoder is not None

    def encode(self, x, y=None, record_history=False):
        """Data space -> the space the diffusion runs in (karrasmodule.py:1192-1214).  The autoencoder is the user's torch
        module (any ``encode(x[, y])`` / ``decode(z[, y])`` pair); it runs on torch, outside the fused path."""
        if record_history:
            return torch.stack([self.encode(xx, y, record_history=False) for xx in x], dim=0)
        if self.latent_model:
            if self.autoencoder_conditional and self.encode_y:
                x, y = self.autoencoder.encode(x, y)
            else:
                x = self.autoencoder.encode(x, y) if self.autoencoder_conditional else self.autoencoder.encode(x)
        x = x / self.norm if self.norm != 1.0 else x
        return (x, y) if self.encode_y else x

    def decode(self, x, y=None, record_history=False):
        """karrasmodule.py:1216-1234."""
        if record_history:
            return torch.stack([self.decode(xx, y, record_history=False) for xx in x], dim=0)
        if self.norm != 1.0:
            x = x * self.norm
        if self.latent_model:
            x = self.autoencoder.decode(x, y) if self.autoencoder_conditional else self.autoencoder.decode(x)
        return x

    def _sigma_data(self) -> float:
        sd = getattr(self.config.preconditioner, "sigma_data", None)
        return float(sd) if sd is not None else 0.5

    def _network(self, x: Tensor, c_in: Tensor, cond_noise: Tensor, y=None, guidance: float = 1.0):
        """F = model(c_in * x, c_noise[, y]) with the reference's conditional dispatch (karrasmodule.py:703-716):
        conditional call iff ``self.conditional and guidance != 0``; classifier-free mix (1-g) F_u + g F_c iff additionally
        guidance != 1 -- evaluated as ONE 2B-sample network call on native networks.  Returns (F, act dtype or None)."""
        B = x.shape[0]
        Cc = x.shape[1] if x.ndim > 1 else 1
        S = x.numel() // (B * Cc)
        cond = bool(self.conditional and guidance != 0.0)
        cfg = cond and guidance != 1.0
        # gradients wanted (training, eval-mode fine-tuning, gradient diagnostics): the network's own forward, which takes the
        # autograd seam; otherwise the allocation-free inference plan
        wants_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.model.parameters())
        native = hasattr(self.model, "plan") and getattr(self.model, "engine_native", True) and not wants_grad
        if native and cond and not hasattr(self.model, "conditioning_vector"):
            raise NotImplementedError(f"diffsci_b200: {type(self.model).__name__} has no conditional path")
        if native:
            spatial = tuple(x.shape[2:]) if x.ndim > 2 else tuple(x.shape[1:])
            ychan, ye = None, None
            if cond:
                ychan, rest = self.model.split_condition(y, x)
                if cfg and ychan is not None:      # the reference's unconditional call model(x, t) fails here too (:718)
                    raise TypeError("classifier-free guidance needs an unconditional evaluation, which a channel-"
                                    "conditioned network (PUNetGCond) does not have")
                ye = self.model.conditioning_vector(rest, B)
            Bn = 2 * B if cfg else B
            plan = self.model.plan(Bn, spatial, x.device)
            ld = plan.xin.shape[-1]
            ncond = 0 if ychan is None else int(ychan.shape[1])
            if ld != Cc + ncond + int(getattr(self.model, "ones_channel", 0)):
                raise ValueError(f"network expects {ld} input channels, got {Cc} state + {ncond} conditioning channels")
            check(lib.dsk_precond_scale_cond(ptr(x), ptr(c_in), ptr(plan.xin), B, Cc, S, dt_code(plan.act_dtype), ld,
                                             1 if cfg else 0, stream()))
            if ychan is not None:
                plan.xin.view(B, S, ld)[:, :, Cc:Cc + ncond].copy_(ychan.reshape(B, ncond, S).transpose(1, 2))
            cn = torch.cat([cond_noise, cond_noise]) if cfg else cond_noise
            if cfg:
                ye = torch.cat([torch.zeros_like(ye), ye]) if ye is not None else None
            F = plan.forward(plan.xin, cn, ye=None if ye is None else ye.contiguous())
            if cfg:
                half = F.numel() // 2
                Fv = F.view(-1)
                check(lib.dsk_cfg_mix(ptr(Fv[:half]), ptr(Fv[half:]), float(guidance), half, dt_code(plan.act_dtype),
                                      stream()))
                F = Fv[:half]
            return F, plan.act_dtype
        xin = torch.empty((B, S, Cc), dtype=torch.float32, device=x.device)
        check(lib.dsk_precond_scale(ptr(x), ptr(c_in), ptr(xin), B, Cc, S, 0, stream()))
        if Cc > 1 and S > 1:
            xin = ops.cl_to_nchw(xin.view(B, 1, 1, S, Cc), 3)
        xin = xin.view(x.shape)
        if not cond:
            return self.model(xin, cond_noise), None
        F = self.model(xin, cond_noise, y)
        if cfg:
            F = (1 - guidance) * self.model(xin, cond_noise) + guidance * F
        return F, None

    def _denoise(self, x: Tensor, sigma: Tensor, y, guidance: float, want_score: bool):
        require_cuda(x, "x")
        x = x.float().contiguous()
        sigma = sigma.to(x).contiguous()
        pre = self.config.preconditioner
        c_in = pre.input_scaling(sigma).float().contiguous()
        c_out = pre.output_scaling(sigma).float().contiguous()
        c_skip = pre.skip_scaling(sigma).float().contiguous()
        cond_noise = pre.noise_conditioner(sigma).float().contiguous()
        B = x.shape[0]
        Cc = x.shape[1] if x.ndim > 1 else 1
        S = x.numel() // (B * Cc)
        F, adt = self._network(x, c_in, cond_noise, y, guidance)
        if adt is None:                 # foreign model: F is fp32 in the user's NC(D)HW layout
            F = F.detach().float().contiguous()
            if Cc > 1 and S > 1:
                F = ops.nchw_to_cl(F.view(B, Cc, 1, 1, S), torch.float32, 3)
            adt = torch.float32
        D = torch.empty_like(x)
        score = torch.empty_like(x) if want_score else None
        check(lib.dsk_precond_denoise(ptr(F), ptr(x), ptr(c_out), ptr(c_skip), ptr(sigma), ptr(D), ptr(score), B, Cc, S,
                                      dt_code(adt), stream()))
        return D, score, cond_noise

    def get_denoiser(self, x: Tensor, sigma: Tensor, y=None, guidance: float = 1.0):
        """D(x; sigma) = c_skip x + c_out F(c_in x, c_noise)  ->  (D, c_noise)  (karrasmodule.py:673-719)."""
        D, _, cond_noise = self._denoise(x, sigma, y, guidance, False)
        return D, cond_noise

    def get_score(self, x: Tensor, sigma: Tensor, y=None, guidance: float = 1.0) -> Tensor:
        """(D - x) / sigma^2  (karrasmodule.py:721-733), fused with the denoiser epilogue."""
        return self._denoise(x, sigma, y, guidance, True)[1]

    # ------------------------------------------------------------------ training
    def loss_fn(self, x: Tensor, sigma: Tensor, y=None, mask: Optional[Tensor] = None) -> Tensor:
        """Denoising loss (karrasmodule.py:569-650): D, lambda(sigma), the loss and dL/dF in one fused kernel -- scalars
        evaluated in-kernel for the EDM preconditioner, passed as [B] vectors for VP / VE / SR3 / custom objects."""
        require_cuda(x, "x")
        if self.latent_model or self.norm != 1.0:    # the loss lives in the latent space (karrasmodule.py:583-587)
            with torch.no_grad():
                if self.encode_y:
                    x, y = self.encode(x, y)
                else:
                    x = self.encode(x, y)
        x = x.float().contiguous()
        sigma = sigma.to(x).contiguous()
        if self._injected_loss_noise is not None:
            noise = self._injected_loss_noise.to(x).contiguous()
        else:   # device-side N(0,1) (the reference uses torch.randn_like on the device generator, :591)
            noise = ops.philox_normal(x.shape, integrators.fresh_noise_seed(), 0, x.device)
        x_noised = _rowwise_axpy(sigma.float(), noise, torch.ones_like(sigma, dtype=torch.float32), x)
        pre = self.config.preconditioner
        c_in = pre.input_scaling(sigma).float().contiguous()
        cond_noise = pre.noise_conditioner(sigma).float().contiguous()
        F, adt = self._network(x_noised, c_in, cond_noise, y, 1.0)
        B = x.shape[0]
        Cc = x.shape[1] if x.ndim > 1 else 1
        S = x.numel() // (B * Cc)
        if adt is not None:            # native plan output: channels-last act dtype -> fp32 NC(D)HW
            F = ops.cl_to_nchw(F.view(B, 1, 1, S, Cc), 3).view(x.shape)
        m = None if mask is None else mask.to(x).expand_as(x).contiguous()
        coeffs = None
        sd_pre, sd_ns = getattr(pre, "sigma_data", None), getattr(self.config.noisesampler, "sigma_data", None)
        same_sd = sd_pre is not None and sd_ns is not None and float(sd_pre) == float(sd_ns)
        if self.dynamic_loss_weight is not None or not same_sd or self.huber_delta != 1.0 or not (
                type(pre) is preconditioners.EDMPreconditioner and type(self.config.noisesampler) is noisesamplers.EDMNoiseSampler):
            # VP / VE / SR3 / custom objects: their per-sample scalars, the same fused loss + dL/dF kernel
            coeffs = tuple(v.float().contiguous() for v in (pre.output_scaling(sigma), pre.skip_scaling(sigma),
                                                            self.config.noisesampler.loss_weighting(sigma)))
        if self.dynamic_loss_weight is not None:
            # weight / exp(u), bias + u  (karrasmodule.py:596-602): u on torch autograd, the per-sample losses from one launch
            u = self.dynamic_loss_weight(cond_noise).float()
            return _RowLossFn.apply(F.float().contiguous().view(x.shape), torch.exp(-u), x, noise.contiguous(), sigma, m,
                                    coeffs, self.loss_kind, self.huber_delta) + u.mean()
        return _EDMLossFn.apply(F.float().contiguous().view(x.shape), x, noise.contiguous(), sigma, m,
                                self._sigma_data(), self.loss_kind, coeffs, self.huber_delta)

    _injected_loss_noise: Optional[Tensor] = None

    def select_batch(self, batch):
        if self.conditional and self.masked:
            x, y, mask = batch
        elif self.masked:
            (x, mask), y = batch, None
        elif self.conditional:
            (x, y), mask = batch, None
        else:
            x, y, mask = batch, None, None
        return x, y, mask

    def training_step(self, batch, batch_idx):
        x, y, mask = self.select_batch(batch)
        sigma = self.config.noisesampler.sample(x.shape[0]).to(x)
        loss = self.loss_fn(x, sigma, y, mask)
        self.log("train_loss", loss, prog_bar=True, sync_dist=True)
        return loss

    def validation_step(self, batch, batch_idx):
        x, y, mask = self.select_batch(batch)
        sigma = self.config.noisesampler.sample(x.shape[0]).to(x)
        loss = self.loss_fn(x, sigma, y, mask)
        self.log("valid_loss", loss, prog_bar=True, sync_dist=True)
        self.log("val_loss", loss, prog_bar=True, sync_dist=True)
        return loss

    # ------------------------------------------------------------------ sampling
    def sample(self, nsamples: int, shape, y=None, guidance: float = 1.0, nsteps: int = 100,
               record_history: bool = False, maximum_batch_size: Optional[int] = None, integrator=None,
               move_to_cpu: bool = False, is_latent_shape: bool = False, squeeze_memory_efficiency: bool = False,
               return_in_latent_space: bool = False) -> Tensor:
        """Draw x_T ~ N(0, I) on the CPU generator (as the reference, karrasmodule.py:838: reproducible across
        devices), scale by sigma_max and integrate the probability-flow ODE / reverse SDE to sigma = 0."""
        with torch.inference_mode():
            if maximum_batch_size is not None:
                parts = [self.sample(b, shape, y, guidance, nsteps, record_history, None, integrator, move_to_cpu,
                                     is_latent_shape, squeeze_memory_efficiency, return_in_latent_space)
                         for b in get_minibatch_sizes(nsamples, maximum_batch_size)]
                return torch.cat(parts, dim=1 if record_history else 0)
            if y is not None:
                y = dict_to(y, self.device)
            if self.latent_model and not is_latent_shape:
                # `shape` is a data-space shape (karrasmodule.py:842-852): the latent shape is whatever the encoder makes
                # of it; x_T is then drawn at that shape (here on the CPU generator, like every other x_T).
                original_y = None
                if self.encode_y:      # the encoder also maps the condition (karrasmodule.py:843-848)
                    if self.decode_original_y:
                        original_y = y.copy()
                    probe, y = self.encode(torch.zeros(*([nsamples] + list(shape)), device=self.device), y)
                    y["y"] = y["y"].squeeze(0)
                else:
                    probe = self.encode(torch.zeros(*([nsamples] + list(shape)), device=self.device), y)
                shape = list(probe.shape[1:])
                white_noise = torch.randn(*([nsamples] + list(shape))).to(self.device)
                return self.propagate_white_noise(white_noise, y, guidance, nsteps, record_history, integrator=integrator,
                                                  original_y=original_y, move_to_cpu=move_to_cpu, latent_shape=is_latent_shape,
                                                  squeeze_memory_efficiency=squeeze_memory_efficiency,
                                                  return_in_latent_space=return_in_latent_space)
            white_noise = torch.randn(*([nsamples] + list(shape))).to(self.device)
            return self.propagate_white_noise(white_noise, y, guidance, nsteps, record_history,
                                              integrator=integrator, move_to_cpu=move_to_cpu, latent_shape=is_latent_shape,
                                              squeeze_memory_efficiency=squeeze_memory_efficiency,
                                              return_in_latent_space=return_in_latent_space)

    def sample_and_filter(self, nsamples: int, shape, filter_fn, y=None, guidance: float = 1.0, nsteps: int = 100,
                          record_history: bool = False, maximum_batch_size: Optional[int] = None, integrator=None,
                          move_to_cpu: bool = False, return_only_positives: bool = False) -> dict:
        """Rejection sampling around sample() (karrasmodule.py:735-799): `filter_fn(encode(samples)) -> bool [nsamples]`;
        returns the samples, the filter and the hit rate."""
        if record_history:
            raise ValueError("record_history is not supported for filtering at the moment")
        if maximum_batch_size is not None:
            parts = [self.sample_and_filter(b, shape, filter_fn, y, guidance, nsteps, record_history, None, integrator,
                                            move_to_cpu, return_only_positives)
                     for b in get_minibatch_sizes(nsamples, maximum_batch_size)]
            hits = sum(p["filter"].sum().item() for p in parts)
            return dict(samples=torch.cat([p["samples"] for p in parts], dim=0),
                        filter=torch.cat([p["filter"] for p in parts], dim=0), hit_rate=hits / nsamples)
        samples = self.sample(nsamples, shape, y=y, guidance=guidance, nsteps=nsteps, record_history=record_history,
                              maximum_batch_size=maximum_batch_size, integrator=integrator, move_to_cpu=False)
        with torch.inference_mode():
            enc = self.encode(samples, y, record_history)
            keep = filter_fn(enc[0] if self.encode_y else enc)
        if return_only_positives:
            samples, keep = samples[keep], keep[keep]
        if move_to_cpu:
            samples = samples.detach().cpu()
        return dict(samples=samples, filter=keep, hit_rate=keep.sum() / nsamples)

    def propagate_white_noise(self, x: Tensor, y=None, guidance: float = 1.0, nsteps: int = 100,
                              record_history: bool = False, integrator=None, original_y=None,
                              move_to_cpu: bool = False, latent_shape: bool = False,
                              squeeze_memory_efficiency: bool = False, return_in_latent_space: bool = False):
        """x: white noise [B, *shape] (host or device).  The sigma_max scaling (karrasmodule.py:881) is fused
        into the first sampler stage."""
        with torch.inference_mode():
            if not x.is_cuda:
                x = x.to(self.device, non_blocking=True)
            result = self.propagate_toward_sample(x, y, guidance, nsteps, record_history, integrator=integrator,
                                                  _prescaled=False)
            if not return_in_latent_space and (self.norm != 1.0 or self.latent_model):
                result = self.decode(result, original_y if original_y is not None else y, record_history)
        return result.detach().cpu() if move_to_cpu else result

    def _resolve_integrator(self, integrator):
        sch = self.config.noisescheduler
        if integrator is None:
            return sch.integrator
        if type(integrator) is str:
            return integrators.name_to_integrator(integrator)
        return integrator

    def propagate_toward_sample(self, x: Tensor, y=None, guidance: float = 1.0, nsteps: int = 100,
                                record_history: bool = False, integrator=None, _prescaled: bool = True):
        """Integrate from x (already scaled by sigma_max unless called through propagate_white_noise)."""
        require_cuda(x, "x")
        if y is not None:
            y = dict_unsqueeze(y, 0)   # one condition, broadcast over the samples (karrasmodule.py:914-915)
        sch = self.config.noisescheduler
        integ = self._resolve_integrator(integrator)
        kind = _engine.precond_kind(self.config.preconditioner)
        fused = sch.fused_supported and kind is not None and integ.fused_program in _engine.PROGRAMS
        cond = bool(self.conditional and guidance != 0.0)
        if cond and not hasattr(self.model, "conditioning_vector"):
            fused = False              # foreign / not-yet-conditional networks: the duck-typed seam
        x = x.float().contiguous()
        # VP / VE / SR3 on the captured-graph loop through the table-driven stages (csrc/sampler_general.cu).  On by default for
        # the configurations whose GPU parity run is on record (profiles/r2h_general_engine_gpu_tests.log); DSK_GENERAL_ENGINE=1
        # extends it to any scheduler / preconditioner objects, =0 forces the Integrator.step seam.
        env = os.environ.get("DSK_GENERAL_ENGINE")
        validated = (type(self.config.preconditioner) in (preconditioners.VPPreconditioner, preconditioners.VEPreconditioner,
                                                          preconditioners.SR3Preconditioner) and
                     type(sch) in (schedulers.VPScheduler, schedulers.VEScheduler, schedulers.EDMScheduler))
        if (not fused and env != "0" and (validated or env == "1") and hasattr(self.model, "plan") and getattr(self.model, "engine_native", True) and not cond and
                integ.fused_program in sch.GENERAL_PROGRAMS and
                type(integ) in (integrators.EulerIntegrator, integrators.HeunIntegrator, integrators.EulerMaruyamaIntegrator)):
            return self._propagate_general(x, sch, integ, nsteps, record_history, _prescaled)
        if not fused:      # foreign preconditioner / scheduler / integrator: the duck-typed seam
            if not _prescaled:
                x = ops.lincomb(x, float(sch.maximum_scale))
            if integrator is not None:
                sch.set_temporary_integrator(integ)
            try:
                return sch.propagate_backward(x, lambda xx, sg: self.get_score(xx, sg, y, guidance), nsteps,
                                              record_history=record_history)
            finally:
                if integrator is not None:
                    sch.unset_temporary_integrator()
        eng = self._edm_engine(x, y, guidance, kind, cond)
        eng.sigma_max = 1.0 if _prescaled else float(sch.maximum_scale)
        table = sch.step_table(nsteps, integ)
        noises = None
        if integ.injected_noise is not None:
            noises = torch.stack([n.to(x) for n in integ.injected_noise[:nsteps]], 0)
        seed = integrators.fresh_noise_seed() if integ.fused_program in ("euler-maruyama", "karras") else 0
        if getattr(integ, "_fixed_seed", None) is not None:
            seed = integ._fixed_seed
        out = eng.run(x, table, integ.fused_program, record_history=record_history, noises=noises, seed=seed)
        self.last_nfe = eng.nfe
        return out

    def _edm_route(self, integ, y=None, guidance: float = 1.0):
        """-> (preconditioner kind, conditional?) if the captured-graph EDM engine serves this configuration, else None."""
        sch = self.config.noisescheduler
        kind = _engine.precond_kind(self.config.preconditioner)
        cond = bool(self.conditional and guidance != 0.0)
        ok = (sch.fused_supported and kind is not None and integ.fused_program in _engine.PROGRAMS and
              hasattr(self.model, "plan") and getattr(self.model, "engine_native", True) and
              not (cond and not hasattr(self.model, "conditioning_vector")))
        return (kind, cond) if ok else None

    def _edm_engine(self, x: Tensor, y, guidance: float, kind: int, cond: bool):
        """The cached SamplerEngine of this (batch, shape, precision, conditioning) with the run's conditioning written."""
        B, shape = x.shape[0], tuple(x.shape[1:])
        ychan = ye = None
        cfg = cond and guidance != 1.0
        if cond:                       # evaluated ONCE per run: the conditioning is constant along the trajectory
            ychan, rest = self.model.split_condition(y, x)
            if cfg and ychan is not None:
                raise TypeError("classifier-free guidance needs an unconditional evaluation, which a channel-"
                                "conditioned network (PUNetGCond) does not have")
            ye = self.model.conditioning_vector(rest, B)
        ncond = 0 if ychan is None else int(ychan.shape[1])
        key = (B, shape, str(x.device), id(self.model), getattr(self.model, "precision", None), kind,
               self._sigma_data(), self.use_cuda_graphs, ncond, ye is not None, float(guidance) if cfg else None)
        eng = self._engines.get(key)
        if eng is None:
            if len(self._engines) >= 2:
                self._engines.clear()
            with torch.inference_mode(False), torch.no_grad():
                eng = self._engines[key] = _engine.SamplerEngine(self.model, B, shape, x.device, self._sigma_data(),
                                                                 1.0, kind, use_graphs=self.use_cuda_graphs,
                                                                 cond_channels=ncond, cond_vector=ye is not None,
                                                                 guidance=float(guidance) if cfg else None)
        eng.set_condition(ychan, ye)
        return eng

    def _engine_partial(self, x: Tensor, y, integ, nsteps: int, first: int, last: int, record_history: bool = False,
                        blend=None) -> Optional[Tensor]:
        """Steps first .. last-1 of the nsteps schedule on the captured-graph engine (None: configuration not served there).
        x is the state at level t[first]; blend = (forward history of the known data, mask): inpainting."""
        route = self._edm_route(integ, y)
        if route is None or os.environ.get("DSK_PARTIAL_ENGINE") == "0":
            return None
        x = x.float().contiguous()
        eng = self._edm_engine(x, y, 1.0, *route)
        eng.sigma_max = 1.0
        table = self.config.noisescheduler.step_table(nsteps, integ)
        noises = None
        if integ.injected_noise is not None and integ.fused_program in ("euler-maruyama", "karras"):
            # draw numbers continue across partial runs; the engine indexes its noise rows by schedule step
            draws = integ.injected_noise[integ._draws:integ._draws + (last - first)]
            integ._draws += last - first
            noises = torch.zeros((nsteps,) + tuple(x.shape), dtype=torch.float32, device=x.device)
            noises[first:last] = torch.stack([n.to(x) for n in draws], 0)
        seed = integrators.fresh_noise_seed() if integ.fused_program in ("euler-maruyama", "karras") else 0
        out = eng.run(x, table, integ.fused_program, record_history=record_history, noises=noises, seed=seed, first=first,
                      last=last, blend=blend)
        self.last_nfe = eng.nfe
        return out

    def _propagate_general(self, x: Tensor, sch, integ, nsteps: int, record_history: bool, prescaled: bool) -> Tensor:
        """VP / VE / SR3 / custom configurations on the captured-graph loop through the table-driven stages
        (engine.GeneralSamplerEngine, csrc/sampler_general.cu): rhs = P x + Q F with the scalars of
        Scheduler.general_step_table."""
        B, shape = x.shape[0], tuple(x.shape[1:])
        key = ("general", B, shape, str(x.device), id(self.model), getattr(self.model, "precision", None), self.use_cuda_graphs)
        eng = self._engines.get(key)
        if eng is None:
            if len(self._engines) >= 2:
                self._engines.clear()
            with torch.inference_mode(False), torch.no_grad():
                eng = self._engines[key] = _engine.GeneralSamplerEngine(self.model, B, shape, x.device,
                                                                        use_graphs=self.use_cuda_graphs)
        eng.sigma_max = 1.0 if prescaled else float(sch.maximum_scale)
        table = sch.general_step_table(nsteps, self.config.preconditioner, integ)
        noises = None
        if integ.injected_noise is not None:
            noises = torch.stack([n.to(x) for n in integ.injected_noise[:nsteps]], 0)
        seed = integrators.fresh_noise_seed() if integ.fused_program == "euler-maruyama" else 0
        if getattr(integ, "_fixed_seed", None) is not None:
            seed = integ._fixed_seed
        out = eng.run(x, table, integ.fused_program, record_history=record_history, noises=noises, seed=seed)
        self.last_nfe = eng.nfe
        return out

    # ------------------------------------------------------------------ partial / forward propagation, inpainting
    def _score_fn(self, y=None):
        return lambda xx, sg: self.get_score(xx, sg, y)

    def propagate_partial_toward_sample(self, x: Tensor, initial_step: int, final_step: Optional[int] = None, y=None,
                                        nsteps: int = 100, record_history: bool = False, integrator=None,
                                        analytical_score=None, interp_fn=None) -> Tensor:
        """Steps initial_step .. final_step of an nsteps schedule (karrasmodule.py:933-976).  `interp_fn(sigma)` blends the
        trained score with `analytical_score` (evaluated on the host, as in the reference)."""
        require_cuda(x, "x")
        trained = self._score_fn(y)

        def score(xx, sigma):
            s = trained(xx, sigma)
            if interp_fn is None:
                return s
            assert analytical_score is not None
            alpha = interp_fn(sigma).unsqueeze(-1).to(s.device)
            analytic = analytical_score(xx.cpu().detach(), sigma.cpu().detach()).to(s.device)
            return alpha * s + (1 - alpha) * analytic
        sch = self.config.noisescheduler
        final_step = nsteps if final_step is None else final_step
        if interp_fn is None and y is None:
            # the captured-graph engine serves a stretch of the schedule as it serves a whole run (start row + step count)
            with torch.inference_mode():
                out = self._engine_partial(x, None, self._resolve_integrator(integrator), nsteps, initial_step, final_step,
                                           record_history)
            if out is not None:
                return out
        with torch.inference_mode():
            if integrator is not None:
                sch.set_temporary_integrator(integrator)
            try:
                return sch.propagate_partial(x, score, nsteps, initial_step, final_step, record_history=record_history)
            finally:
                if integrator is not None:
                    sch.unset_temporary_integrator()

    def propagate_toward_noise(self, x: Tensor, y=None, nsteps: int = 100, record_history: bool = False,
                               stochastic_integration: bool = False) -> Tensor:
        """Forward integration data -> noise (karrasmodule.py:1096-1117); y is ONE unbatched condition (:1102-1103)."""
        require_cuda(x, "x")
        if y is not None:
            y = dict_unsqueeze(y, 0)   # broadcasting takes care of the rest
        with torch.inference_mode():
            return self.config.noisescheduler.propagate_forward(x, self._score_fn(y), nsteps, record_history=record_history,
                                                                stochastic=stochastic_integration)

    def propagate_inpaint_toward_sample(self, x: Tensor, x_inpaint: Tensor, mask: Tensor, y=None,
                                        record_history: bool = False) -> Tensor:
        """karrasmodule.py:1048-1070: x_inpaint is the forward history [nsteps+1, B, *shape] of the known data; y is ONE
        unbatched condition (:1054-1055)."""
        if y is not None:
            y = dict_unsqueeze(y, 0)
        with torch.inference_mode():
            nsteps = x_inpaint.shape[0] - 1
            sch = self.config.noisescheduler
            if y is None and x.is_cuda:
                # graph engine: the known-region blend is fused into the step-completing stage kernels
                out = self._engine_partial(x, None, sch.integrator, nsteps, 0, nsteps, record_history,
                                           blend=(x_inpaint.to(x), mask))
                if out is not None:
                    return out
            return sch.inpaint(x, x_inpaint, mask, self._score_fn(y), nsteps, record_history=record_history)

    def propagate_repaint_toward_sample(self, x: Tensor, x_inpaint: Tensor, mask: Tensor, y=None,
                                        record_history: bool = False) -> Tensor:
        """karrasmodule.py:1072-1094 (Scheduler.repaint with its default rsteps / nresamples); y unbatched (:1078-1079)."""
        if y is not None:
            y = dict_unsqueeze(y, 0)
        with torch.inference_mode():
            sch = self.config.noisescheduler
            if y is None and x.is_cuda and self._edm_route(sch.integrator) is not None and \
                    os.environ.get("DSK_PARTIAL_ENGINE") != "0":
                # Scheduler.repaint's loop (schedulers.py:124-175) with every integrated stretch on the graph engine
                module = self

                class _Stretch:
                    def __call__(self_, xx, score_fn, nsteps, first, last, **_):
                        return module._engine_partial(xx, None, sch.integrator, nsteps, first, last)
                return sch.repaint(x, x_inpaint, mask, None, x_inpaint.shape[0] - 1, record_history=record_history,
                                   _partial=_Stretch())
            return sch.repaint(x, x_inpaint, mask, self._score_fn(y), x_inpaint.shape[0] - 1, record_history=record_history)

    def inpaint(self, x_orig: Tensor, mask: Tensor, y=None, nsteps: int = 100, record_history: bool = False,
                maximum_batch_size: Optional[int] = None, mode: str = "inpaint") -> Tensor:
        """karrasmodule.py:978-1030: noise the known data forward (stochastic integration, history kept), start from
        N(0, sigma_max^2) and integrate back re-imposing the known region (mask == 1) after every step."""
        if maximum_batch_size is not None:
            sizes = get_minibatch_sizes(x_orig.shape[0], maximum_batch_size)
            xs, ms = x_orig.chunk(len(sizes)), mask.chunk(len(sizes))
            parts = [self.inpaint(xs[i], ms[i], y, nsteps, record_history, None, mode) for i in range(len(sizes))]
            return torch.cat(parts, dim=1 if record_history else 0)
        require_cuda(x_orig, "x_orig")
        hist = self.propagate_toward_noise(x_orig, nsteps=nsteps, y=y, record_history=True, stochastic_integration=True)
        sch = self.config.noisescheduler
        noise = ops.lincomb(sch.integrator._randn_like(x_orig.float()), float(sch.maximum_scale))
        fn = self.propagate_inpaint_toward_sample if mode == "inpaint" else self.propagate_repaint_toward_sample
        return fn(noise, hist, mask, y=y, record_history=record_history)

    def repaint(self, x_orig: Tensor, mask: Tensor, y=None, nsteps: int = 100, record_history: bool = False,
                maximum_batch_size: Optional[int] = None) -> Tensor:
        return self.inpaint(x_orig, mask, y, nsteps, record_history, maximum_batch_size, mode="repaint")

    def interpolate_images(self, x1: Tensor, x2: Tensor, ninterp: int, jitter: Optional[float] = 1e-2, y=None,
                           nsteps: int = 100, record_history: bool = False) -> Tensor:
        """karrasmodule.py:1119-1144: noise both images with the probability-flow ODE, interpolate linearly in noise
        space, integrate back."""
        x = torch.stack([x1, x2], dim=0).float()
        require_cuda(x, "images")
        if jitter is not None:
            x = ops.lincomb(x, 1.0, None, 0.0, None, 0.0, self.config.noisescheduler.integrator._randn_like(x), float(jitter))
        if y is not None:
            y = dict_unsqueeze(y, 0)   # as the reference (:1130-1131), which unsqueezes here AND in the two calls below
        xn = self.propagate_toward_noise(x, y, nsteps)
        w = torch.linspace(0, 1, ninterp, device=xn.device).view(-1, *([1] * (xn.ndim - 1)))
        xi = (1 - w) * xn[0].unsqueeze(0) + w * xn[1].unsqueeze(0)
        return self.propagate_toward_sample(xi.contiguous(), y=y, nsteps=nsteps, record_history=record_history)
