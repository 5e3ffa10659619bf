"""EnsembleKarrasModule / EnsembleKarrasModuleConfig -- SURVEY.md 8(f)-4: the ensemble training loss of
diffsci.models.karras.karrasmodule_new (reference karras/karrasmodule_new.py:963-1149) on the fused CUDA path.

What the reference does for ``n_ensemble = E > 1``: draw E noises per sample, evaluate the denoiser ONCE on the B*E rows,
reshape to [B, E, ...] and hand the whole ensemble to an ensemble-aware metric (custom_losses.py:536-690, 765-865) that
returns a scalar; the scalar is multiplied by the batch MEAN of lambda(sigma) (``weight.mean() * loss``,
karrasmodule_new.py:1136-1138).  Here: one launch builds the B*E noisy rows (``dsk_ensemble_noise_add``), the network runs
on them as for any batch, and one launch (``dsk_ensemble_loss_fwd_bwd``) evaluates D, the metric ("huber", "mse",
"CRPS"), its reduction and dL/dF.  Every constant of the reference's reductions is folded into two per-sample scale vectors
on the host side (`_ensemble_scales`).

Kept from the reference's class: the constructor / config signatures, ``loss_fn(x, sigma, y, mask, n_ensemble)``,
``old_loss_fn``, ``training_step`` / ``validation_step`` using ``config.ensemble_size_{train,val}``, the EMA hook
(``on_before_zero_grad``, karrasmodule_new.py:2155-2157).  Replay loss, layer freezing, L2-SP regularisation,
autoregressive mixins, multi-space and windowed / indicator metrics are training recipes outside SURVEY section 8 and
raise NotImplementedError.
"""
from __future__ import annotations

from contextlib import contextmanager
from typing import Any, Dict, Optional, Union

import torch
from torch import Tensor

from ... import ops
from ..._lib import lib, check, ptr, stream, require_cuda
from .ema import ModelEMA
from .karrasmodule import KarrasModule, KarrasModuleConfig

_ENSEMBLE_KINDS = {"huber": 0, "mse": 1, "CRPS": 2}
MAX_ENSEMBLE = 16


class EnsembleKarrasModuleConfig(KarrasModuleConfig):
    """karrasmodule_new.py:32-166.  The replay / freeze / regularisation keys are accepted so that existing constructor
    calls keep working, and must stay at their inert defaults."""

    def __init__(self, preconditioner, noisesampler, noisescheduler, loss_metric: Union[str, Dict[str, Any]] = "huber",
                 tag: str = "custom", has_edm_batch_norm: bool = False, dynamic_loss_weight: Optional[int] = None,
                 extra_args: Optional[dict] = None, ensemble_size_train: int = 1, ensemble_size_val: int = 1,
                 ensemble_size_test: int = 1, ema_enabled: bool = False, ema_type: str = "traditional",
                 ema_decay: float = 0.999, ema_halflife_steps: Optional[float] = None,
                 ema_rampup_ratio: Optional[float] = None, ema_power_function_stds=None,
                 ema_use_for_validation: bool = True, ema_use_for_sampling: bool = True, ema_device: Optional[str] = None,
                 ema_profile_index: int = 0, freeze_layer_patterns=None, freeze_layer_strict: bool = True,
                 replay_enabled: bool = False, replay_loss_weight: float = 0.1, replay_loss_schedule=None,
                 replay_validation_enabled: bool = False, pretrained_weight_regularization=None, **common):
        super().__init__(preconditioner, noisesampler, noisescheduler, loss_metric=loss_metric, tag=tag,
                         has_edm_batch_norm=has_edm_batch_norm, dynamic_loss_weight=dynamic_loss_weight,
                         extra_args=extra_args, **common)
        if freeze_layer_patterns or replay_enabled or replay_validation_enabled or pretrained_weight_regularization:
            raise NotImplementedError("diffsci_b200: layer freezing, replay loss and L2-SP regularisation are training "
                                      "recipes outside SURVEY.md section 8")
        self.ensemble_size_train, self.ensemble_size_val, self.ensemble_size_test = (
            int(ensemble_size_train), int(ensemble_size_val), int(ensemble_size_test))
        self.ema_enabled, self.ema_type, self.ema_decay = bool(ema_enabled), ema_type, ema_decay
        self.ema_halflife_steps, self.ema_rampup_ratio = ema_halflife_steps, ema_rampup_ratio
        self.ema_power_function_stds = ema_power_function_stds
        self.ema_use_for_validation, self.ema_use_for_sampling = ema_use_for_validation, ema_use_for_sampling
        self.ema_device, self.ema_profile_index = ema_device, ema_profile_index
        self.freeze_layer_patterns, self.freeze_layer_strict = None, freeze_layer_strict
        self.replay_enabled = self.replay_validation_enabled = False
        self.replay_loss_weight, self.replay_loss_schedule = replay_loss_weight, replay_loss_schedule
        self.pretrained_weight_regularization = None

    _EMA_KEYS = ("ema_enabled", "ema_type", "ema_decay", "ema_halflife_steps", "ema_rampup_ratio", "ema_power_function_stds",
                 "ema_use_for_validation", "ema_use_for_sampling", "ema_device", "ema_profile_index")
    _INERT_KEYS = ("freeze_layer_patterns", "freeze_layer_strict", "replay_enabled", "replay_loss_weight",
                   "replay_loss_schedule", "replay_validation_enabled", "pretrained_weight_regularization")

    @classmethod
    def _with_own(cls, base_ctor, kwargs):
        """The reference's factories take the EMA keys as **ema_kwargs and nothing else beyond their named arguments --
        the ensemble sizes are constructor arguments / attributes only (karrasmodule_new.py:204-211, 238-261)."""
        ema = {k: kwargs.pop(k) for k in list(kwargs) if k in cls._EMA_KEYS}
        inert = {k: kwargs.pop(k) for k in list(kwargs) if k in cls._INERT_KEYS}
        import inspect
        unknown = set(kwargs) - set(inspect.signature(base_ctor).parameters)
        if unknown:
            raise TypeError(f"Unexpected EMA config key(s): {', '.join(sorted(unknown))}")
        if any(inert.get(k) for k in ("freeze_layer_patterns", "replay_enabled", "replay_validation_enabled",
                                      "pretrained_weight_regularization")):
            raise NotImplementedError("diffsci_b200: layer freezing, replay loss and L2-SP regularisation are training "
                                      "recipes outside SURVEY.md section 8")
        cfg = base_ctor(**kwargs)
        for k, v in ema.items():
            setattr(cfg, k, v)
            cfg.extra_args[k] = v
        return cfg

    @classmethod
    def from_edm(cls, **kwargs):          # karrasmodule_new.py:238-351
        return cls._with_own(super().from_edm, kwargs)

    @classmethod
    def from_vp(cls, **kwargs):           # karrasmodule_new.py:353-438
        return cls._with_own(super().from_vp, kwargs)

    @classmethod
    def from_ve(cls, **kwargs):           # karrasmodule_new.py:440-515
        return cls._with_own(super().from_ve, kwargs)


def _ensemble_scales(kind: int, lam_mean: Tensor, B: int, E: int, C: int, S: int, mask: Optional[Tensor],
                     single: bool = False):
    """The constants of the reference's reductions as per-sample vectors (s1: target term, s2: CRPS pair term) and the mask
    the kernel applies per element (None for CRPS, whose mask only rescales samples).

    huber  (custom_losses.py:623-690): mean_b [ sum_{e,c,s} l (1-m) / clamp(sum(1 - m_b as stored), 1) ], no mask: /(E C S)
    mse    (custom_losses.py:543-560): mean_e [ sum_{b,c,s} l (1-m) / clamp(sum(1 - m as stored), 1) ],   no mask: /(B C S)
    CRPS   (custom_losses.py:817-865): mean_b [ f_b (mean_{e,c,s}|D_e - x| - 0.5 mean_{i<j} mean_{c,s}|D_i - D_j|) ],
           f_b = fraction of pixels with mask == 0 (1 without a mask).
    single: a 4-D prediction through the same metric objects (n_ensemble <= 1 on an ensemble-configured module,
           custom_losses.py:610-621): Huber then normalises by the GLOBAL count of unmasked entries, like mse."""
    N = C * S
    dev = lam_mean.device
    ones = torch.ones(B, dtype=torch.float32, device=dev)
    if kind == 2:
        f = ones
        if mask is not None:
            mexp = mask.expand(B, C, *mask.shape[2:]) if mask.shape[1] == 1 else mask
            f = (~mexp.bool()).reshape(B, -1).sum(1).float().clamp(min=1) / N
        s1 = lam_mean * f / (B * E * N)
        npairs = max(E * (E - 1) / 2, 1)
        s2 = 0.5 * lam_mean * f / (B * npairs * N)
        return s1.contiguous(), s2.contiguous(), None
    if mask is None:
        return (lam_mean * ones / (B * E * N)).contiguous(), None, None
    keep = 1.0 - mask.float()
    if kind == 0 and not single:
        cnt = keep.reshape(B, -1).sum(1).clamp(min=1)
        s1 = lam_mean / (B * cnt)
    else:
        s1 = lam_mean * ones / (E * keep.sum().clamp(min=1))
    return s1.contiguous(), None, mask.float().contiguous()


class _EnsembleLossFn(torch.autograd.Function):
    """Scalar ensemble loss; backward = dL/dF [B*E, ...] written by the same launch (csrc/train.cu: ens_loss_kernel)."""

    @staticmethod
    def forward(ctx, F, x, noise, sigma, c_out, c_skip, s1, s2, mask, kind, E):
        B = x.shape[0]
        Cc = x.shape[1] if x.ndim > 1 else 1
        S = x.numel() // (B * Cc)
        loss = torch.zeros((), dtype=torch.float32, device=x.device)
        dF = torch.empty_like(F, dtype=torch.float32)
        mask_c = 0 if mask is None else int(mask.shape[1])
        check(lib.dsk_ensemble_loss_fwd_bwd(ptr(F), ptr(x), ptr(noise), ptr(sigma), ptr(c_out), ptr(c_skip), ptr(s1), ptr(s2),
                                            ptr(mask), mask_c, ptr(loss), ptr(dF), B, int(E), Cc, S, int(kind), stream()))
        ctx.save_for_backward(dF)
        return loss

    @staticmethod
    def backward(ctx, g):
        (dF,) = ctx.saved_tensors
        return (dF * g,) + (None,) * 10


class EnsembleKarrasModule(KarrasModule):
    def __init__(self, model: torch.nn.Module, config: EnsembleKarrasModuleConfig, conditional: bool = False,
                 masked: bool = False, autoencoder: Optional[torch.nn.Module] = None,
                 autoencoder_conditional: bool = False, encode_y: bool = False, decode_original_y: bool = False):
        for k in ("ensemble_size_train", "ensemble_size_val", "ensemble_size_test"):
            if not hasattr(config, k):
                setattr(config, k, 1)
        super().__init__(model, config, conditional, masked, autoencoder, autoencoder_conditional, encode_y,
                         decode_original_y)
        self.start_ema()
        self._ema_scope_depth = 0
        self._ema_validation_backup = None

    # ------------------------------------------------------------------ loss configuration
    def set_loss_metric(self):
        """karrasmodule_new.py:832-961.  "CRPS" only exists on the ensemble-aware side of the reference's switch (any
        ensemble size != 1); with all sizes at 1 the plain metrics apply."""
        lm = self.config.loss_metric
        name = lm
        if isinstance(lm, dict):
            if "losses" in lm:
                raise NotImplementedError("diffsci_b200: multi-space losses are out of scope (SURVEY.md section 2 #10)")
            name, params = next(iter(lm.items()))
            if name == "huber" and (params or {}).get("delta", 1.0) != 1.0:
                raise NotImplementedError("diffsci_b200: Huber delta != 1 is not fused yet")
        elif not isinstance(lm, str):
            raise ValueError(f"loss_metric must be string or dict, got {type(lm)}")
        ensemble = not (self.config.ensemble_size_train == 1 and self.config.ensemble_size_test == 1)
        known = ("mse", "huber", "weighted_gaussian", "smoothed_indicator") + (("CRPS",) if ensemble else ())
        if name not in known:
            raise ValueError(f"loss_type {name} not recognized")
        if name not in _ENSEMBLE_KINDS:
            raise NotImplementedError(f"diffsci_b200: loss_metric={name!r} is out of scope (SURVEY.md section 2 #10); "
                                      "'huber', 'mse' and 'CRPS' are fused")
        self.loss_metric = name
        self.ensemble_kind = _ENSEMBLE_KINDS[name]
        self.loss_kind = min(self.ensemble_kind, 1)
        self.ensemble_metrics = ensemble

    def old_loss_fn(self, x: Tensor, sigma: Tensor, y=None, mask: Optional[Tensor] = None) -> Tensor:
        """karrasmodule_new.py:1151-1235.  With every ensemble size at 1 the metric is elementwise and lambda(sigma) weights
        each sample (= KarrasModule.loss_fn); on an ensemble-configured module the metric objects are the ensemble-aware
        ones, which reduce to a scalar first, so the weight is the batch mean of lambda (:1222-1224)."""
        if self.ensemble_metrics:
            return self.loss_fn(x, sigma, y, mask, n_ensemble=1, _force_ensemble=True)
        return KarrasModule.loss_fn(self, x, sigma, y, mask)

    _injected_ensemble_noise: Optional[Tensor] = None     # [B, E, *shape], tests only

    def loss_fn(self, x: Tensor, sigma: Tensor, y=None, mask: Optional[Tensor] = None, n_ensemble: int = 1,
                _force_ensemble: bool = False) -> Tensor:
        """karrasmodule_new.py:963-1149."""
        if n_ensemble <= 1 and not _force_ensemble:
            return self.old_loss_fn(x, sigma, y, mask)
        E = max(int(n_ensemble), 1)
        if self.dynamic_loss_weight is not None:
            # the reference's own ensemble branch fails here (cond_noise.mean(dim=1) on a 1-D tensor, karrasmodule_new.py:1103)
            raise NotImplementedError("diffsci_b200: dynamic_loss_weight with the ensemble-aware metrics")
        if E > MAX_ENSEMBLE:
            raise NotImplementedError(f"diffsci_b200: n_ensemble={E} > {MAX_ENSEMBLE}")
        require_cuda(x, "x")
        if self.latent_model or self.norm != 1.0:
            with torch.no_grad():
                if self.encode_y:
                    x, y = self.encode(x, y)
                else:
                    x = self.encode(x, y)
        x = x.float().contiguous()
        sigma = sigma.to(x).contiguous()
        B = x.shape[0]
        Cc = x.shape[1] if x.ndim > 1 else 1
        S = x.numel() // (B * Cc)
        eshape = (B, E) + tuple(x.shape[1:])
        if self._injected_ensemble_noise is not None:
            noise = self._injected_ensemble_noise.to(x).reshape(eshape).contiguous()
        else:
            noise = ops.philox_normal(eshape, int(torch.randint(0, 2 ** 62, (1,)).item()), 0, x.device)
        x_noised = torch.empty((B * E,) + tuple(x.shape[1:]), dtype=torch.float32, device=x.device)
        check(lib.dsk_ensemble_noise_add(ptr(x), ptr(noise), ptr(sigma), ptr(x_noised), B, E, Cc * S, stream()))
        sig_e = sigma.repeat_interleave(E)
        pre = self.config.preconditioner
        c_in = pre.input_scaling(sig_e).float().contiguous()
        cond_noise = pre.noise_conditioner(sig_e).float().contiguous()
        if y is not None:
            y = self._expand_condition(y, B, E)
        F, adt = self._network(x_noised, c_in, cond_noise, y, 1.0)
        if adt is not None:            # native plan output: channels-last act dtype -> fp32 NC(D)HW
            F = ops.cl_to_nchw(F.view(B * E, 1, 1, S, Cc), 3).view(x_noised.shape)
        lam_mean = self.config.noisesampler.loss_weighting(sigma).float().mean()
        m = None
        if mask is not None:
            m = mask.to(x)
            if m.ndim == x.ndim - 1:
                m = m.unsqueeze(1)
            if m.shape[1] not in (1, Cc) or tuple(m.shape[2:]) != tuple(x.shape[2:]):
                raise ValueError(f"mask shape {tuple(mask.shape)} does not broadcast against {tuple(x.shape)}")
        s1, s2, m = _ensemble_scales(self.ensemble_kind, lam_mean, B, E, Cc, S, m, single=_force_ensemble)
        return _EnsembleLossFn.apply(F.float().contiguous().view(x_noised.shape), x, noise, sigma,
                                     pre.output_scaling(sigma).float().contiguous(),
                                     pre.skip_scaling(sigma).float().contiguous(), s1, s2, m, self.ensemble_kind, E)

    @staticmethod
    def _expand_condition(y, B: int, E: int):
        """[B, ...] -> [B*E, ...], member-minor like the noisy rows (karrasmodule_new.py:1049-1075)."""
        def rep(v):
            return v.unsqueeze(1).expand(B, E, *v.shape[1:]).reshape(B * E, *v.shape[1:])
        if isinstance(y, dict):
            return {k: (rep(v) if isinstance(v, Tensor) else v) for k, v in y.items() if v is not None}
        return rep(y)

    # ------------------------------------------------------------------ steps
    def training_step(self, batch, batch_idx):
        x, y, mask = self.select_batch(batch)
        sigma = self.config.noisesampler.sample(x.shape[0]).to(x)
        loss = self.loss_fn(x, sigma, y, mask, n_ensemble=self.config.ensemble_size_train)
        self.log("train_loss", loss, prog_bar=True, sync_dist=True)
        return loss

    def validation_step(self, batch, batch_idx, dataloader_idx: int = 0):
        x, y, mask = self.select_batch(batch)
        sigma = self.config.noisesampler.sample(x.shape[0]).to(x)
        loss = self.loss_fn(x, sigma, y, mask, n_ensemble=self.config.ensemble_size_val)
        self.log("valid_loss", loss, prog_bar=True, sync_dist=True)
        self.log("val_loss", loss, prog_bar=True, sync_dist=True)
        return loss

    # ------------------------------------------------------------------ EMA (karrasmodule_new.py:2127-2157)
    def start_ema(self):
        self.ema_tracker = None
        if getattr(self.config, "ema_enabled", False):
            g = lambda k, d: getattr(self.config, k, d)  # noqa: E731
            self.ema_tracker = ModelEMA(self.model, ema_type=g("ema_type", "traditional"), decay=g("ema_decay", 0.999),
                                        halflife_steps=g("ema_halflife_steps", None),
                                        rampup_ratio=g("ema_rampup_ratio", None),
                                        power_function_stds=g("ema_power_function_stds", None),
                                        device=g("ema_device", None), profile_index=g("ema_profile_index", 0))

    @property
    def has_ema(self) -> bool:
        return self.ema_tracker is not None

    def on_fit_start(self):
        if self.has_ema:
            self.ema_tracker.reset(self.model)

    def on_before_zero_grad(self, optimizer):
        if self.has_ema:
            self.ema_tracker.update(self.model)

    # EMA weights for validation / eval-time sampling (karrasmodule_new.py:2190-2227).  apply_to / restore copy into the live
    # parameters in place; the sampler engine refreshes its packed weight copies at the start of every run.
    def _should_use_ema_for_validation(self) -> bool:
        return self.has_ema and getattr(self.config, "ema_use_for_validation", True) and self._ema_scope_depth == 0

    def _should_use_ema_for_sampling(self) -> bool:
        return (self.has_ema and getattr(self.config, "ema_use_for_sampling", True) and not self.training and
                self._ema_scope_depth == 0)

    @contextmanager
    def ema_scope(self, enabled: bool = True):
        if not enabled or not self.has_ema or self._ema_scope_depth > 0:
            yield
            return
        with torch.inference_mode(False), torch.no_grad():
            backup = self.ema_tracker.apply_to(self.model)
        self._ema_scope_depth += 1
        try:
            yield
        finally:
            with torch.inference_mode(False), torch.no_grad():
                self.ema_tracker.restore(self.model, backup)
            self._ema_scope_depth = max(self._ema_scope_depth - 1, 0)

    def on_validation_epoch_start(self):
        if self._should_use_ema_for_validation():
            with torch.no_grad():
                self._ema_validation_backup = self.ema_tracker.apply_to(self.model)
            self._ema_scope_depth += 1

    def on_validation_epoch_end(self):
        if self._ema_validation_backup is not None:
            self.ema_tracker.restore(self.model, self._ema_validation_backup)
            self._ema_validation_backup = None
            self._ema_scope_depth = max(self._ema_scope_depth - 1, 0)

    def sample(self, nsamples: int, shape, y=None, guidance: float = 1.0, nsteps: int = 100, record_history: bool = False,
               maximum_batch_size: Optional[int] = None, integrator=None, move_to_cpu: bool = False,
               is_latent_shape: bool = False, squeeze_memory_efficiency: bool = False,
               return_in_latent_space: bool = False, use_ema: Optional[bool] = None) -> Tensor:
        """karrasmodule_new.py:1386-1471: KarrasModule.sample, on the EMA weights when the module is in eval mode and the
        config says so (or use_ema=True)."""
        if use_ema is None:
            use_ema = self._should_use_ema_for_sampling()
        with self.ema_scope(enabled=bool(use_ema)):
            return super().sample(nsamples, shape, y, guidance, nsteps, record_history, maximum_batch_size, integrator,
                                  move_to_cpu, is_latent_shape, squeeze_memory_efficiency, return_in_latent_space)
