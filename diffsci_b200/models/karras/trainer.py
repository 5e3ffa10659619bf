"""EDMTrainer -- one EDM training iteration of a native score network as a fixed sequence of libdiffsci_b200 launches.

What the reference does per iteration (Lightning loop around ``KarrasModule.training_step``, karras/karrasmodule.py:
1146-1155, 569-662; optimizer :497-507; EMA hook karrasmodule_new.py:2127-2227 + karras/ema.py:127-156):

    sigma ~ EDMNoiseSampler (CPU RNG)           noisesamplers.py:35-41
    x_n   = x + sigma * randn_like(x)           karrasmodule.py:591-594
    F     = model(c_in x_n, c_noise)            :690-716          (ATen forward, autograd tape)
    L     = mean(lambda * huber(c_out F + c_skip x_n, x))        :596-648
    L.backward() ; DDP all-reduce ; AdamW.step() ; ema.update()   (~10^3 ATen launches + one lerp per tensor)

Here: Philox noise, fused noising + c_in scaling straight into the network's channels-last input, the static forward
launch list (TrainGraph), ONE fused loss + dL/dF kernel, the hand-written backward launch list with the bucketed gradient
all-reduce overlapped (distributed.GradBucketer), and ONE fused multi-tensor AdamW + EMA kernel.  No autograd tape, no
allocation, no host synchronisation inside the step; the loss comes back as a device scalar.
``KarrasModule.training_step`` + any torch optimizer remains available (autograd seam, nets/graph.py:NetFunction) and is
tested to produce the same gradients.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor

from ... import ops
from ..._lib import lib, check, ptr, stream, dt_code, require_cuda
from ...distributed import GradBucketer
from . import preconditioners
from .ema import ModelEMA


class EDMTrainer:
    def __init__(self, module, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-4,
                 ema: Optional[ModelEMA] = None, process_group=None, bucket_mb: float = 32.0, seed: int = 0xD1FF5C1,
                 data_parallel: bool = True):
        """Defaults = the reference's default optimizer (karrasmodule.py:497-500).  `ema`: a ModelEMA over
        ``module.model`` (its first profile is updated inside the AdamW kernel, further profiles by dsk_ema_update)."""
        net = module.model
        if not hasattr(net, "train_graph"):
            raise TypeError("EDMTrainer drives the native networks (PUNetG / ADM); foreign torch modules train through "
                            "KarrasModule.training_step + a torch optimizer")
        # any KarrasPreconditioner / NoiseSampler pair (EDM, VP, VE, SR3, custom): the EDM pair has its scalars evaluated inside
        # the loss kernel, the others hand it per-sample (c_out, c_skip, lambda) vectors from their own objects
        if module.conditional or getattr(net, "conditional_embedding", None) is not None or getattr(net, "cond_drop", None) is not None:
            raise NotImplementedError("EDMTrainer: conditional models train through KarrasModule.training_step + a torch "
                                      "optimizer (the native backward returns d loss / d embedding to autograd)")
        if getattr(module, "dynamic_loss_weight", None) is not None:
            raise NotImplementedError("EDMTrainer: the learned loss weighting (its own torch parameters) trains through "
                                      "KarrasModule.training_step + a torch optimizer")
        self.module, self.net, self.ema, self.group = module, net, ema, process_group
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(weight_decay)
        self.bucket_bytes = int(bucket_mb * (1 << 20))
        self.seed, self.nstep, self.data_parallel = int(seed), 0, bool(data_parallel)
        self._state = None        # (graph id, tables...)
        self.params = [p for p in net.parameters()]
        self.exp_avg = [torch.zeros_like(p) for p in self.params]
        self.exp_avg_sq = [torch.zeros_like(p) for p in self.params]

    # ------------------------------------------------------------------ per-graph tables
    def _tables(self, g):
        key = (id(g), tuple(p.data_ptr() for p in self.params))
        if self._state is None or self._state[0] != key:
            dev = g.device
            mk = lambda ts: torch.tensor([t.data_ptr() for t in ts], dtype=torch.int64, device=dev)  # noqa: E731
            numels = [p.numel() for p in self.params]
            ready = [g.grad_ready_pos[id(p)] for p in self.params]
            bucketer = GradBucketer(g.flat_grad, numels, ready, self.bucket_bytes, self.group, self.data_parallel)
            self._state = (key, mk(self.params), mk(g.grads()), mk(self.exp_avg), mk(self.exp_avg_sq),
                           torch.tensor(numels, dtype=torch.int64, device=dev), max(numels), bucketer)
        return self._state

    # ------------------------------------------------------------------ one iteration
    def step(self, x: Tensor, sigma: Optional[Tensor] = None, noise: Optional[Tensor] = None,
             mask: Optional[Tensor] = None) -> Tensor:
        """x: fp32 [B, C, *S] on the device.  Returns the loss as a 0-dim device tensor (no host sync)."""
        require_cuda(x, "training batch")
        mod, net = self.module, self.net
        if getattr(mod, "latent_model", False):          # latent diffusion: the loss lives on encode(x) (karrasmodule.py:583-587)
            with torch.no_grad():
                x = mod.encode(x)
        x = x.float().contiguous()
        B = x.shape[0]
        Cc = x.shape[1]
        S = x.numel() // (B * Cc)
        with torch.no_grad():
            g = net.train_graph(B, tuple(x.shape[2:]), x.device)
            _, ptab, gtab, mtab, vtab, ntab, max_numel, bucketer = self._tables(g)
            if sigma is None:
                sigma = mod.config.noisesampler.sample(B)              # CPU generator, as the reference
            sigma = sigma.to(x).contiguous()
            if noise is None:
                noise = ops.philox_normal(x.shape, self.seed, self.nstep, x.device)
            else:
                noise = noise.to(x).contiguous()
            pre = mod.config.preconditioner
            c_in = pre.input_scaling(sigma).float().contiguous()
            g.t_in.copy_(pre.noise_conditioner(sigma).float().reshape(-1))
            # x_n = x + sigma n ; network input = c_in x_n, written channels-last in the activation dtype
            x_n = torch.empty_like(x)
            check(lib.dsk_precond_denoise(ptr(noise), ptr(x), ptr(sigma), ptr(torch.ones_like(sigma)), None, ptr(x_n), None, B, 1,
                                          Cc * S, 0, stream()))
            check(lib.dsk_precond_scale(ptr(x_n), ptr(c_in), ptr(g.x_in.t), B, Cc, S, dt_code(g.act_dtype), stream()))
            g.run_forward()
            F = ops.cl_to_nchw(g.output.t, g.ndim)
            loss = torch.zeros((), dtype=torch.float32, device=x.device)
            dF = torch.empty_like(x)
            m = None if mask is None else mask.to(x).expand_as(x).contiguous()
            smp = mod.config.noisesampler
            sd_p, sd_s = getattr(pre, "sigma_data", None), getattr(smp, "sigma_data", None)
            edm_pair = (type(pre) is preconditioners.EDMPreconditioner and type(smp).__name__ == "EDMNoiseSampler" and
                        sd_p is not None and sd_s is not None and float(sd_p) == float(sd_s) and
                        float(getattr(mod, "huber_delta", 1.0)) == 1.0)
            if edm_pair:
                check(lib.dsk_edm_loss_fwd_bwd(ptr(F), ptr(x), ptr(noise), ptr(sigma), ptr(m), ptr(loss), ptr(dF), B, Cc, S,
                                               float(mod._sigma_data()), int(mod.loss_kind), stream()))
            else:
                c_out, c_skip, lam = (v.float().contiguous() for v in (pre.output_scaling(sigma), pre.skip_scaling(sigma),
                                                                       smp.loss_weighting(sigma)))
                check(lib.dsk_precond_loss_fwd_bwd_huber(ptr(F), ptr(x), ptr(noise), ptr(sigma), ptr(c_out), ptr(c_skip), ptr(lam),
                                                         ptr(m), ptr(loss), ptr(dF), B, Cc, S, int(mod.loss_kind),
                                                         float(getattr(mod, "huber_delta", 1.0)), stream()))
            g.backward_nchw(dF, bucketer.hooks())
            gscale = bucketer.finish()
            self.nstep += 1
            sh, ema_beta = None, 0.0
            if self.ema is not None:
                nxt = self.ema.num_updates + 1
                _, shadow_tabs, _, _, _, _ = self.ema._pointer_tables(net)
                sh = shadow_tabs[0]
                ema_beta = self.ema._beta_for_profile(self.ema.profiles[0], nxt)
                self.ema.profiles[0]["last_beta"] = ema_beta
            check(lib.dsk_adamw_ema_step(ptr(ptab), ptr(gtab), ptr(mtab), ptr(vtab), ptr(sh), ptr(ntab), len(self.params),
                                         max_numel, self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay,
                                         self.nstep, ema_beta, gscale, stream()))
            ops.bump_weight_epoch()     # the kernel wrote the parameters behind PyTorch's version counters
            if self.ema is not None:
                _, shadow_tabs, ptab_e, ntab_e, max_e, nt = self.ema._pointer_tables(net)
                for pr, stab in list(zip(self.ema.profiles, shadow_tabs))[1:]:
                    beta = self.ema._beta_for_profile(pr, nxt)
                    pr["last_beta"] = beta
                    check(lib.dsk_ema_update(ptr(stab), ptr(ptab_e), ptr(ntab_e), nt, max_e, beta, stream()))
                self.ema.num_updates = nxt
                self.ema.last_beta = self.ema.selected_profile().get("last_beta")
        return loss
