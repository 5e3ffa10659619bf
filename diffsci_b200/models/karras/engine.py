"""CUDA-graph sampler engine: the hot loop of ``KarrasModule.sample`` (reference call stack
karrasmodule.py:801-931 -> schedulers.py:48-89 -> integrators.py:29-113 -> karrasmodule.py:673-733).

One integrator step = {network evaluation(s)} + {one fused elementwise stage per evaluation}
(csrc/sampler.cu).  Step scalars live in a device table indexed by a device-side step counter, so
ONE captured graph is replayed for every step (plus a second graph for the final step, whose
``t + dt == 0`` branch the reference resolves with a host sync, integrators.py:45-50).

Works with
  * native networks (``PUNetG`` / ``ADM`` / ``MLPUncond`` of this package): their ``plan(...)`` exposes
    static buffers, the step is captured into a CUDA graph;
  * any other ``torch.nn.Module`` honouring ``model(x_scaled, c_noise) -> F`` (the reference's
    denoiser-net seam): the same fused stages run eagerly around the foreign forward.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from ... import ops
from ..._lib import (lib, check, ptr, stream, dt_code, require_cuda, launch_count, STAGE_INIT, STAGE_EULER, STAGE_HEUN_MID,
                     STAGE_HEUN_FIN, STAGE_HEUN_LAST, STAGE_EM, STAGE_KARRAS_MID, STAGE_KARRAS_FIN,
                     STAGE_KARRAS_LAST)
from . import preconditioners

# program -> (stages of a regular step, stages of the final step); one network evaluation precedes each stage
PROGRAMS = {
    "euler": ((STAGE_EULER,), (STAGE_EULER,)),
    "heun": ((STAGE_HEUN_MID, STAGE_HEUN_FIN), (STAGE_HEUN_LAST,)),
    "euler-maruyama": ((STAGE_EM,), (STAGE_EM,)),
    "karras": ((STAGE_KARRAS_MID, STAGE_KARRAS_FIN), (STAGE_KARRAS_LAST,)),
}


def precond_kind(precond) -> Optional[int]:
    if type(precond) is preconditioners.EDMPreconditioner:
        return 0
    if type(precond) is preconditioners.NullPreconditioner:
        return 1
    return None


class SamplerEngine:
    def __init__(self, model: torch.nn.Module, B: int, shape: Sequence[int], device, sigma_data: float,
                 sigma_max: float, kind: int, use_graphs: bool = True, cond_channels: int = 0,
                 cond_vector: bool = False, guidance: Optional[float] = None):
        """cond_channels: channel-concatenated conditioning (PUNetGCond) held in the network-input buffer; cond_vector:
        a [B, M] conditioning vector is added to the time embedding; guidance (not None): classifier-free guidance, the
        network runs a 2B batch per evaluation (rows [0,B) unconditional, [B,2B) conditional) -- native networks only."""
        self.model, self.B, self.shape = model, int(B), tuple(int(s) for s in shape)
        self.cond_channels, self.cond_vector, self.guidance = int(cond_channels), bool(cond_vector), guidance
        self.cfg = guidance is not None
        self.Bn = 2 * self.B if self.cfg else self.B
        self.device = torch.device(device)
        self.sigma_data, self.sigma_max, self.kind = float(sigma_data), float(sigma_max), int(kind)
        self.Cc = self.shape[0]
        self.S = 1
        for s in self.shape[1:]:
            self.S *= s
        f32 = dict(dtype=torch.float32, device=self.device)
        full = (self.B,) + self.shape
        self.x = torch.empty(full, **f32)
        self.x_aux = torch.empty(full, **f32)
        self.r1 = torch.empty(full, **f32)
        self.cnoise = torch.empty((self.Bn,), **f32)
        self.row = torch.zeros((4,), dtype=torch.int32, device=self.device)   # [step, seed_lo, seed_hi, -]
        self.native = hasattr(model, "plan") and getattr(model, "engine_native", True)
        self.xin_ld = self.Cc + self.cond_channels + int(getattr(model, "ones_channel", 0))   # + the bias=False ones channel
        self.ye = None
        if (self.cond_channels or self.cond_vector or self.cfg) and not self.native:
            raise NotImplementedError("SamplerEngine: the conditional path needs a native network")
        if self.native:
            self.plan = model.plan(self.Bn, self.shape[1:], self.device)
            self.act_dtype = self.plan.act_dtype
            self.xin = self.plan.xin
            if self.xin.shape[-1] != self.xin_ld:
                raise ValueError(f"network expects {self.xin.shape[-1]} input channels, got {self.Cc} state + "
                                 f"{self.cond_channels} conditioning channels")
            if self.cond_vector:
                cdim = getattr(model, "cond_dim", None) or model.config.model_channels
                self.ye = torch.zeros((self.Bn, cdim), **f32)   # rows [0,B) stay 0 under CFG
        else:
            self.plan = None
            self.act_dtype = torch.float32
            self.xin = torch.empty((self.B, self.S, self.Cc), **f32)
        self.use_graphs = bool(use_graphs and self.native)
        self._graphs = {}
        self._graph_launches = {}
        self._last_graph_launches = 0
        self._tab = None
        self._noise = None
        self._hist = None
        self._F = None
        self._blend_y = None      # inpainting: forward history of the known data [nsteps + 1, N] and the mask (static buffers)
        self._blend_mask = None
        self._blend_on = False
        self.seed = 0
        self.nfe = 0

    programs = PROGRAMS

    def _program(self, program: str, table: torch.Tensor):
        """-> (stages of a regular step, stages of the final step)."""
        return PROGRAMS[program]

    # ------------------------------------------------------------------ pieces
    def _net(self):
        self.nfe += 1
        if self.native:
            self._F = self.plan.forward(self.xin, self.cnoise, ye=self.ye) if self.ye is not None else \
                self.plan.forward(self.xin, self.cnoise)
            return
        xin = self.xin
        if self.Cc > 1 and self.S > 1:
            xin = ops.cl_to_nchw(xin.view(self.B, 1, 1, self.S, self.Cc), 3).view((self.B,) + self.shape)
        else:
            xin = xin.view((self.B,) + self.shape)
        F = self.model(xin, self.cnoise).float().contiguous()
        if self.Cc > 1 and self.S > 1:
            F = ops.nchw_to_cl(F.view(self.B, self.Cc, 1, 1, self.S), torch.float32, 3)
        self._F = F

    def _stage(self, stage: int):
        by = self._blend_y if self._blend_on else None
        check(lib.dsk_sampler_stage_blend(stage, ptr(self.x), ptr(self.x_aux), ptr(self.r1), ptr(self._F), ptr(self.xin),
                                          ptr(self.cnoise), ptr(self._tab), ptr(self.row), ptr(self._noise),
                                          C.c_uint64(0), ptr(self._hist), self.B, self.Cc,
                                          self.S, self.sigma_data, self.sigma_max, self.kind, dt_code(self.act_dtype),
                                          self.xin_ld, 1 if self.cfg else 0,
                                          float(self.guidance) if self.cfg else 1.0, ptr(by),
                                          ptr(self._blend_mask) if by is not None else None,
                                          self._blend_mask.numel() if by is not None else 0,
                                          self._blend_y.shape[0] if by is not None else 0, stream()))

    def set_condition(self, ychan: Optional[torch.Tensor], ye: Optional[torch.Tensor]):
        """Write the run's conditioning into the static buffers the captured graphs read: channel conditioning
        [B, Cy, *S] into channels [C, C+Cy) of the network-input rows, the vector [B, M] into the (conditional half of
        the) time-embedding addend.  Once per run -- the reference recomputes both at every network evaluation."""
        if (ychan is None) != (self.cond_channels == 0) or (ye is None) != (not self.cond_vector):
            raise ValueError("SamplerEngine.set_condition: conditioning does not match the engine")
        with torch.inference_mode(False), torch.no_grad():
            if ychan is not None:
                src = ychan.detach().reshape(self.B, self.cond_channels, self.S).transpose(1, 2)
                self.xin.view(self.Bn, self.S, self.xin_ld)[:self.B, :, self.Cc:self.Cc + self.cond_channels].copy_(src)
            if ye is not None:
                self.ye[self.Bn - self.B:].copy_(ye.detach())

    def _step(self, stages):
        for st in stages:
            self._net()
            self._stage(st)
        check(lib.dsk_sampler_advance(ptr(self.row), stream()))

    def _graph_for(self, stages, key):
        """Capture one step; buffers, table pointer, noise/history pointers are all static."""
        g = self._graphs.get(key)
        if g is None:
            self.plan.prepare()
            self._stage(STAGE_INIT)
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):      # warm-up outside capture (lazy module loads, packed weights)
                self._step(stages)
            torch.cuda.current_stream(self.device).wait_stream(side)
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            c0 = launch_count()
            with torch.cuda.graph(g):
                self._step(stages)
            self._graph_launches[key] = launch_count() - c0     # kernels recorded into this graph
            self.row.zero_()
            self._graphs[key] = g
        return g

    def graph_launches_per_run(self) -> int:
        """Kernel launches replayed by the graphs of the most recent run()."""
        return self._last_graph_launches

    # ------------------------------------------------------------------ run
    def run(self, white_noise: torch.Tensor, table: torch.Tensor, program: str, record_history: bool = False,
            noises: Optional[torch.Tensor] = None, seed: int = 0, first: int = 0, last: Optional[int] = None,
            blend=None) -> torch.Tensor:
        """white_noise: fp32 [B, *shape] on the device; table: CPU fp32 [nsteps+1, 8] (Scheduler.step_table).
        first / last: run steps first .. last-1 of the table only, from the state `white_noise` at level t[first] -- scaled by
        self.sigma_max, so callers of a partial run set it to 1 (Scheduler.propagate_partial, schedulers.py:177-217); the
        history then has last - first + 1 rows.  blend = (y, mask): Scheduler.inpaint (schedulers.py:91-122) -- y fp32
        [nsteps + 1, B, *shape] the forward history of the known data, mask broadcastable to the state; the blend is fused
        into the step-completing stage kernels, so an inpainting run replays the same captured graphs."""
        require_cuda(white_noise, "white noise")
        if program not in self.programs:
            raise ValueError(f"Unknown integrator program: {program}")
        with torch.inference_mode(False), torch.no_grad():
            return self._run(white_noise, table, program, record_history, noises, seed, first, last, blend)

    def _run(self, white_noise, table, program, record_history, noises, seed, first=0, last=None, blend=None):
        nsteps = table.shape[0] - 1
        last = nsteps if last is None else int(last)
        first = int(first)
        if not (0 <= first < last <= nsteps):
            raise ValueError(f"steps [{first}, {last}) outside the schedule of {nsteps} steps")
        regular, final = self._program(program, table)
        N = self.B * self.Cc * self.S
        # static-address inputs of the captured graphs: refill in place, re-capture only on shape change
        if self._tab is None or self._tab.shape != table.shape:
            self._tab = torch.empty_like(table, device=self.device)
            self._graphs.clear()
        self._tab.copy_(table, non_blocking=True)
        want_hist = (nsteps + 1, N) if record_history else None
        if (self._hist is None) != (want_hist is None) or (want_hist and tuple(self._hist.shape) != want_hist):
            self._hist = torch.empty(want_hist, dtype=torch.float32, device=self.device) if want_hist else None
            self._graphs.clear()
        want_noise = None if noises is None else (nsteps + 1, N)
        if (self._noise is None) != (want_noise is None) or (want_noise and tuple(self._noise.shape) != want_noise):
            self._noise = torch.zeros(want_noise, dtype=torch.float32, device=self.device) if want_noise else None
            self._graphs.clear()
        if noises is not None:
            self._noise[:nsteps].copy_(noises.reshape(nsteps, N))
        if blend is not None:
            if type(self) is not SamplerEngine:
                raise NotImplementedError("inpainting on the graph engine: EDM stages only")
            y, mask = blend
            if tuple(y.shape) != (nsteps + 1, self.B) + self.shape:
                raise ValueError(f"inpainting history has shape {tuple(y.shape)}, expected {(nsteps + 1, self.B) + self.shape}")
            m = mask.to(self.x)
            if not (m.ndim <= self.x.ndim and tuple(self.x.shape[self.x.ndim - m.ndim:]) == tuple(m.shape)):
                m = m.expand_as(self.x)              # general broadcasting (size-1 batch / channel dimensions)
            m = m.contiguous().reshape(-1)
            if self._blend_y is None or tuple(self._blend_y.shape) != (nsteps + 1, N) or self._blend_mask.numel() != m.numel():
                self._blend_y = torch.empty((nsteps + 1, N), dtype=torch.float32, device=self.device)
                self._blend_mask = torch.empty_like(m)
                self._graphs.clear()
            self._blend_y.copy_(y.reshape(nsteps + 1, N))
            self._blend_mask.copy_(m)
        if self._blend_on != (blend is not None):
            self._blend_on = blend is not None
            self._graphs.clear()                     # the blend pointers are baked into the captured stage launches
        self.seed = int(seed)
        if self.native:       # captured graphs read the PACKED weight copies: refresh them (in place, same addresses) when the
            self.plan.prepare()   # parameters changed since the last run (optimizer steps, EMA apply_to, load_state_dict)
        if self.use_graphs:   # capture (with a throw-away warm-up step on zero state) before the real run starts
            self.x.zero_()
            self.row.zero_()
            g_reg = self._graph_for(regular, (program, "regular"))
            g_fin = self._graph_for(final, (program, "final"))
        self.x.copy_(white_noise.reshape(self.x.shape))
        sd = self.seed & 0xFFFFFFFFFFFFFFFF
        as_i32 = lambda v: v - (1 << 32) if v >= (1 << 31) else v  # noqa: E731
        self.row.copy_(torch.tensor([first, as_i32(sd & 0xFFFFFFFF), as_i32(sd >> 32), 0], dtype=torch.int32))
        self.nfe = 0
        self._stage(STAGE_INIT)
        n_reg = (last - first - 1) if last == nsteps else (last - first)     # the table's final step has its own stage program
        n_fin = 1 if last == nsteps else 0
        if self.use_graphs:
            for _ in range(n_reg):
                g_reg.replay()
            if n_fin:
                g_fin.replay()
            self.nfe = n_reg * len(regular) + n_fin * len(final)
            self._last_graph_launches = (n_reg * self._graph_launches[(program, "regular")] +
                                         n_fin * self._graph_launches[(program, "final")])
        else:
            for _ in range(n_reg):
                self._step(regular)
            if n_fin:
                self._step(final)
        if record_history:
            return self._hist.view((nsteps + 1, self.B) + self.shape)[first:last + 1].clone()
        return self.x.clone()


# ---------------------------------------------------------------------------------------------------------------------
GSTAGE_INIT, GSTAGE_STEP1, GSTAGE_HEUN_MID, GSTAGE_HEUN_FIN = 0, 1, 2, 3      # include/diffsci_b200.h: dsk_gstage
GENERAL_PROGRAMS = {
    "euler": ((GSTAGE_STEP1,), (GSTAGE_STEP1,)),
    "heun": ((GSTAGE_HEUN_MID, GSTAGE_HEUN_FIN), (GSTAGE_STEP1,)),       # final: only when the last step ends at t = 0
    "euler-maruyama": ((GSTAGE_STEP1,), (GSTAGE_STEP1,)),
}


class GeneralSamplerEngine(SamplerEngine):
    """The same captured-graph loop for ANY scheduler / preconditioner pair (VP / VE / SR3 / custom): the stages
    of csrc/sampler_general.cu read rhs = P x + Q F and the network-input scale / c_noise of every evaluation point from
    Scheduler.general_step_table, so nothing family-specific is left in the kernel.

    The table and the stage program are pinned on the CPU against the oracle (tests/test_host_logic.py:
    test_general_step_table_program_equals_the_oracle) and on the GPU against the Integrator.step seam and the fp64 budget
    (tests/test_gpu_precond.py: test_general_engine_equals_the_step_seam)."""
    programs = GENERAL_PROGRAMS

    def __init__(self, model, B, shape, device, use_graphs: bool = True):
        super().__init__(model, B, shape, device, 0.5, 1.0, 0, use_graphs=use_graphs)
        if not self.native:
            raise NotImplementedError("GeneralSamplerEngine drives the native networks")

    def _program(self, program: str, table: torch.Tensor):
        regular, final = GENERAL_PROGRAMS[program]
        if program == "heun" and float(table[table.shape[0] - 2, 9]) != 0.0:     # G_HAS2 of the last step
            final = regular
        return regular, final

    def _stage(self, stage: int):
        gstage = GSTAGE_INIT if stage == STAGE_INIT else stage     # _run's initial stage constant is the EDM engine's (== 0)
        check(lib.dsk_sampler_stage_general(gstage, ptr(self.x), ptr(self.x_aux), ptr(self.r1), ptr(self._F), ptr(self.xin),
                                            ptr(self.cnoise), ptr(self._tab), ptr(self.row), ptr(self._noise), ptr(self._hist),
                                            self.B, self.Cc, self.S, self.sigma_max, dt_code(self.act_dtype), self.xin_ld,
                                            stream()))
