#!/usr/bin/env python
"""bench.py -- headline benchmark of the Karras/EDM hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload c4|c2|c5|c1]

A "step" is ONE full sampling pass of the hot path over one batch of synthetic white noise
(workload c4: PUNetG-3D, 1x64^3 volumes, EDM Heun 64 steps = 127 denoiser evaluations per sample).
`value` = samples/s over all ranks with x_T already resident in HBM; `e2e` = the same metric through
the public API (KarrasModule.propagate_white_noise) with pinned HOST buffers, H2D and D2H inside the
timed region.  Inputs per step (B x 1 MiB fp32 volumes) are tiny next to the >100 GB of activation
traffic of one step, so L2 (126 MB) holds no useful state between steps ("inputs larger than L2").
One rank per GPU; no data-path collective (samples are independent): scaling = weak.
"""
from __future__ import annotations

import argparse
import contextlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (net kind, config kwargs, sample shape, nsteps, integrator, default per-GPU batch)
    "c4": ("punetg", dict(dimension=3), (1, 64, 64, 64), 64, "heun", 16),     # batch sweep on B200: 8 -> 7.09, 16 -> 7.26 samples/s
    "c2": ("punetg", dict(dimension=2, model_channels=128), (1, 28, 28), 40, "heun", 256),
    "c5": ("punetg", dict(dimension=2), (1, 256, 256), 256, "euler-maruyama", 32),   # sweep: 8 -> 11.4, 16 -> 12.5, 32 -> 12.9, 64 -> 13.0
    "c1": ("mlp", dict(dim=2, hidden_dims=[128, 128, 128]), (2,), 18, "heun", 65536),
}
# training workloads: (net kind, config kwargs, sample shape, loss metric, per-GPU batch, forward GFLOP/sample or None)
TRAIN_WORKLOADS = {
    "c2train": ("punetg", dict(dimension=2, model_channels=128), (1, 28, 28), "huber", 256, None),
    "c3train": ("adm", dict(input_channels=3, output_channels=3), (3, 128, 128), "huber", 32, 59.44),   # SURVEY 8a row a9 probe
    "c4train": ("punetg", dict(dimension=3), (1, 64, 64, 64), "huber", 2, None),
}
TRAIN_NAMES = {"c2train": "PUNetG-2D(mc=128) 1x28x28 EDM training iteration, batch 256/GPU (BASELINE configs[1])",
               "c3train": "ADM-2D(mc=64,[2,4]) 3x128x128 EDM training iteration + EMA(0.999), batch 32/GPU, data-parallel (BASELINE configs[2])",
               "c4train": "PUNetG-3D(mc=64,[2,4]) 1x64^3 EDM training iteration, batch 2/GPU"}
NAMES = {"c4": "PUNetG-3D(mc=64,[2,4]) 1x64^3 EDM Heun-64 sampling (BASELINE configs[3])",
         "c2": "PUNetG-2D(mc=128) 1x28x28 EDM Heun-40 sampling (BASELINE configs[1])",
         "c5": "PUNetG-2D(mc=64) 1x256x256 Euler-Maruyama-256 sampling (BASELINE configs[4])",
         "c1": "MLPUncond(2,[128]*3,SiLU) toy Heun-18 sampling (BASELINE configs[0])"}


@contextlib.contextmanager
def stdout_to_stderr():
    """stdout carries exactly ONE JSON line.  NCCL writes "NCCL version ..." to the process's stdout (fd 1) when the
    communicator is created under NCCL_DEBUG=VERSION / INFO (the GPU boxes export VERSION), so fd 1 points at stderr while the
    process group comes up."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        yield
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)


def nfe_per_sample(nsteps, integrator):
    return 2 * nsteps - 1 if integrator in ("heun", "karras") else nsteps


def punetg_conv_flops(cfg, spatial):
    """Algorithmic forward FLOPs of one network evaluation per sample (conv + attention), SURVEY 8d."""
    nd = cfg.dimension
    M = cfg.model_channels
    mult = cfg.extended_channel_expansion
    k = cfg.kernel_size ** nd
    S = [1]
    for s in spatial:
        S[0] *= s
    for _ in cfg.channel_expansion:
        S.append(S[-1] // (2 ** nd))
    fl = 2 * S[0] * k * (cfg.input_channels * M + M * cfg.output_channels)
    nl = len(cfg.channel_expansion)
    for l in range(nl):
        c, cn = mult[l] * M, mult[l + 1] * M
        fl += (cfg.number_resnet_downward_block + cfg.number_resnet_upward_block) * 2 * (2 * S[l] * k * c * c)
        fl += 2 * S[l + 1] * k * c * cn + 2 * S[l] * k * cn * c          # down conv, up conv
    cb = mult[-1] * M
    nb = cfg.number_resnet_before_attn_block + cfg.number_resnet_attn_block + cfg.number_resnet_after_attn_block
    fl += nb * 2 * (2 * S[nl] * k * cb * cb)
    na = cfg.number_resnet_attn_block - 1
    fl += na * (8 * S[nl] * cb * cb + 4 * S[nl] * S[nl] * cb)
    return fl


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 7:
                    self.rows.append(parts)
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def summary(self):
        self._stop_evt.set()
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        reasons = [n for i, n in ((3, "hw_slowdown"), (4, "hw_thermal_slowdown"), (5, "sw_thermal_slowdown"),
                                  (6, "sw_power_cap")) if any(r[i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "power_w_max": max(float(r[2]) for r in self.rows), "samples": len(self.rows)}


KARRAS_CONFIG = ["edm"]      # --karras-config (sweeps of the VP / VE route; the headline workloads are EDM)


def build_workload(name, device, precision):
    import torch
    import diffsci_b200 as d
    kind, kw, shape, nsteps, integ, batch = WORKLOADS[name]
    torch.manual_seed(0)
    if kind == "punetg":
        if os.environ.get("DSK_BENCH_CIRCULAR"):     # periodic variant of the same network, for A/B runs
            kw = dict(kw, convolution_type="circular")
        cfg = d.PUNetGConfig(**kw)
        net = d.PUNetG(cfg, precision=precision)
        flops = punetg_conv_flops(cfg, shape[1:])
    else:
        cfg = None
        net = d.MLPUncond(kw["dim"], kw["hidden_dims"], torch.nn.SiLU())
        flops = 2 * sum(p.numel() for n, p in net.named_parameters() if n.endswith("weight"))
    net = net.to(device).eval() if device is not None else net.eval()
    kcfg = {"edm": d.KarrasModuleConfig.from_edm, "vp": d.KarrasModuleConfig.from_vp, "ve": d.KarrasModuleConfig.from_ve}[KARRAS_CONFIG[0]]
    module = d.KarrasModule(net, kcfg())
    return module, net, cfg, shape, nsteps, integ, batch, flops


def live_reference_arm(name, steps, warmup, max_seconds=25.0):
    """The UNMODIFIED reference -- its own PUNetG / MLPUncond modules (default init under torch.manual_seed(0)) inside its own
    KarrasModule.propagate_white_noise (karras/karrasmodule.py:867-931) -- on the host cores, imported from oracle/_ref (a copy
    of the reference's package, oracle/build_ref.py) or /root/reference.  Bounded sample: B = 1, two integrator steps,
    extrapolated linearly in NFE.  Returns None when the reference is not importable (then the oracle port is timed).
    No diffsci_b200 import, no .so load in this arm."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    try:
        import refload
        if refload.reference_root() is None:
            return None
        with stdout_to_stderr():
            refload.load_reference()
            import diffsci.models as M
            from diffsci.models.nets.mlp import MLPUncond
            from diffsci.models.nets.punetg import PUNetG
            from diffsci.models.nets.punetg_config import PUNetGConfig
    except Exception as e:                      # a missing optional dependency of the reference on this box
        print(f"bench.py: live reference unavailable ({type(e).__name__}: {e}); timing the oracle port", file=sys.stderr)
        return None
    kind, kw, shape, nsteps, integ, _ = WORKLOADS[name]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    net = PUNetG(PUNetGConfig(**kw)) if kind == "punetg" else MLPUncond(kw["dim"], kw["hidden_dims"], nonlinearity=torch.nn.SiLU())
    mod = M.KarrasModule(net.eval(), M.KarrasModuleConfig.from_edm())
    mod.eval()
    B, sub = (1, 2) if kind == "punetg" else (4096, nsteps)
    if name == "c2":
        B = 8
    torch.manual_seed(1234)
    wn = torch.randn(B, *shape)
    nfe_sub = nfe_per_sample(sub, integ)
    vals = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        mod.propagate_white_noise(wn, nsteps=sub, integrator=integ)
        dt = time.perf_counter() - t0
        if i >= warmup:
            vals.append(dt)
        if sum(vals) > max_seconds and vals:
            break
    per_nfe = (sum(vals) / len(vals)) / (nfe_sub * B)
    value = 1.0 / (per_nfe * nfe_per_sample(nsteps, integ))
    sample = (f"the reference itself (diffsci.models.KarrasModule.propagate_white_noise, torch {torch.__version__} CPU fp32, "
              f"{cores} threads), B={B}, {sub} {integ} steps = {nfe_sub} NFE timed {len(vals)}x, extrapolated linearly to "
              f"{nfe_per_sample(nsteps, integ)} NFE")
    return value, cores, sample, sum(vals) / len(vals), "reference", len(vals)


def cpu_reference_arm(name, steps, warmup, max_seconds=25.0, ref_device="cpu", ref_mode="fp32", ref_batch=0):
    """The oracle port (torch-CPU restatement of the reference, oracle/*.py) timed on the host cores on a
    BOUNDED sample of the workload: B=1, a few integrator steps, extrapolated linearly in NFE.

    ref_device="cuda" (never the default, never what the driver runs) times the SAME port on torch's own CUDA path
    (cuDNN / cuBLAS; ref_mode fp32 = TF32 off, tf32 = TF32 on, bf16 = autocast) -- the "reference on the box's
    PyTorch-CUDA path" SURVEY 8(d) asks for beside the CPU number."""
    import torch
    if ref_device == "cpu" and not os.environ.get("DSK_BENCH_PORT_ONLY"):
        live = live_reference_arm(name, steps, warmup, max_seconds)
        if live is not None:
            return live
    from oracle import karras_oracle as K, nets_oracle as N
    kind, kw, shape, nsteps, integ, _ = WORKLOADS[name]
    module, net, cfg, *_ = build_workload(name, None, "fp32")
    sd = {k: v.detach() for k, v in net.state_dict().items()}
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    if ref_device == "cuda":
        return cuda_port_arm(name, steps, warmup, sd, cfg, ref_mode, ref_batch, max_seconds)
    if kind == "punetg":
        fn = lambda x, t: N.punetg_forward(sd, cfg, x, t)  # noqa: E731
        B, sub = 1, 2
    else:
        fn = lambda x, t: N.mlp_uncond_forward(sd, x, t, "silu")  # noqa: E731
        B, sub = 4096, nsteps
    if name == "c2":
        B = 8
    torch.manual_seed(1234)
    wn = torch.randn(B, *shape)
    noises = [torch.randn(B, *shape) for _ in range(sub)]
    nfe_sub = nfe_per_sample(sub, integ)
    vals = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            K.sample_from_white_noise(fn, wn, sub, integ, noises=noises)
            dt = time.perf_counter() - t0
            if i >= warmup:
                vals.append(dt)
            if sum(vals) > max_seconds and len(vals) >= 1:
                break
    per_nfe = (sum(vals) / len(vals)) / (nfe_sub * B)                # seconds per sample-NFE
    value = 1.0 / (per_nfe * nfe_per_sample(nsteps, integ))          # samples/s, extrapolated linearly in NFE
    sample = (f"oracle port of the reference (torch {torch.__version__} CPU fp32), B={B}, {sub} {integ} steps = {nfe_sub} NFE "
              f"timed {len(vals)}x, extrapolated linearly to {nfe_per_sample(nsteps, integ)} NFE")
    return value, cores, sample, sum(vals) / len(vals), "port", len(vals)


def cuda_port_arm(name, steps, warmup, sd, cfg, mode, batch, max_seconds):
    """Oracle port on torch CUDA (library kernels: cuDNN convolutions, cuBLAS / SDPA attention).  Side measurement only."""
    import contextlib
    import torch
    from oracle import karras_oracle as K, nets_oracle as N
    kind, kw, shape, nsteps, integ, wl_batch = WORKLOADS[name]
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.allow_tf32 = mode != "fp32"
    torch.backends.cuda.matmul.allow_tf32 = mode != "fp32"
    torch.backends.cudnn.benchmark = True
    sd = {k: v.to(dev) for k, v in sd.items()}
    B = batch or wl_batch
    sub = 2 if kind == "punetg" else nsteps
    if kind == "punetg":
        fn = lambda x, t: N.punetg_forward(sd, cfg, x, t).float()  # noqa: E731
    else:
        fn = lambda x, t: N.mlp_uncond_forward(sd, x, t, "silu")  # noqa: E731
    ctx = (lambda: torch.autocast("cuda", dtype=torch.bfloat16)) if mode == "bf16" else contextlib.nullcontext
    torch.manual_seed(1234)
    wn = torch.randn(B, *shape).to(dev)
    noises = [torch.randn(B, *shape).to(dev) for _ in range(sub)]
    nfe_sub = nfe_per_sample(sub, integ)
    vals = []
    with torch.no_grad(), ctx():
        for i in range(max(3, warmup) + steps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            K.sample_from_white_noise(fn, wn, sub, integ, noises=noises)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if i >= max(3, warmup):
                vals.append(dt)
            if sum(vals) > max_seconds and vals:
                break
    per_nfe = (sum(vals) / len(vals)) / (nfe_sub * B)
    value = 1.0 / (per_nfe * nfe_per_sample(nsteps, integ))
    sample = (f"oracle port of the reference on torch {torch.__version__} CUDA ({mode}: cuDNN {torch.backends.cudnn.version()}, "
              f"benchmark=True), B={B}, {sub} {integ} steps = {nfe_sub} NFE timed {len(vals)}x, extrapolated linearly to "
              f"{nfe_per_sample(nsteps, integ)} NFE; {per_nfe * B * 1e3:.2f} ms per batched evaluation")
    return value, 0, sample, sum(vals) / len(vals), "port-on-torch-cuda", len(vals)


def build_train_workload(name, device, precision):
    import torch
    import diffsci_b200 as d
    kind, kw, shape, metric, batch, gf = TRAIN_WORKLOADS[name]
    torch.manual_seed(0)
    if kind == "punetg":
        if os.environ.get("DSK_BENCH_CIRCULAR"):     # periodic variant of the same network, for A/B runs
            kw = dict(kw, convolution_type="circular")
        cfg = d.PUNetGConfig(**kw)
        net = d.PUNetG(cfg, precision=precision)
        flops = punetg_conv_flops(cfg, shape[1:])
    else:
        cfg = d.ADMConfig(**kw)
        net = d.ADM(cfg, precision=precision)
        flops = gf * 1e9
    if device is not None:
        net = net.to(device)
    net.train()
    module = d.KarrasModule(net, d.KarrasModuleConfig.from_edm(loss_metric=metric)).train()
    return module, net, cfg, shape, metric, batch, flops


def cpu_train_arm(name, steps, warmup, max_seconds=25.0):
    """Oracle port of one training iteration (forward, EDM loss, autograd backward, AdamW, EMA lerp) on the host
    cores, on a reduced batch; it/s scaled linearly to the workload's per-GPU batch."""
    import torch
    from oracle import karras_oracle as K, nets_oracle as N
    kind, kw, shape, metric, batch, _ = TRAIN_WORKLOADS[name]
    module, net, cfg, *_ = build_train_workload(name, None, "fp32")
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    names = [k for k, _ in net.named_parameters()]
    fwd = N.punetg_forward if kind == "punetg" else N.adm_forward
    Bs = {"c2train": 16, "c3train": 2, "c4train": 1}[name]
    m = {k: torch.zeros_like(sd[k]) for k in names}
    v = {k: torch.zeros_like(sd[k]) for k in names}
    sh = {k: sd[k].clone() for k in names}
    torch.manual_seed(7)
    vals = []
    for i in range(warmup + steps):
        x = torch.randn(Bs, *shape) * 0.5
        sigma = torch.exp(torch.randn(Bs) * 1.2 - 1.2)
        noise = torch.randn(Bs, *shape)
        t0 = time.perf_counter()
        leaves = {k: sd[k].clone().requires_grad_(True) for k in names}
        full = dict(sd, **leaves)
        L = K.edm_loss(lambda xx, tt: fwd(full, cfg, xx, tt), x, sigma, noise, loss_metric=metric)
        L.backward()
        with torch.no_grad():
            for k in names:
                sd[k], m[k], v[k] = K.adamw_step(sd[k], leaves[k].grad, m[k], v[k], i + 1)
                sh[k] = K.ema_update(sh[k], sd[k], 0.999)
        dt = time.perf_counter() - t0
        if i >= warmup:
            vals.append(dt)
        if sum(vals) > max_seconds and vals:
            break
    per_sample = (sum(vals) / len(vals)) / Bs
    value = 1.0 / (per_sample * batch)
    sample = (f"oracle port (torch {torch.__version__} CPU fp32 autograd + AdamW + EMA), batch {Bs} timed {len(vals)}x, scaled "
              f"linearly to batch {batch}")
    return value, cores, sample, sum(vals) / len(vals)


def train_arm(args, rank, world, local_rank):
    """`--workload c2train|c3train|c4train`: one step = one EDM training iteration through EDMTrainer.step (noising,
    forward, fused loss, backward, bucketed gradient all-reduce over NCCL for N>1, fused AdamW + EMA)."""
    kind, kw, shape, metric, batch, _ = TRAIN_WORKLOADS[args.workload]
    B = args.batch or batch
    if args.impl == "reference":
        if rank != 0:
            return
        value, cores, sample, secs = cpu_train_arm(args.workload, max(1, args.steps), min(1, args.warmup))
        line = {"impl": "reference", "metric": "EDM train it/s", "value": value, "unit": "it/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": TRAIN_NAMES[args.workload], "per_gpu_batch": batch},
                "cpu_baseline": {"value": value, "unit": "it/s", "cores": cores, "kind": "port", "sample": sample},
                "e2e": {"value": value, "unit": "it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return
    import torch
    import torch.distributed as dist
    import diffsci_b200 as d
    from diffsci_b200 import _lib

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        with stdout_to_stderr():        # eager communicator creation (device_id) + the first collective
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
    precision = args.precision
    if precision == "auto":
        precision = "bf16" if d.TC_CONV_ENABLED else "fp32"
    module, net, cfg, shape, metric, _, flops = build_train_workload(args.workload, dev, precision)
    ema = d.ModelEMA(net, ema_type="traditional", decay=0.999)
    tr = d.EDMTrainer(module, ema=ema)
    nparams = sum(p.numel() for p in net.parameters())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    torch.manual_seed(2 + rank)
    nb = 4                                                   # distinct synthetic batches, rotated
    x_host = [(torch.randn(B, *shape) * 0.5).pin_memory() for _ in range(nb)]
    x_dev = [x.to(dev) for x in x_host]
    losses = []
    for i in range(args.warmup):
        losses.append(tr.step(x_dev[i % nb]))
    barrier()
    clocks = ClockSampler(local_rank) if rank == 0 else None
    if clocks:
        clocks.start()
    n0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(args.steps):
        loss = tr.step(x_dev[i % nb])
    ev1.record()
    barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms)
    clk = clocks.summary() if clocks else None
    launches = _lib.launch_count() - n0
    value = args.steps / (total_ms / 1e3)                    # iterations/s of the whole data-parallel job
    # e2e: pinned host batch -> device, step, loss read back on the host, every iteration
    barrier()
    t0 = time.perf_counter()
    last = 0.0
    for i in range(args.steps):
        xb = x_host[i % nb].to(dev, non_blocking=True)
        last = float(tr.step(xb))
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    import math
    assert math.isfinite(last), "training loss is not finite"
    N_el = B
    for s_ in shape:
        N_el *= s_
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("bf16_tflops_sustained") or peaks.get("bf16_tflops") or 1590.0
    tfl = 3.0 * flops * B * world * value / 1e12
    line = {"metric": "EDM train it/s", "value": value, "unit": "it/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if precision == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": TRAIN_NAMES[args.workload], "per_gpu_batch": B, "global_batch": world * B,
                       "loss": metric, "optimizer": "AdamW(1e-3,(0.9,0.999),wd=1e-4) fused with EMA(0.999)",
                       "parameters": nparams, "precision": precision,
                       "parallelism": f"data-parallel x{world}, bucketed NCCL all-reduce of the flat fp32 gradient",
                       "l2_policy": "4 rotating input batches; activations per iteration exceed L2"},
            "samples_per_s": value * B * world, "model_tflops": tfl,
            "e2e": {"value": args.steps / float(e2e_s), "unit": "it/s", "h2d_bytes_per_step": N_el * 4, "d2h_bytes_per_step": 4},
            "gpu_launches": int(launches), "clocks": clk, "final_loss": last,
            "roofline": {"bound": "tensor", "kernel": "whole iteration (fwd + dgrad + wgrad = 3x forward FLOPs)", "achieved": tfl / world,
                         "peak": peak, "unit": "TFLOP/s", "frac": tfl / world / peak, "traffic": None,
                         "peak_source": "MEASURED_PEAKS.json (sustained)" if peaks else "fallback (B200_PROFILING.md)"}}
    if not args.no_cpu_baseline and world == 1:
        v, cores, sample, _ = cpu_train_arm(args.workload, 1, 1)
        line["cpu_baseline"] = {"value": v, "unit": "it/s", "cores": cores, "kind": "port", "sample": sample}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("DSK_BENCH_WORKLOAD", "c4"), choices=list(WORKLOADS) + list(TRAIN_WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (0 = workload default)")
    ap.add_argument("--nsteps", type=int, default=0, help="integrator steps (0 = workload default)")
    ap.add_argument("--karras-config", default="edm", choices=["edm", "vp", "ve"],
                    help="KarrasModuleConfig.from_edm / from_vp / from_ve (vp / ve: the table-driven general engine; "
                         "DSK_GENERAL_ENGINE=0 times the Integrator.step seam instead)")
    ap.add_argument("--integrator", default="", choices=["", "euler", "heun", "euler-maruyama", "karras"],
                    help="override the workload's integrator (sweeps; e.g. the Karras-churn variant of c5)")
    ap.add_argument("--precision", default=os.environ.get("DSK_BENCH_PRECISION", "auto"))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the `train` object (c3train it/s in the same run)")
    ap.add_argument("--no-modes", action="store_true", help="skip `precision_modes` / `parity` (other precision modes + checks)")
    ap.add_argument("--no-library-baseline", action="store_true", help="skip `gpu_library_baseline` (oracle port on torch CUDA)")
    ap.add_argument("--ref-device", default="cpu", choices=["cpu", "cuda"],
                    help="--impl reference only: cuda = the oracle port on torch's cuDNN/cuBLAS path (side measurement)")
    ap.add_argument("--ref-mode", default="fp32", choices=["fp32", "tf32", "bf16"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload in TRAIN_WORKLOADS:
        return train_arm(args, rank, world, local_rank)
    KARRAS_CONFIG[0] = args.karras_config
    if args.integrator:
        WORKLOADS[args.workload] = WORKLOADS[args.workload][:4] + (args.integrator,) + WORKLOADS[args.workload][5:]
    kind, kw, shape, nsteps, integ, batch = WORKLOADS[args.workload]
    nsteps = args.nsteps or nsteps
    B = args.batch or batch
    nfe = nfe_per_sample(nsteps, integ)

    if args.impl == "reference":
        if rank != 0:
            return
        value, cores, sample, secs, rkind, reps = cpu_reference_arm(args.workload, max(1, args.steps), min(1, args.warmup),
                                                                    ref_device=args.ref_device, ref_mode=args.ref_mode,
                                                                    ref_batch=args.batch)
        # `steps` = the repetitions actually timed (the arm stops after ~25 s of CPU work); ms_per_step = their mean
        line = {"impl": "reference", "metric": "EDM Heun samples/sec", "value": value, "unit": "samples/s",
                "n_gpus": args.gpus, "steps": reps, "steps_requested": args.steps, "warmup": min(1, args.warmup),
                "ms_per_step": secs * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": NAMES[args.workload], "nsteps": nsteps, "nfe_per_sample": nfe},
                "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": rkind, "sample": sample},
                "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "nfe_per_s": value * nfe}
        print(json.dumps(line), flush=True)
        return

    import torch
    import torch.distributed as dist
    import diffsci_b200 as d
    from diffsci_b200 import _lib, ops

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        with stdout_to_stderr():        # eager communicator creation (device_id) + the first collective
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
    precision = args.precision
    auto = precision == "auto"
    if auto:
        # the fastest mode that meets north_star's 16-bit tolerance (denoiser within 1e-3 of the reference's fp32 result) on THIS
        # network, checked against the CPU oracle before anything is timed (AUTO_MODES, fastest first, each has to stay under
        # 8e-4): plain fp16 operands with fp32 storage ("fp16s32"), activations split hi + lo (2 tcgen05 MMAs per k-step) in the
        # >= 128-channel layers ("fp16x2m"), or everywhere ("fp16x2").  The other modes are measured in the same run
        # (`precision_modes`).
        precision = "fp16x2" if d.TC_CONV_ENABLED else "fp32_ffma"
    module, net, cfg, shape, _, integ, _, flops_per_nfe = build_workload(args.workload, dev, precision)
    precheck = None
    if auto and kind == "punetg" and d.TC_CONV_ENABLED:
        choice = [precision]
        if rank == 0:
            precheck = denoiser_check(module, net, cfg, shape, dev, AUTO_MODES)
            choice = [next((m for m in AUTO_MODES[:-1] if precheck["max_rel"][m] <= AUTO_MAX_REL), AUTO_MODES[-1])]
        if world > 1:
            dist.broadcast_object_list(choice, src=0)
        precision = net.precision = choice[0]
    table_integrator = d.name_to_integrator(integ)
    N_el = B
    for s in shape:
        N_el *= s
    from diffsci_b200 import distributed as dsk_dist

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def one_pass(wn):
        """One step: the local shard through the public API, then the path's ONE collective -- the final all_gather of the
        samples (distributed.gather_samples) -- so that every rank holds the global batch."""
        local = module.propagate_white_noise(wn, nsteps=nsteps, integrator=table_integrator)
        return dsk_dist.gather_samples(local, world * B)

    # ---------------------------------------------------------------- resident-input arm (`value`)
    torch.manual_seed(1234 + rank)
    wn_host = torch.randn(B, *shape).pin_memory()
    wn_dev = wn_host.to(dev)
    for _ in range(args.warmup):
        one_pass(wn_dev)
    barrier()
    clocks = ClockSampler(local_rank) if rank == 0 else None
    if clocks:
        clocks.start()
    n0 = _lib.launch_count()
    eng_before = next(iter(module._engines.values()), None)      # None: the Integrator.step seam (no engine, no graphs)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        one_pass(wn_dev)
    ev1.record()
    barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms)
    clk = clocks.summary() if clocks else None
    eager_launches = _lib.launch_count() - n0
    graph_launches = eng_before.graph_launches_per_run() * args.steps if (eng_before is not None and eng_before.use_graphs) else 0
    value = world * B * args.steps / (total_ms / 1e3)

    # ---------------------------------------------------------------- end-to-end arm (`e2e`)
    out_host = torch.empty(world * B, *shape).pin_memory()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = one_pass(wn_host)                                   # H2D of the local white noise inside
        out_host.copy_(res, non_blocking=True)                    # D2H of the gathered samples inside
        torch.cuda.synchronize(dev)
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / float(e2e_s)

    # ---------------------------------------------------------------- training throughput in the same run (`train`)
    train = None
    if args.workload == "c4" and not args.no_train:
        train = train_measure(dev, rank, world, precision="bf16")

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------------------------------------------------------- roofline of the dominant kernel
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    roofline = None
    if kind == "punetg":
        roofline = dominant_conv_roofline(cfg, shape, B, precision, dev, peaks)

    line = {"metric": "EDM Heun samples/sec", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": DTYPE_NAME.get(precision, "f32"), "data": "synthetic",
            "config": {"workload": NAMES[args.workload], "per_gpu_batch": B, "global_batch": world * B,
                       "integrator": integ, "nsteps": nsteps, "nfe_per_sample": nfe, "precision": precision,
                       "karras_config": args.karras_config,
                       "parallelism": f"batch-sharded x{world} (distributed.py: no collective while integrating, ONE final "
                                      f"all_gather of the samples, inside the timed region)",
                       "l2_policy": "inputs larger than L2: >100 GB of activation traffic per step vs 126 MB L2"},
            "nfe_per_s": value * nfe,
            "model_tflops": value * nfe * flops_per_nfe / 1e12,
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": N_el * 4,
                    "d2h_bytes_per_step": world * N_el * 4},
            "gpu_launches": int(eager_launches + graph_launches),
            "clocks": clk, "roofline": roofline}
    if train is not None:
        line["train"] = train
    if world == 1 and kind == "punetg" and not args.no_modes:
        try:
            line["precision_modes"], line["parity"] = precision_modes_and_parity(args, module, net, cfg, shape, B, nsteps, nfe,
                                                                                 integ, precision, value, dev, precheck)
        except Exception as e:                     # the side measurements must never cost the headline line
            line["precision_modes_error"] = f"{type(e).__name__}: {e}"
        if not args.no_library_baseline:
            try:
                line["gpu_library_baseline"] = library_baseline(args.workload, net, cfg)
            except Exception as e:
                line["gpu_library_baseline_error"] = f"{type(e).__name__}: {e}"
    if not args.no_cpu_baseline and world == 1:
        v, cores, sample, _, rkind, _ = cpu_reference_arm(args.workload, 1, 1)
        line["cpu_baseline"] = {"value": v, "unit": "samples/s", "cores": cores, "kind": rkind, "sample": sample}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


AUTO_MODES = ("fp16s32", "fp16x2m", "fp16x2")      # --precision auto: the first whose denoiser max-rel is <= AUTO_MAX_REL, else the last
AUTO_MAX_REL = 8e-4
DTYPE_NAME = {"bf16": "bf16", "fp16": "f16", "fp16x2": "f16 (activations hi+lo, 2 MMAs per k-step; fp32 storage)",
              "fp16s32": "f16 (fp16 tensor-core operands, 1 MMA per k-step; fp32 accumulation and fp32 storage of the tensors between kernels -- except, in 3-D, the conv1 outputs of the full-resolution blocks, read once by the next norm, stored as fp16)",
              "fp16x2m": "f16 (activations hi+lo in the >=128-channel layers: 2 MMAs per k-step there, 1 elsewhere; fp32 storage)",
              "fp32": "f32 (split-f16 x3 on tcgen05)", "fp32_ffma": "f32"}


def train_measure(dev, rank, world, precision="bf16", workload="c3train", steps=10, warmup=3):
    """BASELINE.json's "train it/s" in the same run: ADM 3x128x128 EDM training iterations (configs[2]) through
    EDMTrainer.step on every rank -- noising, forward, fused loss, backward with the bucketed NCCL all-reduce of the flat fp32
    gradient (N > 1), fused AdamW + EMA -- timed on the device, max over ranks.  Weak scaling: batch 32 per GPU."""
    import torch
    import torch.distributed as dist
    import diffsci_b200 as d
    kind, kw, shape, metric, batch, _ = TRAIN_WORKLOADS[workload]
    module, net, cfg, shape, metric, _, flops = build_train_workload(workload, dev, precision)
    tr = d.EDMTrainer(module, ema=d.ModelEMA(net, ema_type="traditional", decay=0.999))
    torch.manual_seed(2 + rank)
    xs = [(torch.randn(batch, *shape) * 0.5).to(dev) for _ in range(4)]
    for i in range(warmup):
        tr.step(xs[i % 4])
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        loss = tr.step(xs[i % 4])
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    its = steps / (float(ms) / 1e3)
    out = {"metric": "EDM train it/s", "workload": TRAIN_NAMES[workload], "value": its, "unit": "it/s", "steps": steps,
           "warmup": warmup, "ms_per_step": float(ms) / steps, "per_gpu_batch": batch, "global_batch": batch * world,
           "samples_per_s": its * batch * world, "precision": precision, "final_loss": float(loss),
           "model_tflops": 3.0 * flops * batch * world * its / 1e12,
           "parallelism": f"data-parallel x{world}, bucketed NCCL all-reduce of the flat fp32 gradient overlapped with backward"}
    del tr, module, net
    torch.cuda.empty_cache()
    return out


def denoiser_check(module, net, cfg, shape, dev, modes):
    """D(x; sigma) of the given precision modes on ONE full-size sample against the CPU oracle's fp32 evaluation (= the
    reference's arithmetic; checker use of oracle/, outside every timed region)."""
    import types
    import torch
    from oracle import karras_oracle as K, nets_oracle as N
    sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    ocfg = types.SimpleNamespace(**cfg.export_description())
    torch.manual_seed(99)
    sigma = torch.tensor([1.0])
    x = torch.randn(1, *shape) * 1.5
    torch.set_num_threads(os.cpu_count() or 1)
    ref = K.denoiser(lambda xx, tt: N.punetg_forward(sd, ocfg, xx, tt), x, sigma)
    keep = net.precision
    out = {"x": x, "sigma": sigma, "ref": ref, "max_rel": {}, "rel_l2": {}}
    for m in modes:
        net.precision = m
        with torch.no_grad():
            D, _ = module.get_denoiser(x.to(dev), sigma.to(dev))
        e = D.cpu().double() - ref.double()
        out["max_rel"][m] = float(e.abs().max() / ref.double().abs().max())
        out["rel_l2"][m] = float(e.norm() / ref.double().norm())
    net.precision = keep
    module._engines.clear()
    net._plans.clear()
    torch.cuda.empty_cache()
    return out


def precision_modes_and_parity(args, module, net, cfg, shape, B, nsteps, nfe, integ, default_prec, default_value, dev,
                               precheck=None):
    """(1) samples/s of the same workload in the other precision modes (one warm pass + one timed pass each);
    (2) `parity`: the denoiser D(x; sigma) of every mode against the CPU oracle's fp32 evaluation (= the reference's
    arithmetic) on one full-size sample, and the sampled field of the benchmarked mode after the workload's real step count
    against the tensor-core fp32 mode on the same x_T, per pixel."""
    import types
    import torch
    import diffsci_b200 as d
    from oracle import karras_oracle as K, nets_oracle as N
    integrator = d.name_to_integrator(integ)
    modes = [{"precision": default_prec, "value": default_value, "unit": "samples/s", "timed_passes": args.steps}]
    torch.manual_seed(4321)
    wn = torch.randn(B, *shape).to(dev)
    for prec in ("bf16", "fp16", "fp16s32", "fp16x2m", "fp16x2", "fp32"):
        if prec == default_prec:
            continue
        net.precision = prec
        module.propagate_white_noise(wn, nsteps=nsteps, integrator=integrator)         # plan + graph capture + warm-up
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        module.propagate_white_noise(wn, nsteps=nsteps, integrator=integrator)
        e1.record()
        torch.cuda.synchronize(dev)
        modes.append({"precision": prec, "value": B / (e0.elapsed_time(e1) / 1e3), "unit": "samples/s", "timed_passes": 1})
        module._engines.clear()
        net._plans.clear()
        torch.cuda.empty_cache()
    # ---- parity: denoiser
    chk = denoiser_check(module, net, cfg, shape, dev, [m["precision"] for m in modes])
    den = chk["max_rel"]
    for m in modes:
        m["denoiser_max_rel"], m["denoiser_rel_l2"] = chk["max_rel"][m["precision"]], chk["rel_l2"][m["precision"]]
    # ---- parity: sampled field after the real step count, benchmarked mode vs tensor-core fp32 mode, same x_T, B = 1
    torch.manual_seed(77)
    w1 = torch.randn(1, *shape).to(dev)
    fields = {}
    for prec in ("fp32", default_prec):
        net.precision = prec
        it_ = d.name_to_integrator(integ)
        it_.reset_noise(seed=12345)
        fields[prec] = module.propagate_white_noise(w1, nsteps=nsteps, integrator=it_).double().cpu()
    net.precision = default_prec
    diff = (fields[default_prec] - fields["fp32"]).abs()
    frms = float(fields["fp32"].pow(2).mean().sqrt())
    parity = {"denoiser": {"reference": "CPU oracle (torch CPU fp32 restatement of the reference, pinned against the live "
                                        "reference by tests/golden), one full-size sample, sigma = 1",
                           "max_rel": den, "tolerance": {"fp32": 1e-5, "fp16x2": 1e-3, "fp16x2m": 1e-3, "fp16s32": 1e-3},
                           "mode_selection": None if precheck is None else
                           {"rule": f"the first of {list(AUTO_MODES)} whose denoiser max-rel is <= {AUTO_MAX_REL} on this network, else the last "
                                    "(checked before timing)",
                            "measured": precheck["max_rel"]},
                           "benchmarked_mode_within_tolerance": bool(den[default_prec] <= 1e-3)},
              "sampled_field": {"what": f"{integ}-{nsteps} ({nfe} NFE), B = 1, precision {default_prec} vs the tensor-core fp32 mode "
                                        f"on the same x_T (that mode vs the LIVE reference at full length: "
                                        f"tests/test_gpu_fullsteps.py)",
                                "per_pixel_max_abs": float(diff.max()), "per_pixel_rms": float(diff.pow(2).mean().sqrt()),
                                "field_rms": frms, "rms_over_field_rms": float(diff.pow(2).mean().sqrt()) / frms}}
    return modes, parity


def library_baseline(workload, net, cfg):
    """The oracle port of the reference on torch's OWN CUDA path on this GPU (cuDNN convolutions, cuBLAS / SDPA attention;
    TF32 and autocast-bf16) -- SURVEY 8d's "honest GPU baseline", measured in the same run on a bounded sample."""
    import types
    sd = {k: v.detach() for k, v in net.state_dict().items()}
    ocfg = types.SimpleNamespace(**cfg.export_description())
    out = {}
    for mode in ("tf32", "bf16"):
        v, _, sample, _, _, _ = cuda_port_arm(workload, 2, 3, sd, ocfg, mode, 8, 10.0)
        out[mode] = {"value": v, "unit": "samples/s", "sample": sample}
    return out


KERNEL_KIND = {}


def dominant_conv_roofline(cfg, shape, B, precision, dev, peaks, reps=20):
    """Time the dominant kernel (the full-resolution C->C 3^d convolution: 53% of C4's FLOPs) live with CUDA
    events on the launching stream: `reps` back-to-back launches on buffers larger than L2."""
    import torch
    import diffsci_b200 as d
    from diffsci_b200 import ops
    nd, M = cfg.dimension, cfg.model_channels
    sp = (1,) + tuple(shape[1:]) if nd == 2 else tuple(shape[1:])
    KERNEL_KIND.update({torch.float32: "CUDA-core FFMA implicit GEMM", torch.bfloat16: "tcgen05 implicit GEMM, bf16 operands",
                        torch.float16: "tcgen05 implicit GEMM, fp16 operands" + (
                            ", activations split hi+lo: 2 MMAs per k-step" if precision == "fp16x2" else
                            ", fp32 output (mixed mode: this 64-channel layer takes plain fp16 activations)" if precision == "fp16x2m" else
                            ", fp32 output" if precision == "fp16s32" else ""),
                        ops.SPLIT: "tcgen05 implicit GEMM, split-fp16 operands: 3 MMAs per k-step, fp32-parity mode; "
                                   "achieved = algorithmic FLOPs, the tensor pipe executes 3x"})
    from diffsci_b200.models.nets.punetg import _ACT_DTYPE, _W_DTYPE
    adt = _ACT_DTYPE[precision]
    wd = _W_DTYPE.get(precision) if d.TC_CONV_ENABLED else None
    wd = wd or torch.float32
    from diffsci_b200.models.nets.punetg import MIXED_MIN_CIN
    split = precision in ("fp32", "fp16x2") and wd != torch.float32     # split-fp16 activations (hi | lo), fp32 output
    mixed_plain = (precision == "fp16x2m" and M < MIXED_MIN_CIN) or precision == "fp16s32"   # this layer takes plain fp16 activations
    w = torch.randn((M, M) + (cfg.kernel_size,) * nd, device=dev) * 0.02
    pc = ops.PackedConv(w, torch.zeros(M, device=dev), nd, wd)
    nbuf = 4                                      # rotate inputs so consecutive launches do not hit in L2
    xs = [torch.randn((B,) + sp + (M,), device=dev).to(adt) for _ in range(nbuf)]
    out = torch.empty_like(xs[0])
    if split or (precision == "fp16x2m" and not mixed_plain):
        xs = [ops.split_f16(x) for x in xs]
    elif mixed_plain:
        xs = [x.half() for x in xs]
    for i in range(3):
        ops.conv(xs[i % nbuf], pc, out=out)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        ops.conv(xs[i % nbuf], pc, out=out)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / reps
    S = 1
    for s in sp:
        S *= s
    flops = 2.0 * B * S * M * M * cfg.kernel_size ** nd
    achieved = flops / (ms * 1e-3) / 1e12
    # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of this kernel at this shape, from the committed
    # `ncu --set full` capture (profiles/r1q_conv_tc2_64x64_64cube_B8_ncu_full.txt); scaled by batch (the kernel reads its
    # input once and writes its output once: traffic is linear in the number of samples); null for other shapes
    traffic, traffic_src = None, None
    try:            # measured, not hard-coded: profiles/conv_traffic.json is written by tools/conv_traffic_from_ncu.py from the
        #             committed `ncu --set full` captures of THIS kernel at HEAD (per precision mode)
        tj = json.load(open(os.path.join(ROOT, "profiles", "conv_traffic.json")))
        e = tj.get(f"conv{nd}d_{M}x{M}_" + "x".join(map(str, sp[-nd:])), {}).get(precision)
        if e:
            traffic = (e["dram_bytes_read"] + e["dram_bytes_write"]) * B / e["batch"]
            traffic_src = e["source"]
    except Exception:
        pass
    peak = peaks.get("bf16_tflops")
    src = "measured (MEASURED_PEAKS.json bf16_tflops, burst)"
    if peak is None:
        peak, src = 1590.0, "fallback (B200_PROFILING.md)"
    return {"bound": "tensor", "kernel": f"conv{nd}d {M}->{M} k{cfg.kernel_size} @ {'x'.join(map(str, sp[-nd:]))} "
            f"({KERNEL_KIND.get(wd, 'tcgen05 implicit GEMM')})",
            "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
            "traffic_source": traffic_src, "peak_source": src, "ms_per_launch": ms, "flops_per_launch": flops}


if __name__ == "__main__":
    main()
