"""Sampled-field parity at BASELINE.json's REAL step counts (north_star: "sampled fields within a stated per-pixel
tolerance after N steps"): the whole sampling loop -- KarrasModule.propagate_white_noise through the CUDA-graph engine -- against
fields recorded from the LIVE reference's own loop (karrasmodule.py:867-931, schedulers.py:48-89, integrators.py:29-113) on the
same weights, the same x_T and, for the SDE sampler, the same step noise (tests/golden/fullsteps_*.pt, oracle/make_goldens.py
--only fullsteps; fp32 and fp64 runs of the reference):

  c1  MLPUncond(2, [128]*3, SiLU)            Heun 18 steps (35 NFE),   B = 64
  c2  PUNetG 2-D mc = 128, 1 x 28 x 28       Heun 40 steps (79 NFE),   B = 2
  c4  PUNetG 3-D mc = 64, 1 x 64^3           Heun 64 steps (127 NFE),  B = 1
  c5  PUNetG 2-D mc = 64, 1 x 256 x 256      Euler-Maruyama 256 steps (256 NFE), B = 1, injected Brownian increments

Per-pixel error = |ours - reference|, reported as max-abs and RMS, absolute and relative to the RMS of the reference field.
STATED TOLERANCES
  * fp32 modes: the reference's own fp32 result differs from its fp64 result by d0 (rounding amplified through up to 256 chained
    evaluations of a random-weight network); two fp32 implementations that sum in different orders differ by a few d0 (the
    CUDA-core FFMA mode and the tensor-core split mode land at the same distance: c5 5.4 x d0 and 5.6 x d0); ours must stay
    within 8 x d0 + 2e-5 x field RMS of the fp64 field, per pixel (max) and in RMS.
  * 16-bit modes vs the reference's fp32 field, RMS relative to the field RMS: fp16x2 <= 3e-3, fp16x2m / fp16s32 <= 4e-3, fp16 <= 2e-2, bf16 <= 1.5e-1
    for c2; the long runs of the random-weight default-width networks (c4: 127 chained evaluations, c5: 256 SDE steps) amplify
    rounding chaotically for EVERY arithmetic (c5: fp32 vs fp64 of the reference itself differ by 1.3e-3 of the field RMS; c4:
    the two fp32 modes differ from the reference's fp32 by 7e-5 and 2e-4): 16-bit modes <= 1.5e-1 there (measured 1e-2 ... 9e-2)
    (measured values are printed; DESIGN.md section 2 quotes them).
"""
import os
import types

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RMS_TOL = {"fp16x2": 3e-3, "fp16x2m": 4e-3, "fp16s32": 4e-3, "fp16": 2e-2, "bf16": 1.5e-1}


def _errs(a, b):
    d = (a.double() - b.double()).abs()
    return float(d.max()), float(d.pow(2).mean().sqrt())


def run_case(name, modes):
    import diffsci_b200 as d
    from oracle import nets_oracle as N
    path = os.path.join(GOLD, f"fullsteps_{name}.pt")
    if not os.path.exists(path):
        pytest.skip(f"{path} not generated")
    g = torch.load(path, weights_only=False)
    sd = N.synth_state_dict(g["manifest"], g["wseed"])
    if g["net"] == "mlp":
        net = d.MLPUncond(g["kw"]["dim"], g["kw"]["hidden_dims"], nonlinearity=torch.nn.SiLU())
    else:
        net = d.PUNetG(d.PUNetGConfig(**g["kw"]), precision="fp32")
    net.load_state_dict(sd)
    net = net.to(DEV).eval()
    mod = d.KarrasModule(net, d.KarrasModuleConfig.from_edm())
    torch.manual_seed(g["xseed"])
    wn = torch.randn(*g["shape"])
    noises = None
    if g.get("nseed") is not None:
        gen = torch.Generator().manual_seed(g["nseed"])
        noises = [torch.randn(tuple(g["shape"]), generator=gen) for _ in range(g["nsteps"])]
    ref32, ref64 = g["out_f32"], g["out_f64"]
    truth = ref64 if ref64 is not None else ref32
    field_rms = float(truth.double().pow(2).mean().sqrt())
    d0 = _errs(ref32, ref64) if ref64 is not None else (0.0, 0.0)
    out = {}
    for prec in modes:
        if g["net"] != "mlp":
            net.precision = prec
        integ = d.name_to_integrator(g["integrator"])
        if noises is not None:
            integ.reset_noise(injected=noises)
        y = mod.propagate_white_noise(wn.to(DEV), nsteps=g["nsteps"], integrator=integ).cpu()
        assert torch.isfinite(y).all()
        out[prec] = (_errs(y, ref32), _errs(y, truth))
    print(f"\n{name}: {g['integrator']}-{g['nsteps']} ({mod.last_nfe} NFE), field RMS {field_rms:.3e}; reference fp32 vs fp64: "
          f"max-abs {d0[0]:.2e} rms {d0[1]:.2e}")
    for prec, ((m32, r32), (m64, r64)) in out.items():
        print(f"  {prec:10s} vs reference fp32: max-abs {m32:.2e} rms {r32:.2e} (rms / field rms {r32 / field_rms:.2e});  "
              f"vs fp64: max-abs {m64:.2e} rms {r64:.2e}")
    return out, d0, field_rms


@pytest.mark.parametrize("name", ["c1", "c2", "c4", "c5"])
def test_full_length_sampling_vs_live_reference(name):
    modes = ["fp32"] if name == "c1" else ["fp32", "fp32_ffma", "fp16x2", "fp16x2m", "fp16s32", "fp16", "bf16"]
    out, d0, field_rms = run_case(name, modes)
    (_, _), (m64, r64) = out["fp32"]
    assert m64 <= 8 * d0[0] + 2e-5 * field_rms, (m64, d0)
    assert r64 <= 8 * d0[1] + 2e-5 * field_rms, (r64, d0)
    for prec, tol in RMS_TOL.items():
        if prec in out:
            tol = 1.5e-1 if name in ("c4", "c5") else tol
            assert out[prec][0][1] <= tol * field_rms, (prec, out[prec][0], field_rms)
