"""dsk_plan_* / dsk_denoiser_fwd: the network evaluation as a tape replayed from C (include/diffsci_b200.h, csrc/tape.cu).
CPU: the generated call table is current and covers the header, the tape parser validates what it is given.
GPU: a recorded tape replays bit-identically in-process and from a pure-C host (examples/denoise_host.c), and matches
KarrasModule.get_denoiser and the CPU oracle."""
import ctypes as C
import os
import re
import struct
import subprocess
import sys
import types

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEV = "cuda:0"


def test_dispatch_table_is_current_and_covers_the_header():
    assert subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_tape_dispatch.py"), "--check"]).returncode == 0, \
        "diffsci_b200/csrc/tape_dispatch.inc is stale: run python tools/gen_tape_dispatch.py"
    header = open(os.path.join(ROOT, "include", "diffsci_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", " ", header, flags=re.S)
    launches = set(re.findall(r"\bint\s+(dsk_[a-z0-9_]+)\s*\(", header))
    table = set(re.findall(r'TAPE_FN\(\d+, "(dsk_[a-z0-9_]+)"', open(os.path.join(ROOT, "diffsci_b200", "csrc", "tape_dispatch.inc")).read()))
    not_ops = {"dsk_plan_create_from_tape", "dsk_plan_load", "dsk_plan_bind", "dsk_plan_destroy", "dsk_denoiser_fwd"}
    assert table == launches - not_ops


def _tape(ops, buffers=((64, None),), blob=b"", relocs=()):
    content = bytearray()
    brec = []
    for nbytes, data in buffers:
        if data is None:
            brec.append((nbytes, 0xFFFFFFFFFFFFFFFF))
        else:
            brec.append((nbytes, len(content)))
            content += data + bytes((-len(data)) % 8)
    out = bytearray(struct.pack("<8sIIIIQQiiq", b"DSKTAPE1", 1, len(brec), len(ops), len(relocs), len(blob), len(content), 2, 1, 16))
    for b in brec:
        out += struct.pack("<QQ", *b)
    for name, args in ops:
        out += struct.pack("<48sII", name.encode(), len(args), 0)
        for a in args:
            out += struct.pack("<IIQ", *a)
    for r in relocs:
        out += struct.pack("<IIQQ", *r)
    return bytes(out + blob + content)


def test_tape_parser_validates():
    """Host-only: no GPU is touched by dsk_plan_create_from_tape / dsk_plan_info / dsk_plan_destroy."""
    from diffsci_b200 import _lib
    lib = _lib.lib

    def load(tape):
        h = C.c_void_p()
        rc = lib.dsk_plan_create_from_tape(tape, len(tape), C.byref(h))
        return rc, h
    INT, FLT, BUF, NUL, BLOB, EXT, STREAM = range(7)
    add = ("dsk_add", [(EXT, 0, 0), (BUF, 0, 0), (EXT, 2, 0), (INT, 0, 16), (INT, 0, 0), (STREAM, 0, 0)])
    rc, h = load(_tape([add], buffers=((64, bytes(64)), (4096, None))))
    assert rc == 0
    assert lib.dsk_plan_info(h, 0) == 256 + 4096 and lib.dsk_plan_info(h, 1) == 2 and lib.dsk_plan_info(h, 4) == 1
    # an unbound plan refuses to run
    assert lib.dsk_denoiser_fwd(h, C.c_void_p(256), C.c_void_p(256), C.c_void_p(256), None) != 0
    assert "not bound" in _lib.last_error()
    assert lib.dsk_plan_destroy(h) == 0
    for bad, why in [(b"NOTATAPE" + bytes(64), "DSKTAPE1"),
                     (_tape([("dsk_no_such_entry", [])]), "unknown entry point"),
                     (_tape([("dsk_add", add[1][:-1])]), "arguments"),
                     (_tape([("dsk_add", [(INT, 0, 0)] + add[1][1:])]), "kind"),
                     (_tape([("dsk_add", [(BUF, 7, 0)] + add[1][1:])]), "outside its buffer"),
                     (_tape([add], relocs=((0, 0, 0, 0),)), "relocation"),
                     (_tape([add])[:100], "truncated")]:
        rc, _ = load(bad)
        assert rc != 0 and why in _lib.last_error(), (why, _lib.last_error())


@pytest.mark.gpu
@pytest.mark.parametrize("precision,dim,mc,size", [("fp16s32", 3, 64, 16), ("fp32_ffma", 2, 8, 12), ("bf16", 2, 64, 32)])
def test_tape_replays_the_denoiser(precision, dim, mc, size, tmp_path):
    import diffsci_b200 as d
    from diffsci_b200 import tape
    from oracle import karras_oracle as K, nets_oracle as N
    torch.manual_seed(0)
    cfg = d.PUNetGConfig(dimension=dim, model_channels=mc)
    net = d.PUNetG(cfg, precision=precision).to(DEV).eval()
    module = d.KarrasModule(net, d.KarrasModuleConfig.from_edm())
    B, shape = 2, (1,) + (size,) * dim
    path = str(tmp_path / "net.tape")
    blob, x, sigma, D = tape.export_denoiser(module, B, shape, path)
    # (1) the module's own denoiser (EDM scalars by torch) vs the recorded evaluation (scalars by dsk_edm_coeffs)
    with torch.no_grad():
        Dm, _ = module.get_denoiser(x, sigma)
    assert float((Dm - D).abs().max()) <= 2e-6 * float(Dm.abs().max())
    # (2) replay in-process through the C API on OTHER inputs, inside a fresh workspace: bit-identical to the module's plan
    plan = tape.TapePlan(blob, torch.device(DEV))
    assert plan.batch == B and plan.launches > 20
    assert torch.equal(plan.denoise(x, sigma), D)
    x2, s2 = torch.randn_like(x) * 2, torch.tensor([0.7, 5.0], device=DEV)
    D2 = plan.denoise(x2, s2)
    with torch.no_grad():
        Dm2, _ = module.get_denoiser(x2, s2)
    assert float((Dm2 - D2).abs().max()) <= 2e-6 * float(Dm2.abs().max())
    # (3) against the CPU oracle (the reference's arithmetic)
    sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    ocfg = types.SimpleNamespace(**cfg.export_description())
    ref = K.denoiser(lambda xx, tt: N.punetg_forward(sd, ocfg, xx, tt), x2.cpu(), s2.cpu())
    tol = {"fp16s32": 2e-3, "fp32_ffma": 5e-5, "bf16": 5e-2}[precision]
    assert float((D2.cpu() - ref).abs().max() / ref.abs().max()) < tol
    # (4) a pure-C host process: examples/denoise_host.c linked against the library, no Python
    exe = str(tmp_path / "denoise_host")
    cuda = "/usr/local/cuda"
    subprocess.run(["gcc", "-O2", os.path.join(ROOT, "examples", "denoise_host.c"), "-I" + os.path.join(ROOT, "include"),
                    f"-I{cuda}/include", "-L" + os.path.join(ROOT, "diffsci_b200"), "-ldiffsci_b200", f"-L{cuda}/lib64", "-lcudart",
                    "-Wl,-rpath," + os.path.join(ROOT, "diffsci_b200"), "-o", exe], check=True)
    x2.cpu().numpy().tofile(str(tmp_path / "x.f32"))
    s2.cpu().numpy().tofile(str(tmp_path / "s.f32"))
    r = subprocess.run([exe, path, str(tmp_path / "x.f32"), str(tmp_path / "s.f32"), str(tmp_path / "o.f32")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    import numpy as np
    got = torch.from_numpy(np.fromfile(str(tmp_path / "o.f32"), dtype=np.float32)).view(D2.shape)
    assert torch.equal(got, D2.cpu()), r.stdout
