"""Tapes: the launch list of a network evaluation, recorded from the Python plan and replayable from C.

The reference has no FFI; this repo's boundary is Python duck-typing over a C ABI.  What a non-Python host needs on
top of the per-kernel entry points is the ORDER in which one network evaluation calls them -- which lives in
models/nets/punetg.py / adm.py.  `export_denoiser` runs D(x; sigma) = c_skip x + c_out F(c_in x, c_noise)
(KarrasModule.get_denoiser, reference karras/karrasmodule.py:673-719) once with the recorder of `_lib._Lib` switched on and writes

  * every entry point called, in order, with its arguments: integers / floats by value, `dsk_conv_desc` by value, pointers as
    (buffer, byte offset), the three tensors of the call (x, sigma, D) as external slots, the stream as a marker;
  * the size of every device buffer the pointers fall into (torch storages), and the CONTENTS of those the evaluation does not
    write (packed weights, parameters, pointer tables) -- found by comparing checksums across a second evaluation on other inputs;
  * relocations for device pointers stored inside constant buffers (the grouped time-MLP kernels read pointer tables).

csrc/tape.cu (dsk_plan_load / dsk_plan_bind / dsk_denoiser_fwd) replays the file inside one caller-provided workspace;
examples/denoise_host.c is a complete C host.  Format: see the header comment of csrc/tape.cu.
"""
from __future__ import annotations

import ctypes as C
import gc
import struct
from typing import Optional

import torch

from . import _lib as L
from ._lib import lib, check, ptr, stream, dt_code, require_cuda

A_INT, A_FLT, A_BUF, A_NULL, A_BLOB, A_EXT, A_STREAM = range(7)
EXT_X, EXT_SIGMA, EXT_OUT = range(3)
NO_CONTENT = 0xFFFFFFFFFFFFFFFF
RELOC_SCAN_MAX_BYTES = 1 << 20      # pointer tables are tiny; weights are not scanned for embedded pointers


class Recorder:
    """Context manager: collects (entry point, raw ctypes arguments) for every launch made through `_lib.lib`."""

    def __init__(self):
        self.calls = []

    def append(self, item):
        self.calls.append(item)

    def __enter__(self):
        assert lib.recorder is None, "a tape is already being recorded"
        lib.recorder = self
        return self

    def __exit__(self, *exc):
        lib.recorder = None
        return False


def _raw_pointer(arg):
    """ctypes argument in a pointer position -> (kind, payload): an address, None, or the struct behind byref()."""
    if arg is None:
        return "null", None
    if isinstance(arg, int):
        return ("addr", arg) if arg else ("null", None)
    if isinstance(arg, C.c_void_p):
        return ("addr", arg.value) if arg.value else ("null", None)
    obj = getattr(arg, "_obj", None)                      # ctypes.byref(struct)
    if isinstance(obj, C.Structure):
        return "blob", bytes(obj)
    if isinstance(arg, C.Structure):
        return "blob", bytes(arg)
    raise TypeError(f"cannot record pointer argument {arg!r}")


def _live_cuda_storages(device):
    """{base address: (nbytes, storage)} of every CUDA storage alive in this process (plan buffers, packed weights, tables)."""
    import warnings
    out = {}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")          # isinstance() on some lazily-deprecated module attributes warns
        objs = gc.get_objects()
    for o in objs:
        try:
            if type(o) is not torch.Tensor and not isinstance(o, torch.nn.Parameter):
                if not (type(o).__module__.startswith("torch") and isinstance(o, torch.Tensor)):
                    continue
            if o.is_cuda and o.device == device:
                st = o.untyped_storage()
                if st.nbytes() > 0:
                    out[st.data_ptr()] = (st.nbytes(), st)
        except Exception:       # objects that raise on isinstance / attribute access
            continue
    return out


def _checksum(storage) -> int:
    t = torch.empty(0, dtype=torch.uint8, device=storage.device).set_(storage)
    n8 = t.numel() // 8
    s = int(t[:n8 * 8].view(torch.int64).sum().item()) if n8 else 0
    return s * 1000003 + int(t[n8 * 8:].to(torch.int64).sum().item())


def export_denoiser(module, batch: int, shape: tuple, path: Optional[str] = None, seed: int = 0):
    """Record D(x; sigma) of an unconditional EDM `KarrasModule` over a native network (PUNetG / ADM plan) for inputs
    x: [batch, *shape] and write the tape to `path` (or return the bytes).  Returns (tape bytes, x, sigma, D) -- the recorded
    evaluation itself, for checking a replay."""
    from .models.karras import preconditioners
    net = module.model
    pre = module.config.preconditioner
    if type(pre) is not preconditioners.EDMPreconditioner:
        raise NotImplementedError("tape export: the EDM preconditioner's scalars are computed on the device (dsk_edm_coeffs); "
                                  "other preconditioners evaluate theirs with torch")
    if getattr(module, "conditional", False) or not hasattr(net, "plan"):
        raise NotImplementedError("tape export: unconditional modules over a native network (PUNetG / ADM)")
    dev = next(net.parameters()).device
    g = torch.Generator().manual_seed(seed)
    B, Cc = batch, shape[0]
    spatial = tuple(shape[1:])
    S = 1
    for v in spatial:
        S *= v
    f32 = dict(dtype=torch.float32, device=dev)

    def inputs():
        return (torch.randn(B, *shape, generator=g).to(dev).contiguous(), torch.exp(torch.randn(B, generator=g) * 1.2 - 1.2).to(dev))
    x, sigma = inputs()
    require_cuda(x, "x")
    D = torch.empty_like(x)
    c_in, c_out, c_skip, c_noise = (torch.empty(B, **f32) for _ in range(4))
    sd = float(pre.sigma_data)
    plan = net.plan(B, spatial, dev)
    ld = plan.xin.shape[-1]
    adt = plan.act_dtype

    def evaluate(x_, sigma_, D_):
        check(lib.dsk_edm_coeffs(ptr(sigma_), sd, ptr(c_in), ptr(c_out), ptr(c_skip), ptr(c_noise), B, stream()))
        check(lib.dsk_precond_scale_cond(ptr(x_), ptr(c_in), ptr(plan.xin), B, Cc, S, dt_code(adt), ld, 0, stream()))
        F = plan.forward(plan.xin, c_noise)
        check(lib.dsk_precond_denoise(ptr(F), ptr(x_), ptr(c_out), ptr(c_skip), ptr(sigma_), ptr(D_), None, B, Cc, S, dt_code(adt),
                                      stream()))
    with torch.no_grad():
        evaluate(x, sigma, D)                         # warm-up: packs weights, builds tables (not part of the tape)
        torch.cuda.synchronize(dev)
        with Recorder() as rec:
            evaluate(x, sigma, D)
        torch.cuda.synchronize(dev)
        storages = _live_cuda_storages(dev)
        bases = sorted(storages)

        def locate(addr):
            import bisect
            k = bisect.bisect_right(bases, addr) - 1
            if k >= 0 and addr < bases[k] + storages[bases[k]][0]:
                return bases[k], addr - bases[k]
            raise RuntimeError(f"tape export: pointer {addr:#x} is not inside any live CUDA tensor")
        ext = {x.untyped_storage().data_ptr(): EXT_X, sigma.untyped_storage().data_ptr(): EXT_SIGMA,
               D.untyped_storage().data_ptr(): EXT_OUT}
        buf_id, ops, blob = {}, [], bytearray()
        for name, args in rec.calls:
            sig = L.SIGNATURES[name]
            assert len(sig) == len(args), (name, len(sig), len(args))
            rec_args = []
            for j, (ty, a) in enumerate(zip(sig, args)):
                if ty is L.p or ty is C.c_char_p or (isinstance(ty, type) and issubclass(ty, C._Pointer)):
                    kind, val = _raw_pointer(a)
                    if j == len(sig) - 1:
                        rec_args.append((A_STREAM, 0, 0))
                    elif kind == "null":
                        rec_args.append((A_NULL, 0, 0))
                    elif kind == "blob":
                        while len(blob) % 8:
                            blob.append(0)
                        rec_args.append((A_BLOB, 0, len(blob)))
                        blob += val
                    else:
                        base, off = locate(val)
                        if base in ext:
                            rec_args.append((A_EXT, ext[base], off))
                        else:
                            rec_args.append((A_BUF, buf_id.setdefault(base, len(buf_id)), off))
                elif ty is L.f32 or ty is C.c_double:
                    rec_args.append((A_FLT, 0, struct.unpack("<Q", struct.pack("<d", float(a)))[0]))
                else:
                    rec_args.append((A_INT, 0, int(a) & 0xFFFFFFFFFFFFFFFF))
            ops.append((name, rec_args))
        # buffers reached only THROUGH a pointer table (the grouped time-MLP kernels read their operands' addresses from device
        # tables): scan the small buffers for words that are addresses of live storages, transitively
        import bisect

        def pointers_in(base):
            nbytes = storages[base][0]
            if nbytes > RELOC_SCAN_MAX_BYTES or nbytes < 8:
                return []
            data = torch.empty(0, dtype=torch.uint8, device=dev).set_(storages[base][1])[:nbytes // 8 * 8].view(torch.int64).cpu().tolist()
            found = []
            for w, v in enumerate(data):
                if v >= (1 << 32):
                    k = bisect.bisect_right(bases, v) - 1
                    if k >= 0 and v < bases[k] + storages[bases[k]][0]:
                        found.append((w * 8, bases[k], v - bases[k]))
            return found
        tables, todo = {}, list(buf_id)
        while todo:
            b = todo.pop()
            tables[b] = pointers_in(b)
            for _, target, _ in tables[b]:
                if target not in buf_id and target not in ext:
                    buf_id[target] = len(buf_id)
                    todo.append(target)
        # constant buffers = those a second evaluation on other inputs leaves unchanged
        order = sorted(buf_id, key=buf_id.get)
        before = {b: _checksum(storages[b][1]) for b in order}
        x2, sigma2 = inputs()
        evaluate(x2, sigma2, torch.empty_like(D))
        torch.cuda.synchronize(dev)
        const = {b for b in order if _checksum(storages[b][1]) == before[b]}
        if getattr(net, "ones_channel", 0):          # the constant input channel lives in a buffer the evaluation also writes
            const.add(plan.xin.untyped_storage().data_ptr())
        evaluate(x, sigma, D)                         # leave the recorded evaluation's result in D
        torch.cuda.synchronize(dev)
        snapshot = {b: torch.empty(0, dtype=torch.uint8, device=dev).set_(storages[b][1]).cpu().numpy().tobytes() for b in const}
    content, buffers, relocs = bytearray(), [], []
    for b in order:
        nbytes = storages[b][0]
        if b in const:
            while len(content) % 8:
                content.append(0)
            buffers.append((nbytes, len(content)))
            data = snapshot[b]
            content += data
            for off, target, toff in tables.get(b, []):      # device pointers stored in the buffer (pointer tables)
                if target in ext:
                    raise RuntimeError("tape export: a pointer table refers to an external tensor")
                relocs.append((buf_id[b], buf_id[target], off, toff))
        else:
            buffers.append((nbytes, NO_CONTENT))
    out = bytearray()
    out += struct.pack("<8sIIIIQQiiq", b"DSKTAPE1", 1, len(buffers), len(ops), len(relocs), len(blob) + (-len(blob)) % 8,
                       len(content), B, Cc, Cc * S)
    for nbytes, co in buffers:
        out += struct.pack("<QQ", nbytes, co)
    for name, rec_args in ops:
        out += struct.pack("<48sII", name.encode(), len(rec_args), 0)
        for kind, buf, val in rec_args:
            out += struct.pack("<IIQ", kind, buf, val)
    for r in relocs:
        out += struct.pack("<IIQQ", *r)
    out += blob + bytes((-len(blob)) % 8)
    out += content
    if path is not None:
        with open(path, "wb") as f:
            f.write(out)
    return bytes(out), x, sigma, D


class TapePlan:
    """Python handle on the C replay (dsk_plan_*): used by the tests to run a tape in-process exactly as a C host would."""

    def __init__(self, tape: bytes, device):
        self._h = C.c_void_p()
        self._tape = tape
        check(lib.dsk_plan_create_from_tape(tape, len(tape), C.byref(self._h)))
        self.workspace = torch.empty(int(lib.dsk_plan_info(self._h, 0)) + 256, dtype=torch.uint8, device=device)
        base = (self.workspace.data_ptr() + 255) & ~255
        check(lib.dsk_plan_bind(self._h, C.c_void_p(base), stream()))
        self.batch, self.channels = int(lib.dsk_plan_info(self._h, 1)), int(lib.dsk_plan_info(self._h, 2))
        self.launches = int(lib.dsk_plan_info(self._h, 4))

    def denoise(self, x: torch.Tensor, sigma: torch.Tensor) -> torch.Tensor:
        out = torch.empty_like(x)
        check(lib.dsk_denoiser_fwd(self._h, ptr(x), ptr(sigma), ptr(out), stream()))
        return out

    def __del__(self):
        if getattr(self, "_h", None):
            lib.dsk_plan_destroy(self._h)
            self._h = None
