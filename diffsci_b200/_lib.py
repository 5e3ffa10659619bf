"""ctypes binding of libdiffsci_b200.so (the C ABI declared in include/diffsci_b200.h).

The shared library is built IN-TREE by ``__graft_entry__.build()`` / ``make -C diffsci_b200/csrc``.
There is no fallback: if the library is missing, importing this module raises, and every
compute entry point raises ``RuntimeError`` with the library's own error text on failure.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdiffsci_b200.so")

F32, BF16, F16, SPLIT_F16 = 0, 1, 2, 3
TAB_COLS = 8
TAB_T, TAB_DT, TAB_THAT, TAB_LANG, TAB_NOISE, TAB_SQDT, TAB_TNEXT, TAB_CHURN = range(8)
(STAGE_INIT, STAGE_EULER, STAGE_HEUN_MID, STAGE_HEUN_FIN, STAGE_HEUN_LAST, STAGE_EM, STAGE_KARRAS_MID,
 STAGE_KARRAS_FIN, STAGE_KARRAS_LAST) = range(9)

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(or `make -C diffsci_b200/csrc`).  diffsci_b200 has no CPU / PyTorch fallback.")

lib = C.CDLL(LIB_PATH)

p, i32, i64, u64, f32 = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_float


class ConvDesc(C.Structure):
    _fields_ = [(n, i32) for n in ("B", "D", "H", "W", "Cin", "Cout", "ksize", "ndim", "up2", "w_dtype", "in_dtype",
                                   "out_dtype", "out_nchw_f32", "circular", "res_dtype", "operand16")]


# name -> argtypes (every function returns int unless listed in _RESTYPE)
SIGNATURES = {
    "dsk_version": [],
    "dsk_last_error": [],
    "dsk_launch_count": [],
    "dsk_check_device": [i32],
    "dsk_precond_scale": [p, p, p, i32, i32, i64, i32, p],
    "dsk_precond_denoise": [p, p, p, p, p, p, p, i32, i32, i64, i32, p],
    "dsk_sampler_stage": [i32, p, p, p, p, p, p, p, p, p, u64, p, i32, i32, i64, f32, f32, i32, i32, p],
    "dsk_sampler_advance": [p, p],
    "dsk_sampler_stage_general": [i32, p, p, p, p, p, p, p, p, p, p, i32, i32, i64, f32, i32, i32, p],
    "dsk_sampler_stage_cond": [i32, p, p, p, p, p, p, p, p, p, u64, p, i32, i32, i64, f32, f32, i32, i32, i32, i32, f32, p],
    "dsk_sampler_stage_blend": [i32, p, p, p, p, p, p, p, p, p, u64, p, i32, i32, i64, f32, f32, i32, i32, i32, i32, f32, p, p, i64,
                                i32, p],
    "dsk_precond_scale_cond": [p, p, p, i32, i32, i64, i32, i32, i32, p],
    "dsk_cfg_mix": [p, p, f32, i64, i32, p],
    "dsk_lincomb": [p, i64, p, f32, p, f32, p, f32, p, f32, p],
    "dsk_mask_blend": [p, p, p, p, i64, i64, p],
    "dsk_philox_normal": [p, i64, u64, C.c_uint32, p],
    "dsk_dropout": [p, p, p, i64, f32, u64, C.c_uint32, i32, p],
    "dsk_conv_fwd": [C.POINTER(ConvDesc), p, p, p, p, p, p, p],
    "dsk_conv_stats_supported": [C.POINTER(ConvDesc)],
    "dsk_conv_stats_slots": [],
    "dsk_pad_circular": [p, p, i32, i32, i32, i32, i32, i32, i32, p],
    "dsk_conv_pad_ws_bytes": [C.POINTER(ConvDesc)],
    "dsk_conv_fwd_circ": [C.POINTER(ConvDesc), p, p, p, p, p, p, p, p, p],
    "dsk_conv_fwd_stats": [C.POINTER(ConvDesc), p, p, p, p, p, p, p, p],
    "dsk_norm_act_prestat": [p, p, p, p, p, p, p, i32, p, i32, i64, i32, i32, i32, i32, i32, i32, p],
    "dsk_upsample2x": [p, p, i32, i32, i32, i32, i32, i32, i32, p],
    "dsk_pack_upconv_weight": [p, p, i32, i32, i32, i32, p],
    "dsk_pack_conv_weight": [p, p, i32, i32, i32, i32, p],
    "dsk_pack_conv_weights_multi": [p, i32, i32, p],
    "dsk_gemm_f32": [p, p, p, p, i32, i32, i32, i32, i32, i32, i64, i64, i64, i32, i32, f32, i32, p],
    "dsk_gemm_bf16_tc": [p, p, p, p, i32, p, i32, i32, i32, i64, i64, i64, i64, i64, i64, i32, f32, i32, i32, i32, p],
    "dsk_gemm_tc": [p, p, p, p, i32, p, i32, i32, i32, i32, i64, i64, i64, i64, i64, i64, i32, f32, i32, i32, i32, i32, i32, i32, p],
    "dsk_attn_softmax_qk_h16": [p, p, p, p, i32, i32, i64, i64, i64, i64, i32, f32, i32, p],
    "dsk_softmax_rows_h16": [p, p, i64, i32, i32, p],
    "dsk_attn_flash": [p, p, p, p, i32, i32, i64, i64, i64, i64, i64, i64, i64, i64, i32, f32, i32, i32, p],
    "dsk_softmax_bwd_rows_bf16": [p, p, p, i64, i32, p],
    "dsk_attn_softmax_ws_bytes": [i32, i32],
    "dsk_attn_softmax_qk": [p, p, p, p, i32, i32, i64, i64, i64, i64, i32, f32, p],
    "dsk_softmax_rows_bf16": [p, p, i64, i32, p],
    "dsk_norm_ws_bytes": [i32, i64, i32],
    "dsk_norm_act": [p, p, p, p, p, p, p, i32, i64, i32, i32, i32, i32, i32, i32, p],
    "dsk_norm_apply_padded": [p, p, p, i32, i32, i32, i32, i32, i32, i32, i32, i32, p],
    "dsk_pool2x": [p, p, i32, i32, i32, i32, i32, i32, i32, i32, p],
    "dsk_pool2x_f32": [p, p, i32, i32, i32, i32, i32, i32, i32, i32, p],
    "dsk_add": [p, p, p, i64, i32, p],
    "dsk_nchw_to_cl": [p, p, i32, i32, i64, i32, p],
    "dsk_cl_to_nchw": [p, p, i32, i32, i64, i32, p],
    "dsk_cast": [p, p, i64, i32, i32, p],
    "dsk_split_f16": [p, p, i64, i32, p],
    "dsk_concat_channels": [p, p, p, i64, i32, i32, i32, p],
    "dsk_fourier": [p, p, p, i32, i32, p],
    "dsk_grouped_linear": [p, p, p, p, p, p, p, i32, i32, i32, i32, p],
    "dsk_grouped_linear_bwd": [p, p, p, p, p, p, p, p, p, p, i32, i32, i32, i32, i32, i32, i32, p],
    "dsk_grouped_gemm_f32": [p, i32, i32, i32, i32, i32, i32, p],
    "dsk_grouped_dz_bias": [p, p, p, p, p, i32, i32, i32, i32, p],
    "dsk_softmax_rows": [p, i64, i32, p],
    "dsk_edm_loss_fwd_bwd": [p, p, p, p, p, p, p, i32, i32, i64, f32, i32, p],
    "dsk_precond_loss_fwd_bwd": [p, p, p, p, p, p, p, p, p, p, i32, i32, i64, i32, p],
    "dsk_precond_loss_rows": [p, p, p, p, p, p, p, p, p, p, i32, i32, i64, i32, p],
    "dsk_precond_loss_fwd_bwd_huber": [p, p, p, p, p, p, p, p, p, p, i32, i32, i64, i32, f32, p],
    "dsk_precond_loss_rows_huber": [p, p, p, p, p, p, p, p, p, p, i32, i32, i64, i32, f32, p],
    "dsk_ensemble_noise_add": [p, p, p, p, i32, i32, i64, p],
    "dsk_ensemble_loss_fwd_bwd": [p, p, p, p, p, p, p, p, p, i32, p, p, i32, i32, i32, i64, i32, p],
    "dsk_ema_update": [p, p, p, i32, i64, f32, p],
    "dsk_adamw_ema_step": [p, p, p, p, p, p, i32, i64, f32, f32, f32, f32, f32, i32, f32, f32, p],
    "dsk_gemm_f32_ex": [p, p, p, p, i32, i32, i32, i32, i32, i32, i64, i64, i64, i32, i32, i32, f32, f32, i32, p],
    "dsk_pack_conv_weight_dgrad": [p, p, i32, i32, i32, i32, p],
    "dsk_conv_wgrad_ws_bytes": [C.POINTER(ConvDesc)],
    "dsk_conv_wgrad": [C.POINTER(ConvDesc), p, p, p, p, i32, p],
    "dsk_bwd_ws_bytes": [i32, i64, i32],
    "dsk_channel_sum": [p, p, p, i32, i64, i32, i32, i32, p],
    "dsk_colsum_f32": [p, p, i64, i32, i32, p],
    "dsk_norm_act_bwd": [p, p, p, p, p, p, p, p, p, p, p, p, p, i32, i64, i32, i32, i32, i32, i32, p],
    "dsk_pool2x_bwd": [p, p, p, p, i32, i32, i32, i32, i32, i32, i32, i32, p],
    "dsk_upsample2x_bwd": [p, p, p, i32, i32, i32, i32, i32, i32, i32, p],
    "dsk_softmax_bwd_rows": [p, p, i64, i32, p],
    "dsk_silu_fwd": [p, p, i64, p],
    "dsk_silu_bwd": [p, p, p, i64, p],
    "dsk_relu_fwd": [p, p, i64, p],
    "dsk_relu_bwd": [p, p, p, i64, p],
    "dsk_add_ex": [p, i32, p, i32, p, i32, i64, p],
    "dsk_split_channels": [p, p, p, p, p, i64, i32, i32, i32, p],
    "dsk_edm_coeffs": [p, f32, p, p, p, p, i32, p],
    "dsk_plan_create_from_tape": [p, i64, p],
    "dsk_plan_load": [C.c_char_p, p],
    "dsk_plan_info": [p, i32],
    "dsk_plan_bind": [p, p, p],
    "dsk_denoiser_fwd": [p, p, p, p, p],
    "dsk_plan_destroy": [p],
}
_RESTYPE = {"dsk_conv_pad_ws_bytes": i64, "dsk_last_error": C.c_char_p, "dsk_launch_count": u64, "dsk_norm_ws_bytes": i64,
            "dsk_conv_wgrad_ws_bytes": i64, "dsk_bwd_ws_bytes": i64, "dsk_attn_softmax_ws_bytes": i64, "dsk_plan_info": i64}

for _name, _args in SIGNATURES.items():
    _fn = getattr(lib, _name)          # AttributeError here == header/library mismatch: fail loudly
    _fn.argtypes = _args
    _fn.restype = _RESTYPE.get(_name, i32)


class _Lib:
    """The library object every module imports.  Entry points are the ctypes functions themselves, except while a tape is being
    recorded (diffsci_b200/tape.py): then every call of an `int dsk_*` entry point is also appended -- name + raw arguments -- to
    the active recorder, which is how the launch list the Python plan assembles becomes replayable from C (dsk_denoiser_fwd)."""
    _NOT_LAUNCHES = ("dsk_plan_", "dsk_denoiser_fwd", "dsk_check_device", "dsk_version")

    def __init__(self, cdll):
        self._cdll = cdll
        self.recorder = None

    def __getattr__(self, name):
        fn = getattr(self._cdll, name)
        if name not in SIGNATURES or _RESTYPE.get(name, i32) is not i32 or name.startswith(self._NOT_LAUNCHES):
            setattr(self, name, fn)
            return fn

        def call(*args, _fn=fn, _name=name):
            rc = _fn(*args)
            if self.recorder is not None:
                self.recorder.append((_name, args))
            return rc
        call.__name__ = name
        call.argtypes, call.restype = fn.argtypes, fn.restype
        setattr(self, name, call)
        return call


lib = _Lib(lib)


def last_error() -> str:
    return lib.dsk_last_error().decode()


def check(rc: int) -> None:
    if rc != 0:
        raise RuntimeError(f"libdiffsci_b200 error {rc}: {last_error()}")


def launch_count() -> int:
    return int(lib.dsk_launch_count())


_checked_devices: set[int] = set()


def require_cuda(t: torch.Tensor, what: str = "tensor") -> None:
    """The product path is CUDA-only.  Raise (never fall back) for CPU tensors / non-B200 devices.  The library launches on the
    CURRENT device's current stream (it never calls cudaSetDevice), so a tensor on another GPU is an error too: select the
    device first (``torch.cuda.set_device`` / ``with torch.cuda.device(...)``), as one-process-per-GPU code does."""
    if not t.is_cuda:
        raise RuntimeError(f"diffsci_b200: {what} is on {t.device}; the hot path is sm_100a CUDA only (no CPU fallback)")
    cur = torch.cuda.current_device()
    dev = t.device.index if t.device.index is not None else cur
    if dev != cur:
        raise RuntimeError(f"diffsci_b200: {what} is on cuda:{dev} but the current device is cuda:{cur}; kernels launch on the "
                           f"current device's stream -- call torch.cuda.set_device({dev}) (or use `with torch.cuda.device({dev})`)")
    if dev not in _checked_devices:
        check(lib.dsk_check_device(dev))
        _checked_devices.add(dev)


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def dt_code(dtype: torch.dtype) -> int:
    if dtype == torch.float32:
        return F32
    if dtype == torch.bfloat16:
        return BF16
    if dtype == torch.float16:
        return F16
    raise TypeError(f"diffsci_b200: unsupported dtype {dtype}")
