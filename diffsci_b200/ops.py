"""Thin tensor-level wrappers over the C ABI (include/diffsci_b200.h).

PyTorch is used here only for device memory (torch.empty), streams and dtype bookkeeping; all
arithmetic is done by the hand-written kernels in libdiffsci_b200.so.  Activations are
CHANNELS-LAST tensors of shape [B, D, H, W, C] (D == 1 for 2-D nets), fp32 or bf16.
Every function raises RuntimeError for CPU tensors: there is no fallback path.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib as L
from ._lib import lib, check, ptr, stream, dt_code, require_cuda


_weight_epoch = [0]

# Weight format marker of the split tensor-core kernels (DSK_SPLIT_F16): packed rows are [hi | lo] fp16, 2 * Cin long.  A split
# ACTIVATION tensor is a torch.float16 tensor whose last dimension is 2 * C (channels [0, C) = fp16(v), [C, 2C) = fp16(v - hi)).
SPLIT = "split_f16"
H16 = (torch.bfloat16, torch.float16)


def is_tc_dtype(w_dtype) -> bool:
    """Weights packed for the tcgen05 kernels (16-bit K-major rows) as opposed to the fp32 FFMA layout."""
    return w_dtype in H16 or w_dtype == SPLIT


def w_code(w_dtype) -> int:
    return L.SPLIT_F16 if w_dtype == SPLIT else dt_code(w_dtype)


def act_code(x: torch.Tensor, channels: int) -> int:
    """dtype code of a channels-last activation tensor that carries `channels` logical channels (split tensors hold 2x)."""
    if x.dtype == torch.float16 and x.shape[-1] == 2 * channels:
        return L.SPLIT_F16
    return dt_code(x.dtype)


def bump_weight_epoch() -> None:
    """Invalidate every packed weight copy: called by code that updates parameters with a library kernel (fused AdamW),
    which PyTorch's per-tensor version counters cannot see."""
    _weight_epoch[0] += 1


class PackedConv:
    """Device-side packed copy of a reference-layout conv weight [Cout, Cin, k(,k)(,k)].

    torch.float32      : fp32 [taps][Cin][Cout]  (CUDA-core FFMA implicit GEMM)
    bfloat16 / float16 : 16-bit [taps][Cout][Cin]  (tcgen05 implicit GEMM, K-major B operand)
    SPLIT              : fp16 [taps][Cout][2 Cin] = hi | lo (tcgen05 with split operands: the tensor-core fp32-parity mode)
    Rebuilt whenever the source parameter's version counter changes (optimizer step / load).
    """

    def __init__(self, weight: torch.Tensor, bias: Optional[torch.Tensor], ndim: int, w_dtype: torch.dtype,
                 subpixel: bool = False, dgrad: bool = False, circular: bool = False):
        """subpixel=True (bf16 only): pack for the phase-decomposed conv(nearest_up2(x)) of the tcgen05 UpSampler path.
        dgrad=True: the weights of the data-gradient convolution dX = conv_same(dY, flip(W)^T) (Cin and Cout exchanged)."""
        assert not subpixel or (is_tc_dtype(w_dtype) and int(weight.shape[-1]) == 3)
        assert not (dgrad and (subpixel or bias is not None))
        self.subpixel, self.dgrad = subpixel, dgrad
        self.circular = bool(circular)   # circular padding on every spatial axis (CircularConv2d/3d, commonlayers.py:918-1032)
        self.weight, self.bias, self.ndim = weight, bias, ndim
        self.cout, self.cin = int(weight.shape[0]), int(weight.shape[1])
        if dgrad:
            self.cout, self.cin = self.cin, self.cout
        self.ksize = int(weight.shape[-1])
        self.taps = self.ksize ** ndim
        self.w_dtype = w_dtype
        self._packed = None
        self._version = None

    def packed(self) -> torch.Tensor:
        w = self.weight
        key = (w._version, w.data_ptr(), _weight_epoch[0])
        if self._packed is None or self._version != key or self._packed.device != w.device:
            require_cuda(w, "conv weight")
            with torch.inference_mode(False), torch.no_grad():
                return self._repack(w, key)
        return self._packed

    def _repack(self, w, key):
        if True:
            src = w.detach().float().contiguous()
            ntap = (4 ** self.ndim) if self.subpixel else self.taps
            tc = is_tc_dtype(self.w_dtype)
            split = self.w_dtype == SPLIT
            rows = 16 if (tc and self.cout <= 16 and not self.dgrad) else self.cout   # convout: zero-padded to N = 16
            if self._packed is None or self._packed.device != w.device:
                self._packed = torch.empty(ntap * self.cin * rows * (2 if split else 1),
                                           dtype=torch.float16 if split else self.w_dtype, device=w.device)
            if self.subpixel:
                check(lib.dsk_pack_upconv_weight(ptr(src), ptr(self._packed), self.cout, self.cin, self.ndim, w_code(self.w_dtype),
                                                 stream()))
            elif self.dgrad:   # reference weight is [Cout_w = self.cin, Cin_w = self.cout, taps]
                check(lib.dsk_pack_conv_weight_dgrad(ptr(src), ptr(self._packed), self.cin, self.cout, self.taps,
                                                     w_code(self.w_dtype), stream()))
            else:
                check(lib.dsk_pack_conv_weight(ptr(src), ptr(self._packed), self.cout, self.cin, self.taps,
                                               w_code(self.w_dtype), stream()))
            self._version = key
        return self._packed


class MultiPacker:
    """Re-packs many PackedConv weights in ONE launch (dsk_pack_conv_weights_multi): the training graphs re-pack every
    convolution's forward and data-gradient layouts after each optimizer step.  Weights the tiled kernel does not take
    (sub-pixel up-convolutions, few-channel layers, fp32 layouts) keep their own PackedConv.packed() path."""

    def __init__(self, packs):
        self.multi, self.rest = [], []
        for pc in packs:
            ok = (isinstance(pc, PackedConv) and is_tc_dtype(pc.w_dtype) and not pc.subpixel and pc.taps <= 27 and
                  (pc.dgrad or pc.cout > 16) and pc.cout * pc.cin * pc.taps >= 16384 and pc.weight.dtype == torch.float32 and
                  pc.weight.is_contiguous())
            (self.multi if ok else self.rest).append(pc)
        self._table = None
        self._sig = None
        self._total = 0

    def _build(self):
        import struct
        dev = self.multi[0].weight.device
        raw, block0 = b"", 0
        for pc in self.multi:
            split = pc.w_dtype == SPLIT
            n = pc.taps * pc.cin * pc.cout * (2 if split else 1)
            if pc._packed is None or pc._packed.device != dev or pc._packed.numel() != n:
                pc._packed = torch.empty(n, dtype=torch.float16 if split else pc.w_dtype, device=dev)
            # reference weight [Cout_w, Cin_w, taps]; a dgrad pack swaps the roles (PackedConv stores cout / cin swapped)
            co, ci = (pc.cin, pc.cout) if pc.dgrad else (pc.cout, pc.cin)
            raw += struct.pack("2Q6i", pc.weight.data_ptr(), pc._packed.data_ptr(), co, ci, pc.taps, int(pc.dgrad),
                               w_code(pc.w_dtype), block0)
            block0 += (ci if pc.dgrad else co) * (((co if pc.dgrad else ci) + 63) // 64)
        self._table = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(dev)
        self._total = block0
        self._sig = tuple((pc.weight.data_ptr(), pc._packed.data_ptr()) for pc in self.multi)

    def pack(self):
        for pc in self.rest:
            pc.packed()
        if not self.multi:
            return
        keys = [(pc.weight._version, pc.weight.data_ptr(), _weight_epoch[0]) for pc in self.multi]
        if all(pc._packed is not None and pc._version == k for pc, k in zip(self.multi, keys)):
            return
        require_cuda(self.multi[0].weight, "conv weight")
        with torch.inference_mode(False), torch.no_grad():
            if self._table is None or self._sig != tuple((pc.weight.data_ptr(), 0 if pc._packed is None else pc._packed.data_ptr())
                                                         for pc in self.multi):
                self._build()
            check(lib.dsk_pack_conv_weights_multi(ptr(self._table), len(self.multi), self._total, stream()))
        for pc, k in zip(self.multi, keys):
            pc._version = k


def _conv_desc_of(x, pc, out, residual, up2, out_nchw, D, H, W):
    # res_dtype = 1 (DSK_RES_F32): an fp32 residual added into a 16-bit output (the operand copy an fp32-storage mode writes
    # when the block's output is read by one convolution only)
    res_f32 = residual is not None and not out_nchw and residual.dtype == torch.float32 and out.dtype != torch.float32
    return L.ConvDesc(x.shape[0], D, H, W, pc.cin, pc.cout, pc.ksize, pc.ndim, int(up2), w_code(pc.w_dtype), act_code(x, pc.cin),
                      dt_code(residual.dtype if (out_nchw and residual is not None) else
                              (torch.float32 if out_nchw else out.dtype)), int(out_nchw), int(pc.circular), int(res_f32))


def conv_pad_ws_bytes(x_shape, x_dtype, pc: "PackedConv", up2: bool = False) -> int:
    """Bytes of the halo-padded input copy a circular convolution needs (0: zero padding, or a wrapping CUDA-core kernel)."""
    if not pc.circular:
        return 0
    B, D, H, W, Cin = x_shape
    if up2:
        D, H, W = (D * 2 if pc.ndim == 3 else D), H * 2, W * 2
    code = L.SPLIT_F16 if (x_dtype == torch.float16 and Cin == 2 * pc.cin) else dt_code(x_dtype)
    d = L.ConvDesc(B, D, H, W, pc.cin, pc.cout, pc.ksize, pc.ndim, int(up2), w_code(pc.w_dtype), code, code, 0, 1)
    return int(lib.dsk_conv_pad_ws_bytes(C.byref(d)))


def pad_circular(x: torch.Tensor, ndim: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[B, D, H, W, C] -> [B, D+2 (3-D only), H+2, W+2, C], wrapped by one pixel per spatial axis (dsk_pad_circular)."""
    require_cuda(x, "pad input")
    B, D, H, W, Cc = x.shape
    if out is None:
        out = torch.empty((B, D + 2 if ndim == 3 else D, H + 2, W + 2, Cc), dtype=x.dtype, device=x.device)
    check(lib.dsk_pad_circular(ptr(x), ptr(out), B, D, H, W, Cc, ndim, dt_code(x.dtype), stream()))
    return out


def conv_stats_supported(x_shape, x_dtype, pc: "PackedConv", up2: bool = False, out_dtype=None) -> bool:
    """Can the convolution of an input of this shape emit fused norm statistics (dsk_conv_stats_supported)?  A split input is
    recognised by its doubled channel count (fp16, last dimension 2 * pc.cin); out_dtype defaults to the input dtype."""
    B, D, H, W, Cin = x_shape
    if up2:
        D, H, W = (D * 2 if pc.ndim == 3 else D), H * 2, W * 2
    code = L.SPLIT_F16 if (x_dtype == torch.float16 and Cin == 2 * pc.cin) else dt_code(x_dtype)
    d = L.ConvDesc(B, D, H, W, pc.cin, pc.cout, pc.ksize, pc.ndim, int(up2), w_code(pc.w_dtype), code,
                   dt_code(out_dtype or x_dtype), 0)
    return bool(lib.dsk_conv_stats_supported(C.byref(d)))


def conv_stats_buffer(B: int, cout: int, device) -> torch.Tensor:
    """[B, slots, Cout, 2] fp32 buffer for the statistics a convolution epilogue leaves for the following norm."""
    return torch.empty((B, int(lib.dsk_conv_stats_slots()), cout, 2), dtype=torch.float32, device=device)


def conv(x: torch.Tensor, pc: PackedConv, out: Optional[torch.Tensor] = None, chan_bias: Optional[torch.Tensor] = None,
         residual: Optional[torch.Tensor] = None, up2: bool = False, out_dtype: Optional[torch.dtype] = None,
         out_nchw: bool = False, stats: Optional[torch.Tensor] = None, pad_ws=None, prepadded: bool = False,
         operand16: Optional[torch.dtype] = None) -> torch.Tensor:
    """y = conv_same(x) + bias + chan_bias[b, :] + residual  (dsk_conv_fwd).  `stats` (conv_stats_buffer): also leave the
    per-(sample, channel) statistics of y for norm_act(..., conv_stats=stats) (dsk_conv_fwd_stats).
    pc.circular: circular instead of zero padding (dsk_conv_fwd_circ); `pad_ws` (tensor, or callable returning one) is the
    preallocated workspace of the padded copy the tcgen05 path reads -- allocated here if missing (not graph-safe).
    prepadded: x IS the halo-padded tensor [B, D+2 (3-D), H+2, W+2, Cin] (norm_apply_padded): no padding pass.
    operand16 (fp32 x, fp32 weights, fp32 out, Cin <= 4: the first layer of an fp32-storage mode): run on the tensor cores
    (dsk_conv_desc.operand16) where the im2col kernel takes the shape -- a 16-bit dtype: operands rounded to it; SPLIT: fp16
    operands split hi + lo inside the im2col row, an fp32-class result."""
    require_cuda(x, "conv input")
    B, D, H, W, Cin = x.shape
    if prepadded:
        assert pc.circular and pc.w_dtype in H16, "a pre-padded input is the layout of circular tcgen05 convolutions"
        D, H, W = (D - 2 if pc.ndim == 3 else D), H - 2, W - 2
    split_in = x.dtype == torch.float16 and Cin == 2 * pc.cin          # split-fp16 activations (hi | lo)
    assert Cin == pc.cin or split_in, (Cin, pc.cin)
    assert not split_in or pc.w_dtype in (SPLIT, torch.float16), "split activations need split or fp16 weights"
    assert up2 == pc.subpixel or pc.w_dtype == torch.float32, "16-bit weights of an up2 conv must be sub-pixel packed"
    if up2:
        D, H, W = (D * 2 if pc.ndim == 3 else D), H * 2, W * 2
    out_dtype = out_dtype or (torch.float32 if split_in else x.dtype)
    if out is None:
        shape = (B, pc.cout, D, H, W) if out_nchw else (B, D, H, W, pc.cout)
        if out_nchw and pc.ndim == 2:
            shape = (B, pc.cout, H, W)
        out = torch.empty(shape, dtype=torch.float32 if out_nchw else out_dtype, device=x.device)
    d = _conv_desc_of(x, pc, out, residual, up2, out_nchw, D, H, W)
    if operand16 is not None and x.dtype == torch.float32 and pc.w_dtype == torch.float32 and out.dtype == torch.float32:
        d.operand16 = w_code(operand16)          # torch.float16 / torch.bfloat16: rounded operands; SPLIT: exact (hi + lo rows)
    bias = pc.bias.detach() if pc.bias is not None else None
    if prepadded:
        d.circular = 2
    if pc.circular:
        need = int(lib.dsk_conv_pad_ws_bytes(C.byref(d)))
        ws = pad_ws() if callable(pad_ws) else pad_ws
        if need > 0 and (ws is None or ws.numel() * ws.element_size() < need):
            ws = torch.empty(need, dtype=torch.uint8, device=x.device)
        check(lib.dsk_conv_fwd_circ(C.byref(d), ptr(x), ptr(pc.packed()), ptr(bias), ptr(chan_bias), ptr(residual), ptr(out),
                                    ptr(stats), ptr(ws) if need > 0 else None, stream()))
        return out
    if stats is not None:
        check(lib.dsk_conv_fwd_stats(C.byref(d), ptr(x), ptr(pc.packed()), ptr(bias), ptr(chan_bias), ptr(residual), ptr(out),
                                     ptr(stats), stream()))
    else:
        check(lib.dsk_conv_fwd(C.byref(d), ptr(x), ptr(pc.packed()), ptr(bias), ptr(chan_bias), ptr(residual), ptr(out),
                               stream()))
    return out


def upsample2x(x: torch.Tensor, ndim: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Nearest x2 upsample (bf16 channels-last), input of the tcgen05 UpSampler conv (dsk_upsample2x)."""
    require_cuda(x, "upsample input")
    B, D, H, W, Cc = x.shape
    if out is None:
        out = torch.empty((B, D * 2 if ndim == 3 else D, H * 2, W * 2, Cc), dtype=x.dtype, device=x.device)
    check(lib.dsk_upsample2x(ptr(x), ptr(out), B, D, H, W, Cc, ndim, dt_code(x.dtype), stream()))
    return out


def gemm(A: torch.Tensor, Bm: torch.Tensor, out: torch.Tensor, *, M: int, N: int, K: int, lda: int, ldb: int, ldc: int,
         bias: Optional[torch.Tensor] = None, transB: bool = True, alpha: float = 1.0, act: int = 0, batch: int = 1,
         strideA: int = 0, strideB: int = 0, strideC: int = 0, a_off: int = 0, b_off: int = 0) -> torch.Tensor:
    """Batched fp32 GEMM on raw buffers with element offsets (dsk_gemm_f32)."""
    require_cuda(A, "gemm A")
    es = 4
    a = C.c_void_p(A.data_ptr() + a_off * es)
    b = C.c_void_p(Bm.data_ptr() + b_off * es)
    check(lib.dsk_gemm_f32(a, b, ptr(out), ptr(bias), M, N, K, lda, ldb, ldc, strideA, strideB, strideC, batch,
                           int(transB), alpha, act, stream()))
    return out


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], act: int = 0,
           out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y = act(x W^T + b) for fp32 x [M, K], W [N, K]."""
    M, K = x.shape
    N = weight.shape[0]
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=x.device)
    return gemm(x, weight.detach(), out, M=M, N=N, K=K, lda=K, ldb=K, ldc=N,
                bias=bias.detach() if bias is not None else None, transB=True, act=act)


def norm_act(x: torch.Tensor, gamma, beta, G: int, mode: int, silu: bool, out: Optional[torch.Tensor] = None,
             film_scale=None, film_shift=None, out_dtype: Optional[torch.dtype] = None, ws=None,
             conv_stats: Optional[torch.Tensor] = None, table_only: bool = False) -> torch.Tensor:
    """Group LayerNorm (mode 0) / RMS norm (mode 1) + affine (+FiLM) + SiLU  (dsk_norm_act).  `conv_stats`: statistics the
    convolution that produced x left behind (ops.conv(..., stats=...)): the statistics pass over x is skipped."""
    require_cuda(x, "norm input")
    B, Cc = x.shape[0], x.shape[-1]
    S = x.numel() // (B * Cc)
    if table_only:       # statistics + folded scale/shift table into `ws` only; norm_apply_padded writes the output
        assert ws is not None
        out = None
    elif out is None:
        out = torch.empty(x.shape, dtype=out_dtype or x.dtype, device=x.device)
    if ws is None:
        ws = torch.empty(int(lib.dsk_norm_ws_bytes(B, S, Cc)), dtype=torch.uint8, device=x.device)
    g = gamma.detach() if gamma is not None else None
    b = beta.detach() if beta is not None else None
    ocode = dt_code(x.dtype) if out is None else act_code(out, Cc)     # `out` [.., 2C] fp16: split-fp16 output (fp32 input)
    if conv_stats is not None:
        check(lib.dsk_norm_act_prestat(ptr(x), ptr(out), ptr(g), ptr(b), ptr(film_scale), ptr(film_shift), ptr(conv_stats),
                                       conv_stats.shape[1], ptr(ws), B, S, Cc, G, mode, int(silu), dt_code(x.dtype), ocode,
                                       stream()))
        return out
    check(lib.dsk_norm_act(ptr(x), ptr(out), ptr(g), ptr(b), ptr(film_scale), ptr(film_shift), ptr(ws), B, S, Cc, G,
                           mode, int(silu), dt_code(x.dtype), ocode, stream()))
    return out


def split_f16(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 [.., C] -> split-fp16 [.., 2C] (hi | lo): the operand form of the split tensor-core kernels (dsk_split_f16)."""
    require_cuda(x, "split input")
    assert x.dtype == torch.float32 and x.is_contiguous()
    Cc = x.shape[-1]
    if out is None:
        out = torch.empty(x.shape[:-1] + (2 * Cc,), dtype=torch.float16, device=x.device)
    assert out.dtype == torch.float16 and out.numel() == 2 * x.numel()
    check(lib.dsk_split_f16(ptr(x), ptr(out), x.numel() // Cc, Cc, stream()))
    return out


def norm_apply_padded(x: torch.Tensor, ws: torch.Tensor, out: torch.Tensor, ndim: int, silu: bool = True) -> torch.Tensor:
    """The apply pass of norm_act(..., table_only=True) writing the halo-padded input layout of a circular tcgen05
    convolution: out [B, D+2 (3-D), H+2, W+2, C] = wrap(act(x * scale + shift))  (dsk_norm_apply_padded)."""
    require_cuda(x, "norm input")
    B, D, H, W, Cc = x.shape
    assert tuple(out.shape) == (B, D + 2 if ndim == 3 else D, H + 2, W + 2, Cc), (out.shape, x.shape)
    check(lib.dsk_norm_apply_padded(ptr(x), ptr(out), ptr(ws), B, D, H, W, Cc, ndim, int(silu), dt_code(x.dtype),
                                    dt_code(out.dtype), stream()))
    return out


def norm_ws(B: int, S: int, Cc: int, device) -> torch.Tensor:
    return torch.empty(int(lib.dsk_norm_ws_bytes(B, S, Cc)), dtype=torch.uint8, device=device)


def pool2x(x: torch.Tensor, ndim: int, is_max: bool, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    require_cuda(x, "pool input")
    B, D, H, W, Cc = x.shape
    if out is None:
        out = torch.empty((B, D // 2 if ndim == 3 else 1, H // 2, W // 2, Cc), dtype=x.dtype, device=x.device)
    if x.dtype == torch.float32 and Cc % 4 == 0:      # float4 kernel; `out` may be a 16-bit operand copy (fp32-storage modes)
        check(lib.dsk_pool2x_f32(ptr(x), ptr(out), B, D, H, W, Cc, ndim, int(is_max), dt_code(out.dtype), stream()))
        return out
    assert out.dtype == x.dtype
    check(lib.dsk_pool2x(ptr(x), ptr(out), B, D, H, W, Cc, ndim, int(is_max), dt_code(x.dtype), stream()))
    return out


def add(a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    require_cuda(a, "add input")
    if out is None:
        out = torch.empty_like(a)
    check(lib.dsk_add(ptr(a), ptr(b), ptr(out), a.numel(), dt_code(a.dtype), stream()))
    return out


def cast(x: torch.Tensor, dtype: torch.dtype, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    require_cuda(x, "cast input")
    if out is None:
        out = torch.empty(x.shape, dtype=dtype, device=x.device)
    check(lib.dsk_cast(ptr(x), ptr(out), x.numel(), dt_code(x.dtype), dt_code(out.dtype), stream()))
    return out


def concat_channels(a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    require_cuda(a, "concat input")
    Ca, Cb = a.shape[-1], b.shape[-1]
    rows = a.numel() // Ca
    if out is None:
        out = torch.empty(a.shape[:-1] + (Ca + Cb,), dtype=a.dtype, device=a.device)
    check(lib.dsk_concat_channels(ptr(a), ptr(b), ptr(out), rows, Ca, Cb, dt_code(a.dtype), stream()))
    return out


def nchw_to_cl(x: torch.Tensor, dtype: torch.dtype, ndim: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 [B, C, *S] -> channels-last [B, D, H, W, C]."""
    require_cuda(x, "input")
    x = x.contiguous()
    B, Cc = x.shape[0], x.shape[1]
    sp = tuple(x.shape[2:])
    if ndim == 2:
        sp = (1,) + sp
    S = sp[0] * sp[1] * sp[2]
    if out is None:
        out = torch.empty((B,) + sp + (Cc,), dtype=dtype, device=x.device)
    check(lib.dsk_nchw_to_cl(ptr(x), ptr(out), B, Cc, S, dt_code(dtype), stream()))
    return out


def cl_to_nchw(x: torch.Tensor, ndim: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    require_cuda(x, "input")
    B, D, H, W, Cc = x.shape
    if out is None:
        out = torch.empty((B, Cc, H, W) if ndim == 2 else (B, Cc, D, H, W), dtype=torch.float32, device=x.device)
    check(lib.dsk_cl_to_nchw(ptr(x), ptr(out), B, Cc, D * H * W, dt_code(x.dtype), stream()))
    return out


def fourier(t: torch.Tensor, W: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    require_cuda(t, "time input")
    B, half = t.shape[0], W.shape[0]
    if out is None:
        out = torch.empty((B, 2 * half), dtype=torch.float32, device=t.device)
    check(lib.dsk_fourier(ptr(t), ptr(W), ptr(out), B, half, stream()))
    return out


def _ptr_table(ts, dev) -> torch.Tensor:
    return torch.tensor([0 if t is None else t.data_ptr() for t in ts], dtype=torch.int64, device=dev)


def _gemm_table(descs, dev) -> torch.Tensor:
    """Device table of GroupedGemmDesc {A, B, C, bias, Z, M, N, K, lda, ldb, ldc} (csrc/gemm_ffma.cu)."""
    import struct
    raw = b"".join(struct.pack("5Q6i", *(0 if t is None else t.data_ptr() for t in d[:5]), *d[5:]) for d in descs)
    return torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(dev)


class GroupedLinear:
    """One launch for many small fp32 linears y_g = act(x_g W_g^T + b_g) (built once; graph-capturable).  `zs` (optional):
    buffers that receive the pre-activations (training).  Small batches run one warp per output feature
    (dsk_grouped_linear); from GEMM_MIN_B rows on, a grouped smem-tiled GEMM (dsk_grouped_gemm_f32)."""
    GEMM_MIN_B = 32

    def __init__(self, xs, ws, bs, ys, act: int, zs=None):
        dev = ws[0].device
        self.keep = (xs, ws, bs, ys, zs)
        self.xs, self.ws = xs, ws
        self.B = int(xs[0].shape[0])
        self.X, self.W, self.Bi, self.Y = (_ptr_table(t, dev) for t in (xs, ws, bs, ys))
        self.Z = _ptr_table(zs, dev) if zs is not None else None
        self.in_dim = torch.tensor([w.shape[1] for w in ws], dtype=torch.int32, device=dev)
        self.out_dim = torch.tensor([w.shape[0] for w in ws], dtype=torch.int32, device=dev)
        self.max_out = max(int(w.shape[0]) for w in ws)
        self.max_in = max(int(w.shape[1]) for w in ws)
        self.n = len(ws)
        self.act = act
        self.sig = tuple(int(w.data_ptr()) for w in ws)
        self.gemm = self.B >= self.GEMM_MIN_B
        if self.gemm:
            B = self.B
            self.fwd_table = _gemm_table([(x, w, y, b, None if zs is None else zs[i], B, w.shape[0], w.shape[1], w.shape[1],
                                           w.shape[1], w.shape[0]) for i, (x, w, b, y) in enumerate(zip(xs, ws, bs, ys))], dev)

    def run(self):
        if self.gemm:
            check(lib.dsk_grouped_gemm_f32(ptr(self.fwd_table), self.n, self.B, self.max_out, 0, 1, self.act, stream()))
            return
        check(lib.dsk_grouped_linear(ptr(self.X), ptr(self.W), ptr(self.Bi), ptr(self.Y), ptr(self.Z), ptr(self.in_dim),
                                     ptr(self.out_dim), self.n, self.max_out, self.B, self.act, stream()))

    def backward_tables(self, dys, dzs, dws, dbs, dxs, shared_dx: bool = False, accumulate_dx: bool = False):
        """Bind the gradient buffers of this layer; returns the launch closure (dsk_grouped_linear_bwd, or for large
        batches dsk_grouped_dz_bias + two dsk_grouped_gemm_f32 [+ a fixed-order sum over the groups for a shared input])."""
        dev = self.W.device
        keep = [dys, dzs, dws, dbs, dxs]
        dY, dZ, dW, dB = (_ptr_table(t, dev) for t in (dys, dzs, dws, dbs))
        dX = _ptr_table(dxs, dev) if dxs is not None else None
        B = self.B
        if not self.gemm:
            def run():
                _ = keep
                check(lib.dsk_grouped_linear_bwd(ptr(dY), ptr(self.Z), ptr(self.X), ptr(self.W), ptr(dZ), ptr(dW), ptr(dB), ptr(dX),
                                                 ptr(self.in_dim), ptr(self.out_dim), self.n, self.max_out, self.max_in, B,
                                                 self.act, int(shared_dx), int(accumulate_dx), stream()))
            return run
        assert not accumulate_dx
        # dW_g [N, K] = dZ_g^T X_g : A = dZ_g stored [B][N] (transA), B = X_g stored [B][K]
        wtab = _gemm_table([(dz, x, dw, None, None, w.shape[0], w.shape[1], B, w.shape[0], w.shape[1], w.shape[1])
                            for dz, x, dw, w in zip(dzs, self.xs, dws, self.ws)], dev)
        xtab = part = None
        if dxs is not None:
            K = self.max_in
            if shared_dx:                       # per-group partials dZ_g W_g, then a fixed-order sum over the groups
                part = torch.empty((self.n, B * K), dtype=torch.float32, device=dev)
                outs = [part[i] for i in range(self.n)]
            else:
                outs = dxs
            xtab = _gemm_table([(dz, w, o, None, None, B, w.shape[1], w.shape[0], w.shape[0], w.shape[1], w.shape[1])
                                for dz, w, o in zip(dzs, self.ws, outs)], dev)
        keep += [wtab, xtab, part]

        def run():
            _ = keep
            check(lib.dsk_grouped_dz_bias(ptr(dY), ptr(self.Z), ptr(dZ), ptr(dB), ptr(self.out_dim), self.n, self.max_out, B, self.act,
                                          stream()))
            check(lib.dsk_grouped_gemm_f32(ptr(wtab), self.n, self.max_out, self.max_in, 1, 0, 0, stream()))
            if xtab is not None:
                check(lib.dsk_grouped_gemm_f32(ptr(xtab), self.n, B, self.max_in, 0, 0, 0, stream()))
                if part is not None:
                    colsum(part, dxs[0].view(-1))
        return run


def softmax_rows(S: torch.Tensor, rows: int, cols: int) -> torch.Tensor:
    require_cuda(S, "scores")
    check(lib.dsk_softmax_rows(ptr(S), rows, cols, stream()))
    return S


def self_attention_f32(tok: torch.Tensor, in_w, in_b, out_w, out_b, bufs: dict, residual: bool) -> torch.Tensor:
    """nn.MultiheadAttention(C, 1 head) self-attention on fp32 tokens [B, L, C] (nets/attention.py:54-72).

    fp32-parity path: packed QKV projection GEMM, batched QK^T, row softmax, batched PV, output
    projection -- all through dsk_gemm_f32 / dsk_softmax_rows.  `bufs` holds preallocated scratch.
    """
    B, Lq, Cc = tok.shape
    qkv, sc, ao, out = bufs["qkv"], bufs["scores"], bufs["ao"], bufs["out"]
    gemm(tok, in_w.detach(), qkv, M=B * Lq, N=3 * Cc, K=Cc, lda=Cc, ldb=Cc, ldc=3 * Cc, bias=in_b.detach(), transB=True)
    gemm(qkv, qkv, sc, M=Lq, N=Lq, K=Cc, lda=3 * Cc, ldb=3 * Cc, ldc=Lq, transB=True, alpha=Cc ** -0.5, batch=B,
         strideA=Lq * 3 * Cc, strideB=Lq * 3 * Cc, strideC=Lq * Lq, a_off=0, b_off=Cc)
    softmax_rows(sc, B * Lq, Lq)
    gemm(sc, qkv, ao, M=Lq, N=Cc, K=Lq, lda=Lq, ldb=3 * Cc, ldc=Cc, transB=False, batch=B, strideA=Lq * Lq,
         strideB=Lq * 3 * Cc, strideC=Lq * Cc, b_off=2 * Cc)
    gemm(ao, out_w.detach(), out, M=B * Lq, N=Cc, K=Cc, lda=Cc, ldb=Cc, ldc=Cc, bias=out_b.detach(), transB=True)
    if residual:
        add(out, tok, out)
    return out


def attention_core_f32(qkv: torch.Tensor, sc: torch.Tensor, ao: torch.Tensor, B: int, Lq: int, Cc: int) -> torch.Tensor:
    """softmax(Q K^T / sqrt(C)) V per sample on fp32 packed projections qkv [B*L, 3C] -> ao [B*L, C] (the middle of
    self_attention_f32, for callers that run the projections elsewhere)."""
    gemm(qkv, qkv, sc, M=Lq, N=Lq, K=Cc, lda=3 * Cc, ldb=3 * Cc, ldc=Lq, transB=True, alpha=Cc ** -0.5, batch=B,
         strideA=Lq * 3 * Cc, strideB=Lq * 3 * Cc, strideC=Lq * Lq, a_off=0, b_off=Cc)
    softmax_rows(sc, B * Lq, Lq)
    gemm(sc, qkv, ao, M=Lq, N=Cc, K=Lq, lda=Lq, ldb=3 * Cc, ldc=Cc, transB=False, batch=B, strideA=Lq * Lq,
         strideB=Lq * 3 * Cc, strideC=Lq * Cc, b_off=2 * Cc)
    return ao


def lincomb(x=None, a0: float = 0.0, r1=None, a1: float = 0.0, r2=None, a2: float = 0.0, z=None, a3: float = 0.0,
            out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out = a0*x + a1*r1 + a2*r2 + a3*z on fp32 tensors (dsk_lincomb); None operands are skipped."""
    ref = next(t for t in (x, r1, r2, z) if t is not None)
    require_cuda(ref, "integrator state")
    ts = [None if t is None else t.float().contiguous() for t in (x, r1, r2, z)]
    if out is None:
        out = torch.empty(ref.shape, dtype=torch.float32, device=ref.device)
    check(lib.dsk_lincomb(ptr(out), out.numel(), ptr(ts[0]), a0, ptr(ts[1]), a1, ptr(ts[2]), a2, ptr(ts[3]), a3,
                          stream()))
    return out


def mask_blend(x: torch.Tensor, y: torch.Tensor, mask: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x * (1 - mask) + y * mask with torch broadcasting of the mask (dsk_mask_blend)."""
    require_cuda(x, "inpainting state")
    x = x.float().contiguous()
    y = y.to(x).contiguous()
    assert y.shape == x.shape, (y.shape, x.shape)
    m = mask.to(x)
    if not (m.ndim <= x.ndim and tuple(x.shape[x.ndim - m.ndim:]) == tuple(m.shape)):
        m = m.expand_as(x)                   # general broadcasting (e.g. a size-1 batch or channel dimension)
    m = m.contiguous()
    if out is None:
        out = torch.empty_like(x)
    check(lib.dsk_mask_blend(ptr(out), ptr(x), ptr(y), ptr(m), x.numel(), m.numel(), stream()))
    return out


def philox_normal(shape, seed: int, stream_id: int, device) -> torch.Tensor:
    """N(0,1) tensor from the library's counter-based Philox4x32-10 (dsk_philox_normal)."""
    out = torch.empty(tuple(shape), dtype=torch.float32, device=device)
    require_cuda(out, "noise")
    check(lib.dsk_philox_normal(ptr(out), out.numel(), seed & 0xFFFFFFFFFFFFFFFF, stream_id & 0xFFFFFFFF, stream()))
    return out


def dropout(x: torch.Tensor, p: float, seed: int, stream_id: int, out: Optional[torch.Tensor] = None,
            dres: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out = x * keep / (1 - p) (+ dres); the mask is a function of (seed, stream_id, element index) only (dsk_dropout), so a
    backward launch with the same pair applies the mask of its forward site."""
    require_cuda(x, "x")
    if out is None:
        out = torch.empty_like(x)
    check(lib.dsk_dropout(ptr(x), ptr(dres), ptr(out), x.numel(), float(p), seed & 0xFFFFFFFFFFFFFFFF, stream_id & 0xFFFFFFFF,
                          dt_code(x.dtype), stream()))
    return out


class PackedLinear:
    """16-bit device copy of an fp32 [N, K] weight (K-major B/A operand of dsk_gemm_tc), version tracked.  dtype: bfloat16 |
    float16 | SPLIT ([N, 2K] fp16, hi | lo)."""

    def __init__(self, weight: torch.Tensor, dtype=torch.bfloat16):
        self.weight = weight
        self.dtype = dtype
        self._packed = None
        self._version = None

    def packed(self) -> torch.Tensor:
        w = self.weight
        key = (w._version, w.data_ptr(), _weight_epoch[0])
        if self._packed is None or self._version != key or self._packed.device != w.device:
            require_cuda(w, "linear weight")
            with torch.inference_mode(False), torch.no_grad():
                if self.dtype == SPLIT:
                    if self._packed is None or self._packed.device != w.device:
                        self._packed = torch.empty((w.shape[0], 2 * w.shape[1]), dtype=torch.float16, device=w.device)
                    split_f16(w.detach().float().contiguous(), out=self._packed)
                else:
                    if self._packed is None or self._packed.device != w.device:
                        self._packed = torch.empty(w.shape, dtype=self.dtype, device=w.device)
                    cast(w.detach().float().contiguous(), self.dtype, out=self._packed)
                self._version = key
        return self._packed


def gemm_bf16_tc(A: torch.Tensor, Bm: torch.Tensor, out: torch.Tensor, *, M: int, N: int, K: int, lda: int, ldb: int,
                 ldc: int, bias: Optional[torch.Tensor] = None, bias_rows: bool = False,
                 residual: Optional[torch.Tensor] = None, alpha: float = 1.0, batch: int = 1, strideA: int = 0,
                 strideB: int = 0, strideC: int = 0, a_off: int = 0, b_off: int = 0, c_off: int = 0, transA: bool = False,
                 transB: bool = False) -> torch.Tensor:
    """Batched bf16 tensor-core GEMM C = alpha op(A) op(B)^T + bias (+ residual) on raw buffers (dsk_gemm_bf16_tc).
    transA / transB: the operand is stored [K, M] / [K, N].  Offsets are in elements."""
    require_cuda(A, "gemm A")
    assert A.dtype in H16 and Bm.dtype == A.dtype
    a = C.c_void_p(A.data_ptr() + a_off * 2)
    b = C.c_void_p(Bm.data_ptr() + b_off * 2)
    c = C.c_void_p(out.data_ptr() + c_off * out.element_size())
    rf32 = residual is not None and residual.dtype == torch.float32
    r = None if residual is None else C.c_void_p(residual.data_ptr() + c_off * residual.element_size())
    check(lib.dsk_gemm_tc(a, b, c, ptr(bias), int(bias_rows), r, int(rf32), M, N, K, lda, ldb, ldc, strideA, strideB, strideC,
                          batch, alpha, int(out.dtype == torch.float32), int(transA), int(transB), dt_code(A.dtype), 0, 0, stream()))
    return out


def gemm_split_tc(A: torch.Tensor, Bm: torch.Tensor, out: torch.Tensor, *, M: int, N: int, K: int, lda: int, ldb: int, ldc: int,
                  a_lo: int, b_lo: int, bias: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None,
                  alpha: float = 1.0, batch: int = 1, strideA: int = 0, strideB: int = 0, strideC: int = 0, a_off: int = 0,
                  b_off: int = 0, transA: bool = False, transB: bool = False) -> torch.Tensor:
    """C (fp32) = alpha op(A) op(B)^T + bias (+ fp32 residual) with split-fp16 operands (dsk_gemm_tc, DSK_SPLIT_F16): the lo
    half of an operand lies a_lo / b_lo elements after its hi half along the operand's contiguous coordinate; lda / ldb are
    the full row lengths.  Three tcgen05 MMAs per k-step (hi*hi + hi*lo + lo*hi) -- fp32-class products on the tensor cores."""
    require_cuda(A, "gemm A")
    assert A.dtype == torch.float16 and Bm.dtype == torch.float16 and out.dtype in (torch.float32, torch.float16)
    assert residual is None or residual.dtype == torch.float32
    a = C.c_void_p(A.data_ptr() + a_off * 2)
    b = C.c_void_p(Bm.data_ptr() + b_off * 2)
    check(lib.dsk_gemm_tc(a, b, ptr(out), ptr(bias), 0, ptr(residual), 1, M, N, K, lda, ldb, ldc, strideA, strideB, strideC, batch,
                          alpha, int(out.dtype == torch.float32), int(transA), int(transB), L.SPLIT_F16, a_lo, b_lo, stream()))
    return out


def attention_tc_buffers(B: int, Lq: int, Cc: int, device, dtype=torch.bfloat16) -> dict:
    """Scratch of self_attention_tc: packed Q|K|V, the 16-bit probabilities, the attention output and the row-statistics
    workspace of dsk_attn_softmax_qk.  (No fp32 score tensor: the scores never leave the SM.)"""
    bf = dict(dtype=dtype, device=device)
    return dict(qkv=torch.empty((B * Lq, 3 * Cc), **bf), probs=torch.empty((B, Lq, Lq), **bf), ao=torch.empty((B * Lq, Cc), **bf),
                rowstat=torch.empty(int(lib.dsk_attn_softmax_ws_bytes(B, Lq)), dtype=torch.uint8, device=device))


def attn_softmax_qk(qkv: torch.Tensor, probs: torch.Tensor, ws: torch.Tensor, B: int, Lq: int, Cc: int) -> torch.Tensor:
    """probs[b] = softmax(Q[b] K[b]^T / sqrt(C)) (bf16) from packed projections qkv [B*L, 3C] (dsk_attn_softmax_qk)."""
    q = C.c_void_p(qkv.data_ptr())
    k = C.c_void_p(qkv.data_ptr() + Cc * 2)
    check(lib.dsk_attn_softmax_qk_h16(q, k, ptr(probs), ptr(ws), Lq, Cc, 3 * Cc, 3 * Cc, Lq * 3 * Cc, Lq * 3 * Cc, B, Cc ** -0.5,
                                      dt_code(qkv.dtype), stream()))
    return probs


def attention_split_buffers(B: int, Lq: int, Cc: int, device) -> dict:
    """Scratch of self_attention_split (tensor-core fp32-parity attention): split tokens, fp32 + split packed Q|K|V, fp32
    scores, split probabilities, fp32 + split attention output."""
    h = dict(dtype=torch.float16, device=device)
    f = dict(dtype=torch.float32, device=device)
    return dict(tok_s=torch.empty((B * Lq, 2 * Cc), **h), qkv=torch.empty((B * Lq, 3 * Cc), **f),
                qkv_s=torch.empty((B * Lq, 6 * Cc), **h), scores=torch.empty((B, Lq, Lq), **f),
                probs_s=torch.empty((B, Lq, 2 * Lq), **h), ao=torch.empty((B * Lq, Cc), **f), ao_s=torch.empty((B * Lq, 2 * Cc), **h))


def self_attention_split(tok: torch.Tensor, w_in: "PackedLinear", in_b, w_out: "PackedLinear", out_b, bufs: dict,
                         out: torch.Tensor, residual: bool) -> torch.Tensor:
    """nn.MultiheadAttention(C, 1 head) on fp32 tokens [B, L, C] (nets/attention.py:54-72) at fp32-class accuracy on the tensor
    cores: every product is a split-operand tcgen05 GEMM (hi*hi + hi*lo + lo*hi, fp32 accumulate); the scores, the softmax and
    every stored tensor stay fp32, operands are re-split (dsk_split_f16 / dsk_softmax_rows_h16) in front of each GEMM."""
    B, Lq, Cc = tok.shape
    M = B * Lq
    ts = split_f16(tok.reshape(M, Cc), out=bufs["tok_s"])
    gemm_split_tc(ts, w_in.packed(), bufs["qkv"], M=M, N=3 * Cc, K=Cc, lda=2 * Cc, ldb=2 * Cc, ldc=3 * Cc, a_lo=Cc, b_lo=Cc,
                  bias=in_b.detach())
    qs = split_f16(bufs["qkv"], out=bufs["qkv_s"])                      # rows [q k v | q_lo k_lo v_lo], lo half 3C after hi
    gemm_split_tc(qs, qs, bufs["scores"], M=Lq, N=Lq, K=Cc, lda=6 * Cc, ldb=6 * Cc, ldc=Lq, a_lo=3 * Cc, b_lo=3 * Cc,
                  alpha=Cc ** -0.5, batch=B, strideA=Lq * 6 * Cc, strideB=Lq * 6 * Cc, strideC=Lq * Lq, b_off=Cc)
    check(lib.dsk_softmax_rows_h16(ptr(bufs["scores"]), ptr(bufs["probs_s"]), B * Lq, Lq, L.SPLIT_F16, stream()))
    # P V in K slices of 1024 keys accumulated through the fp32 residual operand: tcgen05 truncates when it adds into its
    # accumulators, so a 4096-long chain in one accumulator would cost more than the split operands gain (the [L, C] output is
    # tiny next to the operands)
    for k0 in range(0, Lq, 1024):
        gemm_split_tc(bufs["probs_s"], qs, bufs["ao"], M=Lq, N=Cc, K=min(1024, Lq - k0), lda=2 * Lq, ldb=6 * Cc, ldc=Cc, a_lo=Lq,
                      b_lo=3 * Cc, batch=B, strideA=Lq * 2 * Lq, strideB=Lq * 6 * Cc, strideC=Lq * Cc, a_off=k0,
                      b_off=2 * Cc + k0 * 6 * Cc, transB=True, residual=bufs["ao"] if k0 else None)
    aos = split_f16(bufs["ao"], out=bufs["ao_s"])
    gemm_split_tc(aos, w_out.packed(), out.view(M, Cc), M=M, N=Cc, K=Cc, lda=2 * Cc, ldb=2 * Cc, ldc=Cc, a_lo=Cc, b_lo=Cc,
                  bias=out_b.detach(), residual=tok.reshape(M, Cc) if residual else None)
    return out


def attn_flash_supported(Lq: int, Cc: int) -> bool:
    """Shapes dsk_attn_flash takes: the output accumulator of a 128-query tile occupies C TMEM columns next to two 128-column
    score buffers (C = 128 | 256); below 128 tokens a query tile is mostly padding and the GEMM path is kept."""
    return Cc in (128, 256) and Lq >= 128


def attn_flash(qkv: torch.Tensor, out: torch.Tensor, B: int, Lq: int, Cc: int) -> torch.Tensor:
    """out[b] = softmax(Q[b] K[b]^T / sqrt(C)) V[b] from packed 16-bit projections qkv [B*L, 3C] in one flash-style tcgen05
    kernel (dsk_attn_flash; nets/attention.py:93-102).  out: [B*L, C] in qkv's dtype or fp32, or a split-fp16 [B*L, 2C] tensor
    (hi | lo: the A operand of a split output projection)."""
    require_cuda(qkv, "attention qkv")
    assert qkv.dtype in H16 and qkv.is_contiguous() and qkv.shape[-1] == 3 * Cc
    ldo = out.shape[-1]
    if out.dtype == torch.float32:
        mode = 1
    elif out.dtype == torch.float16 and ldo == 2 * Cc:
        mode = 2
    else:
        assert out.dtype == qkv.dtype and ldo == Cc
        mode = 0
    base = qkv.data_ptr()
    check(lib.dsk_attn_flash(C.c_void_p(base), C.c_void_p(base + 2 * Cc), C.c_void_p(base + 4 * Cc), ptr(out), Lq, Cc, 3 * Cc, 3 * Cc,
                             3 * Cc, ldo, Lq * 3 * Cc, Lq * 3 * Cc, Lq * 3 * Cc, Lq * ldo, B, Cc ** -0.5, dt_code(qkv.dtype), mode,
                             stream()))
    return out


def attention_flash_buffers(B: int, Lq: int, Cc: int, device, dtype=torch.bfloat16, split: bool = False) -> dict:
    """Scratch of self_attention_flash (16-bit modes) / self_attention_flash_split (fp32 storage, split projections): packed
    16-bit Q|K|V and the attention output -- nothing of size L x L."""
    h = dict(dtype=torch.float16 if split else dtype, device=device)
    if split:
        return dict(tok_s=torch.empty((B * Lq, 2 * Cc), **h), qkv=torch.empty((B * Lq, 3 * Cc), **h),
                    ao_s=torch.empty((B * Lq, 2 * Cc), **h))
    return dict(qkv=torch.empty((B * Lq, 3 * Cc), **h), ao=torch.empty((B * Lq, Cc), **h))


def self_attention_flash(tok: torch.Tensor, w_in: "PackedLinear", in_b: torch.Tensor, w_out: "PackedLinear", out_b: torch.Tensor,
                         bufs: dict, out: torch.Tensor, residual: bool) -> torch.Tensor:
    """nn.MultiheadAttention(C, 1 head) on 16-bit tokens [B, L, C] (nets/attention.py:54-72) in three launches: packed Q|K|V
    projection, dsk_attn_flash, output projection (+ residual).  `bufs`: attention_flash_buffers."""
    B, Lq, Cc = tok.shape
    qkv, ao = bufs["qkv"], bufs["ao"]
    gemm_bf16_tc(tok, w_in.packed(), qkv, M=B * Lq, N=3 * Cc, K=Cc, lda=Cc, ldb=Cc, ldc=3 * Cc, bias=in_b.detach())
    attn_flash(qkv, ao, B, Lq, Cc)
    gemm_bf16_tc(ao, w_out.packed(), out, M=B * Lq, N=Cc, K=Cc, lda=Cc, ldb=Cc, ldc=Cc, bias=out_b.detach(),
                 residual=tok if residual else None)
    return out


def self_attention_flash_split(tok: torch.Tensor, w_in: "PackedLinear", in_b, w_out: "PackedLinear", out_b, bufs: dict,
                               out: torch.Tensor, residual: bool) -> torch.Tensor:
    """The attention block of the fp16x2 / fp16x2m modes (fp32 tokens [B, L, C] in and out): both projections are split-operand
    GEMMs (tokens and weights hi + lo, fp32 accumulate), Q|K|V are rounded ONCE to fp16 by the projection's epilogue and the
    core is dsk_attn_flash on plain fp16 operands, its output written in split form for the output projection.  Measured
    against running the core on split operands as well (oracle/split_budget.py emulation, 3-D mc = 64): network max-rel
    8.3e-4 -> 8.8e-4, L2 4.98e-4 -> 5.08e-4."""
    B, Lq, Cc = tok.shape
    M = B * Lq
    ts = split_f16(tok.reshape(M, Cc), out=bufs["tok_s"])
    gemm_split_tc(ts, w_in.packed(), bufs["qkv"], M=M, N=3 * Cc, K=Cc, lda=2 * Cc, ldb=2 * Cc, ldc=3 * Cc, a_lo=Cc, b_lo=Cc,
                  bias=in_b.detach())
    attn_flash(bufs["qkv"], bufs["ao_s"], B, Lq, Cc)
    gemm_split_tc(bufs["ao_s"], w_out.packed(), out.view(M, Cc), M=M, N=Cc, K=Cc, lda=2 * Cc, ldb=2 * Cc, ldc=Cc, a_lo=Cc, b_lo=Cc,
                  bias=out_b.detach(), residual=tok.reshape(M, Cc) if residual else None)
    return out


def self_attention_tc(tok: torch.Tensor, w_in: PackedLinear, in_b: torch.Tensor, w_out: PackedLinear, out_b: torch.Tensor,
                      bufs: dict, out: torch.Tensor, residual: bool) -> torch.Tensor:
    """nn.MultiheadAttention(C, 1 head) on bf16 tokens [B, L, C] with every product on the tensor cores
    (reference nets/attention.py:54-72):  packed Q|K|V projection, P = softmax(QK^T/sqrt(C)) written once as bf16 (two
    QK^T passes, row statistics in the epilogue of the first), O = P V with V as it lies (MN-major operand), output
    projection (+ residual) written straight into `out` ([B, L, C] bf16).  `bufs`: attention_tc_buffers."""
    B, Lq, Cc = tok.shape
    qkv, pr, ao = bufs["qkv"], bufs["probs"], bufs["ao"]
    gemm_bf16_tc(tok, w_in.packed(), qkv, M=B * Lq, N=3 * Cc, K=Cc, lda=Cc, ldb=Cc, ldc=3 * Cc, bias=in_b.detach())
    attn_softmax_qk(qkv, pr, bufs["rowstat"], B, Lq, Cc)
    gemm_bf16_tc(pr, qkv, ao, M=Lq, N=Cc, K=Lq, lda=Lq, ldb=3 * Cc, ldc=Cc, batch=B, strideA=Lq * Lq, strideB=Lq * 3 * Cc,
                 strideC=Lq * Cc, b_off=2 * Cc, transB=True)
    gemm_bf16_tc(ao, w_out.packed(), out, M=B * Lq, N=Cc, K=Cc, lda=Cc, ldb=Cc, ldc=Cc, bias=out_b.detach(),
                 residual=tok if residual else None)
    return out


# ------------------------------------------------------------------------------------------------ backward (K2)
def conv_desc(B, D, H, W, cin, cout, ksize, ndim, up2, w_dtype, in_dtype, out_dtype, circular: bool = False) -> "L.ConvDesc":
    return L.ConvDesc(B, D, H, W, cin, cout, ksize, ndim, int(up2), dt_code(w_dtype), dt_code(in_dtype), dt_code(out_dtype), 0,
                      int(circular))


def conv_wgrad_ws_bytes(desc) -> int:
    return int(lib.dsk_conv_wgrad_ws_bytes(C.byref(desc)))


def conv_wgrad(desc, x: torch.Tensor, dy: torch.Tensor, dw: torch.Tensor, ws: torch.Tensor, accumulate: bool = False):
    """dw[Cout, Cin, k..] (fp32, reference layout) (+)= sum_pixels dy * shifted x  (dsk_conv_wgrad)."""
    require_cuda(x, "wgrad input")
    assert dw.dtype == torch.float32 and dw.is_contiguous()
    assert ws.numel() * ws.element_size() >= conv_wgrad_ws_bytes(desc)
    check(lib.dsk_conv_wgrad(C.byref(desc), ptr(x), ptr(dy), ptr(dw), ptr(ws), int(accumulate), stream()))
    return dw


def bwd_ws_bytes(B: int, S: int, Cc: int) -> int:
    return int(lib.dsk_bwd_ws_bytes(B, S, Cc))


def channel_sum(dy: torch.Tensor, out: torch.Tensor, ws: Optional[torch.Tensor], per_sample: bool):
    """out[b, c] (per_sample) or out[c] = sum over the spatial (and batch) positions of channels-last dy."""
    require_cuda(dy, "channel_sum input")
    B, Cc = dy.shape[0], dy.shape[-1]
    S = dy.numel() // (B * Cc)
    check(lib.dsk_channel_sum(ptr(dy), ptr(out), ptr(ws), B, S, Cc, dt_code(dy.dtype), int(per_sample), stream()))
    return out


def colsum(x: torch.Tensor, out: torch.Tensor, ld: Optional[int] = None):
    """out[c] = sum_r x[r, c] for an fp32 matrix with row stride `ld` (default: contiguous)."""
    require_cuda(x, "colsum input")
    rows, cols = x.shape
    check(lib.dsk_colsum_f32(ptr(x), ptr(out), rows, cols, cols if ld is None else ld, stream()))
    return out


def softmax_rows_bf16(S: torch.Tensor, P: torch.Tensor, rows: int, cols: int) -> torch.Tensor:
    check(lib.dsk_softmax_rows_bf16(ptr(S), ptr(P), rows, cols, stream()))
    return P


def softmax_bwd_rows_bf16(P: torch.Tensor, dP: torch.Tensor, dS: torch.Tensor, rows: int, cols: int) -> torch.Tensor:
    """dS = P * (dP - rowsum(dP * P)): P bf16, dP fp32, dS bf16 (dsk_softmax_bwd_rows_bf16)."""
    check(lib.dsk_softmax_bwd_rows_bf16(ptr(P), ptr(dP), ptr(dS), rows, cols, stream()))
    return dS


def norm_act_bwd(x, dy, dx, gamma, beta, G: int, mode: int, silu: bool, fwd_ws, ws, dgamma=None, dbeta=None, dres=None,
                 film_scale=None, dfilm_scale=None, dfilm_shift=None):
    require_cuda(x, "norm input")
    B, Cc = x.shape[0], x.shape[-1]
    S = x.numel() // (B * Cc)
    assert x.dtype == dy.dtype == dx.dtype and (dres is None or dres.dtype == x.dtype)
    g = gamma.detach() if gamma is not None else None
    b = beta.detach() if beta is not None else None
    check(lib.dsk_norm_act_bwd(ptr(x), ptr(dy), ptr(dres), ptr(dx), ptr(g), ptr(b), ptr(film_scale), ptr(fwd_ws), ptr(dgamma),
                               ptr(dbeta), ptr(dfilm_scale), ptr(dfilm_shift), ptr(ws), B, S, Cc, G, mode, int(silu),
                               dt_code(x.dtype), stream()))
    return dx


def pool2x_bwd(x, dy, dx, ndim: int, is_max: bool, dres=None):
    require_cuda(dy, "pool grad")
    B, D, H, W, Cc = dx.shape
    check(lib.dsk_pool2x_bwd(ptr(x), ptr(dy), ptr(dres), ptr(dx), B, D, H, W, Cc, ndim, int(is_max), dt_code(dx.dtype), stream()))
    return dx


def upsample2x_bwd(dy, dx, ndim: int, dres=None):
    require_cuda(dy, "upsample grad")
    B, D, H, W, Cc = dx.shape
    check(lib.dsk_upsample2x_bwd(ptr(dy), ptr(dres), ptr(dx), B, D, H, W, Cc, ndim, dt_code(dx.dtype), stream()))
    return dx


def add_ex(a: torch.Tensor, b: Optional[torch.Tensor], out: torch.Tensor):
    """out = a (+ b), any mix of fp32 / bf16 operands (dsk_add_ex); out may alias a or b."""
    require_cuda(a, "add input")
    assert b is None or b.numel() == a.numel()
    check(lib.dsk_add_ex(ptr(a), dt_code(a.dtype), ptr(b), dt_code(b.dtype) if b is not None else 0, ptr(out), dt_code(out.dtype),
                         a.numel(), stream()))
    return out


def split_channels(dy, da, db, ra=None, rb=None):
    require_cuda(dy, "concat grad")
    ref = da if da is not None else db
    Ct = dy.shape[-1]
    Ca = da.shape[-1] if da is not None else Ct - db.shape[-1]
    rows = dy.numel() // Ct
    check(lib.dsk_split_channels(ptr(dy), ptr(ra), ptr(rb), ptr(da), ptr(db), rows, Ca, Ct - Ca, dt_code(ref.dtype), stream()))


def gemm_ex(A: torch.Tensor, Bm: torch.Tensor, out: torch.Tensor, *, M: int, N: int, K: int, lda: int, ldb: int, ldc: int,
            bias: Optional[torch.Tensor] = None, transA: bool = False, transB: bool = True, alpha: float = 1.0,
            beta: float = 0.0, act: int = 0, batch: int = 1, strideA: int = 0, strideB: int = 0, strideC: int = 0,
            a_off: int = 0, b_off: int = 0, c_off: int = 0) -> torch.Tensor:
    """Batched fp32 GEMM C = act(alpha op(A) op(B) + bias) + beta C on raw buffers with element offsets (dsk_gemm_f32_ex)."""
    require_cuda(A, "gemm A")
    a = C.c_void_p(A.data_ptr() + a_off * 4)
    b = C.c_void_p(Bm.data_ptr() + b_off * 4)
    c = C.c_void_p(out.data_ptr() + c_off * 4)
    check(lib.dsk_gemm_f32_ex(a, b, c, ptr(bias), M, N, K, lda, ldb, ldc, strideA, strideB, strideC, batch, int(transA),
                              int(transB), alpha, beta, act, stream()))
    return out


def silu_fwd(z: torch.Tensor, out: torch.Tensor):
    require_cuda(z, "silu input")
    check(lib.dsk_silu_fwd(ptr(z), ptr(out), z.numel(), stream()))
    return out


def act_fwd(z: torch.Tensor, out: torch.Tensor, act: int):
    """out = SiLU(z) (act 1) or ReLU(z) (act 2) on fp32 vectors."""
    require_cuda(z, "activation input")
    check((lib.dsk_silu_fwd if act == 1 else lib.dsk_relu_fwd)(ptr(z), ptr(out), z.numel(), stream()))
    return out


def act_bwd(z: torch.Tensor, da: torch.Tensor, out: torch.Tensor, act: int):
    require_cuda(z, "activation input")
    check((lib.dsk_silu_bwd if act == 1 else lib.dsk_relu_bwd)(ptr(z), ptr(da), ptr(out), z.numel(), stream()))
    return out


def silu_bwd(z: torch.Tensor, da: torch.Tensor, out: torch.Tensor):
    require_cuda(z, "silu input")
    check(lib.dsk_silu_bwd(ptr(z), ptr(da), ptr(out), z.numel(), stream()))
    return out


def softmax_bwd_rows(P: torch.Tensor, dP: torch.Tensor, rows: int, cols: int):
    require_cuda(P, "softmax probabilities")
    check(lib.dsk_softmax_bwd_rows(ptr(P), ptr(dP), rows, cols, stream()))
    return dP
