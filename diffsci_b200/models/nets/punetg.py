"""PUNetG -- drop-in for diffsci.models.nets.punetg.PUNetG (reference nets/punetg.py:10-416).

Same constructor signature, same ``forward(x, t, y=None)`` contract (fp32 NC(D)HW in/out) and the
same ``state_dict()`` keys/shapes, so reference checkpoints load unchanged.  The forward pass is a
sequence of hand-written sm_100a kernels (diffsci_b200.ops -> libdiffsci_b200.so) on channels-last
activations held in a per-(batch, shape, precision) plan of preallocated buffers, which makes the
whole evaluation CUDA-graph capturable (no allocation, no host sync, no pointer-table upload).

Fusions relative to the reference graph:
  * conv epilogue adds bias, the time-embedding vector y + timeblock(te)[:, :, None...] and the
    ResNet identity residual (commonlayers.py:824-833) -- no separate elementwise passes;
  * UpSampler's F.interpolate(x2, nearest) is folded into the conv's input gather and the additive
    U-Net skip (punetg.py:373) into its epilogue -- the 8x upsampled tensor never exists;
  * ``skips.append(x.clone())`` (punetg.py:363) is a buffer hand-off, not a copy;
  * the 3-layer time MLPs of all ResNet blocks run as 3 grouped launches.
"""
from __future__ import annotations

import os
from typing import Any, Optional

import torch
from torch import nn

from ... import ops
from ..._lib import require_cuda
from .layers import (AttentionParams, ConditionDrop, FourierParams, NormParams, TimeBlockParams, _Holder, make_conv)
from .punetg_config import PUNetGConfig

_NORM_MODE = {"GroupLN": 0, "GroupRMS": 1}
DEFAULT_PRECISION = "fp32"

# Precision modes of the inference plans (DESIGN.md section 2 has the measured error of each against the reference's fp32):
#   "fp32"      fp32 tensors between kernels; every tensor-core-eligible convolution / attention product runs on tcgen05 with
#               SPLIT fp16 operands (hi + lo, 3 MMAs per k-step, fp32 accumulate): fp32-class results (~4e-6 per network
#               evaluation) at a third of the 16-bit tensor-core rate.  Layers the tensor-core kernels do not take (first conv,
#               channel counts that are not multiples of 64) run the CUDA-core FFMA kernels.
#   "fp32_ffma" the same storage with every contraction on the CUDA-core FFMA kernels (the round-1 parity mode).
#   "fp16x2"    as "fp32" with plain fp16 weights (activations split only, 2 MMAs per k-step).
#   "fp16x2m"   "mixed": as "fp16x2" for the contractions with >= 128 input channels; the 64-channel full-resolution layers
#               take plain fp16 activations (1 MMA per k-step, fp32 storage).  The many deep layers contribute most of the
#               operand-rounding error and the few full-resolution ones most of the time (oracle/split_budget.py --mixed).
#   "fp16s32"   fp16 operands everywhere (1 MMA per k-step), fp32 STORAGE of every tensor between kernels (residual stream,
#               skips, conv outputs, norm statistics from the fp32 accumulators); only the attention projections stay split.
#               Measured on the full-size networks (tools/sweep_mixed_min_cin.py): the fp16 mode's error (2.7e-3 on C4) comes
#               from rounding the STORED tensors, not the operands -- with fp32 storage plain fp16 operands give 6.2e-4 on C4
#               (fp16x2m 5.4e-4) and 6.0e-4 on C5 at 0.77x / 0.79x the evaluation time of fp16x2m.
#               One exception (3-D networks): conv1's output of the FULL-RESOLUTION ResNet blocks is stored as fp16 -- it is read
#               exactly once, by the second norm, whose statistics come from the fp32 accumulators either way.  Measured on C4
#               (tools/check_mode.py, B = 8): denoiser max-rel 7.1e-4 -> 6.8e-4, rel-L2 5.6e-4 -> 5.9e-4, one evaluation
#               8.74 -> 8.29 ms (the level holds 8x the bytes of the next one; at every level: 6.8e-4 / 6.0e-4 for 8.54 ms --
#               the 16-bit epilogue costs more than the bytes save on the small levels).  DSK_Y16 = 0: fp32 everywhere,
#               1: fp16 at every level.
#   "fp16"      fp16 storage and operands, 1 MMA per k-step: the throughput mode (3 more mantissa bits than bf16).
#   "bf16"      bf16 storage and operands (the training format; fp32's exponent range).
PRECISIONS = ("fp32", "fp32_ffma", "fp16x2", "fp16x2m", "fp16s32", "fp16", "bf16")
_ACT_DTYPE = {"fp32": torch.float32, "fp32_ffma": torch.float32, "fp16x2": torch.float32, "fp16x2m": torch.float32,
              "fp16s32": torch.float32, "fp16": torch.float16, "bf16": torch.bfloat16}
_W_DTYPE = {"fp32": ops.SPLIT, "fp16x2": torch.float16, "fp16x2m": torch.float16, "fp16s32": torch.float16, "fp16": torch.float16,
            "bf16": torch.bfloat16}   # tcgen05 weight format
SPLIT_MODES = ("fp32", "fp16x2", "fp16x2m", "fp16s32")   # fp32 storage; tensor-core operands are 16-bit COPIES: split-fp16
#                                                          (hi | lo) tensors, or plain fp16 below the mode's split_min_cin
_ATTN_FLASH = os.environ.get("DSK_ATTN_FLASH", "1") != "0"   # 0: keep the multi-launch GEMM attention (A/B measurements)
MIXED_MIN_CIN = 128                              # "fp16x2m": contractions with fewer input channels take plain fp16 activations


def split_min_cin(precision: str) -> int:
    """Input-channel count from which a contraction of this mode takes split (hi + lo) activations."""
    return MIXED_MIN_CIN if precision == "fp16x2m" else (1 << 30) if precision == "fp16s32" else 0


def _tc_eligible(cin: int, cout: int, ksize: int = 3, few_out_ok: bool = False) -> bool:
    """Shapes the tcgen05 implicit-GEMM kernel takes (see csrc/conv_tc.cu): Cin % 64 == 0 and Cout % 64 == 0, or -- for a
    plain conv without epilogue operands (convout) -- the few-output-channel tile Cout <= 16."""
    import diffsci_b200
    return (diffsci_b200.TC_CONV_ENABLED and ksize == 3 and cin % 64 == 0 and
            (cout % 64 == 0 or (few_out_ok and cout <= 16)))


class ResnetBlockParams(_Holder):
    """ResnetBlockC as PUNetG builds it (punetg.py:238-261): C_in == C_out, own time MLP."""

    def __init__(self, channels: int, embed: int, ndim: int, ksize: int, affine: bool, bias: bool,
                 convolution_type: str = "default"):
        super().__init__()
        self.channels = channels
        self.gnorm1 = NormParams(channels, affine)
        self.gnorm2 = NormParams(channels, affine)
        self.conv1 = make_conv(channels, channels, ksize, ndim, bias, convolution_type)
        self.conv2 = make_conv(channels, channels, ksize, ndim, bias, convolution_type)
        self.timeblock = TimeBlockParams(embed, channels)


class SamplerParams(_Holder):
    """DownSampler / UpSampler (commonlayers.py:25-158): one conv; pooling / upsampling is weight-free."""

    def __init__(self, cin: int, cout: int, ndim: int, ksize: int, bias: bool, convolution_type: str = "default"):
        super().__init__()
        self.conv = make_conv(cin, cout, ksize, ndim, bias, convolution_type)


class PUNetG(nn.Module):
    def __init__(self, config: PUNetGConfig, conditional_embedding: Optional[nn.Module] = None,
                 extra_residual: Optional[nn.Module] = None, *, precision: Optional[str] = None):
        super().__init__()
        c = self.config = config
        unsupported = []
        if c.convolution_type not in ("default", "circular"):
            unsupported.append(f"convolution_type={c.convolution_type!r}")
        if c.in_embedding:
            unsupported.append("in_embedding=True")
        if c.first_resblock_norm not in _NORM_MODE or c.second_resblock_norm not in _NORM_MODE:
            unsupported.append("norms other than GroupLN/GroupRMS")
        if c.attn_type != "default":
            unsupported.append(f"attn_type={c.attn_type!r}")
        if extra_residual is not None:
            unsupported.append("extra_residual")
        if c.dimension not in (2, 3) or c.transition_scale_factor != 2:
            unsupported.append("dimension not in {2,3} or transition_scale_factor != 2")
        if c.kernel_size not in (1, 3) or c.in_out_kernel_size not in (1, 3) or c.transition_kernel_size not in (1, 3):
            unsupported.append("kernel sizes other than 1/3")
        if unsupported:
            raise NotImplementedError("diffsci_b200.PUNetG: not built yet (SURVEY.md 8f): " + ", ".join(unsupported))
        self.precision = precision or DEFAULT_PRECISION
        nd, M = c.dimension, c.model_channels
        mult = c.extended_channel_expansion
        ct = c.convolution_type          # "circular": every conv pads periodically (punetg.py:221-232, commonlayers.py:84-88)

        def blocks(m, n):
            return nn.ModuleList([ResnetBlockParams(m * M, M, nd, c.kernel_size, c.affine_norm, c.bias, ct) for _ in range(n)])

        self.time_projection = FourierParams(M, c.time_projection_scale)
        self.extra_residual = None
        self.conditional_embedding = conditional_embedding      # any torch module: y -> [B, M] (punetg.py:93, 400-404)
        # bias=False: no conv carries a bias; the network input gets a constant ones channel instead (punetg.py:190-193, 390-394)
        self.ones_channel = 0 if c.bias else 1
        self.convin = make_conv(c.input_channels + self.ones_channel, M, c.in_out_kernel_size, nd, c.bias, ct)
        self.convout = make_conv(M, c.output_channels, c.in_out_kernel_size, nd, c.bias, ct)
        self.downward_blocks = nn.ModuleList([blocks(m, c.number_resnet_downward_block) for m in mult[:-1]])
        self.downsamplers = nn.ModuleList([SamplerParams(a * M, b * M, nd, c.transition_kernel_size, c.bias, ct)
                                           for a, b in zip(mult[:-1], mult[1:])])
        rev = mult[::-1]
        self.upward_blocks = nn.ModuleList([blocks(m, c.number_resnet_upward_block) for m in rev[1:]])
        self.upsamplers = nn.ModuleList([SamplerParams(a * M, b * M, nd, c.transition_kernel_size, c.bias, ct)
                                         for a, b in zip(rev[:-1], rev[1:])])
        self.before_block = blocks(mult[-1], c.number_resnet_before_attn_block)
        self.after_block = blocks(mult[-1], c.number_resnet_after_attn_block)
        self.attn_resnet_block = blocks(mult[-1], c.number_resnet_attn_block)
        self.attn_block = nn.ModuleList([AttentionParams(mult[-1] * M) for _ in range(c.number_resnet_attn_block - 1)])
        self.cond_dropout = nn.Dropout(c.cond_dropout)
        if c.cond_drop is not None and c.cond_drop > 0:
            self.cond_drop = ConditionDrop(p=c.cond_drop, hidden_dim=M, null_is_learnable=c.cond_drop_learnable)
        else:
            self.cond_drop = None
        self._plans: dict[Any, "_Plan"] = {}

    # ------------------------------------------------------------------ reference API
    def export_description(self) -> dict[str, Any]:
        emb = self.conditional_embedding
        cemb_args = emb.export_description() if getattr(emb, "export_description", None) else None
        return dict(config=self.config.export_description(), conditional_embedding_args=cemb_args,
                    has_conditional_embedding=emb is not None)

    def set_conditional_embedding(self, conditional_embedding: Optional[nn.Module] = None):
        self.conditional_embedding = conditional_embedding

    # ------------------------------------------------------------------ conditioning
    def native_parameters(self) -> list:
        """The parameters the CUDA path owns (everything except the user's embedder and ConditionDrop's null embedding,
        which act on the [B, M] conditioning vector before it enters the network)."""
        return [p for n, p in self.named_parameters()
                if not (n.startswith("conditional_embedding.") or n.startswith("cond_drop."))]

    def conditioning_vector(self, y, B: int) -> Optional[torch.Tensor]:
        """ye of punetg.py:400-410: cond_dropout(cond_drop(conditional_embedding(y))) as fp32 [B, M] (None if y is None)."""
        if y is None:
            return None
        ye = y if self.conditional_embedding is None else self.conditional_embedding(y)
        if not torch.is_tensor(ye):
            raise TypeError("diffsci_b200.PUNetG: y must be a tensor when the network has no conditional_embedding")
        if self.cond_drop is not None:
            ye = self.cond_drop(ye)
        ye = self.cond_dropout(ye)
        M = self.config.model_channels
        if ye.ndim == 1:
            ye = ye.unsqueeze(0)
        if ye.ndim != 2 or ye.shape[-1] != M or ye.shape[0] not in (1, B):
            raise NotImplementedError(f"diffsci_b200.PUNetG: conditioning of shape {tuple(ye.shape)} (only [B or 1, "
                                      f"model_channels={M}] vectors added to the time embedding are built)")
        return ye.float().expand(B, M)

    def split_condition(self, y, x: torch.Tensor):
        """-> (channel conditioning fp32 [B, Cy, *S] or None, remaining y).  PUNetG has none (PUNetGCond overrides)."""
        return None, y

    def forward(self, x: torch.Tensor, t: Optional[torch.Tensor] = None, y=None) -> torch.Tensor:
        """x: fp32 [B, Cin, *S]; t: [B] (= c_noise); returns fp32 [B, Cout, *S] (punetg.py:389-416)."""
        require_cuda(x, "PUNetG input")
        B = x.shape[0]
        if self.ones_channel:
            x = torch.cat([x, torch.ones_like(x[:, :1])], dim=1)
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            # gradients wanted (training, or eval-mode fine-tuning / diagnostics): static forward/backward launch lists with
            # hand-written backward kernels (graph.py); dropout acts only in training mode, as torch.nn.Dropout does
            from .graph import NetFunction
            ye = self.conditioning_vector(y, B)
            graph = self.train_graph(B, tuple(x.shape[2:]), x.device, cond=ye is not None, dropout=self.training)
            return NetFunction.apply(graph, x, t, None if ye is None else ye.contiguous(), *self.native_parameters())
        if self.training and float(getattr(self.config, "dropout", 0.0)) > 0.0:
            raise NotImplementedError("dropout > 0 acts on the training path (gradients enabled); call .eval() for inference")
        ye = self.conditioning_vector(y, B)
        plan = self.plan(B, tuple(x.shape[2:]), x.device)
        xin = ops.nchw_to_cl(x.float(), plan.act_dtype, self.config.dimension, out=plan.xin)
        tt = torch.zeros(B, device=x.device) if t is None else t.float().contiguous()
        out = torch.empty((B, self.config.output_channels) + tuple(x.shape[2:]), dtype=torch.float32, device=x.device)
        plan.forward(xin, tt, out_nchw=out, zero_time=t is None, ye=None if ye is None else ye.contiguous())
        return out

    # ------------------------------------------------------------------ plans
    def plan(self, B: int, spatial: tuple, device, precision: Optional[str] = None) -> "_Plan":
        precision = precision or self.precision
        key = (B, tuple(spatial), str(device), precision)
        sig = tuple(p.data_ptr() for p in self.native_parameters())
        plan = self._plans.get(key)
        if plan is None or plan.sig != sig:
            if len(self._plans) >= 4:   # bound the HBM held by cached plans
                self._plans.clear()
            with torch.inference_mode(False), torch.no_grad():   # persistent buffers must be normal tensors
                plan = self._plans[key] = _Plan(self, B, tuple(spatial), device, precision, sig)
        return plan

    def train_graph(self, B: int, spatial: tuple, device, precision: Optional[str] = None, cond: bool = False,
                    dropout: bool = True):
        """The TrainGraph (forward + backward launch lists, saved activations, flat gradient buffer) of this shape.
        cond: the graph takes a conditioning vector ye [B, M] (te + ye, punetg.py:410) and returns its gradient.
        dropout=False: the graph of an eval-mode module (config.dropout inactive)."""
        from .graph import build_punetg
        precision = precision or self.precision
        dropout = bool(dropout) and float(getattr(self.config, "dropout", 0.0)) > 0.0
        key = ("train", B, tuple(spatial), str(device), precision, bool(cond), dropout)
        sig = tuple(p.data_ptr() for p in self.native_parameters())
        g = self._plans.get(key)
        if g is None or g.sig != sig:
            for k in [k for k in self._plans if k[0] == "train"]:
                del self._plans[k]                                  # one training shape resident at a time
            with torch.inference_mode(False), torch.no_grad():
                g = self._plans[key] = build_punetg(self, B, tuple(spatial), device, precision, cond=cond, dropout=dropout)
        return g

    def _apply(self, fn, *a, **k):
        self._plans = {}
        return super()._apply(fn, *a, **k)


class _Plan:
    """Preallocated buffers + packed weights + pointer tables for one (B, shape, precision)."""

    def __init__(self, net: PUNetG, B: int, spatial: tuple, device, precision: str, sig):
        if precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {PRECISIONS}, got {precision!r}")
        c = net.config
        self.net, self.B, self.sig, self.precision = net, B, sig, precision
        self.ndim = nd = c.dimension
        self.act_dtype = adt = _ACT_DTYPE[precision]
        wdt = _W_DTYPE.get(precision)                       # None: no tensor-core kernels in this mode
        self.split = split = precision in SPLIT_MODES      # conv / GEMM inputs are split-fp16 tensors (hi | lo)
        self.split_min_cin = smin = split_min_cin(precision)
        dev = self.device = torch.device(device)
        if len(spatial) != nd:
            raise ValueError(f"PUNetG(dimension={nd}) got spatial shape {spatial}")
        M, mult = c.model_channels, c.extended_channel_expansion
        self.nlev = nlev = len(c.channel_expansion)
        sp = [(1,) + tuple(spatial) if nd == 2 else tuple(spatial)]
        for _ in range(nlev):
            d, h, w = sp[-1]
            if h % 2 or w % 2 or (nd == 3 and d % 2):
                raise ValueError(f"PUNetG: spatial size {sp[-1]} is not divisible by 2 at a down-sampling level "
                                 "(the reference fails at `x + skip` for such shapes)")
            sp.append((d // 2 if nd == 3 else 1, h // 2, w // 2))
        self.sp = sp
        ch = [m * M for m in mult]

        def buf(l, cc, dtype=adt):
            return torch.empty((B,) + sp[l] + (cc,), dtype=dtype, device=dev)

        def pack(cp, subpixel=False, few_out_ok=False):
            tc = wdt is not None and _tc_eligible(cp.cin, cp.cout, cp.ksize, few_out_ok)
            return ops.PackedConv(cp.weight, cp.bias, nd, wdt if tc else torch.float32, subpixel and tc, circular=cp.circular)

        def nbuf(l):      # norm + SiLU output == conv input: a split-fp16 tensor where the block convs run on the tensor cores
            if split and _tc_eligible(ch[l], ch[l], c.kernel_size):
                if ch[l] < smin:          # mixed mode: plain fp16 activations for the narrow layers
                    return buf(l, ch[l], torch.float16)
                return torch.empty((B,) + sp[l] + (2 * ch[l],), dtype=torch.float16, device=dev)
            return buf(l, ch[l])

        self.xin = buf(0, net.convin.cin)
        if net.ones_channel:          # constant last channel; the fused stages / scale kernels only write the state channels
            self.xin[..., -1] = 1.0
        self.X = [buf(l, ch[l]) for l in range(nlev + 1)]          # encoder / bottom state (doubles as skip)
        self.XU = [buf(l, ch[l]) for l in range(nlev)]             # decoder state
        self.N = [nbuf(l) for l in range(nlev + 1)]                # norm+SiLU output == conv input
        # conv1 output: read once, by the second norm, whose statistics come from the fp32 accumulators either way
        # fp16s32, 3-D: fp16 at the full-resolution level (default, "l0"; see the mode table above); DSK_Y16 = 0 / 1: nowhere / everywhere
        y16e = os.environ.get("DSK_Y16", "l0" if nd == 3 else "0") if precision == "fp16s32" else "0"
        y16 = [y16e == "1" or (y16e == "l0" and l == 0) for l in range(nlev + 1)]
        self.Y = [buf(l, ch[l], torch.float16 if (y16[l] and _tc_eligible(ch[l], ch[l], c.kernel_size)) else adt)
                  for l in range(nlev + 1)]
        self.P = [buf(l + 1, ch[l]) for l in range(nlev)]          # pooled (re-typed below where only a 16-bit operand is needed)
        self.XA = buf(nlev, ch[nlev])
        self.XA2 = buf(nlev, ch[nlev])
        self.F = buf(0, c.output_channels)
        self.WS = [ops.norm_ws(B, sp[l][0] * sp[l][1] * sp[l][2], ch[l], dev) for l in range(nlev + 1)]
        # circular padding: one halo-padded input copy for the tcgen05 convs (TMA boxes cannot wrap), sized for the largest
        self.pad_ws = None
        self.NP = [None] * (nlev + 1)      # norm+SiLU outputs in the halo-padded layout (the norm's apply pass writes the halos)
        if c.convolution_type == "circular" and wdt is not None:
            halo = lambda s: (s[0] + (2 if nd == 3 else 0)) * (s[1] + 2) * (s[2] + 2)  # noqa: E731
            self.pad_ws = torch.empty(max(B * halo(sp[l]) * ch[l] * (4 if split else 2) for l in range(nlev + 1)),
                                      dtype=torch.uint8, device=dev)
            for l in range(nlev + 1):
                if _tc_eligible(ch[l], ch[l], c.kernel_size) and not split:   # split mode: a padding pass in front of each conv
                    d_, h_, w_ = sp[l]
                    self.NP[l] = torch.empty((B, d_ + 2 if nd == 3 else d_, h_ + 2, w_ + 2, ch[l]), dtype=adt, device=dev)
        # fused norm statistics (conv epilogue -> following per-channel norm): one buffer per level, consumed by the very
        # next norm.  Only where the convolution kernel can emit them and the norms are per channel (G == C: the PUNetG norms).
        self.ST = [None] * (nlev + 1)
        self.pc_in, self.pc_out = pack(net.convin), pack(net.convout, few_out_ok=True)
        self.pc_down = [pack(s.conv) for s in net.downsamplers]
        self.pc_up = [pack(s.conv, subpixel=True) for s in net.upsamplers]   # conv(up2(x)) in sub-pixel form on tcgen05

        # every ResNet block in execution order, with its level
        self.blocks = []
        for l in range(nlev):
            self.blocks += [(b, l) for b in net.downward_blocks[l]]
        for grp in (net.before_block, net.attn_resnet_block, net.after_block):
            self.blocks += [(b, nlev) for b in grp]
        for i in range(nlev):
            self.blocks += [(b, nlev - 1 - i) for b in net.upward_blocks[i]]
        self.pc = {id(b): (pack(b.conv1), pack(b.conv2)) for b, _ in self.blocks}
        # fp32-storage modes: a tensor whose ONLY reader is a tensor-core convolution taking plain fp16 activations is produced
        # directly as that fp16 operand (the value is rounded exactly as the cast pass would round it; an fp32 copy plus a cast
        # launch per such tensor otherwise):  the pooled tensor (reader: the DownSampler conv) and the output of the last ResNet
        # block in front of an UpSampler conv / convout (conv2 adds the fp32 residual and writes fp16: DSK_RES_F32).
        def plain16(pc):
            return split and ops.is_tc_dtype(pc.w_dtype) and pc.cin < smin and os.environ.get("DSK_OPERAND16", "1") != "0"
        for l in range(nlev):
            if plain16(self.pc_down[l]):
                self.P[l] = buf(l + 1, ch[l], torch.float16)
        self.X16 = [None] * (nlev + 1)      # per level: fp16 output of the last block before pc_up / pc_out
        for l in range(nlev + 1):
            reader = self.pc_out if l == 0 else self.pc_up[nlev - l]
            last = (net.upward_blocks[nlev - 1 - l] if l < nlev else net.after_block)
            if plain16(reader) and len(last) > 0 and ops.is_tc_dtype(self.pc[id(last[-1])][1].w_dtype):
                self.X16[l] = buf(l, ch[l], torch.float16)
        # split mode: the convolutions NOT fed by a norm (down / up samplers, convout) read a split copy of their fp32 input
        # (dsk_split_f16) held in one scratch buffer
        self.S16 = None
        if split:
            need = [self.P[l].numel() for l in range(nlev) if ops.is_tc_dtype(self.pc_down[l].w_dtype)]
            need += [self.X[nlev - i].numel() for i in range(nlev) if ops.is_tc_dtype(self.pc_up[i].w_dtype)]
            if ops.is_tc_dtype(self.pc_out.w_dtype):
                need.append(self.X[0].numel())
            if need:
                self.S16 = torch.empty(2 * max(need), dtype=torch.float16, device=dev)
        if wdt is not None:
            def sup(shape, pc, up2=False):
                if not ops.is_tc_dtype(pc.w_dtype):
                    return False
                if split and pc.cin >= smin:
                    shape = tuple(shape[:-1]) + (2 * shape[-1],)
                return ops.conv_stats_supported(shape, torch.float16 if split else adt, pc, up2=up2, out_dtype=adt)
            lvl_pc = {l: self.pc[id(b)][0] for b, l in self.blocks}
            self.st_ok = [sup(self.X[l].shape, lvl_pc[l]) for l in range(nlev + 1)]
            self.st_down = [sup(self.P[l].shape, self.pc_down[l]) for l in range(nlev)]
            self.st_up = [sup(self.X[nlev - i].shape, self.pc_up[i], up2=True) for i in range(nlev)]
            for l in range(nlev + 1):
                if self.st_ok[l] or (l > 0 and self.st_down[l - 1]) or (l < nlev and self.st_up[nlev - 1 - l]):
                    self.ST[l] = ops.conv_stats_buffer(B, ch[l], dev)
        else:
            self.st_ok, self.st_down, self.st_up = [False] * (nlev + 1), [False] * nlev, [False] * nlev

        # time embedding: Fourier features + 3 grouped launches for all blocks' MLPs
        f32 = dict(dtype=torch.float32, device=dev)
        self.te = torch.empty((B, M), **f32)
        self.tvec = {id(b): torch.empty((B, b.channels), **f32) for b, _ in self.blocks}
        h1 = [torch.empty((B, 4 * M), **f32) for _ in self.blocks]
        h2 = [torch.empty((B, 4 * M), **f32) for _ in self.blocks]
        net_ = [b.timeblock.net for b, _ in self.blocks]
        self.tmlp = [
            ops.GroupedLinear([self.te] * len(net_), [n[0].weight for n in net_], [n[0].bias for n in net_], h1, 1),
            ops.GroupedLinear(h1, [n[2].weight for n in net_], [n[2].bias for n in net_], h2, 1),
            ops.GroupedLinear(h2, [n[4].weight for n in net_], [n[4].bias for n in net_],
                              [self.tvec[id(b)] for b, _ in self.blocks], 0),
        ]
        # attention scratch (fp32 tokens)
        Lq = sp[nlev][0] * sp[nlev][1] * sp[nlev][2]
        Cb = ch[nlev]
        self.attn = None
        self.attn_tc = (adt in ops.H16 and _tc_eligible(Cb, Cb) and Lq % 8 == 0 and Lq <= 8192)
        self.attn_split = (split and _tc_eligible(Cb, Cb) and Lq % 64 == 0 and Lq <= 8192)
        self.attn_hybrid = self.attn_hybrid_split = False
        # flash-style core (dsk_attn_flash) for every 16-bit-operand mode; the fp32 mode keeps the split-operand GEMM chain
        self.attn_flash = (_ATTN_FLASH and ops.attn_flash_supported(Lq, Cb) and
                           ((self.attn_split and precision != "fp32") or (self.attn_tc and not split)))
        if len(net.attn_block) > 0 and self.attn_flash:
            self.attn = ops.attention_flash_buffers(B, Lq, Cb, dev, adt, split=split)
            wfmt = ops.SPLIT if split else adt
            self.attn_w = [(ops.PackedLinear(a.mhattn.in_proj_weight, wfmt), ops.PackedLinear(a.mhattn.out_proj.weight, wfmt))
                           for a in net.attn_block]
        elif len(net.attn_block) > 0 and self.attn_split:
            self.attn = ops.attention_split_buffers(B, Lq, Cb, dev)
            self.attn_w = [(ops.PackedLinear(a.mhattn.in_proj_weight, ops.SPLIT), ops.PackedLinear(a.mhattn.out_proj.weight, ops.SPLIT))
                           for a in net.attn_block]
        elif len(net.attn_block) > 0 and self.attn_tc:
            self.attn = ops.attention_tc_buffers(B, Lq, Cb, dev, adt)
            self.attn_w = [(ops.PackedLinear(a.mhattn.in_proj_weight, adt), ops.PackedLinear(a.mhattn.out_proj.weight, adt))
                           for a in net.attn_block]
        elif len(net.attn_block) > 0:
            self.attn = dict(qkv=torch.empty((B * Lq, 3 * Cb), **f32), scores=torch.empty((B, Lq, Lq), **f32),
                             ao=torch.empty((B * Lq, Cb), **f32), out=torch.empty((B, Lq, Cb), **f32))
            if split and _tc_eligible(Cb, Cb) and (B * Lq) % 8 == 0:
                # fp32-storage modes at token counts the tensor-core score paths do not take (the 7 x 7 bottom level of MNIST): the
                # two projections -- 99 % of the block's FLOPs at small L -- as split-operand tcgen05 GEMMs, the L x L core in fp32
                # (the all-FFMA block was 18 % of a C2 evaluation in fp16s32)
                self.attn_hybrid_split = True
                self.attn["tok_s"] = torch.empty((B * Lq, 2 * Cb), dtype=torch.float16, device=dev)
                self.attn["ao_s"] = torch.empty((B * Lq, 2 * Cb), dtype=torch.float16, device=dev)
                self.attn_w = [(ops.PackedLinear(a.mhattn.in_proj_weight, ops.SPLIT), ops.PackedLinear(a.mhattn.out_proj.weight, ops.SPLIT))
                               for a in net.attn_block]
            if adt != torch.float32:
                self.attn["tok"] = torch.empty((B, Lq, Cb), **f32)
                if _tc_eligible(Cb, Cb) and (B * Lq) % 8 == 0:
                    self.attn_hybrid = True
                    self.attn["ao16"] = torch.empty((B * Lq, Cb), dtype=adt, device=dev)
                    self.attn_w = [(ops.PackedLinear(a.mhattn.in_proj_weight, adt), ops.PackedLinear(a.mhattn.out_proj.weight, adt))
                                   for a in net.attn_block]
        self.Lq, self.Cb = Lq, Cb
        self.prepare()

    def prepare(self):
        """Materialise packed weights (must happen outside CUDA-graph capture)."""
        for pc in [self.pc_in, self.pc_out, *self.pc_down, *self.pc_up, *[p for pair in self.pc.values() for p in pair]]:
            pc.packed()
        if (self.attn_tc or self.attn_hybrid or self.attn_hybrid_split or self.attn_split or self.attn_flash) and len(self.net.attn_block) > 0:
            for wi, wo in self.attn_w:
                wi.packed()
                wo.packed()

    def _conv(self, x, pc, **k):
        """ops.conv; in split mode a tensor-core convolution whose input is still an fp32 tensor (not the split output of a
        norm) gets its split copy first."""
        if self.split and ops.is_tc_dtype(pc.w_dtype) and x.dtype == torch.float32:
            if pc.cin < self.split_min_cin:      # mixed mode: plain fp16 operand
                x = ops.cast(x, torch.float16, out=self.S16[:x.numel()].view(x.shape))
            else:
                x = ops.split_f16(x, out=self.S16[:2 * x.numel()].view(x.shape[:-1] + (2 * x.shape[-1],)))
        return ops.conv(x, pc, pad_ws=self.pad_ws, **k)

    # ------------------------------------------------------------------ forward
    def _resblock(self, x, blk, l, out, xs=None):
        """-> (block output, its fused norm statistics or None).  `xs`: statistics of x left by the conv that produced it.
        A 16-bit `out` in an fp32-storage mode (self.X16): the block's output is only the operand of the next convolution --
        no statistics, fp32 residual added in the epilogue."""
        c = self.net.config
        C = blk.channels
        pc1, pc2 = self.pc[id(blk)]
        st = self.ST[l] if self.st_ok[l] else None
        st2 = None if (out.dtype != x.dtype) else st
        if self.NP[l] is not None:      # periodic net on the tcgen05 path: the norm's apply pass writes the padded conv input
            ops.norm_act(x, blk.gnorm1.weight, blk.gnorm1.bias, C, _NORM_MODE[c.first_resblock_norm], True, ws=self.WS[l],
                         conv_stats=xs, table_only=True)
            n = ops.norm_apply_padded(x, self.WS[l], self.NP[l], self.ndim)
            y = self._conv(n, pc1, out=self.Y[l], chan_bias=self.tvec[id(blk)], stats=st, prepadded=True)
            ops.norm_act(y, blk.gnorm2.weight, blk.gnorm2.bias, C, _NORM_MODE[c.second_resblock_norm], True, ws=self.WS[l],
                         conv_stats=st, table_only=True)
            n = ops.norm_apply_padded(y, self.WS[l], self.NP[l], self.ndim)
            return self._conv(n, pc2, out=out, residual=x, stats=st2, prepadded=True), st2
        n = ops.norm_act(x, blk.gnorm1.weight, blk.gnorm1.bias, C, _NORM_MODE[c.first_resblock_norm], True,
                         out=self.N[l], ws=self.WS[l], conv_stats=xs)
        y = self._conv(n, pc1, out=self.Y[l], chan_bias=self.tvec[id(blk)], stats=st)
        n = ops.norm_act(y, blk.gnorm2.weight, blk.gnorm2.bias, C, _NORM_MODE[c.second_resblock_norm], True,
                         out=self.N[l], ws=self.WS[l], conv_stats=st)
        return self._conv(n, pc2, out=out, residual=x, stats=st2), st2

    def _attention(self, x, attn, out, index=0):
        a = self.attn
        m = attn.mhattn
        B, Lq, Cb = self.B, self.Lq, self.Cb
        if self.attn_flash:
            wi, wo = self.attn_w[index]
            fn = ops.self_attention_flash_split if self.split else ops.self_attention_flash
            fn(x.view(B, Lq, Cb), wi, m.in_proj_bias, wo, m.out_proj.bias, a, out.view(B, Lq, Cb), self.net.config.attn_residual)
            return out
        if self.attn_split:
            wi, wo = self.attn_w[index]
            ops.self_attention_split(x.view(B, Lq, Cb), wi, m.in_proj_bias, wo, m.out_proj.bias, a, out.view(B, Lq, Cb),
                                     self.net.config.attn_residual)
            return out
        if self.attn_tc:
            wi, wo = self.attn_w[index]
            ops.self_attention_tc(x.view(B, Lq, Cb), wi, m.in_proj_bias, wo, m.out_proj.bias, a, out.view(B, Lq, Cb),
                                  self.net.config.attn_residual)
            return out
        if self.attn_hybrid_split:
            wi, wo = self.attn_w[index]
            M = B * Lq
            ts = ops.split_f16(x.view(M, Cb), out=a["tok_s"])
            ops.gemm_split_tc(ts, wi.packed(), a["qkv"], M=M, N=3 * Cb, K=Cb, lda=2 * Cb, ldb=2 * Cb, ldc=3 * Cb, a_lo=Cb, b_lo=Cb,
                              bias=m.in_proj_bias.detach())
            ops.attention_core_f32(a["qkv"], a["scores"], a["ao"], B, Lq, Cb)
            aos = ops.split_f16(a["ao"], out=a["ao_s"])
            ops.gemm_split_tc(aos, wo.packed(), out.view(M, Cb), M=M, N=Cb, K=Cb, lda=2 * Cb, ldb=2 * Cb, ldc=Cb, a_lo=Cb, b_lo=Cb,
                              bias=m.out_proj.bias.detach(), residual=x.view(M, Cb) if self.net.config.attn_residual else None)
            return out
        if x.dtype == torch.float32:
            tok = x.view(B, Lq, Cb)
        elif self.attn_hybrid:
            # token counts the tensor-core score path does not take (L % 8 != 0, e.g. the 7x7 bottom level of MNIST): the two
            # projections -- 99% of the block's FLOPs at small L -- still run on tcgen05, the tiny per-sample L x L part in fp32
            wi, wo = self.attn_w[index]
            M = B * Lq
            ops.gemm_bf16_tc(x.view(M, Cb), wi.packed(), a["qkv"], M=M, N=3 * Cb, K=Cb, lda=Cb, ldb=Cb, ldc=3 * Cb,
                             bias=m.in_proj_bias.detach())
            ops.attention_core_f32(a["qkv"], a["scores"], a["ao"], B, Lq, Cb)
            ao16 = ops.cast(a["ao"], x.dtype, out=a["ao16"])
            ops.gemm_bf16_tc(ao16, wo.packed(), out.view(M, Cb), M=M, N=Cb, K=Cb, lda=Cb, ldb=Cb, ldc=Cb,
                             bias=m.out_proj.bias.detach(), residual=x.view(M, Cb) if self.net.config.attn_residual else None)
            return out
        else:
            tok = ops.cast(x.view(B, Lq, Cb), torch.float32, out=a["tok"])
        o = ops.self_attention_f32(tok, m.in_proj_weight, m.in_proj_bias, m.out_proj.weight, m.out_proj.bias, a,
                                   self.net.config.attn_residual)
        ops.cast(o, out.dtype, out=out.view(B, Lq, Cb))
        return out

    def forward(self, xin: torch.Tensor, cnoise: torch.Tensor, out_nchw: Optional[torch.Tensor] = None,
                zero_time: bool = False, ye: Optional[torch.Tensor] = None) -> torch.Tensor:
        """xin: channels-last [B, D, H, W, Cin] (act dtype); cnoise: fp32 [B]; ye: fp32 [B, M] conditioning vector added to
        the time embedding (punetg.py:410) or None.
        Returns F channels-last (self.F), or writes fp32 NC(D)HW into `out_nchw`."""
        net, c, nlev = self.net, self.net.config, self.nlev
        if zero_time:
            self.te.zero_()       # te = zeros when t is None (punetg.py:398-399)
        else:
            ops.fourier(cnoise, net.time_projection.W, out=self.te)
        if ye is not None:
            ops.add_ex(self.te, ye, self.te)
        for g in self.tmlp:
            g.run()
        # first layer of the 16-bit-operand fp32-storage modes: im2col on the tensor cores (convin_tc) with the input AND the
        # weights split hi + lo inside the im2col row -- an fp32-class product.  (Plain fp16 operands here, DSK_CONVIN_TC16=1,
        # measured on C4: the rounding of the network input and of the 27 first-layer weights alone moves the denoiser error of
        # fp16s32 from 6.2e-4 to 7.0e-4 -- every later layer sees it.)  The fp32 mode keeps the CUDA-core kernel.
        op16 = None
        if self.split and self.precision != "fp32" and os.environ.get("DSK_CONVIN_TC", "1") != "0":
            op16 = torch.float16 if os.environ.get("DSK_CONVIN_TC16", "0") == "1" else ops.SPLIT
        x, xs = self._conv(xin, self.pc_in, out=self.X[0], operand16=op16), None
        for l in range(nlev):
            for blk in net.downward_blocks[l]:
                x, xs = self._resblock(x, blk, l, x, xs)
            p = ops.pool2x(x, self.ndim, True, out=self.P[l])            # MaxPool (commonlayers.py:60-63)
            xs = self.ST[l + 1] if self.st_down[l] else None
            x = self._conv(p, self.pc_down[l], out=self.X[l + 1], stats=xs)
        for blk in net.before_block:
            x, xs = self._resblock(x, blk, nlev, x, xs)
        xa, xas = x, xs
        for r, blk in enumerate(net.attn_resnet_block):
            xa, xas = self._resblock(xa, blk, nlev, self.XA, xas)
            if r < len(net.attn_block):
                xa, xas = self._attention(xa, net.attn_block[r], self.XA2, r), None
                # next block reads XA2 and writes XA (its residual input is XA2)
        x, xs = ops.add(x, xa, out=x), None
        for r, blk in enumerate(net.after_block):
            o16 = self.X16[nlev] if r == len(net.after_block) - 1 else None
            x, xs = self._resblock(x, blk, nlev, x if o16 is None else o16, xs)
        for i in range(nlev):
            l = nlev - 1 - i
            # conv(F.interpolate(x, 2)) + skip, fused (commonlayers.py:145; punetg.py:372-373)
            xs = self.ST[l] if self.st_up[i] else None
            x = self._conv(x, self.pc_up[i], out=self.XU[l], residual=self.X[l], up2=True, stats=xs)
            for r, blk in enumerate(net.upward_blocks[i]):
                o16 = self.X16[l] if r == len(net.upward_blocks[i]) - 1 else None
                x, xs = self._resblock(x, blk, l, x if o16 is None else o16, xs)
        if out_nchw is not None:
            return self._conv(x, self.pc_out, out=out_nchw, out_nchw=True)
        return self._conv(x, self.pc_out, out=self.F)


class PUNetGCond(PUNetG):
    """PUNetGCond (reference nets/punetg.py:633-735): the items of ``y`` named in ``channel_conditional_items`` are
    concatenated to x along the channel axis (``config.input_channels`` counts them); what is left of ``y`` goes to the
    conditional embedding.  In the sampler engine the conditioning channels are written into the network-input buffer
    ONCE per run and the fused stage kernels fill only the state channels (dsk_sampler_stage_cond, xin_ld)."""

    def __init__(self, config: PUNetGConfig, conditional_embedding: Optional[nn.Module] = None,
                 channel_conditional_items=False, extra_residual: Optional[nn.Module] = None, *,
                 precision: Optional[str] = None):
        super().__init__(config, conditional_embedding, extra_residual=extra_residual, precision=precision)
        self.channel_conditional_items = channel_conditional_items

    def export_description(self) -> dict[str, Any]:
        args = super().export_description()
        args["channel_conditional_items"] = self.channel_conditional_items
        return args

    def split_condition(self, y, x: torch.Tensor):
        items = self.channel_conditional_items
        y_channels = [y[item] for item in items]           # y is None -> TypeError, as in the reference (:718-719)
        rest = {k: v for k, v in y.items() if k not in items}
        y_cat = torch.cat(y_channels, dim=1)
        if y_cat.shape[0] == 1 and x.shape[0] > 1:
            y_cat = y_cat.expand((x.shape[0],) + tuple(y_cat.shape[1:]))
        return y_cat.to(x), (rest if len(rest) else None)

    def forward(self, x: torch.Tensor, t: Optional[torch.Tensor] = None, y=None) -> torch.Tensor:
        y_cat, rest = self.split_condition(y, x)
        return super().forward(torch.cat([x, y_cat], dim=1), t, rest)
