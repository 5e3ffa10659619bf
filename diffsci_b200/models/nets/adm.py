"""ADM -- drop-in for diffsci.models.nets.adm.ADM / ADMConfig (reference nets/adm.py:8-216, blocks :219-443,
time embedding :1024-1053).  Same constructor signature, ``forward(x, t, y=None)`` contract and state-dict layout.

Block (ADMBaseBlock.forward, adm.py:292-343) as executed here:
    n  = SiLU(GroupNorm(G, C)(x))                       one fused norm kernel
    [n, x' = pool/upsample(n), pool/upsample(x)]         avg-pool kernel / upsample folded into the conv gather
    y  = conv1(n)                                        implicit-GEMM conv (tcgen05 in bf16 mode)
    h  = SiLU(GroupRMSNorm(G, C')(y) * te1 + te2)        one fused norm kernel (FiLM without the '1 +')
    r  = conv1x1(x')                                     GEMM
    o  = conv2(h) + r                                    residual fused into the conv epilogue
    [o = o + MHA(o)]                                     tensor-core attention
The per-block ``embed_linear`` of every block runs as ONE grouped launch per evaluation.
"""
from __future__ import annotations

from typing import Any, Optional

import torch
from torch import nn

from ... import ops
from ..._lib import require_cuda
from .layers import AttentionParams, ConvParams, FourierParams, LinearParams, NormParams, _Act, _Holder, make_conv
from .punetg import _NORM_MODE, _tc_eligible, DEFAULT_PRECISION, PRECISIONS, _ACT_DTYPE, _W_DTYPE, SPLIT_MODES, MIXED_MIN_CIN, _ATTN_FLASH, split_min_cin


class ADMConfig:
    def __init__(self, input_channels: int = 1, output_channels: int = 1, dimension: int = 2, model_channels: int = 64,
                 time_embed_dim: int = 64, output_embed_dim: int = 256, channel_expansion: list[int] = [2, 4],
                 number_resnet_downward_block: int = 2, number_resnet_upward_block: int = 2,
                 number_resnet_attn_block: int = 2, number_resnet_before_attn_block: int = 2,
                 number_resnet_after_attn_block: int = 2, kernel_size: int = 3, time_projection_scale: float = 30.0,
                 transition_scale_factor: int = 2, transition_kernel_size: int = 3, dropout: float = 0.0,
                 cond_dropout: float = 0.0, first_resblock_norm: str = "GroupLN", second_resblock_norm: str = "GroupRMS",
                 affine_norm: bool = True, convolution_type: str = "default", num_groups: int = 1,
                 skip_integration_type: str = "concat", attn_residual: bool = True, decoder_type: int = 1):
        for name, value in list(locals().items()):
            if name != "self":
                setattr(self, name, value)

    @property
    def middle_channel(self):
        return self.model_channels * self.channel_expansion[-1]

    @property
    def extended_channel_expansion(self):
        return [1] + list(self.channel_expansion)

    @property
    def middle_block_attn_config(self):
        return ([False] * self.number_resnet_before_attn_block + [True] * (self.number_resnet_attn_block - 1) + [False] +
                [False] * self.number_resnet_after_attn_block)

    @property
    def num_blocks_middle_block(self):
        return (self.number_resnet_before_attn_block + self.number_resnet_attn_block +
                self.number_resnet_after_attn_block)

    def export_description(self) -> dict[str, Any]:
        import inspect
        return {n: getattr(self, n) for n in inspect.signature(type(self).__init__).parameters if n != "self"}


class ADMBlockParams(_Holder):
    """Parameters of one ADMBaseBlock in the reference's registration order."""

    def __init__(self, cin: int, cout: int, cembed: int, ndim: int, attn: bool, sample: Optional[str],
                 convolution_type: str = "default"):
        super().__init__()
        self.cin, self.cout, self.sample, self.has_attn = cin, cout, sample, attn
        self.norm1 = NormParams(cin)
        self.norm2 = NormParams(cout)
        # convolution_type="circular": the three block convolutions are CircularConv2d (adm.py:427-441; keys <name>.conv.*);
        # the 1x1 residual projection has no halo, so periodic and zero padding coincide for it
        self.conv1 = make_conv(cin, cout, 3, ndim, True, convolution_type)
        self.conv2 = make_conv(cout, cout, 3, ndim, True, convolution_type)
        self.embed_linear = LinearParams(cembed, 2 * cout)
        self.convresidual = make_conv(cin, cout, 1, ndim, True, convolution_type)
        if attn:
            self.attn = AttentionParams(cout)


class _BlockList(_Holder):
    def __init__(self, name: str, blocks):
        super().__init__()
        setattr(self, name, nn.ModuleList(blocks))


class _Layers(_Holder):
    def __init__(self, layers):
        super().__init__()
        self.layers = nn.ModuleList(layers)


class _TimeEmbedding(_Holder):
    def __init__(self, embed_dim, output_dim, scale):
        super().__init__()
        self.projection = FourierParams(embed_dim, scale)
        self.mlp = nn.Sequential(LinearParams(embed_dim, output_dim), _Act(), LinearParams(output_dim, output_dim))


class ADM(nn.Module):
    def __init__(self, config: ADMConfig, conditional_embedding: Optional[nn.Module] = None, *,
                 precision: Optional[str] = None):
        super().__init__()
        c = self.config = config
        bad = []
        if c.dimension != 2:
            bad.append("dimension != 2 (the reference hard-codes Conv2d input/output layers, adm.py:189-196)")
        if c.convolution_type not in ("default", "circular"):
            bad.append(f"convolution_type={c.convolution_type!r}")
        if c.first_resblock_norm not in _NORM_MODE or c.second_resblock_norm not in _NORM_MODE:
            bad.append("norms other than GroupLN/GroupRMS")
        if c.decoder_type != 1 or c.skip_integration_type not in ("concat", "add"):
            bad.append("decoder_type != 1 / unknown skip_integration_type")
        if c.transition_scale_factor != 2 or c.kernel_size != 3:
            bad.append("transition_scale_factor != 2 or kernel_size != 3")
        if bad:
            raise NotImplementedError("diffsci_b200.ADM: not built yet (SURVEY.md 8f): " + "; ".join(bad))
        self.precision = precision or DEFAULT_PRECISION
        self.conditional_embedding = conditional_embedding     # any torch module: y -> [B, output_embed_dim] (adm.py:199-203)
        M, E, nd = c.model_channels, c.output_embed_dim, c.dimension
        ct = c.convolution_type          # the input / output layers stay zero-padded torch.nn.Conv2d in the reference (adm.py:189-196)
        mult = c.extended_channel_expansion
        self.time_embedding = _TimeEmbedding(c.time_embed_dim, E, c.time_projection_scale)
        enc = []
        for i in range(len(mult) - 1):
            cin, cout, nb = M * mult[i], M * mult[i + 1], c.number_resnet_downward_block
            enc.append(_BlockList("input_blocks", [ADMBlockParams(cin, cout if r == nb - 1 else cin, E, nd, False,
                                                                  "down" if r == nb - 1 else None, ct) for r in range(nb)]))
        self.encoder = _Layers(enc)
        mc = c.middle_channel
        self.middle_block = _BlockList("middle_blocks", [ADMBlockParams(mc, mc, E, nd, a, None, ct)
                                                         for a in c.middle_block_attn_config])
        rev = mult[::-1]
        dec = []
        for i in range(len(rev) - 1):
            cin, cout, nb = M * rev[i], M * rev[i + 1], c.number_resnet_upward_block
            cc = 2 * cin if c.skip_integration_type == "concat" else cin
            dec.append(_BlockList("input_blocks", [ADMBlockParams(cc, cout if r == nb - 1 else cc, E, nd, False,
                                                                  "up" if r == nb - 1 else None, ct) for r in range(nb)]))
        self.decoder = _Layers(dec)
        self.input_layer = ConvParams(c.input_channels, M, c.kernel_size, 2)
        self.output_layer = ConvParams(M, c.output_channels, c.kernel_size, 2)
        self.cond_dropout = nn.Dropout(c.cond_dropout)
        self._plans: dict[Any, "_ADMPlan"] = {}

    # ------------------------------------------------------------------ conditioning
    @property
    def cond_dim(self) -> int:
        return self.config.output_embed_dim

    def native_parameters(self) -> list:
        """The parameters the CUDA path owns (everything except the user's conditional embedding)."""
        return [p for n, p in self.named_parameters() if not n.startswith("conditional_embedding.")]

    def conditioning_vector(self, y, B: int) -> Optional[torch.Tensor]:
        """ye of adm.py:199-209: cond_dropout(conditional_embedding(y)) as fp32 [B, output_embed_dim]; None when y is None
        (the reference adds zeros then, which is the same thing)."""
        if y is None:
            return None
        if self.conditional_embedding is None:
            raise TypeError("'NoneType' object is not callable (ADM got y but has no conditional_embedding, adm.py:201)")
        ye = self.cond_dropout(self.conditional_embedding(y))
        E = self.config.output_embed_dim
        if ye.ndim == 1:
            ye = ye.unsqueeze(0)
        if ye.ndim != 2 or ye.shape[-1] != E or ye.shape[0] not in (1, B):
            raise NotImplementedError(f"diffsci_b200.ADM: conditioning of shape {tuple(ye.shape)} (expected [B or 1, {E}])")
        return ye.float().expand(B, E)

    def split_condition(self, y, x: torch.Tensor):
        return None, y

    def forward(self, x: torch.Tensor, t: torch.Tensor, y=None) -> torch.Tensor:
        require_cuda(x, "ADM input")
        B = x.shape[0]
        ye = self.conditioning_vector(y, B)
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            from .graph import NetFunction        # dropout acts only in training mode, as torch.nn.Dropout does
            graph = self.train_graph(B, tuple(x.shape[2:]), x.device, cond=ye is not None, dropout=self.training)
            return NetFunction.apply(graph, x, t, None if ye is None else ye.contiguous(), *self.native_parameters())
        if self.training and float(getattr(self.config, "dropout", 0.0)) > 0.0:
            raise NotImplementedError("dropout > 0 acts on the training path (gradients enabled); call .eval() for inference")
        plan = self.plan(B, tuple(x.shape[2:]), x.device)
        xin = ops.nchw_to_cl(x.float(), plan.act_dtype, 2, out=plan.xin)
        out = torch.empty((B, self.config.output_channels) + tuple(x.shape[2:]), dtype=torch.float32, device=x.device)
        plan.forward(xin, t.float().contiguous(), out_nchw=out, ye=None if ye is None else ye.contiguous())
        return out

    def plan(self, B: int, spatial: tuple, device, precision: Optional[str] = None) -> "_ADMPlan":
        precision = precision or self.precision
        key = (B, tuple(spatial), str(device), precision)
        sig = tuple(p.data_ptr() for p in self.native_parameters())
        plan = self._plans.get(key)
        if plan is None or plan.sig != sig:
            if len(self._plans) >= 4:
                self._plans.clear()
            with torch.inference_mode(False), torch.no_grad():
                plan = self._plans[key] = _ADMPlan(self, B, tuple(spatial), device, precision, sig)
        return plan

    def train_graph(self, B: int, spatial: tuple, device, precision: Optional[str] = None, cond: bool = False,
                    dropout: bool = True):
        from .graph import build_adm
        precision = precision or self.precision
        dropout = bool(dropout) and float(getattr(self.config, "dropout", 0.0)) > 0.0
        key = ("train", B, tuple(spatial), str(device), precision, bool(cond), dropout)
        sig = tuple(p.data_ptr() for p in self.native_parameters())
        g = self._plans.get(key)
        if g is None or g.sig != sig:
            for k in [k for k in self._plans if k[0] == "train"]:
                del self._plans[k]
            with torch.inference_mode(False), torch.no_grad():
                g = self._plans[key] = build_adm(self, B, tuple(spatial), device, precision, cond=cond, dropout=dropout)
        return g

    def _apply(self, fn, *a, **k):
        self._plans = {}
        return super()._apply(fn, *a, **k)


class _ADMPlan:
    def __init__(self, net: ADM, B: int, spatial: tuple, device, precision: str, sig):
        if precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {PRECISIONS}, got {precision!r}")
        c = net.config
        self.net, self.B, self.sig, self.precision = net, B, sig, precision
        self.act_dtype = adt = _ACT_DTYPE[precision]         # modes: see punetg.PRECISIONS
        wdt = _W_DTYPE.get(precision)
        self.split = precision in SPLIT_MODES                 # tensor-core convs read split-fp16 inputs (hi | lo)
        self.split_min_cin = split_min_cin(precision)
        self.device = dev = torch.device(device)
        nlev = len(c.channel_expansion)
        H, W = spatial
        if H % (2 ** nlev) or W % (2 ** nlev):
            raise ValueError(f"ADM: spatial size {spatial} must be divisible by {2 ** nlev}")
        self.blocks = ([b for layer in net.encoder.layers for b in layer.input_blocks] + list(net.middle_block.middle_blocks) +
                       [b for layer in net.decoder.layers for b in layer.input_blocks])
        self._bufs: dict[Any, torch.Tensor] = {}
        self._pad_ws = None
        self.xin = torch.empty((B, 1, H, W, c.input_channels), dtype=adt, device=dev)
        self.F = torch.empty((B, 1, H, W, c.output_channels), dtype=adt, device=dev)

        def pack(cp, subpixel=False, few_out_ok=False):
            tc = wdt is not None and _tc_eligible(cp.cin, cp.cout, cp.ksize, few_out_ok)
            return ops.PackedConv(cp.weight, cp.bias, 2, wdt if tc else torch.float32, subpixel and tc,
                                  circular=bool(getattr(cp, "circular", False)) and cp.ksize > 1)

        self.pc_in, self.pc_out = pack(net.input_layer), pack(net.output_layer, few_out_ok=True)
        self.pc = {id(b): (pack(b.conv1, b.sample == "up"), pack(b.conv2), pack(b.convresidual)) for b in self.blocks}
        # time embedding + one grouped launch for every block's embed_linear (two groups per block: te1 | te2)
        f32 = dict(dtype=torch.float32, device=dev)
        E = c.output_embed_dim
        self.four = torch.empty((B, c.time_embed_dim), **f32)
        self.h1 = torch.empty((B, E), **f32)
        self.te = torch.empty((B, E), **f32)
        mlp = net.time_embedding.mlp
        self.t_mlp = [ops.GroupedLinear([self.four], [mlp[0].weight], [mlp[0].bias], [self.h1], 1),
                      ops.GroupedLinear([self.h1], [mlp[2].weight], [mlp[2].bias], [self.te], 1)]   # + act_final SiLU
        # conditional: te = SiLU(mlp(fourier(t)) + ye)  (adm.py:1047-1053): second linear without activation, add, SiLU
        self.tez = torch.empty((B, E), **f32)
        self.t_mlp2_noact = ops.GroupedLinear([self.h1], [mlp[2].weight], [mlp[2].bias], [self.tez], 0)
        self.film = {id(b): (torch.empty((B, b.cout), **f32), torch.empty((B, b.cout), **f32)) for b in self.blocks}
        ws, bs, ys = [], [], []
        for b in self.blocks:
            w, bias = b.embed_linear.weight, b.embed_linear.bias
            ws += [w[:b.cout], w[b.cout:]]
            bs += [bias[:b.cout], bias[b.cout:]]
            ys += list(self.film[id(b)])
        self.film_gl = ops.GroupedLinear([self.te] * len(ws), ws, bs, ys, 0)
        self.attn_state = {}
        self.attn_bufs = {}
        self.prepare()

    def buf(self, tag, shape, dtype=None):
        key = (tag, tuple(shape), dtype or self.act_dtype)
        t = self._bufs.get(key)
        if t is None:
            t = self._bufs[key] = torch.empty(tuple(shape), dtype=key[2], device=self.device)
        return t

    def prepare(self):
        for pc in [self.pc_in, self.pc_out] + [p for tpl in self.pc.values() for p in tpl]:
            pc.packed()

    # ------------------------------------------------------------------ pieces
    def _resample(self, x, mode, tag):
        B, _, H, W, C = x.shape
        if mode == "down":
            return ops.pool2x(x, 2, False, out=self.buf(tag, (B, 1, H // 2, W // 2, C)))     # AvgPool2d (adm.py:361-371)
        return x

    def _conv(self, x, pc, out, up: bool, **kw):
        # nearest x2 upsample is never materialised: FFMA folds it into the gather, tcgen05 runs the sub-pixel form
        if self.split and ops.is_tc_dtype(pc.w_dtype) and x.dtype == torch.float32:
            # split mode: a tensor-core convolution reads the split copy (hi | lo) of its fp32 input
            if pc.cin < self.split_min_cin:      # mixed mode: plain fp16 operand for the narrow layers
                x = ops.cast(x, torch.float16, out=self.buf("cast16", x.shape, torch.float16))
            else:
                x = ops.split_f16(x, out=self.buf("split", x.shape[:-1] + (2 * x.shape[-1],), torch.float16))
        if pc.circular:      # the tcgen05 path reads a halo-padded copy (TMA boxes cannot wrap): one workspace, grown on demand
            need = ops.conv_pad_ws_bytes(x.shape, x.dtype, pc, up)
            if need > 0 and (self._pad_ws is None or self._pad_ws.numel() < need):
                self._pad_ws = torch.empty(need, dtype=torch.uint8, device=x.device)
            kw["pad_ws"] = self._pad_ws
        return ops.conv(x, pc, out=out, up2=up, **kw)

    def _attention(self, x, blk, idx):
        B, _, H, W, C = x.shape
        Lq = H * W
        m = blk.attn.mhattn
        res = self.net.config.attn_residual
        out = self.buf(("attn_out", idx), x.shape)
        tc = self.act_dtype in ops.H16 and _tc_eligible(C, C) and Lq % 8 == 0 and Lq <= 8192
        f32 = torch.float32
        split_ok = self.split and _tc_eligible(C, C) and Lq % 64 == 0 and Lq <= 8192
        if _ATTN_FLASH and ops.attn_flash_supported(Lq, C) and ((split_ok and self.precision != "fp32") or (tc and not self.split)):
            # flash-style core (dsk_attn_flash) for every 16-bit-operand mode; the fp32 mode keeps the split-operand GEMM chain
            wfmt = ops.SPLIT if self.split else self.act_dtype
            st = self.attn_state.get(idx)
            if st is None:
                st = self.attn_state[idx] = (ops.PackedLinear(m.in_proj_weight, wfmt), ops.PackedLinear(m.out_proj.weight, wfmt))
                st[0].packed(), st[1].packed()
            bufs = self.attn_bufs.get((B, Lq, C))
            if bufs is None:
                bufs = self.attn_bufs[(B, Lq, C)] = ops.attention_flash_buffers(B, Lq, C, x.device, self.act_dtype, split=self.split)
            fn = ops.self_attention_flash_split if self.split else ops.self_attention_flash
            fn(x.view(B, Lq, C), st[0], m.in_proj_bias, st[1], m.out_proj.bias, bufs, out.view(B, Lq, C), res)
            return out
        if split_ok:
            st = self.attn_state.get(idx)
            if st is None:
                st = self.attn_state[idx] = (ops.PackedLinear(m.in_proj_weight, ops.SPLIT), ops.PackedLinear(m.out_proj.weight, ops.SPLIT))
                st[0].packed(), st[1].packed()
            bufs = self.attn_bufs.get((B, Lq, C))
            if bufs is None:
                bufs = self.attn_bufs[(B, Lq, C)] = ops.attention_split_buffers(B, Lq, C, x.device)
            ops.self_attention_split(x.view(B, Lq, C), st[0], m.in_proj_bias, st[1], m.out_proj.bias, bufs, out.view(B, Lq, C), res)
            return out
        if tc:
            st = self.attn_state.get(idx)
            if st is None:
                st = self.attn_state[idx] = (ops.PackedLinear(m.in_proj_weight, self.act_dtype),
                                             ops.PackedLinear(m.out_proj.weight, self.act_dtype))
            bufs = self.attn_bufs.get((B, Lq, C))
            if bufs is None:
                bufs = self.attn_bufs[(B, Lq, C)] = ops.attention_tc_buffers(B, Lq, C, x.device, self.act_dtype)
            ops.self_attention_tc(x.view(B, Lq, C), st[0], m.in_proj_bias, st[1], m.out_proj.bias, bufs, out.view(B, Lq, C), res)
            return out
        bufs = dict(qkv=self.buf("qkv", (B * Lq, 3 * C), f32), scores=self.buf("scores", (B, Lq, Lq), f32),
                    ao=self.buf("ao32", (B * Lq, C), f32), out=self.buf("out32", (B, Lq, C), f32))
        tok = x.view(B, Lq, C) if x.dtype == f32 else ops.cast(x.view(B, Lq, C), f32, out=self.buf("tok32", (B, Lq, C), f32))
        o = ops.self_attention_f32(tok, m.in_proj_weight, m.in_proj_bias, m.out_proj.weight, m.out_proj.bias, bufs, res)
        ops.cast(o, out.dtype, out=out.view(B, Lq, C))
        return out

    def _block(self, x, blk, idx):
        c = self.net.config
        G = c.num_groups
        B, _, H, W, Cin = x.shape
        pc1, pc2, pcr = self.pc[id(blk)]
        te1, te2 = self.film[id(blk)]
        down, up = blk.sample == "down", blk.sample == "up"
        Ho, Wo = (H // 2, W // 2) if down else ((2 * H, 2 * W) if up else (H, W))
        S_in = H * W
        n = ops.norm_act(x, blk.norm1.weight, blk.norm1.bias, G, _NORM_MODE[c.first_resblock_norm], True,
                         out=self.buf("n1", x.shape), ws=self.buf("ws", (ops.lib.dsk_norm_ws_bytes(B, S_in, Cin),), torch.uint8))
        xr = x
        if down:
            n, xr = self._resample(n, "down", "n1p"), self._resample(x, "down", "xp")
        y = self._conv(n, pc1, self.buf("y", (B, 1, Ho, Wo, blk.cout)), up)
        h = ops.norm_act(y, blk.norm2.weight, blk.norm2.bias, G, _NORM_MODE[c.second_resblock_norm], True,
                         out=self.buf("n2", y.shape), film_scale=te1, film_shift=te2,
                         ws=self.buf("ws", (ops.lib.dsk_norm_ws_bytes(B, Ho * Wo, blk.cout),), torch.uint8))
        r = self._conv(xr, pcr, self.buf("r", y.shape), up)                            # conv1x1([pool|up](x))
        o = self._conv(h, pc2, self.buf(("o", idx), y.shape), False, residual=r)
        if blk.has_attn:
            o = self._attention(o, blk, idx)
        return o

    def forward(self, xin: torch.Tensor, cnoise: torch.Tensor, out_nchw: Optional[torch.Tensor] = None,
                ye: Optional[torch.Tensor] = None, **_):
        net, c = self.net, self.net.config
        ops.fourier(cnoise, net.time_embedding.projection.W, out=self.four)
        if ye is None:
            for g in self.t_mlp:
                g.run()                                                          # te = SiLU(mlp(fourier(t)))  (adm.py:1047-1053)
        else:
            self.t_mlp[0].run()
            self.t_mlp2_noact.run()
            ops.add_ex(self.tez, ye, self.tez)                                   # + ye before the final SiLU (adm.py:1050-1052)
            ops.silu_fwd(self.tez, self.te)
        self.film_gl.run()
        x = ops.conv(xin, self.pc_in, out=self.buf("x0", xin.shape[:-1] + (c.model_channels,)))
        skips, idx = [x], 0
        for layer in net.encoder.layers:
            for blk in layer.input_blocks:
                x = self._block(x, blk, idx)
                idx += 1
            skips.append(x)
        for blk in net.middle_block.middle_blocks:
            x = self._block(x, blk, idx)
            idx += 1
        for layer in net.decoder.layers:
            hskip = skips.pop()
            if c.skip_integration_type == "concat":
                x = ops.concat_channels(x, hskip, out=self.buf(("cat", idx), x.shape[:-1] + (x.shape[-1] + hskip.shape[-1],)))
            else:
                x = ops.add(x, hskip, out=self.buf(("sum", idx), x.shape))
            for blk in layer.input_blocks:
                x = self._block(x, blk, idx)
                idx += 1
        if out_nchw is not None:
            return self._conv(x, self.pc_out, out_nchw, False, out_nchw=True)
        return self._conv(x, self.pc_out, self.F, False)
