"""Static training graph for the native score networks: forward + hand-written backward (K2).

The reference trains with ATen autograd: ``KarrasModule.training_step`` -> ``loss_fn`` -> ``loss.backward()``
(karras/karrasmodule.py:569-662, 1146-1155) through PUNetG / ADM (nets/punetg.py:389-416, nets/adm.py:199-216).
Here a network instance is unrolled ONCE per (batch, shape, precision) into a list of forward launches and a list of
backward launches over preallocated buffers -- no autograd tape, no allocation and no host synchronisation per step,
so a whole training step is a fixed launch sequence (CUDA-graph capturable).  Every launch is a libdiffsci_b200
kernel (diffsci_b200.ops); PyTorch only owns the memory.

Gradient bookkeeping is resolved at build time.  A ``Var`` is a forward buffer plus a gradient slot that is
  NONE  -> nothing has contributed yet,
  ALIAS -> exactly one consumer contributed a pure copy (ResNet identity, U-Net skip, `+`): the slot just points at
           that consumer's own gradient buffer, no pass over memory,
  OWN   -> a buffer of its own; later contributions are accumulated through the kernels' ``dres`` operand
           (out = f(...) + dres, dres may alias out) instead of a separate add pass.
Backward ops are generated in reverse forward order, so when an op asks for the gradient of its output every consumer
has already contributed.  Parameter gradients are fp32 views (reference layout) of ONE flat buffer in
``net.parameters()`` order -- the unit of the data-parallel all-reduce and of the fused AdamW/EMA step.
"""
from __future__ import annotations

import math
from typing import Callable, Optional

import torch

from ... import ops

_NORM_MODE = {"GroupLN": 0, "GroupRMS": 1}
NONE, ALIAS, OWN = 0, 1, 2
HYBRID_ATTENTION = True     # bf16 training, token counts off the tensor-core score path: projections on tcgen05 (A/B switch for tests)


class Var:
    __slots__ = ("t", "needs_grad", "state", "g", "name", "stats")

    def __init__(self, t: torch.Tensor, needs_grad: bool = True, name: str = ""):
        self.t, self.needs_grad, self.state, self.g, self.name = t, needs_grad, NONE, None, name
        self.stats = None      # per-(sample, channel) statistics the producing conv epilogue left for a following per-channel norm


class TrainGraph:
    """Forward/backward launch lists of one network for one (batch, spatial shape, precision)."""

    def __init__(self, net: torch.nn.Module, B: int, device, precision: str, ndim: int):
        if precision == "fp32_ffma":
            precision = "fp32"       # training in fp32 storage runs the CUDA-core kernels either way
        if precision not in ("fp32", "bf16"):
            raise ValueError(f"training graphs run in 'fp32' (CUDA-core parity mode) or 'bf16' (tensor cores; gradients need "
                             f"bf16's exponent range), got {precision!r}; the fp16 modes are inference formats")
        self.net, self.B, self.device, self.precision, self.ndim = net, B, torch.device(device), precision, ndim
        self.act_dtype = torch.float32 if precision == "fp32" else torch.bfloat16
        self.fwd: list[Callable[[], None]] = []
        self._bwd_builders: list[Callable[[], list]] = []
        self.bwd: list[Callable[[], None]] = []
        # the parameters the CUDA path owns (a conditional net's embedder / ConditionDrop stay with torch autograd)
        self.params = list(net.native_parameters()) if hasattr(net, "native_parameters") else [p for p in net.parameters()]
        self.ye_in: Optional[Var] = None          # conditioning vector [B, M], set by the builders
        self.sig = tuple(p.data_ptr() for p in self.params)
        total = sum(p.numel() for p in self.params)
        self.flat_grad = torch.zeros(max(total, 1), dtype=torch.float32, device=self.device)
        self._gviews, off = {}, 0
        for p in self.params:
            self._gviews[id(p)] = self.flat_grad[off:off + p.numel()].view(p.shape)
            off += p.numel()
        self._written: set[int] = set()
        self._written_log: list = []          # parameter ids in the order their gradient writers were generated
        self.grad_ready_pos: dict = {}        # id(param) -> number of backward launches after which its gradient is final
        self._scratch: dict = {}
        self._packs: list = []
        self.nbytes = 0
        self.dropout_sites: list = []          # (module name, Philox stream id, shape) of every dropout op, in forward order
        self.dropout_seed = 0
        self.fixed_dropout_seed: Optional[int] = None      # tests: pin the masks

    # ------------------------------------------------------------------ memory
    def empty(self, shape, dtype=None) -> torch.Tensor:
        t = torch.empty(tuple(shape), dtype=dtype or self.act_dtype, device=self.device)
        self.nbytes += t.numel() * t.element_size()
        return t

    def scratch(self, tag, nbytes_or_shape, dtype=torch.uint8) -> torch.Tensor:
        """Buffers that live only inside one op's launches are shared across ops (execution is sequential)."""
        shape = (int(nbytes_or_shape),) if isinstance(nbytes_or_shape, int) else tuple(nbytes_or_shape)
        n = 1
        for s in shape:
            n *= s
        key = (tag, dtype)
        t = self._scratch.get(key)
        if t is None or t.numel() < n:
            t = self._scratch[key] = torch.empty(max(n, 1), dtype=dtype, device=self.device)
        return t[:n].view(shape)

    def lazy_scratch(self, tag, shape, dtype=torch.uint8):
        """Scratch resolved at call time (the backing buffer may grow while the graph is being built)."""
        self.scratch(tag, shape, dtype)
        return lambda: self.scratch(tag, shape, dtype)

    def grad_view(self, p: torch.Tensor, rows: Optional[slice] = None) -> torch.Tensor:
        g = self._gviews[id(p)]
        key = (id(p), None if rows is None else (rows.start, rows.stop))
        if key in self._written or (id(p), None) in self._written:
            raise RuntimeError("TrainGraph: a parameter is used by more than one op (gradient accumulation across ops "
                               "is not built)")
        self._written.add(key)
        self._written_log.append(id(p))
        return g if rows is None else g[rows]

    # ------------------------------------------------------------------ gradient slots (build time)
    def contribute_compute(self, v: Var):
        """-> (dres, out) for a kernel that writes out = f(...) (+ dres)."""
        if v.state == NONE:
            v.g, v.state = self.empty(v.t.shape, v.t.dtype), OWN
            return None, v.g
        if v.state == ALIAS:
            src = v.g
            v.g, v.state = self.empty(v.t.shape, v.t.dtype), OWN
            return src, v.g
        return v.g, v.g

    def contribute_copy(self, v: Var, src: torch.Tensor) -> list:
        """The contribution IS `src` (same shape/dtype as v)."""
        assert src.shape == v.t.shape and src.dtype == v.t.dtype, (src.shape, v.t.shape, src.dtype, v.t.dtype)
        if v.state == NONE:
            v.g, v.state = src, ALIAS
            return []
        if v.state == ALIAS:
            a = v.g
            v.g, v.state = self.empty(v.t.shape, v.t.dtype), OWN
            out = v.g
            return [lambda: ops.add_ex(a, src, out)]
        out = v.g
        return [lambda: ops.add_ex(out, src, out)]

    @staticmethod
    def grad_of(v: Var) -> Optional[torch.Tensor]:
        return v.g if v.state != NONE else None

    # ------------------------------------------------------------------ ops
    def conv(self, x: Var, cp, chan_bias: Optional[Var] = None, residual: Optional[Var] = None, up2: bool = False,
             few_out_ok: bool = False, want_stats: bool = False) -> Var:
        """y = conv_same([up2](x)) + bias + chan_bias[b, :] + residual  (forward: dsk_conv_fwd; backward: dsk_conv_wgrad,
        dsk_channel_sum, dsk_conv_fwd with dgrad-packed weights [+ dsk_upsample2x_bwd])."""
        from .punetg import _tc_eligible
        import diffsci_b200
        nd = self.ndim
        if (cp.ksize == 1 and self.precision == "bf16" and diffsci_b200.TC_CONV_ENABLED and cp.cin % 8 == 0 and cp.cout % 8 == 0
                and chan_bias is None and not (up2 and residual is not None)):
            return self._conv1x1_tc(x, cp, residual, up2)
        tc = self.precision == "bf16" and _tc_eligible(cp.cin, cp.cout, cp.ksize, few_out_ok)
        wdt = torch.bfloat16 if tc else torch.float32
        # CircularConv (commonlayers.py:918-1032): forward, dgrad and wgrad wrap; a 1x1 kernel has no halo to wrap
        circ = bool(getattr(cp, "circular", False)) and cp.ksize > 1
        pc = ops.PackedConv(cp.weight, cp.bias, nd, wdt, subpixel=bool(up2 and tc), circular=circ)
        self._packs.append(pc)
        pws = self.lazy_scratch("pad_ws", max(ops.conv_pad_ws_bytes(x.t.shape, x.t.dtype, pc, up2), 1))
        B, D, H, W, _ = x.t.shape
        if up2:
            D, H, W = (D * 2 if nd == 3 else D), H * 2, W * 2
        y = Var(self.empty((B, D, H, W, cp.cout)))
        xt, yt = x.t, y.t
        cb = chan_bias.t if chan_bias is not None else None
        rs = residual.t if residual is not None else None
        st = None
        if want_stats and tc and ops.conv_stats_supported(tuple(xt.shape), xt.dtype, pc, up2=up2):
            # norm statistics fused into the conv epilogue (as on the inference plan): the following norm skips its statistics pass
            st = y.stats = ops.conv_stats_buffer(B, cp.cout, self.device)
            self.nbytes += st.numel() * 4
        self.fwd.append(lambda: ops.conv(xt, pc, out=yt, chan_bias=cb, residual=rs, up2=up2, pad_ws=pws, stats=st))

        def build_bwd():
            dy = self.grad_of(y)
            if dy is None:
                return []
            out = []
            gb = self.grad_view(cp.bias) if cp.bias is not None else None
            S = D * H * W
            csws = self.lazy_scratch("bwd_ws", ops.bwd_ws_bytes(B, S, cp.cout))
            if chan_bias is not None:
                dres, dcb = self.contribute_compute(chan_bias)
                assert dres is None, "chan_bias with several consumers is not supported"
                out.append(lambda: ops.channel_sum(dy, dcb, csws(), True))
                if gb is not None:
                    out.append(lambda: ops.colsum(dcb, gb))
            elif gb is not None:
                out.append(lambda: ops.channel_sum(dy, gb, csws(), False))
            gw = self.grad_view(cp.weight)
            # the tcgen05 weight-gradient kernel reads a materialised operand; the CUDA-core one gathers (up2 folded in)
            if up2 and self.precision == "bf16" and _tc_eligible(cp.cin, cp.cout, cp.ksize):
                desc = ops.conv_desc(B, D, H, W, cp.cin, cp.cout, cp.ksize, nd, False, wdt, xt.dtype, dy.dtype, circ)
                wws = self.lazy_scratch("wgrad_ws", ops.conv_wgrad_ws_bytes(desc))
                u = self.lazy_scratch("u_up", (B, D, H, W, cp.cin), xt.dtype)
                out.append(lambda: ops.upsample2x(xt, nd, out=u()))
                out.append(lambda: ops.conv_wgrad(desc, u(), dy, gw, wws()))
            else:
                desc = ops.conv_desc(B, D, H, W, cp.cin, cp.cout, cp.ksize, nd, up2, wdt, xt.dtype, dy.dtype, circ)
                wws = self.lazy_scratch("wgrad_ws", ops.conv_wgrad_ws_bytes(desc))
                out.append(lambda: ops.conv_wgrad(desc, xt, dy, gw, wws()))
            if residual is not None and residual.needs_grad:
                out.extend(self.contribute_copy(residual, dy))
            if x.needs_grad:
                dg_tc = self.precision == "bf16" and _tc_eligible(cp.cout, cp.cin, cp.ksize)
                pd = ops.PackedConv(cp.weight, None, nd, torch.bfloat16 if dg_tc else torch.float32, dgrad=True, circular=circ)
                self._packs.append(pd)
                dpws = self.lazy_scratch("pad_ws", max(ops.conv_pad_ws_bytes(dy.shape, dy.dtype, pd), 1))
                if not up2:
                    dres, dx = self.contribute_compute(x)
                    out.append(lambda: ops.conv(dy, pd, out=dx, residual=dres, pad_ws=dpws))
                else:
                    du = self.lazy_scratch("du", (B, D, H, W, cp.cin), xt.dtype)
                    dres, dx = self.contribute_compute(x)
                    out.append(lambda: ops.conv(dy, pd, out=du(), pad_ws=dpws))
                    out.append(lambda: ops.upsample2x_bwd(du(), dx, nd, dres=dres))
            return out

        self._bwd_builders.append(build_bwd)
        return y

    def _conv1x1_tc(self, x: Var, cp, residual: Optional[Var], up2: bool) -> Var:
        """1x1 convolution (ADM's residual projection, adm.py:433-441) as plain tensor-core GEMMs over the pixel rows:
        y = x W^T + b (+ residual);  dX = dY W;  dW = dY^T X with the pixel range split into slices that are summed in a
        fixed order.  A 1x1 conv commutes with nearest upsampling, so `up2` runs at the input resolution."""
        nd = self.ndim
        B, D, H, W, Cin = x.t.shape
        Cout = cp.cout
        pix = B * D * H * W
        oD, oH, oW = (D * 2 if (up2 and nd == 3) else D), (H * 2 if up2 else H), (W * 2 if up2 else W)
        bf, f32 = torch.bfloat16, torch.float32
        y = Var(self.empty((B, oD, oH, oW, Cout)))
        wl = ops.PackedLinear(cp.weight)
        self._packs.append(wl)
        bias = cp.bias.detach() if cp.bias is not None else None
        xt = x.t.view(pix, Cin)
        G = ops.gemm_bf16_tc
        if up2:
            low = self.lazy_scratch("c1_low", (B, D, H, W, Cout), bf)
            yt = y.t
            self.fwd.append(lambda: G(xt, wl.packed(), low().view(pix, Cout), M=pix, N=Cout, K=Cin, lda=Cin, ldb=Cin, ldc=Cout,
                                      bias=bias))
            self.fwd.append(lambda: ops.upsample2x(low(), nd, out=yt))
        else:
            yt = y.t.view(pix, Cout)
            rs = residual.t.view(pix, Cout) if residual is not None else None
            self.fwd.append(lambda: G(xt, wl.packed(), yt, M=pix, N=Cout, K=Cin, lda=Cin, ldb=Cin, ldc=Cout, bias=bias,
                                      residual=rs))
        # pixel slices of the weight-gradient GEMM: enough CTAs to fill the machine, >= 512 rows each
        tiles = ((Cout + 127) // 128) * ((Cin + 63) // 64)
        nsl = 1
        while tiles * nsl < 148 and pix % (nsl * 2) == 0 and (pix // (nsl * 2)) % 8 == 0 and pix // (nsl * 2) >= 512:
            nsl *= 2
        rows = pix // nsl

        def build_bwd():
            dy = self.grad_of(y)
            if dy is None:
                return []
            out = []
            if up2:
                dl = self.lazy_scratch("c1_dlow", (B, D, H, W, Cout), bf)
                out.append(lambda: ops.upsample2x_bwd(dy, dl(), nd))
                dyl = lambda: dl().view(pix, Cout)  # noqa: E731
            else:
                dyl = lambda: dy.view(pix, Cout)  # noqa: E731
            if cp.bias is not None:
                gb = self.grad_view(cp.bias)
                csws = self.lazy_scratch("bwd_ws", ops.bwd_ws_bytes(B, D * H * W, Cout))
                out.append(lambda: ops.channel_sum(dyl().view(B, D * H * W, Cout), gb, csws(), False))
            gw = self.grad_view(cp.weight)
            wsp = self.lazy_scratch("c1_wsplit", (nsl, Cout * Cin), f32)
            out.append(lambda: G(dyl(), xt, wsp(), M=Cout, N=Cin, K=rows, lda=Cout, ldb=Cin, ldc=Cin, batch=nsl,
                                 strideA=rows * Cout, strideB=rows * Cin, strideC=Cout * Cin, transA=True, transB=True))
            out.append(lambda: ops.colsum(wsp(), gw.view(-1)))
            if residual is not None and residual.needs_grad:
                out.extend(self.contribute_copy(residual, dy))
            if x.needs_grad:
                dres, dx = self.contribute_compute(x)
                dxt = dx.view(pix, Cin)
                out.append(lambda: G(dyl(), wl.packed(), dxt, M=pix, N=Cin, K=Cout, lda=Cout, ldb=Cin, ldc=Cin, transB=True,
                                     residual=dres.view(pix, Cin) if dres is not None else None))
            return out

        self._bwd_builders.append(build_bwd)
        return y

    def norm(self, x: Var, np_, G: int, kind: str, silu: bool, film: Optional[tuple] = None) -> Var:
        mode = _NORM_MODE[kind]
        B, Cc = x.t.shape[0], x.t.shape[-1]
        S = x.t.numel() // (B * Cc)
        y = Var(self.empty(x.t.shape))
        fws = self.empty((int(ops.lib.dsk_norm_ws_bytes(B, S, Cc)),), torch.uint8)    # kept: table + stats feed the backward
        xt, yt = x.t, y.t
        fsc = film[0].t if film is not None else None
        fsh = film[1].t if film is not None else None
        cs = x.stats if G == Cc else None
        self.fwd.append(lambda: ops.norm_act(xt, np_.weight, np_.bias, G, mode, silu, out=yt, film_scale=fsc, film_shift=fsh,
                                             ws=fws, conv_stats=cs))

        def build_bwd():
            dy = self.grad_of(y)
            if dy is None:
                return []
            affine = np_.weight is not None
            dg = self.grad_view(np_.weight) if affine else None
            db = self.grad_view(np_.bias) if affine else None
            dfs = dfh = None
            if film is not None:
                r1, dfs = self.contribute_compute(film[0])
                r2, dfh = self.contribute_compute(film[1])
                assert r1 is None and r2 is None, "FiLM vectors with several consumers are not supported"
            ws = self.lazy_scratch("bwd_ws", ops.bwd_ws_bytes(B, S, Cc))
            if not x.needs_grad:
                raise RuntimeError("norm of a tensor that needs no gradient")
            dres, dx = self.contribute_compute(x)
            return [lambda: ops.norm_act_bwd(xt, dy, dx, np_.weight, np_.bias, G, mode, silu, fws, ws(), dgamma=dg, dbeta=db,
                                             dres=dres, film_scale=fsc, dfilm_scale=dfs, dfilm_shift=dfh)]

        self._bwd_builders.append(build_bwd)
        return y

    def dropout(self, x: Var, p: float, name: str = "") -> Var:
        """Training-mode torch.nn.Dropout(p) (commonlayers.py:792, 830; adm.py:312-313).  Each site owns a Philox stream; the
        seed is redrawn (CPU generator) at every run_forward, and the backward launch regenerates the site's mask from the
        same (seed, stream) instead of reading a stored one."""
        sid = len(self.dropout_sites)
        self.dropout_sites.append((name, sid, tuple(x.t.shape)))
        y = Var(self.empty(x.t.shape))
        xt, yt = x.t, y.t
        self.fwd.append(lambda: ops.dropout(xt, p, self.dropout_seed, sid, out=yt))

        def build_bwd():
            dy = self.grad_of(y)
            if dy is None or not x.needs_grad:
                return []
            dres, dx = self.contribute_compute(x)
            return [lambda: ops.dropout(dy, p, self.dropout_seed, sid, out=dx, dres=dres)]

        self._bwd_builders.append(build_bwd)
        return y

    def pool(self, x: Var, is_max: bool) -> Var:
        nd = self.ndim
        B, D, H, W, Cc = x.t.shape
        y = Var(self.empty((B, D // 2 if nd == 3 else 1, H // 2, W // 2, Cc)))
        xt, yt = x.t, y.t
        self.fwd.append(lambda: ops.pool2x(xt, nd, is_max, out=yt))

        def build_bwd():
            dy = self.grad_of(y)
            if dy is None or not x.needs_grad:
                return []
            dres, dx = self.contribute_compute(x)
            return [lambda: ops.pool2x_bwd(xt, dy, dx, nd, is_max, dres=dres)]

        self._bwd_builders.append(build_bwd)
        return y

    def add(self, a: Var, b: Var) -> Var:
        y = Var(self.empty(a.t.shape))
        at, bt, yt = a.t, b.t, y.t
        self.fwd.append(lambda: ops.add(at, bt, out=yt))

        def build_bwd():
            dy = self.grad_of(y)
            if dy is None:
                return []
            return self.contribute_copy(a, dy) + self.contribute_copy(b, dy)

        self._bwd_builders.append(build_bwd)
        return y

    def add_vec(self, a: Var, b: Var) -> Var:
        """fp32 [B, M] vectors: y = a + b (te + ye, punetg.py:410 / adm.py:1050-1051); dy flows to whichever input needs it."""
        y = Var(self.empty(a.t.shape, torch.float32))
        at, bt, yt = a.t, b.t, y.t
        self.fwd.append(lambda: ops.add_ex(at, bt, yt))

        def build_bwd():
            dy = self.grad_of(y)
            if dy is None:
                return []
            out = []
            for v in (a, b):
                if v.needs_grad:
                    out += self.contribute_copy(v, dy)
            return out

        self._bwd_builders.append(build_bwd)
        return y

    def silu_vec(self, x: Var) -> Var:
        """fp32 vectors: y = SiLU(x) (ADM's act_final after the conditioning add, adm.py:1052)."""
        y = Var(self.empty(x.t.shape, torch.float32))
        xt, yt = x.t, y.t
        self.fwd.append(lambda: ops.silu_fwd(xt, yt))

        def build_bwd():
            dy = self.grad_of(y)
            if dy is None or not x.needs_grad:
                return []
            dres, dx = self.contribute_compute(x)
            assert dres is None, "silu_vec input with several consumers is not supported"
            return [lambda: ops.silu_bwd(xt, dy, dx)]

        self._bwd_builders.append(build_bwd)
        return y

    def concat(self, a: Var, b: Var) -> Var:
        y = Var(self.empty(a.t.shape[:-1] + (a.t.shape[-1] + b.t.shape[-1],)))
        at, bt, yt = a.t, b.t, y.t
        self.fwd.append(lambda: ops.concat_channels(at, bt, out=yt))

        def build_bwd():
            dy = self.grad_of(y)
            if dy is None:
                return []
            ra, da = self.contribute_compute(a)
            rb, db = self.contribute_compute(b)
            return [lambda: ops.split_channels(dy, da, db, ra, rb)]

        self._bwd_builders.append(build_bwd)
        return y

    def fourier(self, t_buf: torch.Tensor, W: torch.Tensor) -> Var:
        y = Var(self.empty((self.B, 2 * W.shape[0]), torch.float32), needs_grad=False)
        yt = y.t
        self.fwd.append(lambda: ops.fourier(t_buf, W, out=yt))
        return y

    def linear(self, x: Var, weight: torch.Tensor, bias: Optional[torch.Tensor], silu: bool, param_w: torch.Tensor,
               param_b: Optional[torch.Tensor], rows: Optional[slice] = None) -> Var:
        """y = [act](x W^T + b) on fp32 [B, K] vectors (time MLPs, MLPUncond); silu: False / True (SiLU) / 2 (ReLU).
        `weight`/`bias` may be row slices of the parameters
        `param_w`/`param_b` (ADM's packed FiLM projection, adm.py:331-343); `rows` names that slice for the gradient."""
        Bn, K = x.t.shape
        N = weight.shape[0]
        f32 = torch.float32
        z = self.empty((Bn, N), f32)
        y = Var(self.empty((Bn, N), f32) if silu else z)
        xt, yt = x.t, y.t
        wd = weight.detach()
        bd = bias.detach() if bias is not None else None
        self.fwd.append(lambda: ops.gemm_ex(xt, wd, z, M=Bn, N=N, K=K, lda=K, ldb=K, ldc=N, bias=bd, transB=True))
        act = int(silu)
        if silu:
            self.fwd.append(lambda: ops.act_fwd(z, yt, act))

        def build_bwd():
            dy = self.grad_of(y)
            if dy is None:
                return []
            out = []
            if silu:
                dz = self.empty((Bn, N), f32)
                out.append(lambda: ops.act_bwd(z, dy, dz, act))
            else:
                dz = dy
            gw = self.grad_view(param_w, rows)
            # dW[n][k] = sum_b dz[b][n] x[b][k]  (A = dz stored [K'=B][M'=N])
            out.append(lambda: ops.gemm_ex(dz, xt, gw, M=N, N=K, K=Bn, lda=N, ldb=K, ldc=K, transA=True, transB=False))
            if param_b is not None:
                gbv = self.grad_view(param_b, rows)
                out.append(lambda: ops.colsum(dz, gbv))
            if x.needs_grad:
                dres, dx = self.contribute_compute(x)
                if dres is not None and dres is not dx:
                    out.append(lambda: ops.add_ex(dres, None, dx))
                beta = 0.0 if dres is None else 1.0
                out.append(lambda: ops.gemm_ex(dz, wd, dx, M=Bn, N=K, K=N, lda=N, ldb=K, ldc=K, transB=False, beta=beta))
            return out

        self._bwd_builders.append(build_bwd)
        return y

    def attention(self, x: Var, mha, residual: bool) -> Var:
        """nn.MultiheadAttention(C, 1 head) self-attention over the spatial positions (nets/attention.py:54-102), fp32
        CUDA-core GEMMs in both directions (the training path keeps Q, K, V and the probabilities for the backward)."""
        B = self.B
        Cc = x.t.shape[-1]
        Lq = x.t.numel() // (B * Cc)
        f32 = torch.float32
        import diffsci_b200
        if (self.precision == "bf16" and diffsci_b200.TC_CONV_ENABLED and Cc % 64 == 0 and Lq % 8 == 0 and Lq <= 8192):
            return self._attention_tc(x, mha, residual)
        if HYBRID_ATTENTION and self.precision == "bf16" and diffsci_b200.TC_CONV_ENABLED and Cc % 64 == 0 and (B * Lq) % 8 == 0:
            return self._attention_hybrid(x, mha, residual)
        y = Var(self.empty(x.t.shape))
        xt, yt = x.t, y.t
        tok = xt.view(B, Lq, Cc) if xt.dtype == f32 else self.empty((B, Lq, Cc), f32)
        bufs = dict(qkv=self.empty((B * Lq, 3 * Cc), f32), scores=self.empty((B, Lq, Lq), f32),
                    ao=self.empty((B * Lq, Cc), f32), out=self.empty((B, Lq, Cc), f32))

        def fwd():
            if tok.data_ptr() != xt.data_ptr():
                ops.cast(xt.view(B, Lq, Cc), f32, out=tok)
            o = ops.self_attention_f32(tok, mha.in_proj_weight, mha.in_proj_bias, mha.out_proj.weight, mha.out_proj.bias, bufs,
                                       residual)
            ops.cast(o, yt.dtype, out=yt.view(B, Lq, Cc))

        self.fwd.append(fwd)

        def build_bwd():
            dy = self.grad_of(y)
            if dy is None:
                return []
            gwi, gbi = self.grad_view(mha.in_proj_weight), self.grad_view(mha.in_proj_bias)
            gwo, gbo = self.grad_view(mha.out_proj.weight), self.grad_view(mha.out_proj.bias)
            wi, wo = mha.in_proj_weight.detach(), mha.out_proj.weight.detach()
            qkv, P, ao = bufs["qkv"], bufs["scores"], bufs["ao"]
            M = B * Lq
            dout = dy.view(M, Cc) if dy.dtype == f32 else self.empty((M, Cc), f32)
            dO = self.empty((M, Cc), f32)
            dP = self.lazy_scratch("attn_dP", (B, Lq, Lq), f32)
            dqkv = self.empty((M, 3 * Cc), f32)
            dtok = self.empty((M, Cc), f32)
            alpha = Cc ** -0.5
            out = []
            if dout.data_ptr() != dy.data_ptr():
                out.append(lambda: ops.cast(dy.view(M, Cc), f32, out=dout))
            # out = ao Wo^T + bo :  dWo = dout^T ao ; dbo = colsum(dout) ; dO = dout Wo
            out.append(lambda: ops.gemm_ex(dout, ao, gwo, M=Cc, N=Cc, K=M, lda=Cc, ldb=Cc, ldc=Cc, transA=True, transB=False))
            out.append(lambda: ops.colsum(dout, gbo))
            out.append(lambda: ops.gemm_ex(dout, wo, dO, M=M, N=Cc, K=Cc, lda=Cc, ldb=Cc, ldc=Cc, transB=False))
            s3 = Lq * 3 * Cc
            # O = P V :  dP = dO V^T ; dV = P^T dO
            out.append(lambda: ops.gemm_ex(dO, qkv, dP(), M=Lq, N=Lq, K=Cc, lda=Cc, ldb=3 * Cc, ldc=Lq, transB=True, batch=B,
                                           strideA=Lq * Cc, strideB=s3, strideC=Lq * Lq, b_off=2 * Cc))
            out.append(lambda: ops.gemm_ex(P, dO, dqkv, M=Lq, N=Cc, K=Lq, lda=Lq, ldb=Cc, ldc=3 * Cc, transA=True, transB=False,
                                           batch=B, strideA=Lq * Lq, strideB=Lq * Cc, strideC=s3, c_off=2 * Cc))
            # P = softmax(alpha Q K^T) :  dS = P (dP - rowsum(dP P)) ; dQ = alpha dS K ; dK = alpha dS^T Q
            out.append(lambda: ops.softmax_bwd_rows(P, dP(), B * Lq, Lq))
            out.append(lambda: ops.gemm_ex(dP(), qkv, dqkv, M=Lq, N=Cc, K=Lq, lda=Lq, ldb=3 * Cc, ldc=3 * Cc, transB=False,
                                           alpha=alpha, batch=B, strideA=Lq * Lq, strideB=s3, strideC=s3, b_off=Cc, c_off=0))
            out.append(lambda: ops.gemm_ex(dP(), qkv, dqkv, M=Lq, N=Cc, K=Lq, lda=Lq, ldb=3 * Cc, ldc=3 * Cc, transA=True,
                                           transB=False, alpha=alpha, batch=B, strideA=Lq * Lq, strideB=s3, strideC=s3, b_off=0,
                                           c_off=Cc))
            # qkv = tok Wi^T + bi :  dWi = dqkv^T tok ; dbi = colsum(dqkv) ; dtok = dqkv Wi
            out.append(lambda: ops.gemm_ex(dqkv, tok.view(M, Cc), gwi, M=3 * Cc, N=Cc, K=M, lda=3 * Cc, ldb=Cc, ldc=Cc, transA=True,
                                           transB=False))
            out.append(lambda: ops.colsum(dqkv, gbi))
            if x.needs_grad:
                out.append(lambda: ops.gemm_ex(dqkv, wi, dtok, M=M, N=Cc, K=3 * Cc, lda=3 * Cc, ldb=Cc, ldc=Cc, transB=False))
                if residual:
                    out.append(lambda: ops.add_ex(dtok, dout, dtok))
                dres, dx = self.contribute_compute(x)
                out.append(lambda: ops.add_ex(dtok, dres.view(M, Cc) if dres is not None else None, dx.view(M, Cc)))
            return out

        self._bwd_builders.append(build_bwd)
        return y

    def _attention_hybrid(self, x: Var, mha, residual: bool) -> Var:
        """Token counts the tensor-core score path does not take (L % 8 != 0: the 7x7 bottom level of MNIST): the two
        projections and their gradients -- 99% of the block's FLOPs at small L -- run on tcgen05 over the flattened B*L token
        axis (weight gradients split over `nb` token slices, summed in a fixed order), the per-sample L x L part stays fp32."""
        B = self.B
        Cc = x.t.shape[-1]
        Lq = x.t.numel() // (B * Cc)
        M = B * Lq
        f32, bf = torch.float32, torch.bfloat16
        y = Var(self.empty(x.t.shape))
        tok, yt = x.t.view(M, Cc), y.t.view(M, Cc)
        qkv = self.empty((M, 3 * Cc), f32)
        P = self.empty((B, Lq, Lq), f32)
        ao = self.empty((M, Cc), f32)
        ao16 = self.empty((M, Cc), bf)
        wi, wo = ops.PackedLinear(mha.in_proj_weight), ops.PackedLinear(mha.out_proj.weight)
        self._packs += [wi, wo]
        bi, bo = mha.in_proj_bias.detach(), mha.out_proj.bias.detach()
        alpha = Cc ** -0.5
        s3, sLL, sLC = Lq * 3 * Cc, Lq * Lq, Lq * Cc
        G = ops.gemm_bf16_tc
        nb = max(n for n in range(1, 65) if M % n == 0 and (M // n) % 8 == 0)     # token slices of the weight gradients
        Kb = M // nb

        def fwd():
            G(tok, wi.packed(), qkv, M=M, N=3 * Cc, K=Cc, lda=Cc, ldb=Cc, ldc=3 * Cc, bias=bi)
            ops.attention_core_f32(qkv, P, ao, B, Lq, Cc)
            ops.cast(ao, bf, out=ao16)
            G(ao16, wo.packed(), yt, M=M, N=Cc, K=Cc, lda=Cc, ldb=Cc, ldc=Cc, bias=bo, residual=tok if residual else None)

        self.fwd.append(fwd)

        def build_bwd():
            dy = self.grad_of(y)
            if dy is None:
                return []
            gwi, gbi = self.grad_view(mha.in_proj_weight), self.grad_view(mha.in_proj_bias)
            gwo, gbo = self.grad_view(mha.out_proj.weight), self.grad_view(mha.out_proj.bias)
            dyt = dy.view(M, Cc)
            dO = self.empty((M, Cc), f32)
            dP = self.lazy_scratch("attn_dP", (B, Lq, Lq), f32)
            dqkv = self.empty((M, 3 * Cc), f32)
            dqkv16 = self.empty((M, 3 * Cc), bf)
            wsp = self.lazy_scratch("attn_wsplit", (nb, 3 * Cc * Cc), f32)
            csws = self.lazy_scratch("bwd_ws", ops.bwd_ws_bytes(B, Lq, 3 * Cc))
            out = []
            # out = ao Wo^T + bo :  dWo = dy^T ao ; dbo ; dO = dy Wo
            out.append(lambda: G(dyt, ao16, wsp()[:, :Cc * Cc], M=Cc, N=Cc, K=Kb, lda=Cc, ldb=Cc, ldc=Cc, batch=nb, strideA=Kb * Cc,
                                 strideB=Kb * Cc, strideC=3 * Cc * Cc, transA=True, transB=True))
            out.append(lambda: ops.colsum(wsp()[:, :Cc * Cc], gwo.view(-1), ld=3 * Cc * Cc))
            out.append(lambda: ops.channel_sum(dy.view(B, Lq, Cc), gbo, csws(), False))
            out.append(lambda: G(dyt, wo.packed(), dO, M=M, N=Cc, K=Cc, lda=Cc, ldb=Cc, ldc=Cc, transB=True))
            # the L x L part in fp32 (as TrainGraph.attention):  dP = dO V^T ; dV = P^T dO ; dS ; dQ = alpha dS K ; dK = alpha dS^T Q
            out.append(lambda: ops.gemm_ex(dO, qkv, dP(), M=Lq, N=Lq, K=Cc, lda=Cc, ldb=3 * Cc, ldc=Lq, transB=True, batch=B,
                                           strideA=sLC, strideB=s3, strideC=sLL, b_off=2 * Cc))
            out.append(lambda: ops.gemm_ex(P, dO, dqkv, M=Lq, N=Cc, K=Lq, lda=Lq, ldb=Cc, ldc=3 * Cc, transA=True, transB=False,
                                           batch=B, strideA=sLL, strideB=sLC, strideC=s3, c_off=2 * Cc))
            out.append(lambda: ops.softmax_bwd_rows(P, dP(), B * Lq, Lq))
            out.append(lambda: ops.gemm_ex(dP(), qkv, dqkv, M=Lq, N=Cc, K=Lq, lda=Lq, ldb=3 * Cc, ldc=3 * Cc, transB=False,
                                           alpha=alpha, batch=B, strideA=sLL, strideB=s3, strideC=s3, b_off=Cc, c_off=0))
            out.append(lambda: ops.gemm_ex(dP(), qkv, dqkv, M=Lq, N=Cc, K=Lq, lda=Lq, ldb=3 * Cc, ldc=3 * Cc, transA=True,
                                           transB=False, alpha=alpha, batch=B, strideA=sLL, strideB=s3, strideC=s3, b_off=0,
                                           c_off=Cc))
            # qkv = tok Wi^T + bi :  dWi = dqkv^T tok ; dbi ; dtok = dqkv Wi
            out.append(lambda: ops.cast(dqkv, bf, out=dqkv16))
            out.append(lambda: G(dqkv16, tok, wsp(), M=3 * Cc, N=Cc, K=Kb, lda=3 * Cc, ldb=Cc, ldc=Cc, batch=nb, strideA=Kb * 3 * Cc,
                                 strideB=Kb * Cc, strideC=3 * Cc * Cc, transA=True, transB=True))
            out.append(lambda: ops.colsum(wsp(), gwi.view(-1)))
            out.append(lambda: ops.colsum(dqkv, gbi))
            if x.needs_grad:
                dres, dx = self.contribute_compute(x)
                dxt = dx.view(M, Cc)
                out.append(lambda: G(dqkv16, wi.packed(), dxt, M=M, N=Cc, K=3 * Cc, lda=3 * Cc, ldb=Cc, ldc=Cc, transB=True,
                                     residual=dyt if residual else (dres.view(M, Cc) if dres is not None else None)))
                if residual and dres is not None:
                    out.append(lambda: ops.add_ex(dx, dres, dx))
            return out

        self._bwd_builders.append(build_bwd)
        return y

    def _attention_tc(self, x: Var, mha, residual: bool) -> Var:
        """The same attention block with every product on the tensor cores (bf16 operands, fp32 accumulation): forward as
        ops.self_attention_tc but keeping Q|K|V, P and the attention output; backward = 9 dsk_gemm_bf16_tc launches (the
        A^T B products take their operands MN-major, as they lie) + the bf16 softmax backward."""
        B = self.B
        Cc = x.t.shape[-1]
        Lq = x.t.numel() // (B * Cc)
        M = B * Lq
        f32, bf = torch.float32, torch.bfloat16
        y = Var(self.empty(x.t.shape))
        tok, yt = x.t.view(M, Cc), y.t.view(M, Cc)
        qkv = self.empty((M, 3 * Cc), bf)
        P = self.empty((B, Lq, Lq), bf)
        ao = self.empty((M, Cc), bf)
        sc = self.lazy_scratch("attn_scores", (B, Lq, Lq), f32)
        wi, wo = ops.PackedLinear(mha.in_proj_weight), ops.PackedLinear(mha.out_proj.weight)
        self._packs += [wi, wo]
        bi, bo = mha.in_proj_bias.detach(), mha.out_proj.bias.detach()
        alpha = Cc ** -0.5
        s3, sLL, sLC = Lq * 3 * Cc, Lq * Lq, Lq * Cc
        G = ops.gemm_bf16_tc

        def fwd():
            G(tok, wi.packed(), qkv, M=M, N=3 * Cc, K=Cc, lda=Cc, ldb=Cc, ldc=3 * Cc, bias=bi)
            G(qkv, qkv, sc(), M=Lq, N=Lq, K=Cc, lda=3 * Cc, ldb=3 * Cc, ldc=Lq, alpha=alpha, batch=B, strideA=s3, strideB=s3,
              strideC=sLL, b_off=Cc)
            ops.softmax_rows_bf16(sc(), P, B * Lq, Lq)
            G(P, qkv, ao, M=Lq, N=Cc, K=Lq, lda=Lq, ldb=3 * Cc, ldc=Cc, batch=B, strideA=sLL, strideB=s3, strideC=sLC,
              b_off=2 * Cc, transB=True)                                              # O = P V, V as it lies ([L, C] = [K, N])
            G(ao, wo.packed(), yt, M=M, N=Cc, K=Cc, lda=Cc, ldb=Cc, ldc=Cc, bias=bo, residual=tok if residual else None)

        self.fwd.append(fwd)

        def build_bwd():
            dy = self.grad_of(y)
            if dy is None:
                return []
            gwi, gbi = self.grad_view(mha.in_proj_weight), self.grad_view(mha.in_proj_bias)
            gwo, gbo = self.grad_view(mha.out_proj.weight), self.grad_view(mha.out_proj.bias)
            dyt = dy.view(M, Cc)
            dAO = self.empty((M, Cc), bf)
            dqkv = self.empty((M, 3 * Cc), bf)
            dP = self.lazy_scratch("attn_scores", (B, Lq, Lq), f32)
            dS = self.lazy_scratch("attn_dS", (B, Lq, Lq), bf)
            wsp = self.lazy_scratch("attn_wsplit", (B, 3 * Cc * Cc), f32)
            csws = self.lazy_scratch("bwd_ws", ops.bwd_ws_bytes(B, Lq, 3 * Cc))
            out = []
            # out = ao Wo^T + bo :  dWo = dy^T ao (split over the batch, summed in a fixed order) ; dbo ; dAO = dy Wo
            out.append(lambda: G(dyt, ao, wsp()[:, :Cc * Cc], M=Cc, N=Cc, K=Lq, lda=Cc, ldb=Cc, ldc=Cc, batch=B, strideA=sLC,
                                 strideB=sLC, strideC=3 * Cc * Cc, transA=True, transB=True))
            out.append(lambda: ops.colsum(wsp()[:, :Cc * Cc], gwo.view(-1), ld=3 * Cc * Cc))
            out.append(lambda: ops.channel_sum(dy.view(B, Lq, Cc), gbo, csws(), False))
            out.append(lambda: G(dyt, wo.packed(), dAO, M=M, N=Cc, K=Cc, lda=Cc, ldb=Cc, ldc=Cc, transB=True))
            # O = P V :  dP = dAO V^T ; dV = P^T dAO
            out.append(lambda: G(dAO, qkv, dP(), M=Lq, N=Lq, K=Cc, lda=Cc, ldb=3 * Cc, ldc=Lq, batch=B, strideA=sLC, strideB=s3,
                                 strideC=sLL, b_off=2 * Cc))
            out.append(lambda: G(P, dAO, dqkv, M=Lq, N=Cc, K=Lq, lda=Lq, ldb=Cc, ldc=3 * Cc, batch=B, strideA=sLL, strideB=sLC,
                                 strideC=s3, c_off=2 * Cc, transA=True, transB=True))
            # P = softmax(alpha Q K^T) :  dS = P (dP - rowsum(dP P)) ; dQ = alpha dS K ; dK = alpha dS^T Q
            out.append(lambda: ops.softmax_bwd_rows_bf16(P, dP(), dS(), B * Lq, Lq))
            out.append(lambda: G(dS(), qkv, dqkv, M=Lq, N=Cc, K=Lq, lda=Lq, ldb=3 * Cc, ldc=3 * Cc, alpha=alpha, batch=B,
                                 strideA=sLL, strideB=s3, strideC=s3, b_off=Cc, c_off=0, transB=True))
            out.append(lambda: G(dS(), qkv, dqkv, M=Lq, N=Cc, K=Lq, lda=Lq, ldb=3 * Cc, ldc=3 * Cc, alpha=alpha, batch=B,
                                 strideA=sLL, strideB=s3, strideC=s3, b_off=0, c_off=Cc, transA=True, transB=True))
            # qkv = tok Wi^T + bi :  dWi = dqkv^T tok ; dbi ; dtok = dqkv Wi
            out.append(lambda: G(dqkv, tok, wsp(), M=3 * Cc, N=Cc, K=Lq, lda=3 * Cc, ldb=Cc, ldc=Cc, batch=B, strideA=s3,
                                 strideB=sLC, strideC=3 * Cc * Cc, transA=True, transB=True))
            out.append(lambda: ops.colsum(wsp(), gwi.view(-1)))
            out.append(lambda: ops.channel_sum(dqkv.view(B, Lq, 3 * Cc), gbi, csws(), False))
            if x.needs_grad:
                dres, dx = self.contribute_compute(x)
                dxt = dx.view(M, Cc)
                out.append(lambda: G(dqkv, wi.packed(), dxt, M=M, N=Cc, K=3 * Cc, lda=3 * Cc, ldb=Cc, ldc=Cc, transB=True,
                                     residual=dyt if residual else (dres.view(M, Cc) if dres is not None else None)))
                if residual and dres is not None:
                    out.append(lambda: ops.add_ex(dx, dres, dx))
            return out

        self._bwd_builders.append(build_bwd)
        return y

    # ------------------------------------------------------------------ finalisation / execution
    def finalize(self, output: Var):
        """Set the gradient slot of the network output and generate the backward launch list."""
        output.g, output.state = self.empty(output.t.shape, output.t.dtype), OWN
        self.output = output
        self.bwd = []
        for build in reversed(self._bwd_builders):
            n0 = len(self._written_log)
            self.bwd.extend(build())
            for pid in self._written_log[n0:]:
                self.grad_ready_pos[pid] = len(self.bwd)
        self._bwd_builders = []
        owned = {id(p) for p in self.params}
        missing = [n for n, p in self.net.named_parameters()
                   if id(p) in owned and p.requires_grad and not any(k[0] == id(p) for k in self._written)]
        if missing:
            raise RuntimeError(f"TrainGraph: no backward op writes the gradient of {missing[:4]}...")
        self.prepare()

    def prepare(self):
        """(Re)pack derived weight layouts; called before every forward so optimizer steps are picked up."""
        if getattr(self, "_packer", None) is None or self._packer_n != len(self._packs):
            self._packer, self._packer_n = ops.MultiPacker(self._packs), len(self._packs)
        self._packer.pack()

    def run_forward(self):
        # the saved activations live in this graph's static buffers: every forward invalidates the previous one's
        self.generation = getattr(self, "generation", 0) + 1
        self.prepare()
        if self.dropout_sites:
            self.dropout_seed = (int(torch.randint(0, 2 ** 62, (1,)).item()) if self.fixed_dropout_seed is None
                                 else int(self.fixed_dropout_seed))
        for f in self.fwd:
            f()

    def run_backward(self, hooks: Optional[dict] = None):
        """hooks: {number of completed backward launches: callable} -- e.g. the per-bucket gradient all-reduce of the
        data-parallel trainer, issued as soon as the bucket's last writer has been enqueued."""
        if not hooks:
            for f in self.bwd:
                f()
            return
        for i, f in enumerate(self.bwd):
            f()
            h = hooks.get(i + 1)
            if h is not None:
                h()

    def grads(self) -> list:
        return [self._gviews[id(p)] for p in self.params]


def grouped_linear_layer(g: TrainGraph, xs: list, specs: list, silu: bool, shared_x: bool = False) -> list:
    """One dsk_grouped_linear launch for many small fp32 linears y_i = [silu](x_i W_i^T + b_i)  (the time MLPs of every
    ResNet block / ADM's FiLM projections), and ONE dsk_grouped_linear_bwd (two kernels) for all their gradients.

    xs: input Vars (one per group; the same Var for every group if `shared_x`).  specs: (weight, bias, param_w, param_b,
    rows) per group -- weight/bias may be row slices of the parameters, `rows` names the slice for the gradient view.
    The builder is registered at the point of the call, so calling this BEFORE the consumers are built makes the backward
    run after every consumer has contributed."""
    f32 = torch.float32
    Bn = xs[0].t.shape[0]
    ws = [sp[0].detach() for sp in specs]
    bs = [sp[1].detach() if sp[1] is not None else None for sp in specs]
    zs = [g.empty((Bn, w.shape[0]), f32) for w in ws] if silu else None
    ys = [Var(g.empty((Bn, w.shape[0]), f32)) for w in ws]
    layer = ops.GroupedLinear([x.t for x in xs], ws, bs, [y.t for y in ys], 1 if silu else 0, zs)
    g.fwd.append(layer.run)

    def build_bwd():
        dys = []
        for y in ys:
            dy = g.grad_of(y)
            if dy is None:
                raise RuntimeError("grouped_linear_layer: an output has no consumer")
            dys.append(dy)
        dzs = [g.empty(dy.shape, f32) for dy in dys] if silu else dys
        dws = [g.grad_view(sp[2], sp[4]) for sp in specs]
        dbs = [g.grad_view(sp[3], sp[4]) if sp[3] is not None else None for sp in specs]
        dxs = None
        if xs[0].needs_grad:
            if shared_x:
                dres, dx = g.contribute_compute(xs[0])
                assert dres is None, "shared input with other consumers is not supported"
                dxs = [dx] + [None] * (len(xs) - 1)
            else:
                dxs = []
                for x in xs:
                    dres, dx = g.contribute_compute(x)
                    assert dres is None, "time-MLP activations have exactly one consumer"
                    dxs.append(dx)
        return [layer.backward_tables(dys, dzs, dws, dbs, dxs, shared_dx=shared_x)]

    g._bwd_builders.append(build_bwd)
    return ys


def time_blocks(g: TrainGraph, te: Var, tbs: list) -> list:
    """ResnetTimeBlock (commonlayers.py:516-550) of EVERY ResNet block at once: Linear-SiLU-Linear-SiLU-Linear on the
    shared Fourier embedding = three grouped launches forward, three grouped backward launches."""
    spec = lambda i: [(tb.net[i].weight, tb.net[i].bias, tb.net[i].weight, tb.net[i].bias, None) for tb in tbs]  # noqa: E731
    h = grouped_linear_layer(g, [te] * len(tbs), spec(0), True, shared_x=True)
    h = grouped_linear_layer(g, h, spec(2), True)
    return grouped_linear_layer(g, h, spec(4), False)


def build_punetg(net, B: int, spatial: tuple, device, precision: str, cond: bool = False, dropout: bool = True) -> TrainGraph:
    """PUNetG.forward (nets/punetg.py:389-416) unrolled into a TrainGraph.  cond: te + ye with ye an input [B, M] whose
    gradient is returned (the conditional embedding that produced it is trained by torch autograd around this graph)."""
    c = net.config
    nd = c.dimension
    if not 0.0 <= c.dropout < 1.0:
        raise ValueError(f"dropout probability has to be in [0, 1), got {c.dropout}")
    g = TrainGraph(net, B, device, precision, nd)
    names = {id(m): n for n, m in net.named_modules()}
    sp = (1,) + tuple(spatial) if nd == 2 else tuple(spatial)
    nlev = len(c.channel_expansion)
    for l in range(nlev):
        if any(s % (2 ** (l + 1)) for s in spatial):
            raise ValueError(f"PUNetG: spatial size {spatial} is not divisible by 2 at a down-sampling level")
    g.t_in = torch.empty(B, dtype=torch.float32, device=g.device)
    g.x_in = Var(g.empty((B,) + sp + (net.convin.cin,)), needs_grad=False)
    te = g.fourier(g.t_in, net.time_projection.W)
    if cond:
        g.ye_in = Var(g.empty((B, c.model_channels), torch.float32), needs_grad=True, name="ye")
        te = g.add_vec(te, g.ye_in)
    blocks = [b for l in range(nlev) for b in net.downward_blocks[l]] + list(net.before_block) + list(net.attn_resnet_block) + \
        list(net.after_block) + [b for i in range(nlev) for b in net.upward_blocks[i]]
    tvs = dict(zip((id(b) for b in blocks), time_blocks(g, te, [b.timeblock for b in blocks])))

    def resblock(x: Var, blk) -> Var:
        C = blk.channels
        tv = tvs[id(blk)]
        n1 = g.norm(x, blk.gnorm1, C, c.first_resblock_norm, True)
        y = g.conv(n1, blk.conv1, chan_bias=tv, want_stats=True)
        n2 = g.norm(y, blk.gnorm2, C, c.second_resblock_norm, True)
        if c.dropout > 0.0 and dropout:         # conv2(dropout(act(gnorm2(y)))), commonlayers.py:829-831; off in eval mode
            n2 = g.dropout(n2, c.dropout, names[id(blk)])
        return g.conv(n2, blk.conv2, residual=x, want_stats=True)

    x = g.conv(g.x_in, net.convin)
    skips = []
    for l in range(nlev):
        for blk in net.downward_blocks[l]:
            x = resblock(x, blk)
        skips.append(x)
        x = g.conv(g.pool(x, True), net.downsamplers[l].conv, want_stats=True)
    for blk in net.before_block:
        x = resblock(x, blk)
    xa = x
    for r, blk in enumerate(net.attn_resnet_block):
        xa = resblock(xa, blk)
        if r < len(net.attn_block):
            xa = g.attention(xa, net.attn_block[r].mhattn, c.attn_residual)
    x = g.add(x, xa)
    for blk in net.after_block:
        x = resblock(x, blk)
    for i in range(nlev):
        x = g.conv(x, net.upsamplers[i].conv, residual=skips.pop(), up2=True, want_stats=True)
        for blk in net.upward_blocks[i]:
            x = resblock(x, blk)
    g.finalize(g.conv(x, net.convout, few_out_ok=True))
    return g


def build_adm(net, B: int, spatial: tuple, device, precision: str, cond: bool = False, dropout: bool = True) -> TrainGraph:
    """ADM.forward (nets/adm.py:199-216; block :292-343) unrolled into a TrainGraph.  cond: te = SiLU(mlp(fourier) + ye) with
    ye an input [B, output_embed_dim] whose gradient is returned (adm.py:1047-1053)."""
    c = net.config
    pdrop = float(getattr(c, "dropout", 0.0))
    if not 0.0 <= pdrop < 1.0:
        raise ValueError(f"dropout probability has to be in [0, 1), got {pdrop}")
    g = TrainGraph(net, B, device, precision, 2)
    names = {id(m): n for n, m in net.named_modules()}
    H, W = spatial
    nlev = len(c.channel_expansion)
    if H % (2 ** nlev) or W % (2 ** nlev):
        raise ValueError(f"ADM: spatial size {spatial} must be divisible by {2 ** nlev}")
    G = c.num_groups
    g.t_in = torch.empty(B, dtype=torch.float32, device=g.device)
    g.x_in = Var(g.empty((B, 1, H, W, c.input_channels)), needs_grad=False)
    four = g.fourier(g.t_in, net.time_embedding.projection.W)
    mlp = net.time_embedding.mlp
    h1 = grouped_linear_layer(g, [four], [(mlp[0].weight, mlp[0].bias, mlp[0].weight, mlp[0].bias, None)], True)[0]
    if cond:
        z = grouped_linear_layer(g, [h1], [(mlp[2].weight, mlp[2].bias, mlp[2].weight, mlp[2].bias, None)], False)[0]
        g.ye_in = Var(g.empty((B, c.output_embed_dim), torch.float32), needs_grad=True, name="ye")
        te = g.silu_vec(g.add_vec(z, g.ye_in))
    else:
        te = grouped_linear_layer(g, [h1], [(mlp[2].weight, mlp[2].bias, mlp[2].weight, mlp[2].bias, None)], True)[0]  # + act_final SiLU (adm.py:1047-1053)
    blocks = [b for layer in net.encoder.layers for b in layer.input_blocks] + list(net.middle_block.middle_blocks) + \
        [b for layer in net.decoder.layers for b in layer.input_blocks]
    specs = []
    for b in blocks:                                   # FiLM projection (adm.py:331-343): rows [:C] scale, [C:] shift
        w, bb, Cc = b.embed_linear.weight, b.embed_linear.bias, b.cout
        specs += [(w[:Cc], bb[:Cc], w, bb, slice(0, Cc)), (w[Cc:], bb[Cc:], w, bb, slice(Cc, 2 * Cc))]
    films = grouped_linear_layer(g, [te] * len(specs), specs, False, shared_x=True)
    film_of = {id(b): (films[2 * i], films[2 * i + 1]) for i, b in enumerate(blocks)}

    def block(x: Var, blk) -> Var:
        down, up = blk.sample == "down", blk.sample == "up"
        n = g.norm(x, blk.norm1, G, c.first_resblock_norm, True)
        xr = x
        if down:
            n, xr = g.pool(n, False), g.pool(x, False)
        y = g.conv(n, blk.conv1, up2=up)
        h = g.norm(y, blk.norm2, G, c.second_resblock_norm, True, film=film_of[id(blk)])
        if pdrop > 0.0 and dropout:             # conv2(dropout(SiLU(FiLM(norm2)))), adm.py:305-313; off in eval mode
            h = g.dropout(h, pdrop, names[id(blk)])
        r = g.conv(xr, blk.convresidual, up2=up)
        o = g.conv(h, blk.conv2, residual=r)
        if blk.has_attn:
            o = g.attention(o, blk.attn.mhattn, c.attn_residual)
        return o

    x = g.conv(g.x_in, net.input_layer)
    skips = [x]
    for layer in net.encoder.layers:
        for blk in layer.input_blocks:
            x = block(x, blk)
        skips.append(x)
    for blk in net.middle_block.middle_blocks:
        x = block(x, blk)
    for layer in net.decoder.layers:
        hskip = skips.pop()
        x = g.concat(x, hskip) if c.skip_integration_type == "concat" else g.add(x, hskip)
        for blk in layer.input_blocks:
            x = block(x, blk)
    g.finalize(g.conv(x, net.output_layer, few_out_ok=True))
    return g


def build_mlp(net, B: int, device) -> TrainGraph:
    """MLPUncond.forward (nets/mlp.py:38-58) unrolled into a TrainGraph: cat[x, t] -> (Linear, act)* -> Linear, fp32
    CUDA-core GEMMs (the toy configuration the reference trains in tests/test_karras_on_toy_dataset.py:86-93)."""
    from .layers import LinearParams
    if net.dropout > 0:
        raise NotImplementedError("diffsci_b200.MLPUncond: training-mode dropout not built")
    g = TrainGraph(net, B, device, "fp32", 2)
    g.flat_io = True
    g.t_in = torch.empty(B, dtype=torch.float32, device=g.device)
    g.x_in = Var(g.empty((B, net.dim), torch.float32), needs_grad=False)
    cat = Var(g.empty((B, net.dim + 1), torch.float32), needs_grad=False)
    xt, tt, ct = g.x_in.t, g.t_in, cat.t
    g.fwd.append(lambda: ops.concat_channels(xt, tt.view(B, 1), out=ct))
    ydim = int(getattr(net, "ydim", 0))
    if ydim:                                            # MLPCond: cat[x, t, y] (nets/mlp.py:118-120); y carries no gradient
        g.y_in = torch.empty((B, ydim), dtype=torch.float32, device=g.device)
        cat2 = Var(g.empty((B, net.dim + 1 + ydim), torch.float32), needs_grad=False)
        yt, c2 = g.y_in, cat2.t
        g.fwd.append(lambda: ops.concat_channels(ct, yt, out=c2))
        cat = cat2
    layers = [m for m in net.net if isinstance(m, LinearParams)]
    h = cat
    for i, m in enumerate(layers):
        last = i == len(layers) - 1
        h = g.linear(h, m.weight, m.bias, False if last else net.act, m.weight, m.bias)
    g.finalize(h)
    return g


class NetFunction(torch.autograd.Function):
    """Autograd seam: F = net(x, t[, ye]) with the hand-written backward; gradients flow to the nn.Parameters and to the
    conditioning vector ye (so that a torch conditional_embedding trains), not to x."""

    @staticmethod
    def forward(ctx, graph: TrainGraph, x: torch.Tensor, t: torch.Tensor, ye: Optional[torch.Tensor], *params):
        ctx.graph = graph
        if x.requires_grad:
            raise NotImplementedError("diffsci_b200: gradients with respect to the network INPUT are not built (the native "
                                      "backward stops at the parameters and the conditioning vector); detach x")
        if (ye is None) != (graph.ye_in is None):
            raise RuntimeError("NetFunction: conditioning vector does not match the graph (build with cond=True)")
        if ye is not None:
            graph.ye_in.t.copy_(ye.detach().float())
        out = graph.forward_nchw(x, t)
        ctx.generation = graph.generation
        return out

    @staticmethod
    def backward(ctx, dF):
        g = ctx.graph
        if ctx.generation != g.generation:
            raise RuntimeError(
                "diffsci_b200: backward of a network call whose saved activations were overwritten by a later forward of the "
                "same (batch, shape) training graph -- the graph keeps ONE set of activation buffers.  Call backward() before "
                "the next forward (e.g. sum losses over separate backward passes), or run the extra forward under no_grad.")
        g.backward_nchw(dF)
        dye = None
        if g.ye_in is not None and ctx.needs_input_grad[3]:
            dye = g.grad_of(g.ye_in).clone()
        # clones: AccumulateGrad may keep the tensors it is handed, and flat_grad is rewritten by the next step
        return (None, None, None, dye) + tuple(v.clone() for v in g.grads())


def _forward_nchw(self: TrainGraph, x: torch.Tensor, t: Optional[torch.Tensor]) -> torch.Tensor:
    """x fp32 [B, C, *S] -> F fp32 [B, Cout, *S] (user layout in/out, channels-last inside)."""
    if t is None:
        raise NotImplementedError("diffsci_b200: training with t=None (zero time embedding) is not built")
    if getattr(self, "flat_io", False):       # MLPUncond: [B, dim] vectors, no layout change
        self.x_in.t.copy_(x.float().reshape(self.x_in.t.shape))
        self.t_in.copy_(t.float().reshape(-1))
        self.run_forward()
        return self.output.t.reshape(x.shape).clone()
    ops.nchw_to_cl(x.float(), self.act_dtype, self.ndim, out=self.x_in.t)
    self.t_in.copy_(t.float().reshape(-1))
    self.run_forward()
    return ops.cl_to_nchw(self.output.t, self.ndim)


def _backward_nchw(self: TrainGraph, dF: torch.Tensor, hooks: Optional[dict] = None) -> None:
    if getattr(self, "flat_io", False):
        self.output.g.copy_(dF.float().reshape(self.output.g.shape))
        return self.run_backward(hooks)
    ops.nchw_to_cl(dF.float().contiguous(), self.output.t.dtype, self.ndim, out=self.output.g)
    self.run_backward(hooks)


TrainGraph.forward_nchw = _forward_nchw
TrainGraph.backward_nchw = _backward_nchw
