"""CPU oracle for the score networks on the hot path (TEST INFRASTRUCTURE ONLY).

Functional, state-dict driven PyTorch-CPU restatements of ``PUNetG``, ``ADM`` and
``MLPUncond`` forward passes.  Each function cites the reference lines it follows.  The
heavy lifting is ATen (``conv2d/3d``, ``group_norm``, ``scaled_dot_product_attention`` via
``multi_head_attention_forward`` semantics), i.e. the same third-party library the
reference dispatches to.  Pinned against the live reference by ``tests/golden/*.pt``
(``tests/test_oracle_vs_golden.py``).  Never imported by the product path.

``cfg`` is any object exposing the reference config attribute names
(``PUNetGConfig`` punetg_config.py:8-38, ``ADMConfig`` adm.py:9-35).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def _conv(x, w, b, dim):
    """torch.nn.Conv{2,3}d(padding='same'), stride 1 (odd kernels -> symmetric zero padding)."""
    pad = w.shape[-1] // 2
    return (F.conv2d if dim == 2 else F.conv3d)(x, w, b, padding=pad)


def _circular_conv(x, w, b, dim):
    """CircularConv2d / CircularConv3d.forward (commonlayers.py:955-972, 1010-1032), all axes circular: pad W, then H
    (then D) with mode='circular', then a padding-free convolution."""
    p = w.shape[-1] // 2
    if dim == 2:
        x = F.pad(x, (p, p, 0, 0), mode="circular")
        x = F.pad(x, (0, 0, p, p), mode="circular")
        return F.conv2d(x, w, b)
    x = F.pad(x, (p, p, 0, 0, 0, 0), mode="circular")
    x = F.pad(x, (0, 0, p, p, 0, 0), mode="circular")
    x = F.pad(x, (0, 0, 0, 0, p, p), mode="circular")
    return F.conv3d(x, w, b)


def _pconv(x, sd, name, cfg):
    """A PUNetG convolution by state-dict name: convolution_type 'default' -> keys <name>.weight/.bias, zero padding;
    'circular' -> keys <name>.conv.weight/.bias (the wrapped Conv), periodic padding (punetg.py:221-232)."""
    if getattr(cfg, "convolution_type", "default") == "circular":
        return _circular_conv(x, sd[name + ".conv.weight"], sd.get(name + ".conv.bias"), cfg.dimension)
    return _conv(x, sd[name + ".weight"], sd.get(name + ".bias"), cfg.dimension)


def _bc(v, x):
    return v.reshape(v.shape + (1,) * (x.ndim - v.ndim))


def fourier(t, W):
    """GaussianFourierProjection.forward (nets/commonlayers.py:175-190)."""
    p = 2 * math.pi * t[..., None] * W
    return torch.cat([torch.sin(p), torch.cos(p)], dim=-1)


def group_rms_norm(x, G, w, b, eps=1e-5):
    """GroupRMSNorm.forward (commonlayers.py:362-384)."""
    B, C = x.shape[:2]
    xv = x.view(B, G, C // G, *x.shape[2:])
    dims = tuple(range(2, xv.dim()))
    xv = xv / torch.sqrt(xv.pow(2).mean(dim=dims, keepdim=True) + eps)
    x = xv.view(B, C, *x.shape[2:])
    if w is not None:
        shp = (1, C) + (1,) * (x.dim() - 2)
        x = x * w.view(shp) + b.view(shp)
    return x


def _norm(kind, x, G, w, b):
    if kind == "GroupLN":
        return F.group_norm(x, G, w, b, 1e-5)
    if kind == "GroupRMS":
        return group_rms_norm(x, G, w, b)
    raise NotImplementedError(kind)


def mha_self_attention(x_nc, sd, prefix, residual):
    """NDimensionalAttention.forward (nets/attention.py:54-102) with nn.MultiheadAttention(C, 1 head).

    x_nc: [B, C, *S] -> tokens [B, L, C]; packed in_proj [3C, C]; softmax(QK^T/sqrt(C)) V; out_proj.
    """
    B, C = x_nc.shape[:2]
    S = x_nc.shape[2:]
    tok = x_nc.reshape(B, C, -1).transpose(1, 2)
    qkv = F.linear(tok, sd[prefix + "mhattn.in_proj_weight"], sd[prefix + "mhattn.in_proj_bias"])
    q, k, v = qkv.chunk(3, dim=-1)
    a = torch.softmax(q @ k.transpose(1, 2) / math.sqrt(C), dim=-1) @ v
    o = F.linear(a, sd[prefix + "mhattn.out_proj.weight"], sd[prefix + "mhattn.out_proj.bias"])
    o = o.transpose(1, 2).reshape(B, C, *S)
    return x_nc + o if residual else o


# ----------------------------------------------------------------------------- PUNetG
def _time_block(te, sd, p):
    """ResnetTimeBlock.forward (commonlayers.py:516-550): Linear-SiLU-Linear-SiLU-Linear."""
    h = F.silu(F.linear(te, sd[p + "net.0.weight"], sd[p + "net.0.bias"]))
    h = F.silu(F.linear(h, sd[p + "net.2.weight"], sd[p + "net.2.bias"]))
    return F.linear(h, sd[p + "net.4.weight"], sd[p + "net.4.bias"])


def resnet_block_c(x, te, sd, p, cfg, dropout_masks=None):
    """ResnetBlockC.forward (commonlayers.py:809-836) as built by PUNetG.resnet_fn (punetg.py:238-261):
    C_in == C_out, identity residual, num_groups == num_channels for both norms."""
    dim = cfg.dimension
    C = x.shape[1]
    aff = getattr(cfg, "affine_norm", True)
    g = lambda n: (sd[p + n + ".weight"], sd[p + n + ".bias"]) if aff else (None, None)  # noqa: E731
    y = _pconv(F.silu(_norm(cfg.first_resblock_norm, x, C, *g("gnorm1"))), sd, p + "conv1", cfg)
    y = y + _bc(_time_block(te, sd, p + "timeblock."), y)
    h = F.silu(_norm(cfg.second_resblock_norm, y, C, *g("gnorm2")))
    if dropout_masks is not None:          # training-mode Dropout (commonlayers.py:829-831) with an explicit, pre-scaled mask
        h = h * dropout_masks[p[:-1]].to(h)
    y = _pconv(h, sd, p + "conv2", cfg)
    return y + x


def porosity_embedder(sd, prefix, porosity):
    """PorosityEmbedder.forward (nets/embedder.py:217-221): y['porosity'] [B, 1] -> squeeze(-1) -> Fourier features ->
    Linear-SiLU-Linear-SiLU-Linear -> [B, dembed]."""
    e = fourier(porosity.squeeze(-1), sd[prefix + "gaussian_proj.W"])
    return _time_block(e, sd, prefix)


def punetg_cond_forward(sd, cfg, x, t, y_channels, ye=None):
    """PUNetGCond.forward (nets/punetg.py:716-735): channel conditioning concatenated to x (a batch-1 condition is
    repeated over the batch), the rest of y reaches PUNetG.forward as the embedding vector."""
    y_cat = torch.cat(list(y_channels), dim=1)
    if y_cat.shape[0] == 1 and x.shape[0] > 1:
        y_cat = torch.cat([y_cat] * x.shape[0], dim=0)
    return punetg_forward(sd, cfg, torch.cat([x, y_cat], dim=1), t, ye)


def punetg_forward(sd, cfg, x, t, ye=None, dropout_masks=None):
    """PUNetG.forward (nets/punetg.py:389-416), default layer types, eval mode.  ye: the conditional embedding vector
    [B or 1, M] added to the time embedding (punetg.py:400-410; cond_drop / cond_dropout are identities in eval)."""
    dim = cfg.dimension
    pool = F.max_pool2d if dim == 2 else F.max_pool3d
    nlev = len(cfg.channel_expansion)
    if not getattr(cfg, "bias", True):
        ones = torch.ones_like(x[:, :1])
        x = torch.cat([x, ones], dim=1)
    x = _pconv(x, sd, "convin", cfg)
    te = fourier(t, sd["time_projection.W"])
    if ye is not None:
        te = te + ye
    skips = []
    for l in range(nlev):                                   # encode, punetg.py:356-365
        for r in range(cfg.number_resnet_downward_block):
            x = resnet_block_c(x, te, sd, f"downward_blocks.{l}.{r}.", cfg, dropout_masks)
        skips.append(x)
        x = _pconv(pool(x, cfg.transition_scale_factor), sd, f"downsamplers.{l}.conv", cfg)
    for r in range(cfg.number_resnet_before_attn_block):    # bottom, punetg.py:378-387
        x = resnet_block_c(x, te, sd, f"before_block.{r}.", cfg, dropout_masks)
    xa = x
    for r in range(cfg.number_resnet_attn_block):
        xa = resnet_block_c(xa, te, sd, f"attn_resnet_block.{r}.", cfg, dropout_masks)
        if r < cfg.number_resnet_attn_block - 1:
            xa = mha_self_attention(xa, sd, f"attn_block.{r}.", getattr(cfg, "attn_residual", False))
    x = x + xa
    for r in range(cfg.number_resnet_after_attn_block):
        x = resnet_block_c(x, te, sd, f"after_block.{r}.", cfg, dropout_masks)
    for l in range(nlev):                                   # decode, punetg.py:367-376
        x = F.interpolate(x, scale_factor=cfg.transition_scale_factor, mode="nearest")
        x = _pconv(x, sd, f"upsamplers.{l}.conv", cfg)
        x = x + skips.pop()
        for r in range(cfg.number_resnet_upward_block):
            x = resnet_block_c(x, te, sd, f"upward_blocks.{l}.{r}.", cfg, dropout_masks)
    return _pconv(x, sd, "convout", cfg)


# ----------------------------------------------------------------------------- ADM
def adm_block(x, te, sd, p, cfg, sample=None, attn=False, dropout_masks=None):
    """ADMBaseBlock.forward (nets/adm.py:292-343): norm1-SiLU-[pool|up]-conv1-norm2, FiLM
    x*te1+te2 (no '1+'), SiLU-conv2, + conv1x1([pool|up](x)), optional attention."""
    dim = cfg.dimension
    G = cfg.num_groups
    fac = cfg.transition_scale_factor

    def resample(z):
        if sample == "down":
            return (F.avg_pool2d if dim == 2 else F.avg_pool3d)(z, fac)
        if sample == "up":
            return F.interpolate(z, scale_factor=fac, mode="nearest")
        return z

    y = F.silu(_norm(cfg.first_resblock_norm, x, G, sd[p + "norm1.weight"], sd[p + "norm1.bias"]))
    y = _pconv(resample(y), sd, p + "conv1", cfg)            # conv_fn: Conv2d | CircularConv2d (adm.py:427-441)
    y = _norm(cfg.second_resblock_norm, y, G, sd[p + "norm2.weight"], sd[p + "norm2.bias"])
    e = F.linear(te, sd[p + "embed_linear.weight"], sd[p + "embed_linear.bias"])
    te1, te2 = torch.chunk(e, 2, dim=-1)
    y = y * _bc(te1, y) + _bc(te2, y)
    y = F.silu(y)
    if dropout_masks is not None:          # ADMBaseBlock.second_block (adm.py:323-329) with an explicit, pre-scaled mask
        y = y * dropout_masks[p[:-1]].to(y)
    y = _pconv(y, sd, p + "conv2", cfg)
    y = y + _pconv(resample(x), sd, p + "convresidual", cfg)
    if attn:
        y = mha_self_attention(y, sd, p + "attn.", cfg.attn_residual)
    return y


def adm_forward(sd, cfg, x, t, ye=None, dropout_masks=None):
    """ADM.forward (nets/adm.py:199-216), decoder_type 1.  ye: conditional embedding vector [B or 1, output_embed_dim],
    added before the final SiLU of the time embedding (adm.py:1047-1053; zeros / None when y is None)."""
    te = fourier(t, sd["time_embedding.projection.W"])
    te = F.linear(F.silu(F.linear(te, sd["time_embedding.mlp.0.weight"], sd["time_embedding.mlp.0.bias"])),
                  sd["time_embedding.mlp.2.weight"], sd["time_embedding.mlp.2.bias"])
    if ye is not None:
        te = te + ye
    te = F.silu(te)                                          # adm.py:1047-1053
    x = F.conv2d(x, sd["input_layer.weight"], sd["input_layer.bias"], padding=cfg.kernel_size // 2)
    nlev = len(cfg.channel_expansion)
    skips = [x]
    for l in range(nlev):                                    # ADMEncoder, adm.py:560-640
        nb = cfg.number_resnet_downward_block
        for r in range(nb):
            x = adm_block(x, te, sd, f"encoder.layers.{l}.input_blocks.{r}.", cfg,
                          sample="down" if r == nb - 1 else None, dropout_masks=dropout_masks)
        skips.append(x)
    flags = cfg.middle_block_attn_config                     # adm.py:73-77
    for r, has_attn in enumerate(flags):
        x = adm_block(x, te, sd, f"middle_block.middle_blocks.{r}.", cfg, attn=has_attn, dropout_masks=dropout_masks)
    assert getattr(cfg, "decoder_type", 1) == 1
    for l in range(nlev):                                    # ADMDecoderLayer1, adm.py:642-740
        h = skips.pop()
        x = torch.cat([x, h], dim=1) if cfg.skip_integration_type == "concat" else x + h
        nb = cfg.number_resnet_upward_block
        for r in range(nb):
            x = adm_block(x, te, sd, f"decoder.layers.{l}.input_blocks.{r}.", cfg,
                          sample="up" if r == nb - 1 else None, dropout_masks=dropout_masks)
    return F.conv2d(x, sd["output_layer.weight"], sd["output_layer.bias"], padding=cfg.kernel_size // 2)


# ----------------------------------------------------------------------------- MLP
def mlp_uncond_forward(sd, x, t, act="relu"):
    """MLPUncond.forward (nets/mlp.py:38-58): cat[x, t] -> (Linear, act)* -> Linear."""
    fn = {"relu": F.relu, "silu": F.silu}[act]
    h = torch.cat([x, t[..., None]], dim=-1)
    idx = sorted({int(k.split(".")[1]) for k in sd if k.startswith("net.")})
    for j, i in enumerate(idx):
        h = F.linear(h, sd[f"net.{i}.weight"], sd[f"net.{i}.bias"])
        if j != len(idx) - 1:
            h = fn(h)
    return h


# ----------------------------------------------------------------------------- helpers
def synth_state_dict(manifest, seed: int, dtype=torch.float32):
    """Deterministic synthetic weights for a (key, shape) manifest; CPU generator => identical
    on every machine with this torch build.  Used by the golden generator (loaded INTO the
    live reference) and by the tests (loaded into the product modules)."""
    out = {}
    for i, (key, shape) in enumerate(manifest):
        g = torch.Generator().manual_seed(seed * 100003 + i)
        shape = tuple(shape)
        leaf = key.rsplit(".", 1)[-1]
        if leaf == "W":
            v = torch.randn(shape, generator=g) * 30.0
        elif len(shape) >= 2:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            v = torch.randn(shape, generator=g) / math.sqrt(fan_in)
        elif leaf == "weight":                              # 1-D weight == norm scale
            v = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:
            v = 0.1 * torch.randn(shape, generator=g)
        out[key] = v.to(dtype)
    return out
