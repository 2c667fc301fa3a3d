"""CPU oracle: functional fp32 restatement of the reference score-UNet (TEST INFRASTRUCTURE).

Follows `/root/reference/sbgm/score_unet.py` (line numbers cited per function) and torchvision's
`BasicBlock` (`torchvision/models/resnet.py:59-103`).  Written as pure functions over a
state-dict so it shares no module structure with either the reference or the CUDA package.
Pinned against the reference's own outputs by `tests/test_oracle_golden.py`.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from .synth import FMAP_CHANNELS, NetConfig, decoder_plan

SIGMA = 25.0  # score_unet.py:932
SD = Dict[str, torch.Tensor]


def marginal_prob_std(t: torch.Tensor, sigma: float = SIGMA, eps: float = 1e-5) -> torch.Tensor:
    """score_unet.py:881-897: sqrt((sigma^(2t) - 1) / (2 ln sigma)), clamped below at eps."""
    t = t.to(torch.float32)
    log_s = torch.log(torch.tensor(sigma, dtype=torch.float32, device=t.device))
    var = (torch.exp((2.0 * t) * log_s) - 1.0) / (2.0 * log_s)
    return torch.clamp(torch.sqrt(var), min=eps)


def diffusion_coeff(t: torch.Tensor, sigma: float = SIGMA) -> torch.Tensor:
    """score_unet.py:916-930: g(t) = sigma^t."""
    return sigma ** t


def fourier_embed(W: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
    """score_unet.py:41-45: cat(sin(2 pi t W), cos(2 pi t W))."""
    proj = t.reshape(-1).to(W.dtype)[:, None] * W[None, :] * (2.0 * torch.pi)
    return torch.cat([proj.sin(), proj.cos()], dim=-1)


def _time_proj(sd: SD, prefix: str, emb: torch.Tensor) -> torch.Tensor:
    """SiLU -> Linear (score_unet.py:373-383, 501-504)."""
    return F.linear(F.silu(emb), sd[f"{prefix}.1.weight"], sd[f"{prefix}.1.bias"])


def _bn(sd: SD, prefix: str, x: torch.Tensor, train: bool) -> torch.Tensor:
    return F.batch_norm(x, None if train else sd[f"{prefix}.running_mean"],
                        None if train else sd[f"{prefix}.running_var"],
                        sd[f"{prefix}.weight"], sd[f"{prefix}.bias"], training=train, eps=1e-5)


def _basic_block(sd: SD, p: str, x: torch.Tensor, stride: int, bn_train: bool) -> torch.Tensor:
    """torchvision resnet.py:89-103."""
    out = F.conv2d(x, sd[f"{p}.conv1.weight"], None, stride=stride, padding=1)
    out = F.relu(_bn(sd, f"{p}.bn1", out, bn_train))
    out = F.conv2d(out, sd[f"{p}.conv2.weight"], None, stride=1, padding=1)
    out = _bn(sd, f"{p}.bn2", out, bn_train)
    if f"{p}.downsample.0.weight" in sd:
        idn = F.conv2d(x, sd[f"{p}.downsample.0.weight"], None, stride=stride)
        idn = _bn(sd, f"{p}.downsample.1", idn, bn_train)
    else:
        idn = x
    return F.relu(out + idn)


def attention_block(sd: SD, p: str, x: torch.Tensor, n_heads: int) -> torch.Tensor:
    """score_unet.py:136-148: h = x + MHA(LN1(x)); y = h + FF(LN2(h)) over flattened pixels."""
    n, c, hh, ww = x.shape
    tok = x.reshape(n, c, hh * ww).permute(0, 2, 1)
    h = F.layer_norm(tok, (c,), sd[f"{p}.ln1.weight"], sd[f"{p}.ln1.bias"], eps=1e-5)
    qkv = F.linear(h, sd[f"{p}.mha.in_proj_weight"], sd[f"{p}.mha.in_proj_bias"])
    q, k, v = qkv.chunk(3, dim=-1)
    d = c // n_heads

    def heads(z: torch.Tensor) -> torch.Tensor:
        return z.reshape(n, -1, n_heads, d).permute(0, 2, 1, 3)

    att = torch.softmax(heads(q) @ heads(k).transpose(-1, -2) / math.sqrt(d), dim=-1) @ heads(v)
    att = att.permute(0, 2, 1, 3).reshape(n, -1, c)
    att = F.linear(att, sd[f"{p}.mha.out_proj.weight"], sd[f"{p}.mha.out_proj.bias"])
    h = tok + att
    g = F.layer_norm(h, (c,), sd[f"{p}.ln2.weight"], sd[f"{p}.ln2.bias"], eps=1e-5)
    g = F.linear(g, sd[f"{p}.ff.0.weight"], sd[f"{p}.ff.0.bias"])
    g = F.linear(F.gelu(g), sd[f"{p}.ff.2.weight"], sd[f"{p}.ff.2.bias"])
    return (h + g).permute(0, 2, 1).reshape(n, c, hh, ww)


def encoder_forward(sd: SD, cfg: NetConfig, x, t, y=None, cond_img=None, lsm_cond=None, topo_cond=None,
                    bn_train: bool = False) -> List[torch.Tensor]:
    """score_unet.py:247-364.  Channel order: x || lsm || topo || cond_img (:273-291)."""
    for name, c in (("lsm_cond", lsm_cond), ("topo_cond", topo_cond)):
        if c is not None:
            if c.shape[0] != x.shape[0]:
                raise ValueError(f"Batch mismatch: x= {x.shape[0]}, {name}={c.shape[0]}.")
            x = torch.cat([x, c], dim=1)
    if cond_img is not None:
        x = torch.cat([x, cond_img], dim=1)
    emb = fourier_embed(sd["encoder.sinusoidal_embedding.W"], t.float())
    if y is not None:
        emb = emb + sd["encoder.label_emb.weight"][y.long()]

    def tadd(f: torch.Tensor, i: int) -> torch.Tensor:
        return f + _time_proj(sd, f"encoder.time_projection_layers.{i}", emb)[:, :, None, None]

    fmaps = []
    f1 = tadd(F.conv2d(x, sd["encoder.conv1.weight"], None, stride=2, padding=3), 0)
    fmaps.append(f1)
    h = F.conv2d(f1, sd["encoder.conv2.weight"], None, stride=2, padding=3)
    h = F.relu(_bn(sd, "encoder.bn1", h, bn_train))
    for li, nblk in enumerate(cfg.block_layers, start=1):
        for b in range(nblk):
            h = _basic_block(sd, f"encoder.layer{li}.{b}", h, 2 if (b == 0 and li > 1) else 1, bn_train)
        h = tadd(h, li)
        if li >= len(FMAP_CHANNELS) - 2:
            h = attention_block(sd, f"encoder.attention_layers.{li}", h, cfg.n_heads)
        fmaps.append(h)
    return fmaps


def _act(name: str, x: torch.Tensor) -> torch.Tensor:
    return {"relu": F.relu, "silu": F.silu, "gelu": F.gelu, "identity": lambda v: v}[name](x)


def _norm(sd: SD, cfg: NetConfig, key: str, x: torch.Tensor) -> torch.Tensor:
    """score_unet.py:483-487."""
    if cfg.norm == "group":
        c = x.shape[1]
        return F.group_norm(x, max(1, min(cfg.gn_groups, c)), sd[f"{key}.weight"], sd[f"{key}.bias"], eps=1e-5)
    return F.instance_norm(x, eps=1e-5)


def decoder_block(sd: SD, cfg: NetConfig, p: str, fmap, skip, t, *, attn: bool, final: bool) -> torch.Tensor:
    """score_unet.py:559-627; `final` = norms / activation replaced by Identity (:726-730)."""
    if cfg.use_resize_conv:
        x = F.interpolate(fmap, scale_factor=2, mode="bilinear", align_corners=False)
        x = F.conv2d(x, sd[f"{p}.conv_up.weight"], sd[f"{p}.conv_up.bias"], padding=1)
    else:
        x = F.conv_transpose2d(fmap, sd[f"{p}.transpose.weight"], sd[f"{p}.transpose.bias"], stride=2)
    if not final:
        x = _norm(sd, cfg, f"{p}.norm1", x)
    x = F.conv2d(x, sd[f"{p}.conv.weight"], sd[f"{p}.conv.bias"], padding=1)
    if not final:
        x = _norm(sd, cfg, f"{p}.norm2", x)
    if skip is not None:
        if skip.shape != x.shape:
            raise AssertionError(f"prev_fmap shape {tuple(skip.shape)} must match output shape {tuple(x.shape)}")
        x = x + skip
    if t is not None:
        emb = fourier_embed(sd[f"{p}.sinusoidal_embedding.W"], t)
        x = x + _time_proj(sd, f"{p}.time_projection_layer", emb)[:, :, None, None]
    x = _act("identity" if final else cfg.activation, x)
    if attn:
        x = attention_block(sd, f"{p}.attention", x, cfg.n_heads)
    return x


def decoder_forward(sd: SD, cfg: NetConfig, fmaps: List[torch.Tensor], t: torch.Tensor) -> torch.Tensor:
    """score_unet.py:733-758."""
    plan = decoder_plan(cfg)
    assert len(fmaps) == len(plan) + 1, f"Decoder expected {len(plan) + 1} feature maps, got {len(fmaps)}"
    rev = list(reversed(fmaps))
    out = rev[0]
    for i, (_, _, attn) in enumerate(plan):
        out = decoder_block(sd, cfg, f"decoder.residual_layers.{i}", out, rev[i + 1], t, attn=attn, final=False)
    return decoder_block(sd, cfg, "decoder.final_layer", out, None, None, attn=False, final=True)


def unet_forward(sd: SD, cfg: NetConfig, x, t, y=None, cond_img=None, lsm_cond=None, topo_cond=None,
                 bn_train: bool = False) -> torch.Tensor:
    """Decoder(Encoder(.)) before the division by the marginal std."""
    t = t.float()
    fmaps = encoder_forward(sd, cfg, x, t, y, cond_img, lsm_cond, topo_cond, bn_train=bn_train)
    return decoder_forward(sd, cfg, fmaps, t)


def score_forward(sd: SD, cfg: NetConfig, x, t, y=None, cond_img=None, lsm_cond=None, topo_cond=None,
                  bn_train: bool = False) -> torch.Tensor:
    """score_unet.py:829-879: network output divided by marginal_prob_std(t)."""
    out = unet_forward(sd, cfg, x, t, y, cond_img, lsm_cond, topo_cond, bn_train=bn_train)
    return out / marginal_prob_std(t.float()).view(-1, 1, 1, 1)


def dsm_loss(sd: SD, cfg: NetConfig, x, random_t, z, y=None, cond_img=None, lsm_cond=None, topo_cond=None,
             sdf_cond=None, bn_train: bool = True) -> torch.Tensor:
    """score_unet.py:936-985 with the two random draws (`random_t`, `z`) injected by the caller."""
    for name, arr in (("cond_img", cond_img), ("lsm_cond", lsm_cond), ("topo_cond", topo_cond), ("y", y)):
        if arr is not None and arr.shape[0] != x.shape[0]:
            raise ValueError(f"Batch size mismatch: x={x.shape[0]}, {name}={arr.shape[0]}")
    std = marginal_prob_std(random_t)
    xt = x + std[:, None, None, None] * z
    score = score_forward(sd, cfg, xt, random_t, y, cond_img, lsm_cond, topo_cond, bn_train=bn_train)
    w = torch.sigmoid(sdf_cond) * 0.5 + 0.5 if sdf_cond is not None else torch.ones_like(x)
    return torch.mean(torch.sum(w * (score * std[:, None, None, None] + z) ** 2, dim=(1, 2, 3)))
