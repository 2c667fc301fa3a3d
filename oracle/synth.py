"""Synthetic weights and inputs of the ORACLE and the parity tests -- TEST INFRASTRUCTURE (only tests/, __graft_entry__.smoke()
and bench.py's baseline legs may import anything under oracle/).

Self-contained on purpose: the oracle's network plan (state-dict key set, shapes, generated values) must not come from
product code.  `sbgm_danra_b200/synth.py` is the product-side twin used by bench.py and tools/ (which may not import oracle/);
tests/test_flop_model.py::test_oracle_and_package_generators_agree pins the two to each other, and the committed golden vectors
(tests/golden/, generated from the REAL reference with these weights) pin both to the reference's checkpoint ABI.

The reference's checkpoint ABI is its state-dict key set (keys probed from
`sbgm/score_unet.py` Encoder :151-229, DecoderBlock :409-512, Decoder :662-730).  A 19 M
parameter state dict is far too big to commit, so weights are *generated*: every tensor is a
deterministic function of (seed, key), independent of constructor order, and can be loaded
into the reference modules, the oracle and the CUDA modules alike.

Inputs follow SURVEY.md section 8(d): z-scored LR fields ~ N(0,1), land/sea value||mask, topography
value||mask in [-1,1], season label in 1..4, SDF in [0,1].
"""
from __future__ import annotations

import hashlib
import math
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch

FMAP_CHANNELS = (64, 64, 128, 256, 512)  # sbgm/score_unet.py:198


@dataclass
class NetConfig:
    """Constructor knobs of Encoder/Decoder that change the state-dict (score_unet.py:158-229, 669-730)."""
    in_channels: int = 1                 # Encoder(input_channels=...): conditioning channels (HR channel is added inside)
    time_embedding: int = 256
    block_layers: Tuple[int, ...] = (2, 2, 2, 2)
    n_heads: int = 4
    num_classes: Optional[int] = None
    last_fmap_channels: int = 512
    first_fmap_channels: int = 64
    out_channels: int = 1
    use_resize_conv: bool = True
    norm: str = "group"                  # "group" | "instance"
    gn_groups: int = 8
    activation: str = "silu"             # "relu" | "silu" | "gelu"

    @property
    def cin_total(self) -> int:
        return self.in_channels + 1


def _attn_schema(prefix: str, c: int, out: "OrderedDict[str, tuple]") -> None:
    out[f"{prefix}.mha.in_proj_weight"] = (3 * c, c)
    out[f"{prefix}.mha.in_proj_bias"] = (3 * c,)
    out[f"{prefix}.mha.out_proj.weight"] = (c, c)
    out[f"{prefix}.mha.out_proj.bias"] = (c,)
    for ln in ("ln1", "ln2"):
        out[f"{prefix}.{ln}.weight"] = (c,)
        out[f"{prefix}.{ln}.bias"] = (c,)
    for i in (0, 2):
        out[f"{prefix}.ff.{i}.weight"] = (c, c)
        out[f"{prefix}.ff.{i}.bias"] = (c,)


def _bn_schema(prefix: str, c: int, out: "OrderedDict[str, tuple]") -> None:
    out[f"{prefix}.weight"] = (c,)
    out[f"{prefix}.bias"] = (c,)
    out[f"{prefix}.running_mean"] = (c,)
    out[f"{prefix}.running_var"] = (c,)
    out[f"{prefix}.num_batches_tracked"] = ()


def decoder_plan(cfg: NetConfig) -> List[Tuple[int, int, bool]]:
    """(in_ch, out_ch, attention?) for the 4 residual decoder blocks (score_unet.py:761-789)."""
    plan = []
    prev_out = None
    for i in range(4):
        in_ch = cfg.last_fmap_channels if i == 0 else prev_out
        out_ch = in_ch // 2 if i != 3 else cfg.first_fmap_channels
        plan.append((in_ch, out_ch, i < 2))
        prev_out = out_ch
    return plan


def param_schema(cfg: NetConfig) -> "OrderedDict[str, tuple]":
    """Every state-dict key of ScoreNet(Encoder, Decoder) with its shape."""
    s: "OrderedDict[str, tuple]" = OrderedDict()
    te = cfg.time_embedding
    s["encoder.conv1.weight"] = (64, cfg.cin_total, 8, 8)
    _bn_schema("encoder.bn1", 64, s)
    inplanes = 64
    for li, (planes, nblk) in enumerate(zip((64, 128, 256, 512), cfg.block_layers), start=1):
        for b in range(nblk):
            stride = 2 if (b == 0 and li > 1) else 1
            p = f"encoder.layer{li}.{b}"
            s[f"{p}.conv1.weight"] = (planes, inplanes, 3, 3)
            _bn_schema(f"{p}.bn1", planes, s)
            s[f"{p}.conv2.weight"] = (planes, planes, 3, 3)
            _bn_schema(f"{p}.bn2", planes, s)
            if stride != 1 or inplanes != planes:
                s[f"{p}.downsample.0.weight"] = (planes, inplanes, 1, 1)
                _bn_schema(f"{p}.downsample.1", planes, s)
            inplanes = planes
    s["encoder.sinusoidal_embedding.W"] = (te // 2,)
    for i, ch in enumerate(FMAP_CHANNELS):
        s[f"encoder.time_projection_layers.{i}.1.weight"] = (ch, te)
        s[f"encoder.time_projection_layers.{i}.1.bias"] = (ch,)
    for i, ch in enumerate(FMAP_CHANNELS):
        if i >= len(FMAP_CHANNELS) - 2:
            _attn_schema(f"encoder.attention_layers.{i}", ch, s)
    s["encoder.conv2.weight"] = (64, 64, 8, 8)
    if cfg.num_classes is not None:
        s["encoder.label_emb.weight"] = (cfg.num_classes + 1, te)

    def block(prefix: str, cin: int, cout: int, attn: bool, norms: bool) -> None:
        if cfg.use_resize_conv:
            s[f"{prefix}.conv_up.weight"] = (cin, cin, 3, 3)
            s[f"{prefix}.conv_up.bias"] = (cin,)
        else:
            s[f"{prefix}.transpose.weight"] = (cin, cin, 2, 2)
            s[f"{prefix}.transpose.bias"] = (cin,)
        if norms and cfg.norm == "group":
            s[f"{prefix}.norm1.weight"] = (cin,)
            s[f"{prefix}.norm1.bias"] = (cin,)
        s[f"{prefix}.conv.weight"] = (cout, cin, 3, 3)
        s[f"{prefix}.conv.bias"] = (cout,)
        if norms and cfg.norm == "group":
            s[f"{prefix}.norm2.weight"] = (cout,)
            s[f"{prefix}.norm2.bias"] = (cout,)
        s[f"{prefix}.sinusoidal_embedding.W"] = (te // 2,)
        s[f"{prefix}.time_projection_layer.1.weight"] = (cout, te)
        s[f"{prefix}.time_projection_layer.1.bias"] = (cout,)
        if attn:
            _attn_schema(f"{prefix}.attention", cout, s)

    plan = decoder_plan(cfg)
    for i, (cin, cout, attn) in enumerate(plan):
        block(f"decoder.residual_layers.{i}", cin, cout, attn, True)
    block("decoder.final_layer", plan[-1][0], cfg.out_channels, False, False)
    return s


def _key_generator(seed: int, key: str) -> torch.Generator:
    h = hashlib.sha256(f"{seed}:{key}".encode()).digest()
    g = torch.Generator(device="cpu")
    g.manual_seed(int.from_bytes(h[:8], "little") & 0x7FFFFFFFFFFFFFFF)
    return g


def synth_state_dict(cfg: NetConfig, seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """Deterministic, well-conditioned random weights for every key of `param_schema(cfg)`.

    Scales mimic a trained network rather than a fresh init so that BatchNorm folding,
    GroupNorm affine terms, biases and the label embedding are all exercised.
    """
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for key, shape in param_schema(cfg).items():
        g = _key_generator(seed, key)
        leaf = key.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            t = torch.tensor(100, dtype=torch.int64)
        elif leaf == "running_mean":
            t = torch.randn(shape, generator=g) * 0.1
        elif leaf == "running_var":
            t = torch.rand(shape, generator=g) + 0.5
        elif leaf == "W":
            t = torch.randn(shape, generator=g) * 30.0          # score_unet.py:37
        elif key.endswith("label_emb.weight"):
            t = torch.randn(shape, generator=g) * 0.5
            t[0].zero_()                                        # null class row, score_unet.py:224-226
        elif len(shape) == 1:
            is_scale = leaf == "weight"                          # norm / BN gamma
            if is_scale:
                t = 1.0 + 0.2 * (torch.rand(shape, generator=g) - 0.5)
            else:
                t = torch.randn(shape, generator=g) * 0.05
        elif leaf in ("in_proj_bias",):
            t = torch.randn(shape, generator=g) * 0.05
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            t = torch.randn(shape, generator=g) * math.sqrt(1.0 / fan_in)
        sd[key] = t.to(torch.float32) if t.dtype != torch.int64 else t
    return sd


@dataclass
class Batch:
    x: torch.Tensor
    t: torch.Tensor
    y: Optional[torch.Tensor] = None
    cond_img: Optional[torch.Tensor] = None
    lsm_cond: Optional[torch.Tensor] = None
    topo_cond: Optional[torch.Tensor] = None
    sdf_cond: Optional[torch.Tensor] = None
    extras: Dict[str, torch.Tensor] = field(default_factory=dict)

    def model_args(self):
        return (self.x, self.t, self.y, self.cond_img, self.lsm_cond, self.topo_cond)


def synth_batch(batch: int, size: int, n_lr: int = 1, geo: bool = False, seasons: bool = False,
                seed: int = 1234, shared_cond: bool = False) -> Batch:
    """Synthetic ERA5/DANRA-shaped inputs.

    `shared_cond=True` broadcasts one conditioning sample to all members (ensemble generation,
    cf. `generate_repeated`, sbgm/evaluate_sbgm/generation.py:269-285).
    """
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    nb = 1 if shared_cond else batch

    def rep(v: torch.Tensor) -> torch.Tensor:
        return v.expand(batch, *v.shape[1:]).contiguous() if shared_cond else v

    x = torch.randn(batch, 1, size, size, generator=g)
    t = torch.rand(batch, generator=g) * (1.0 - 1e-3) + 1e-3
    cond = rep(torch.randn(nb, n_lr, size, size, generator=g)) if n_lr > 0 else None
    lsm = topo = y = None
    if geo:
        lsm_v = (torch.rand(nb, 1, size, size, generator=g) < 0.5).float()
        topo_v = torch.rand(nb, 1, size, size, generator=g) * 2.0 - 1.0
        ones = torch.ones(nb, 1, size, size)
        lsm = rep(torch.cat([lsm_v, ones], dim=1))
        topo = rep(torch.cat([topo_v, ones], dim=1))
    if seasons:
        y = torch.randint(1, 5, (nb,), generator=g)
        y = y.expand(batch).contiguous() if shared_cond else y
    sdf = rep(torch.rand(nb, 1, size, size, generator=g))
    return Batch(x=x, t=t, y=y, cond_img=cond, lsm_cond=lsm, topo_cond=topo, sdf_cond=sdf)


def config_for(n_lr: int = 1, geo: bool = False, seasons: bool = False, **kw) -> NetConfig:
    """NetConfig whose channel count matches `synth_batch(..., n_lr, geo, seasons)`
    (training_utils.py:588-595: n_lr + 2 * n_geo)."""
    return NetConfig(in_channels=n_lr + (4 if geo else 0), num_classes=4 if seasons else None, **kw)
