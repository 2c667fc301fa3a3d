"""Pixel-wise ensemble statistics on the device (mean, std, CRPS) -- BASELINE.json's parity criterion for sampled
ensembles, and SURVEY.md section 8(f) rank 3 (score the ensemble before any device-to-host copy).

    stats = ensemble_statistics(samples, truth)      # samples [M, 1, H, W] (a sampler's return value), truth [1, H, W] / [H, W]
    stats["mean"], stats["std"], stats["crps"]       # each [H, W] fp32 on the samples' device

std uses Bessel's correction (torch.std default); CRPS is the ensemble estimator E|X - y| - 1/2 E|X - X'|."""
from __future__ import annotations

from typing import Dict, Optional

import torch

from ._lib import call
from .engine import _stream


def ensemble_statistics(samples: torch.Tensor, truth: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    if not samples.is_cuda:
        raise RuntimeError("ensemble_statistics: samples must live on a CUDA device (no CPU path)")
    m = samples.shape[0]
    field_shape = tuple(samples.shape[-2:])
    x = samples.reshape(m, -1).contiguous().float()
    pixels = x.shape[1]
    if pixels != field_shape[0] * field_shape[1]:
        raise ValueError(f"samples must be [M, 1, H, W] or [M, H, W], got {tuple(samples.shape)}")
    y = None
    if truth is not None:
        y = truth.to(device=x.device, dtype=torch.float32).reshape(-1).contiguous()
        if y.numel() != pixels:
            raise ValueError(f"truth has {y.numel()} pixels, the ensemble {pixels}")
    mean, std = torch.empty(pixels, device=x.device), torch.empty(pixels, device=x.device)
    crps = torch.empty(pixels, device=x.device) if y is not None else None
    with torch.cuda.device(x.device):
        call("sbgm_ensemble_stats", x.data_ptr(), None if y is None else y.data_ptr(), m, pixels, mean.data_ptr(), std.data_ptr(),
             None if crps is None else crps.data_ptr(), _stream())
    out = {"mean": mean.view(field_shape), "std": std.view(field_shape)}
    if crps is not None:
        out["crps"] = crps.view(field_shape)
    return out
