"""The extreme-value sentinel of the generation / training path on the device.

Reference: `report_precip_extremes` (sbgm/utils.py:1642-1671: per sample the 0.999 quantile and the maximum of the
back-transformed field; "extreme" when max > max(5 p99.9, cap), "negative" when max < 0) and the flow around it in
`generate_and_plot_samples` (sbgm/training.py:700-755: back-transform -> sentinel -> optional clamp to [0, clamp_max_mm]).
The reference moves the samples to the CPU, back-transforms them there and runs torch.quantile (a full sort per sample).
Here the back-transform and the order statistics are ONE kernel per call (`sbgm_back_transform_extremes`: one block per sample,
exact radix select in shared memory); four floats per sample cross to the host, the fields stay on the device.
"""
from __future__ import annotations

import logging
from typing import Callable, Optional

import torch

from ._lib import call
from .engine import _stream

logger = logging.getLogger(__name__)
_INF = float("inf")
_F32_MAX = 3.4028234663852886e38


def _transform_args(bt):
    """(pre, scale, shift, lo, hi, clamp, exp) of a special_transforms back-transform object, or the identity for None."""
    if bt is None:
        return 0.0, 1.0, 0.0, -_INF, _INF, False, False
    from . import special_transforms as st
    if isinstance(bt, st.ZScoreBackTransform):
        scale = float(torch.tensor(st._scalar(bt.std, "std"), dtype=torch.float32) + 1e-8)
        return 0.0, scale, st._scalar(bt.mean, "mean"), -_INF, _INF, False, False
    if isinstance(bt, st.ScaleBackTransform):
        old, new = bt.in_high - bt.in_low, bt.data_max_in - bt.data_min_in
        return -bt.in_low, new / old, bt.data_min_in, -_INF, _INF, False, False
    if isinstance(bt, st.PrcpLogBackTransform):
        clamp = not (bt.lo == -_INF and bt.hi == _INF)
        if bt.scale_type == "log_01":
            return 0.0, bt.glob_max_log - bt.glob_min_log, bt.glob_min_log, bt.lo, bt.hi, clamp, True
        if bt.scale_type == "log_zscore":
            return (0.0, st._scalar(bt.glob_std_log, "glob_std_log") + 1e-8, st._scalar(bt.glob_mean_log, "glob_mean_log"), bt.lo, bt.hi,
                    clamp, True)
        if bt.scale_type == "log_minus1_1":
            return 1.0, 0.5 * (bt.glob_max_log - bt.glob_min_log), bt.glob_min_log, bt.lo, bt.hi, clamp, True
        return 0.0, 1.0, 0.0, bt.lo, bt.hi, clamp, True
    raise TypeError(f"monitoring: {type(bt).__name__} is not a back-transform of sbgm_danra_b200.special_transforms")


def back_transform_with_extremes(x: torch.Tensor, back_transform=None, quantile: float = 0.999):
    """(x in physical units, stats[n][3] = per-sample (quantile, max, min)) from one kernel.  `x`: CUDA tensor [n, ...]."""
    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        raise RuntimeError("monitoring runs on CUDA tensors (the sampler's output); there is no CPU path")
    x = x.contiguous().float()
    n = x.shape[0]
    per = x[0].numel()
    pre, scale, shift, lo, hi, clamp, exp = _transform_args(back_transform)
    f32 = lambda v: max(min(float(v), _F32_MAX), -_F32_MAX)
    y = torch.empty_like(x)
    stats = torch.empty((n, 4), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        call("sbgm_back_transform_extremes", x.data_ptr(), y.data_ptr(), n, per, float(pre), float(scale), float(shift), f32(lo), f32(hi),
             int(clamp), int(exp), float(quantile), stats.data_ptr(), _stream())
    return y, stats[:, :3]


def _verdict(p999, mx, name: str, cap_mm_day: float, log: Callable) -> dict:
    """The decision logic and return shape of report_precip_extremes (sbgm/utils.py:1647-1671), verbatim in behaviour."""
    n_ex, vals_ex, n_b0, vals_b0 = 0, [], 0, []
    for i, (p, m) in enumerate(zip(p999, mx)):
        if m > max(5.0 * p, cap_mm_day):
            log(f"{name} sample {i} has extreme precipitation: max={m:.1f} mm/day > max(5xp99.9={p:.1f} mm/day)")
            n_ex += 1
            vals_ex.append(m)
        if m < 0:
            log(f"{name} sample {i} has negative precipitation: max={m:.1f} mm/day < 0")
            n_b0 += 1
            vals_b0.append(m)
    if n_b0 > 0 and n_ex > 0:
        return {"has_extreme": True, "n_extreme": n_ex, "extreme_values": vals_ex,
                "has_below_zero": True, "n_below_zero": n_b0, "below_zero_values": vals_b0}
    if n_ex > 0:
        return {"has_extreme": True, "n_extreme": n_ex, "extreme_values": vals_ex}
    if n_b0 > 0:
        return {"has_below_zero": True, "n_below_zero": n_b0, "below_zero_values": vals_b0}
    return {"has_extreme": False}


def report_precip_extremes(x_bt: torch.Tensor, name: str, cap_mm_day: float = 500.0, logger=print):
    """Drop-in for sbgm/utils.py:1642-1671 on a CUDA tensor that is already in physical units: same messages, same dict."""
    _, stats = back_transform_with_extremes(x_bt, None)
    host = stats.cpu()
    return _verdict(host[:, 0].tolist(), host[:, 1].tolist(), name, cap_mm_day, logger)


def monitor_generated(samples: torch.Tensor, back_transform=None, threshold_mm: float = 500.0, clamp_in_generation: bool = False,
                      clamp_max_mm: Optional[float] = None, name: str = "generated_hr", log: Callable = logger.warning):
    """The generation-side flow of sbgm/training.py:700-755 without leaving the device: back-transform the sampled fields
    (model space -> mm/day), run the sentinel, and -- if extremes were found and clamping is configured -- clamp to
    [0, clamp_max_mm].  Returns (fields in physical units, the sentinel's dict)."""
    y, stats = back_transform_with_extremes(samples, back_transform)
    host = stats.cpu()
    chk = _verdict(host[:, 0].tolist(), host[:, 1].tolist(), name, float(threshold_mm), log)
    if chk.get("has_extreme", False):
        vals = chk.get("extreme_values", [])
        log(f"[monitor][gen] Extreme precipitation detected in generated samples:")
        log(f"               max={max(vals):.1f} mm/day, count={len(vals)}, threshold={threshold_mm} mm/day")
        if clamp_in_generation:
            cmax = float(threshold_mm if clamp_max_mm is None else clamp_max_mm)
            with torch.cuda.device(y.device):
                call("sbgm_back_transform", y.data_ptr(), y.data_ptr(), y.numel(), 0.0, 1.0, 0.0, 0.0, cmax, 1, 0, _stream())
            log(f"[monitor][gen] Clamped generated samples to max {cmax} mm/day.")
    return y, chk
