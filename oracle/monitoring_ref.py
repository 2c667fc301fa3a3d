"""CPU restatement of the reference's extreme-value sentinel -- TEST INFRASTRUCTURE ONLY (imported by tests/, never by the
product).  Follows sbgm/utils.py:1642-1671 (`report_precip_extremes`) and the clamp of sbgm/training.py:739-748.
Pinned against the reference's own function by tests/golden/make_monitoring_golden.py -> tests/golden/monitoring_golden.json."""
from __future__ import annotations

import torch


def report_precip_extremes(x_bt: torch.Tensor, name: str, cap_mm_day: float = 500.0, logger=print):
    flat = x_bt.flatten(1)                                            # utils.py:1647
    p999 = torch.quantile(flat, 0.999, dim=1)                         # :1648
    mx = torch.max(flat, dim=1).values                                # :1649
    n_ex, vals_ex, n_b0, vals_b0 = 0, [], 0, []
    for i, (p, m) in enumerate(zip(p999.tolist(), mx.tolist())):      # :1654-1663
        if m > max(5.0 * p, cap_mm_day):
            logger(f"{name} sample {i} has extreme precipitation: max={m:.1f} mm/day > max(5xp99.9={p:.1f} mm/day)")
            n_ex += 1
            vals_ex.append(m)
        if m < 0:
            logger(f"{name} sample {i} has negative precipitation: max={m:.1f} mm/day < 0")
            n_b0 += 1
            vals_b0.append(m)
    if n_b0 > 0 and n_ex > 0:                                         # :1664-1671
        return {"has_extreme": True, "n_extreme": n_ex, "extreme_values": vals_ex,
                "has_below_zero": True, "n_below_zero": n_b0, "below_zero_values": vals_b0}
    if n_ex > 0:
        return {"has_extreme": True, "n_extreme": n_ex, "extreme_values": vals_ex}
    if n_b0 > 0:
        return {"has_below_zero": True, "n_below_zero": n_b0, "below_zero_values": vals_b0}
    return {"has_extreme": False}


def quantile_and_max(x_bt: torch.Tensor, q: float = 0.999):
    flat = x_bt.flatten(1)
    return torch.quantile(flat, q, dim=1), flat.max(dim=1).values, flat.min(dim=1).values


def clamp_generated(gen_bt: torch.Tensor, clamp_max: float) -> torch.Tensor:
    return torch.clamp(gen_bt, min=0.0, max=clamp_max)                # training.py:745


def cases():
    """Seeded physical-unit fields (mm/day-like): benign, one with an isolated spike (extreme), one all-negative, both."""
    g = torch.Generator().manual_seed(404)
    out = {}
    base = torch.exp(torch.randn(4, 1, 64, 64, generator=g) * 1.2)            # log-normal "precipitation", max ~ 100
    out["benign"] = base.clone()
    spike = base.clone()
    spike[1, 0, 10, 20] = 2500.0
    spike[3, 0, 0, 0] = 900.0
    out["spikes"] = spike
    neg = base.clone()
    neg[2] = -neg[2] - 0.5
    out["negative_sample"] = neg
    both = spike.clone()
    both[0] = -both[0] - 1.0
    out["both"] = both
    out["small_3x5"] = torch.randn(3, 1, 3, 5, generator=g) * 10.0          # 15 values: the quantile interpolates the top two
    return out
