"""CPU oracle: the inverse transforms of `/root/reference/sbgm/special_transforms.py` restated in numpy float64
(TEST INFRASTRUCTURE).  ZScoreBackTransform :187-237, ScaleBackTransform :103-138, PrcpLogBackTransform :360-462.
Pinned against the reference classes themselves by tests/golden/transforms_golden.npz (make_transforms_golden.py)."""
from __future__ import annotations

import numpy as np


def zscore_back(x, mean, std):
    return np.asarray(x, dtype=np.float64) * (std + 1e-8) + mean


def scale_back(x, in_low, in_high, data_min, data_max):
    return ((np.asarray(x, dtype=np.float64) - in_low) * (data_max - data_min)) / (in_high - in_low) + data_min


def prcp_log_back(x, scale_type, glob_mean_log=None, glob_std_log=None, glob_min_log=None, glob_max_log=None, buffer_frac=0.5,
                  clamp_log_min=None, clamp_log_max=None):
    x = np.asarray(x, dtype=np.float64)
    lo = -np.inf if clamp_log_min is None else float(clamp_log_min)
    hi = np.inf if clamp_log_max is None else float(clamp_log_max)
    if glob_min_log is not None and glob_max_log is not None:        # :392-398: range widened by buffer_frac
        r = glob_max_log - glob_min_log
        glob_min_log, glob_max_log = glob_min_log - (buffer_frac / 2) * r, glob_max_log + (buffer_frac / 2) * r
    if scale_type == "log_01":
        v = x * (glob_max_log - glob_min_log) + glob_min_log
    elif scale_type == "log_zscore":
        v = x * (glob_std_log + 1e-8) + glob_mean_log
    elif scale_type == "log_minus1_1":
        v = 0.5 * (x + 1) * (glob_max_log - glob_min_log) + glob_min_log
    elif scale_type == "log":
        v = x
    else:
        raise ValueError("Invalid scale type. Please choose from ['log_01', 'log_zscore', 'log_minus1_1', 'log'].")
    return np.exp(np.clip(v, lo, hi))


CASES = {
    # name: (kind, kwargs)
    "zscore_t2m": ("zscore", dict(mean=8.69, std=6.19)),
    "scale_01": ("scale", dict(in_low=0, in_high=1, data_min=-3.5, data_max=41.0)),
    "scale_m11": ("scale", dict(in_low=-1, in_high=1, data_min=0.0, data_max=160.0)),
    "log_zscore_clamped": ("log", dict(scale_type="log_zscore", glob_mean_log=-3.0, glob_std_log=3.6, clamp_log_min=-14.0, clamp_log_max=5.0)),
    "log_zscore": ("log", dict(scale_type="log_zscore", glob_mean_log=-1.2, glob_std_log=2.0)),
    "log_01": ("log", dict(scale_type="log_01", glob_min_log=-18.4, glob_max_log=5.1, buffer_frac=0.5)),
    "log_minus1_1": ("log", dict(scale_type="log_minus1_1", glob_min_log=-18.4, glob_max_log=5.1, buffer_frac=0.25, clamp_log_max=6.0)),
    "log_plain": ("log", dict(scale_type="log", clamp_log_min=-20.0, clamp_log_max=4.0)),
}


def apply_case(name, x):
    kind, kw = CASES[name]
    if kind == "zscore":
        return zscore_back(x, **kw)
    if kind == "scale":
        return scale_back(x, **kw)
    return prcp_log_back(x, **kw)


def case_input(n: int = 4096, seed: int = 0) -> np.ndarray:
    rng = np.random.default_rng(seed)
    x = rng.normal(size=n).astype(np.float32) * 1.5
    x[:8] = np.array([0.0, 1.0, -1.0, 0.5, 4.0, -4.0, 2.5, -2.5], dtype=np.float32)
    return x
