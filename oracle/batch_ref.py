"""CPU oracle (TEST INFRASTRUCTURE): the host->device boundary of the reference restated --
`extract_samples` (/root/reference/sbgm/utils.py:405-480) on CPU tensors, and the forward transforms of
/root/reference/sbgm/special_transforms.py (Scale :62-100, ZScoreTransform :143-185, PrcpLogTransform :239-343) in numpy
float32 with the reference's operation order.  Pinned against the reference's own function / classes by
tests/golden/batch_golden.npz (tests/golden/make_batch_golden.py)."""
from __future__ import annotations

import numpy as np
import torch

F = np.float32


def extract_samples_ref(samples: dict):
    """The 9-tuple of utils.py:405-480 as float32 CPU tensors (the device copy is the product's job)."""
    hr_keys = [k for k in samples.keys() if k.endswith("_hr") and not k.endswith("_original")]
    if "lsm_hr" in hr_keys:
        hr_keys.remove("lsm_hr")
    if len(hr_keys) == 0:
        raise ValueError("No HR image found in samples dictionary.")
    hr = samples[hr_keys[0]].float()
    classifier = samples.get("classifier", None)
    lr_keys = [k for k in samples.keys() if k.endswith("_lr") and not k.endswith("_original")]
    if len(lr_keys) == 0:
        lr = None
    elif len(lr_keys) == 1:
        lr = samples[lr_keys[0]].float()
    else:
        lr = torch.cat([samples[k].float() for k in sorted(lr_keys)], dim=1)
    opt = lambda k: None if samples.get(k, None) is None else samples[k].float()
    return hr, classifier, lr, opt("lsm_hr"), opt("lsm"), opt("sdf"), opt("topo"), opt("hr_point"), opt("lr_point")


def scale_fwd(x, in_low, in_high, data_min_in=0, data_max_in=1):
    x = np.asarray(x, dtype=F)
    return ((x - F(data_min_in)) * F(in_high - in_low)) / F(data_max_in - data_min_in) + F(in_low)


def zscore_fwd(x, mean, std):
    x = np.asarray(x, dtype=F)
    return (x - F(mean)) / (F(std) + F(1e-8))


def prcp_log_fwd(x, eps=0.01, scale_type="log_zscore", glob_mean_log=None, glob_std_log=None, glob_min_log=None, glob_max_log=None,
                 buffer_frac=0.5):
    x = np.asarray(x, dtype=F)
    if glob_min_log is not None and glob_max_log is not None:        # :262-266: widened by buffer_frac on each side
        r = glob_max_log - glob_min_log
        glob_min_log, glob_max_log = glob_min_log - buffer_frac * r, glob_max_log + buffer_frac * r
    v = np.log(x + F(eps))
    if scale_type == "log_01":
        return (v - F(glob_min_log)) / F(glob_max_log - glob_min_log)
    if scale_type == "log_zscore":
        return (v - F(glob_mean_log)) / F(glob_std_log + 1e-8)
    if scale_type == "log_minus1_1":
        return F(2) * ((v - F(glob_min_log)) / F(glob_max_log - glob_min_log)) - F(1)
    if scale_type == "log":
        return v
    raise ValueError("Invalid scale type. Please choose 'log_01' or 'log_zscore' or 'log'.")


FWD_CASES = {
    # name: (kind, kwargs, input kind)
    "zscore_t2m": ("zscore", dict(mean=8.69, std=6.19), "temp"),
    "scale_01": ("scale", dict(in_low=0, in_high=1, data_min_in=-3.5, data_max_in=41.0), "temp"),
    "scale_m11": ("scale", dict(in_low=-1, in_high=1, data_min_in=0.0, data_max_in=160.0), "prcp"),
    "log_zscore": ("log", dict(eps=0.01, scale_type="log_zscore", glob_mean_log=-1.2, glob_std_log=2.0), "prcp"),
    "log_01": ("log", dict(eps=0.01, scale_type="log_01", glob_min_log=-4.6, glob_max_log=5.1, buffer_frac=0.5), "prcp"),
    "log_minus1_1": ("log", dict(eps=1e-3, scale_type="log_minus1_1", glob_min_log=-6.9, glob_max_log=5.1, buffer_frac=0.25), "prcp"),
    "log_plain": ("log", dict(eps=0.01, scale_type="log"), "prcp"),
}


def apply_fwd_case(name, x):
    kind, kw, _ = FWD_CASES[name]
    if kind == "zscore":
        return zscore_fwd(x, **kw)
    if kind == "scale":
        return scale_fwd(x, **kw)
    return prcp_log_fwd(x, **kw)


def fwd_case_input(kind: str, n: int = 4096, seed: int = 0) -> np.ndarray:
    rng = np.random.default_rng(seed)
    if kind == "temp":
        x = (rng.normal(size=n) * 9.0 + 8.0).astype(F)
        x[:4] = np.array([0.0, -3.5, 41.0, 8.69], dtype=F)
    else:                                 # precipitation: non-negative, many zeros, a heavy tail
        x = np.where(rng.random(n) < 0.4, 0.0, rng.gamma(0.6, 6.0, size=n)).astype(F)
        x[:4] = np.array([0.0, 0.01, 160.0, 1.0], dtype=F)
    return x


def sample_dict(seed: int = 0, batch: int = 3, size: int = 16, two_lr: bool = True) -> dict:
    """A dataset-shaped sample dict with mixed dtypes (float64 fields as netCDF hands them out, integer / bool masks)."""
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    d = {
        "prcp_hr": r(batch, 1, size, size).double(),
        "prcp_hr_original": r(batch, 1, size, size),
        "classifier": torch.randint(0, 5, (batch, 1), generator=g),
        "temp_lr": r(batch, 1, size, size),
        "temp_lr_original": r(batch, 1, size, size),
        "lsm_hr": (r(batch, 1, size, size) > 0),
        "lsm": (r(batch, 1, size, size) > 0).to(torch.uint8),
        "sdf": r(batch, 1, size, size).half(),
        "topo": r(batch, 1, size, size).bfloat16(),
        "hr_point": torch.randint(0, 500, (batch, 4), generator=g),
        "lr_point": torch.randint(0, 500, (batch, 4), generator=g).int(),
    }
    if two_lr:
        d["prcp_lr"] = r(batch, 2, size, size).double()
    return d
