"""Transforms of `sbgm/special_transforms.py` on the device: the inverse ones: the step after the sampler,
so that a sampled ensemble reaches physical units (and can be scored, `ensemble.py`) before any device-to-host copy.

Same class names, constructor arguments, validation and arithmetic as the reference (`ZScoreBackTransform` :187-237,
`ScaleBackTransform` :103-138, `PrcpLogBackTransform` :360-462, `build_back_transforms` :463-520); `__call__` takes a CUDA
tensor and runs ONE fused kernel (affine -> clamp -> exp).  Scalar statistics only (what the reference's configuration passes);
tensor-valued mean / std raise."""
from __future__ import annotations

import logging

import torch

from ._lib import call
from .engine import _stream

logger = logging.getLogger(__name__)
_INF = float("inf")


def _apply(sample, scale: float, shift: float, lo: float = -_INF, hi: float = _INF, exp: bool = False, pre: float = 0.0) -> torch.Tensor:
    if not isinstance(sample, torch.Tensor):
        raise RuntimeError("back-transforms take CUDA tensors (the sampler's output); there is no CPU path")
    if not sample.is_cuda:
        raise RuntimeError(f"back-transforms run on CUDA tensors only (no CPU fallback); got a tensor on {sample.device}")
    x = sample.contiguous().float()
    out = torch.empty_like(x)
    clamp = not (lo == -_INF and hi == _INF)
    f32 = lambda v: max(min(float(v), 3.4028234663852886e38), -3.4028234663852886e38)
    with torch.cuda.device(x.device):
        call("sbgm_back_transform", x.data_ptr(), out.data_ptr(), x.numel(), float(pre), float(scale), float(shift), f32(lo), f32(hi),
             int(clamp), int(exp), _stream())
    return out


def _scalar(v, what: str) -> float:
    if isinstance(v, torch.Tensor):
        if v.numel() != 1:
            raise NotImplementedError(f"{what}: tensor-valued statistics are not on the CUDA path (scalars only)")
        return float(v)
    return float(v)


class ZScoreBackTransform(object):
    """x * (std + 1e-8) + mean (special_transforms.py:187-237)."""

    def __init__(self, mean, std):
        self.mean, self.std = mean, std

    def __call__(self, sample):
        # the reference forms (std + eps) in float32 before the product
        scale = float(torch.tensor(_scalar(self.std, "std"), dtype=torch.float32) + 1e-8)
        return _apply(sample, scale, _scalar(self.mean, "mean"))


class ScaleBackTransform(object):
    """((x - in_low) * (data_max - data_min)) / (in_high - in_low) + data_min (special_transforms.py:103-138)."""

    def __init__(self, in_low=0, in_high=1, data_min_in=0, data_max_in=1):
        self.in_low, self.in_high, self.data_min_in, self.data_max_in = in_low, in_high, data_min_in, data_max_in

    def __call__(self, sample):
        old, new = self.in_high - self.in_low, self.data_max_in - self.data_min_in
        return _apply(sample, new / old, self.data_min_in, pre=-self.in_low)


class PrcpLogBackTransform(object):
    """exp(clamp(affine(x))) with the affine map chosen by `scale_type` (special_transforms.py:360-462)."""

    def __init__(self, scale_type="log_zscore", glob_mean_log=None, glob_std_log=None, glob_min_log=None, glob_max_log=None,
                 buffer_frac=0.5, clamp_log_min=None, clamp_log_max=None):
        self.scale_type = scale_type
        self.glob_mean_log, self.glob_std_log = glob_mean_log, glob_std_log
        self.glob_min_log, self.glob_max_log = glob_min_log, glob_max_log
        self.buffer_frac, self.clamp_log_min, self.clamp_log_max = buffer_frac, clamp_log_min, clamp_log_max
        self.hi = _INF if clamp_log_max is None else float(clamp_log_max)
        self.lo = -_INF if clamp_log_min is None else float(clamp_log_min)
        if self.glob_min_log is not None and self.glob_max_log is not None:
            log_range = self.glob_max_log - self.glob_min_log          # widened by buffer_frac, as the reference does
            self.glob_min_log = self.glob_min_log - (self.buffer_frac / 2) * log_range
            self.glob_max_log = self.glob_max_log + (self.buffer_frac / 2) * log_range
        if scale_type == "log_zscore":
            if self.glob_mean_log is None or self.glob_std_log is None:
                raise ValueError("Global mean and standard deviation not provided. Using local statistics is not recommended.")
        elif scale_type in ("log_01", "log_minus1_1"):
            if self.glob_min_log is None or self.glob_max_log is None:
                raise ValueError("Min and max log values not provided. Using global statistics is recommended.")
        elif scale_type != "log":
            raise ValueError("Invalid scale type. Please choose from ['log_01', 'log_zscore', 'log_minus1_1', 'log'].")

    def __call__(self, sample):
        if self.scale_type == "log_01":
            return _apply(sample, self.glob_max_log - self.glob_min_log, self.glob_min_log, self.lo, self.hi, exp=True)
        if self.scale_type == "log_zscore":
            return _apply(sample, _scalar(self.glob_std_log, "glob_std_log") + 1e-8, _scalar(self.glob_mean_log, "glob_mean_log"),
                          self.lo, self.hi, exp=True)
        if self.scale_type == "log_minus1_1":
            return _apply(sample, 0.5 * (self.glob_max_log - self.glob_min_log), self.glob_min_log, self.lo, self.hi, exp=True, pre=1.0)
        return _apply(sample, 1.0, 0.0, self.lo, self.hi, exp=True)


# ---- forward transforms (special_transforms.py:62-343), applied by the dataset on the CPU in the reference ------------------
# Each class describes itself as the parameters of the batch-assembly kernel (csrc/batch.cu):
#     v = log ? log(x + eps) : x;   y = (((v - sub) * mul) / div) * post_mul + post_add        (float32, this operation order)
# so that `batch.BatchAssembler` can apply it while the batch is written on the device; called directly on a CUDA tensor it
# runs that kernel as a single job.
class _ForwardTransform(object):
    def kernel_params(self):                       # (log, eps, sub, mul, div, post_mul, post_add)
        raise NotImplementedError

    def __call__(self, sample):
        from .batch import apply_transform
        return apply_transform(sample, self)


class Scale(_ForwardTransform):
    """(((x - data_min_in) * (in_high - in_low)) / (data_max_in - data_min_in)) + in_low (special_transforms.py:62-100)."""

    def __init__(self, in_low, in_high, data_min_in=0, data_max_in=1):
        self.in_low, self.in_high, self.data_min_in, self.data_max_in = in_low, in_high, data_min_in, data_max_in

    def kernel_params(self):
        return (False, 0.0, float(self.data_min_in), float(self.in_high - self.in_low), float(self.data_max_in - self.data_min_in),
                1.0, float(self.in_low))


class ZScoreTransform(_ForwardTransform):
    """(x - mean) / (std + 1e-8) with the statistics held in float32 (special_transforms.py:143-185)."""

    def __init__(self, mean, std):
        self.mean, self.std = mean, std

    def kernel_params(self):
        div = float(torch.tensor(_scalar(self.std, "std"), dtype=torch.float32) + 1e-8)
        return (False, 0.0, float(torch.tensor(_scalar(self.mean, "mean"), dtype=torch.float32)), 1.0, div, 1.0, 0.0)


class PrcpLogTransform(_ForwardTransform):
    """log(x + eps), then the scaling chosen by `scale_type` (special_transforms.py:239-343; the forward class widens the
    log range by buffer_frac on EACH side, the inverse by buffer_frac / 2 -- both as the reference has them)."""

    def __init__(self, eps=0.01, scale_type="log_zscore", glob_mean_log=None, glob_std_log=None, glob_min_log=None, glob_max_log=None,
                 buffer_frac=0.5):
        self.eps, self.scale_type = eps, scale_type
        self.glob_mean_log, self.glob_std_log = glob_mean_log, glob_std_log
        self.glob_min_log, self.glob_max_log, self.buffer_frac = glob_min_log, glob_max_log, buffer_frac
        if self.glob_min_log is not None and self.glob_max_log is not None:
            log_range = self.glob_max_log - self.glob_min_log
            self.glob_min_log = self.glob_min_log - self.buffer_frac * log_range
            self.glob_max_log = self.glob_max_log + self.buffer_frac * log_range
        if scale_type == "log_zscore":
            if self.glob_mean_log is None or self.glob_std_log is None:
                raise ValueError("Global mean and standard deviation not provided. Using local statistics is not recommended.")
        elif scale_type in ("log_01", "log_minus1_1"):
            if self.glob_min_log is None or self.glob_max_log is None:
                raise ValueError("Min and max log values not provided. Using global statistics is recommended.")
        elif scale_type != "log":
            raise ValueError("Invalid scale type. Please choose '01' or 'zscore'.")

    def kernel_params(self):
        eps = float(self.eps)
        if self.scale_type == "log_zscore":
            return (True, eps, float(_scalar(self.glob_mean_log, "glob_mean_log")), 1.0,
                    float(_scalar(self.glob_std_log, "glob_std_log") + 1e-8), 1.0, 0.0)
        if self.scale_type in ("log_01", "log_minus1_1"):
            denom = float(self.glob_max_log - self.glob_min_log)
            if denom == 0:
                raise ValueError("The log-range of data is zero. Cannot scale to [0, 1]. Please check the data.")
            if self.scale_type == "log_01":
                return (True, eps, float(self.glob_min_log), 1.0, denom, 1.0, 0.0)
            return (True, eps, float(self.glob_min_log), 1.0, denom, 2.0, -1.0)
        return (True, eps, 0.0, 1.0, 1.0, 1.0, 0.0)


def build_back_transforms(hr_var, hr_scaling_method, hr_scaling_params, lr_vars, lr_scaling_methods, lr_scaling_params):
    """Plot-key -> inverse transform, as special_transforms.py:463-520."""
    def make(method, prm):
        if method in {"log", "log_01", "log_minus1_1", "log_zscore"}:
            return PrcpLogBackTransform(scale_type=method, glob_mean_log=prm["glob_mean_log"], glob_std_log=prm["glob_std_log"],
                                        glob_min_log=prm["glob_min_log"], glob_max_log=prm["glob_max_log"],
                                        buffer_frac=prm["buffer_frac"], clamp_log_min=prm.get("clamp_log_min", None),
                                        clamp_log_max=prm.get("clamp_log_max", None))
        if method == "zscore":
            return ZScoreBackTransform(prm["glob_mean"], prm["glob_std"])
        if method == "01":
            return ScaleBackTransform(0, 1, prm["glob_min"], prm["glob_max"])
        return None

    bt = {}
    inv = make(hr_scaling_method, hr_scaling_params)
    if inv is None:
        raise ValueError(f"Unknown HR scaling method: {hr_scaling_method}")
    bt[f"{hr_var}_hr"] = inv
    bt["generated"] = inv
    for cond, mth, prm in zip(lr_vars, lr_scaling_methods, lr_scaling_params):
        t = make(mth, prm)
        if t is None:
            raise ValueError(f"Unknown LR scaling method: {mth}")
        bt[f"{cond}_lr"] = t
    return bt
