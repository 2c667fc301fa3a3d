"""CPU oracle: pixel-wise ensemble statistics (TEST INFRASTRUCTURE).

The reference has no ensemble scoring of its own on the hot path (evaluation lives in
`sbgm/evaluate_sbgm/`, out of scope); BASELINE.json's parity criterion for sampled ensembles is
"pixel-wise mean / std and CRPS within 1%".  Definitions used on both sides:
  mean, std   over the member axis, std with Bessel's correction (torch.std default)
  CRPS        the ensemble estimator  E|X - y| - 1/2 E|X - X'|  with both expectations over the M members
              (Gneiting & Raftery 2007, eq. 21), per pixel
"""
from __future__ import annotations

import numpy as np


def ensemble_statistics(members: np.ndarray, truth: np.ndarray | None = None) -> dict:
    """members [M, ...pixels], truth [...pixels] -> {"mean", "std", "crps"} (crps only with truth), float64."""
    x = np.asarray(members, dtype=np.float64)
    m = x.shape[0]
    out = {"mean": x.mean(0), "std": x.std(0, ddof=1) if m > 1 else np.zeros(x.shape[1:])}
    if truth is not None:
        y = np.asarray(truth, dtype=np.float64)
        term1 = np.abs(x - y[None]).mean(0)
        xs = np.sort(x, axis=0)
        k = np.arange(1, m + 1, dtype=np.float64).reshape((m,) + (1,) * (x.ndim - 1))
        pair = ((2.0 * k - m - 1.0) * xs).sum(0)           # sum_{i<j} (x_(j) - x_(i))
        out["crps"] = term1 - pair / (m * m)
    return out


def crps_bruteforce(members: np.ndarray, truth: np.ndarray) -> np.ndarray:
    x = np.asarray(members, dtype=np.float64)
    y = np.asarray(truth, dtype=np.float64)
    return np.abs(x - y[None]).mean(0) - 0.5 * np.abs(x[:, None] - x[None, :]).mean((0, 1))
