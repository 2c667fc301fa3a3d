"""CPU oracle: restatement of the reference VE-SDE samplers with injectable noise
(TEST INFRASTRUCTURE).  Follows `/root/reference/sbgm/score_sampling.py`.

Deliberate deviations from the reference, both documented in SURVEY.md section 0:
  * Euler-Maruyama / ODE initial state uses `img_size` instead of the hard-coded 32
    (score_sampling.py:94, :274) -- the unpatched line crashes for any non-32x32 condition.
  * every Gaussian draw goes through `noise(draw_id, shape)`; with `noise=None` the draws come
    from torch's global generator exactly where the reference draws them.
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np
import torch

from . import philox_ref

ScoreFn = Callable[[torch.Tensor, torch.Tensor], torch.Tensor]
NoiseFn = Optional[Callable[[int, tuple], torch.Tensor]]


def philox_noise(seed: int, first_member: int = 0) -> Callable[[int, tuple], torch.Tensor]:
    """Noise callback reading the Philox stream of `philox_ref` (global element indexing)."""
    def fn(draw: int, shape: tuple) -> torch.Tensor:
        n = int(np.prod(shape))
        per = int(np.prod(shape[1:]))
        return torch.from_numpy(philox_ref.normal(n, seed, draw, first_member * per)).reshape(shape)
    return fn


def _draw(noise: NoiseFn, draw_id: int, like: torch.Tensor) -> torch.Tensor:
    return torch.randn_like(like) if noise is None else noise(draw_id, tuple(like.shape)).to(like)


def guided_score(score_model, x, t, y=None, cond_img=None, lsm_cond=None, topo_cond=None,
                 null_token: int = 0, scale: float = 2.0) -> torch.Tensor:
    """score_sampling.py:10-56: (1+w) s_cond - w s_uncond; null = zero LR image, geo mask channel
    (index 1) zeroed, label -> null token."""
    def strip(v):
        if v is None or v.shape[1] != 2:
            return v
        v = v.clone()
        v[:, 1] = 0.0
        return v
    s_c = score_model(x, t, y, cond_img, lsm_cond, topo_cond)
    s_u = score_model(x, t,
                      torch.full_like(y, null_token) if y is not None else None,
                      torch.zeros_like(cond_img) if cond_img is not None else None,
                      strip(lsm_cond), strip(topo_cond))
    return (1.0 + scale) * s_c - scale * s_u


def euler_maruyama(score: ScoreFn, marginal_prob_std, diffusion_coeff, batch_size: int, num_steps: int,
                   eps: float = 1e-3, img_size: int = 64, noise: NoiseFn = None, device="cpu",
                   trajectory: Optional[list] = None) -> torch.Tensor:
    """score_sampling.py:63-127 (with the img_size fix)."""
    t = torch.ones(batch_size, device=device)
    shape = (batch_size, 1, img_size, img_size)
    z0 = torch.randn(shape, device=device) if noise is None else noise(philox_ref.DRAW_INIT, shape).to(device)
    x = z0 * marginal_prob_std(t)[:, None, None, None]
    ts = torch.linspace(1.0, eps, num_steps, device=device)
    dt = ts[0] - ts[1]
    mean_x = x
    with torch.no_grad():
        for k, tk in enumerate(ts):
            bt = torch.ones(batch_size, device=device) * tk
            g = diffusion_coeff(bt)
            s = score(x, bt)
            mean_x = x + (g ** 2)[:, None, None, None] * s * dt
            x = mean_x + torch.sqrt(dt) * g[:, None, None, None] * _draw(noise, philox_ref.draw_em(k), x)
            if trajectory is not None:
                trajectory.append(mean_x.clone())
    return mean_x


def predictor_corrector(score: ScoreFn, marginal_prob_std, diffusion_coeff, batch_size: int, num_steps: int,
                        snr: float = 0.16, eps: float = 1e-3, img_size: int = 64, noise: NoiseFn = None,
                        device="cpu", trajectory: Optional[list] = None, score_predictor: Optional[ScoreFn] = None) -> torch.Tensor:
    """score_sampling.py:136-230.  `score_predictor`: the score of the predictor half when it differs from the corrector's
    (classifier-free guidance with guidance_scale_max: the corrector clamps the scale :182-186, the predictor does not
    :209-219)."""
    t = torch.ones(batch_size, device=device)
    shape = (batch_size, 1, img_size, img_size)
    z0 = torch.randn(shape, device=device) if noise is None else noise(philox_ref.DRAW_INIT, shape).to(device)
    x = z0 * marginal_prob_std(t)[:, None, None, None]
    ts = np.linspace(1.0, eps, num_steps)
    dt = ts[0] - ts[1]
    x_mean = x
    with torch.no_grad():
        for k, tk in enumerate(ts):
            bt = torch.ones(batch_size, device=device) * tk
            grad = score(x, bt)
            grad_norm = torch.norm(grad.reshape(grad.shape[0], -1), dim=-1).mean()
            noise_norm = np.sqrt(np.prod(x.shape[1:]))
            ls = 2 * (snr * noise_norm / grad_norm) ** 2
            x = x + ls * grad + torch.sqrt(2 * ls) * _draw(noise, philox_ref.draw_pc_corrector(k), x)
            g = diffusion_coeff(bt)
            s = (score_predictor or score)(x, bt)
            x_mean = x + (g ** 2)[:, None, None, None] * s * dt
            x = x_mean + torch.sqrt(g ** 2 * dt)[:, None, None, None] * _draw(noise, philox_ref.draw_pc_predictor(k), x)
            if trajectory is not None:
                trajectory.append(x_mean.clone())
    return x_mean


def ode_rhs(score: ScoreFn, diffusion_coeff, shape, t: float, x_flat: np.ndarray) -> np.ndarray:
    """score_sampling.py:278-293: dx/dt = -1/2 g(t)^2 score(x, t), float64 on the host."""
    x = torch.tensor(x_flat, dtype=torch.float32).reshape(shape)
    tt = torch.tensor(np.ones((shape[0],)) * t, dtype=torch.float32)
    with torch.no_grad():
        s = score(x, tt)
    g = diffusion_coeff(torch.tensor(t)).cpu().numpy()
    return -0.5 * (g ** 2) * s.cpu().numpy().reshape(-1).astype(np.float64)


def ode_solve(score: ScoreFn, marginal_prob_std, diffusion_coeff, batch_size: int, atol=1e-5, rtol=1e-5,
              eps: float = 1e-3, img_size: int = 64, z: Optional[torch.Tensor] = None, noise: NoiseFn = None):
    """score_sampling.py:239-300 (RK45 through scipy), img_size fix applied."""
    from scipy import integrate
    shape = (batch_size, 1, img_size, img_size)
    if z is None:
        z0 = torch.randn(shape) if noise is None else noise(philox_ref.DRAW_INIT, shape)
        init = z0 * marginal_prob_std(torch.ones(batch_size))[:, None, None, None]
    else:
        init = z
    res = integrate.solve_ivp(lambda t, x: ode_rhs(score, diffusion_coeff, tuple(init.shape), t, x), (1.0, eps),
                              init.reshape(-1).cpu().numpy(), rtol=rtol, atol=atol, method="RK45")
    return torch.tensor(res.y[:, -1]).reshape(init.shape), res.nfev
