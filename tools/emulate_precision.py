#!/usr/bin/env python
"""CPU emulation of candidate tensor-core operand formats on the oracle forward (no GPU needed).

Question for the next round ("Next" item 1): can the fp32-class mode drop from three bf16 products per
algorithmic product (bf16x3) to two, or one, and stay inside the 1e-3 score gate?  The oracle's convolutions / Linear
layers are re-run with their operands rounded as the tensor core would see them (fp32 accumulation either way) and the
score is compared with the unrounded fp32 oracle, on the shapes of the golden cases.

    python tools/emulate_precision.py [--size 64] [--batch 2]
    python tools/emulate_precision.py --sampler          # ensemble mean / std / CRPS of a 40-step EM run per format

Calibration: the `bf16` and `bf16x3` rows must reproduce what the kernels measure on the GPU (6.6e-3 and 1.1e-5 in
README.md); the other rows are predictions.  The time projections (2-D Linear inputs) and the final 64->1 convolution stay
fp32, as in the engine.  TEST / ANALYSIS INFRASTRUCTURE: imports the oracle, nothing here is on the product path.
"""
from __future__ import annotations

import argparse
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.nn.functional as TF  # noqa: E402


def _round_mantissa(x: torch.Tensor, bits: int) -> torch.Tensor:
    """Round-to-nearest-even to `bits` explicit mantissa bits (tf32: 10), exponent range of fp32."""
    i = x.contiguous().view(torch.int32)
    drop = 23 - bits
    bias = ((i >> drop) & 1) + ((1 << (drop - 1)) - 1)
    return ((i + bias) & ~((1 << drop) - 1)).view(torch.float32)


def _split(x: torch.Tensor, dt: torch.dtype, planes: int) -> torch.Tensor:
    """x as the sum of `planes` values of dtype `dt` (what a hi|lo operand pair carries)."""
    acc = torch.zeros_like(x)
    for _ in range(planes):
        acc = acc + (x - acc).to(dt).float()
    return acc


MODES = {
    # name: (activation rounding, weight rounding, tensor-core products per algorithmic product)
    "bf16   (x bf16, w bf16)": (lambda x: _split(x, torch.bfloat16, 1), lambda w: _split(w, torch.bfloat16, 1), 1),
    "bf16x3 (x hi|lo bf16, w hi|lo bf16, lo*lo dropped)": ("x3", None, 3),
    "tf32   (x, w 10-bit mantissa)": (lambda x: _round_mantissa(x, 10), lambda w: _round_mantissa(w, 10), "1 at half rate"),
    "fp16   (x fp16, w fp16)": (lambda x: _split(x, torch.float16, 1), lambda w: _split(w, torch.float16, 1), 1),
    "fp16 x, w hi|lo fp16": (lambda x: _split(x, torch.float16, 1), lambda w: _split(w, torch.float16, 2), 2),
    "x hi|lo fp16, w fp16": (lambda x: _split(x, torch.float16, 2), lambda w: _split(w, torch.float16, 1), 2),
}


def emulated_functional(mode):
    qx, qw, _ = MODES[mode]
    ns = types.SimpleNamespace(**{k: getattr(TF, k) for k in dir(TF) if not k.startswith("__")})

    def products(x, w, op):
        if qx == "x3":        # x_hi w_hi + x_hi w_lo + x_lo w_hi, exactly the three products the kernels issue
            xh, wh = x.to(torch.bfloat16).float(), w.to(torch.bfloat16).float()
            xl, wl = (x - xh).to(torch.bfloat16).float(), (w - wh).to(torch.bfloat16).float()
            return op(xh, wh) + op(xh, wl) + op(xl, wh)
        return op(qx(x), qw(w))

    def conv2d(x, w, b=None, **kw):
        if w.shape[0] == 1:                                   # final 64->1 convolution: fp32 in the engine
            return TF.conv2d(x, w, b, **kw)
        out = products(x, w, lambda a, c: TF.conv2d(a, c, None, **kw))
        return out if b is None else out + b.view(1, -1, 1, 1)

    def linear(x, w, b=None):
        if x.dim() == 2:                                      # time projections: fp32 in the engine
            return TF.linear(x, w, b)
        out = products(x, w, lambda a, c: TF.linear(a, c))
        return out if b is None else out + b

    def conv_transpose2d(x, w, b=None, **kw):
        out = products(x, w, lambda a, c: TF.conv_transpose2d(a, c, None, **kw))
        return out if b is None else out + b.view(1, -1, 1, 1)

    ns.conv2d, ns.linear, ns.conv_transpose2d = conv2d, linear, conv_transpose2d
    return ns


def main():
    from oracle import score_ref
    from sbgm_danra_b200.synth import config_for, synth_batch, synth_state_dict
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=64)
    ap.add_argument("--batch", type=int, default=2)
    args = ap.parse_args()
    cases = {"cin2": dict(n_lr=1), "cin7+seasons": dict(n_lr=2, geo=True, seasons=True)}
    print(f"score rel-L2 vs the fp32 oracle, {args.size}x{args.size}, batch {args.batch} (gate for the fp32-class mode: 1e-3)")
    for cname, ck in cases.items():
        cfg = config_for(**ck)
        sd = synth_state_dict(cfg)
        b = synth_batch(batch=args.batch, size=args.size, **ck)
        with torch.no_grad():
            ref = score_ref.score_forward(sd, cfg, *b.model_args())
            for mode, (_, _, products) in MODES.items():
                saved = score_ref.F
                score_ref.F = emulated_functional(mode)
                try:
                    out = score_ref.score_forward(sd, cfg, *b.model_args())
                finally:
                    score_ref.F = saved
                err = float((out - ref).norm() / ref.norm())
                print(f"  {cname:14s} {mode:52s} products {products!s:15s} rel-L2 {err:.2e}")


def sampler_statistics(members: int = 16, size: int = 32, steps: int = 40, seed: int = 99):
    """The second parity criterion (BASELINE.json: sampled-ensemble pixel-wise mean / std and CRPS within 1 %) for each
    emulated format: Euler-Maruyama on the same Philox noise as tests/test_gpu_model.py's ensemble test, statistics against
    the fp32 oracle's ensemble."""
    import numpy as np
    from oracle import samplers_ref, score_ref
    from oracle.ensemble_ref import ensemble_statistics
    from sbgm_danra_b200.synth import config_for, synth_batch, synth_state_dict
    cfg = config_for(n_lr=1)
    sd = synth_state_dict(cfg)
    b = synth_batch(batch=members, size=size, n_lr=1, shared_cond=True)
    truth = synth_batch(batch=1, size=size, n_lr=1, seed=77).x[0, 0].numpy()

    def run():
        score = lambda x, t: score_ref.score_forward(sd, cfg, x, t, None, b.cond_img)
        with torch.no_grad():
            out = samplers_ref.euler_maruyama(score, score_ref.marginal_prob_std, score_ref.diffusion_coeff, members, steps,
                                              img_size=size, noise=samplers_ref.philox_noise(seed))
        return ensemble_statistics(out[:, 0].numpy(), truth)

    want = run()
    print(f"EM ensemble ({members} members, {size}x{size}, {steps} steps, same noise): field rel-L2 of mean / std / CRPS vs the "
          f"fp32 oracle (gate 1e-2)")
    for mode in MODES:
        saved = score_ref.F
        score_ref.F = emulated_functional(mode)
        try:
            got = run()
        finally:
            score_ref.F = saved
        rel = {k: np.linalg.norm(got[k] - want[k]) / np.linalg.norm(want[k]) for k in ("mean", "std", "crps")}
        print(f"  {mode:52s} mean {rel['mean']:.2e}  std {rel['std']:.2e}  crps {rel['crps']:.2e}")


if __name__ == "__main__":
    if "--sampler" in sys.argv:
        sys.argv.remove("--sampler")
        sampler_statistics()
    else:
        main()
