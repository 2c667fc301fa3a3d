#!/usr/bin/env python
"""Stock PyTorch eager (cuDNN / cuBLAS / ATen) on the SAME B200, C2 shape -- the GPU "kernel to beat" of BASELINE.md.

The reference has no GPU code of its own: on a GPU it is torch eager over `sbgm/score_unet.py`.  `/root/reference` does not
exist on the GPU box, so the timed code is the oracle's torch restatement of that forward (`oracle/score_ref.py`, pinned to
the reference by tests/test_oracle_golden.py) moved to cuda:0.  This is a BASELINE measurement (like bench.py's cpu_baseline):
none of this repo's kernels run here, and nothing here is on the product path.

    python tools/bench_torch_eager.py [--members 64] [--em-steps 20]

Prints one JSON line per mode: fp32 with TF32 off (torch's default, the reference's setting), fp32 with TF32 on, and bf16
autocast; forward time (CUDA events, median of 7 after 3 warm-ups), UNet forward TFLOP/s, and fields/s of a short eager
Euler-Maruyama run extrapolated linearly to 500 steps.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SIZE, N_LR, EM_STEPS, FWD_FLOP = 128, 1, 500, 5.146e9


def time_mode(torch, ctx, fwd, x, t, samplers_ref, score_ref, args, dev):
    with torch.no_grad(), ctx:
        for _ in range(3):
            fwd(x, t)
        times = []
        for _ in range(7):
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fwd(x, t)
            e.record()
            e.synchronize()
            times.append(a.elapsed_time(e))
        ms = statistics.median(times)
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        samplers_ref.euler_maruyama(fwd, score_ref.marginal_prob_std, score_ref.diffusion_coeff, args.members, args.em_steps,
                                    img_size=SIZE, device=dev)
        e.record()
        e.synchronize()
        em_ms = a.elapsed_time(e)
    return ms, em_ms


def main():
    import torch
    from oracle import samplers_ref, score_ref
    from sbgm_danra_b200.synth import config_for, synth_batch, synth_state_dict
    ap = argparse.ArgumentParser()
    ap.add_argument("--members", type=int, default=64)
    ap.add_argument("--em-steps", type=int, default=20)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    cfg = config_for(n_lr=N_LR)
    sd = {k: v.to(dev) for k, v in synth_state_dict(cfg, 0).items()}
    b = synth_batch(batch=args.members, size=SIZE, n_lr=N_LR, shared_cond=True)
    x, t, cond = b.x.to(dev), b.t.to(dev), b.cond_img.to(dev)
    torch.backends.cudnn.benchmark = True              # sbgm/training_main.py:111

    def fwd(xx, tt):
        return score_ref.score_forward(sd, cfg, xx, tt, None, cond)

    for mode in ("fp32", "tf32", "bf16-autocast"):
        torch.backends.cuda.matmul.allow_tf32 = mode == "tf32"
        torch.backends.cudnn.allow_tf32 = mode == "tf32"
        ctx = torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "bf16-autocast"))
        try:
            ms, em_ms = time_mode(torch, ctx, fwd, x, t, samplers_ref, score_ref, args, dev)
        except RuntimeError as exc:                    # e.g. an op of the restatement that autocast does not cover
            print(json.dumps({"impl": "torch-eager-on-B200", "mode": mode, "error": str(exc)[:300]}))
            continue
        print(json.dumps({"impl": "torch-eager-on-B200 (oracle port of the reference forward)", "mode": mode,
                          "members": args.members, "fwd_ms": ms, "unet_fwd_tflops": args.members * FWD_FLOP / (ms * 1e-3) / 1e12,
                          "em_fields_per_s": args.members / (em_ms * 1e-3 / args.em_steps * EM_STEPS),
                          "em_sample": f"{args.em_steps} eager EM steps ({em_ms:.0f} ms), extrapolated linearly to {EM_STEPS}",
                          "torch": torch.__version__, "cudnn": torch.backends.cudnn.version()}))


if __name__ == "__main__":
    main()
