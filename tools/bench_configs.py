"""Sampler throughput of the BASELINE.json configurations other than the headline one (not bench lines;
recorded under profiles/).  Usage: python tools/bench_configs.py [C1 C3 C5 ...]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sbgm_danra_b200.synth import config_for, synth_batch, synth_state_dict
from sbgm_danra_b200 import score_sampling as ss
from sbgm_danra_b200._smoke import build_model
from sbgm_danra_b200.score_unet import diffusion_coeff_fn, marginal_prob_std_fn

CONFIGS = {
    # name: (sampler, size, batch, steps, precision, net kwargs, flops/sample/forward, nfe)
    "C1": ("em", 64, 4, 100, "bf16x3", dict(n_lr=1), 1.279e9, 1),
    "C2": ("em", 128, 64, 500, "bf16x3", dict(n_lr=1), 5.146e9, 1),
    "C2-fp16x2": ("em", 128, 64, 500, "fp16x2", dict(n_lr=1), 5.146e9, 1),
    "C2-bf16": ("em", 128, 64, 500, "bf16", dict(n_lr=1), 5.146e9, 1),
    "C3": ("pc", 128, 64, 500, "bf16", dict(n_lr=2, geo=True, seasons=True), 5.313e9, 2),
    "C3-bf16x3": ("pc", 128, 64, 500, "bf16x3", dict(n_lr=2, geo=True, seasons=True), 5.313e9, 2),
    "C3-fp16x2": ("pc", 128, 64, 500, "fp16x2", dict(n_lr=2, geo=True, seasons=True), 5.313e9, 2),
    "C5": ("em", 256, 4, 1000, "bf16x3", dict(n_lr=1), 21.09e9, 1),
    "C5-b32": ("em", 256, 32, 1000, "bf16x3", dict(n_lr=1), 21.09e9, 1),
    "C5-b32-fp16x2": ("em", 256, 32, 1000, "fp16x2", dict(n_lr=1), 21.09e9, 1),
}


def main():
    names = sys.argv[1:] or list(CONFIGS)
    dev = torch.device("cuda:0")
    for name in names:
        kind, size, batch, steps, prec, ck, flop, nfe = CONFIGS[name]
        cfg = config_for(**ck)
        net = build_model(cfg, synth_state_dict(cfg), prec, dev)
        b = synth_batch(batch=batch, size=size, shared_cond=True, **ck)
        fn = ss.Euler_Maruyama_sampler if kind == "em" else ss.pc_sampler
        kw = dict(batch_size=batch, num_steps=steps, device=dev, img_size=size, cond_img=b.cond_img.to(dev),
                  y=None if b.y is None else b.y.to(dev), lsm_cond=None if b.lsm_cond is None else b.lsm_cond.to(dev),
                  topo_cond=None if b.topo_cond is None else b.topo_cond.to(dev))
        ss.clear_sampler_cache()
        for _ in range(2):
            out = fn(net, marginal_prob_std_fn, diffusion_coeff_fn, **kw)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 2
        for _ in range(reps):
            out = fn(net, marginal_prob_std_fn, diffusion_coeff_fn, **kw)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        finite = bool(torch.isfinite(out).all())
        print(f"{name:10s} {kind} {size}x{size} B={batch} steps={steps} {prec:7s}: {dt * 1e3:8.1f} ms/call  "
              f"{batch / dt:8.2f} fields/s  {dt / steps / nfe * 1e3:6.3f} ms/NFE  "
              f"{batch * flop * steps * nfe / dt / 1e12:6.1f} TFLOP/s (algorithmic)  finite={finite}", flush=True)
        del net


if __name__ == "__main__":
    main()
