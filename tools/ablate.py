#!/usr/bin/env python
"""In-graph cost of single operators of the sampler step, by ablation: re-runs tools/bench_configs-style timing of the
C2 Euler-Maruyama step with SBGM_B200_SKIP=<op> (engine.py) and prints ms per network evaluation for each variant.
ncu launch lists are cold-cache and serialised; this is the warm, graph-replayed cost."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, time, torch
sys.path.insert(0, %r)
from sbgm_danra_b200.synth import config_for, synth_batch, synth_state_dict
from sbgm_danra_b200 import score_sampling as ss
from sbgm_danra_b200._smoke import build_model
from sbgm_danra_b200.score_unet import diffusion_coeff_fn, marginal_prob_std_fn
cfg = config_for(n_lr=1)
net = build_model(cfg, synth_state_dict(cfg), sys.argv[1], "cuda:0")
b = synth_batch(batch=64, size=128, n_lr=1, shared_cond=True)
kw = dict(batch_size=64, num_steps=100, device="cuda:0", img_size=128, cond_img=b.cond_img.cuda())
for _ in range(2):
    ss.Euler_Maruyama_sampler(net, marginal_prob_std_fn, diffusion_coeff_fn, **kw)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    ss.Euler_Maruyama_sampler(net, marginal_prob_std_fn, diffusion_coeff_fn, **kw)
e1.record()
torch.cuda.synchronize()
print("MS_PER_NFE", e0.elapsed_time(e1) / 300)
''' % ROOT


def main():
    precision = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
    base = None
    variants = ["", "attn", "attn_core", "ln", "gn", "upsample", "c64", "conv_tc", "c64,conv_tc,attn,gn,upsample"]
    if len(sys.argv) > 2:       # e.g. tools/ablate.py bf16x3 tc_1x1 tc_s2 tc_big tc_8 tc_4  (classes of the generic conv kernel)
        variants = [""] + sys.argv[2:]
    for skip in variants:
        env = dict(os.environ, SBGM_B200_SKIP=skip)
        r = subprocess.run([sys.executable, "-c", CHILD, precision], env=env, capture_output=True, text=True)
        line = [l for l in r.stdout.splitlines() if l.startswith("MS_PER_NFE")]
        if not line:
            print(skip or "baseline", "FAILED", r.stderr[-400:])
            continue
        ms = float(line[0].split()[1])
        base = ms if base is None else base
        print(f"skip={skip or '-':10s} {ms:.4f} ms/NFE   delta vs baseline {base - ms:+.4f} ms ({100 * (base - ms) / base:+.1f}%)")


if __name__ == "__main__":
    main()
