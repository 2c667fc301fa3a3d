"""Per-kernel GPU time of one captured Euler-Maruyama step (C2 shape) as replayed by the sampler: torch.profiler CUDA activity
records of the graph replays (warm caches, in-graph overlap), aggregated per kernel and divided by the number of steps.

    python tools/profile_sampler.py [--precision fp16x2] [--steps 20] [--members 64] [--size 128]"""
import argparse
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="fp16x2")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--members", type=int, default=64)
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--list", default="", help="comma-separated kernel-name substrings: print every matching launch of the LAST step, in order")
    a = ap.parse_args()
    from sbgm_danra_b200.synth import config_for, synth_batch, synth_state_dict
    from sbgm_danra_b200 import score_sampling as ss
    from sbgm_danra_b200._smoke import build_model
    from sbgm_danra_b200.score_unet import diffusion_coeff_fn, marginal_prob_std_fn
    dev = torch.device("cuda:0")
    cfg = config_for(n_lr=1)
    net = build_model(cfg, synth_state_dict(cfg), a.precision, dev)
    b = synth_batch(batch=a.members, size=a.size, shared_cond=True, n_lr=1)
    kw = dict(batch_size=a.members, num_steps=a.steps, device=dev, img_size=a.size, cond_img=b.cond_img.to(dev))
    for _ in range(2):
        ss.Euler_Maruyama_sampler(net, marginal_prob_std_fn, diffusion_coeff_fn, **kw)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        ss.Euler_Maruyama_sampler(net, marginal_prob_std_fn, diffusion_coeff_fn, **kw)
        torch.cuda.synchronize()
    evs = sorted((ev for ev in prof.events() if ev.device_type == torch.autograd.DeviceType.CUDA), key=lambda ev: ev.time_range.start)
    agg = collections.defaultdict(lambda: [0.0, 0])
    for ev in evs:
        name = ev.name.split("(")[0].replace("void sbgm::", "").replace("sbgm::", "")[:90]
        agg[name][0] += ev.time_range.end - ev.time_range.start
        agg[name][1] += 1
    span = evs[-1].time_range.end - evs[0].time_range.start
    busy = sum(v[0] for v in agg.values())
    print(f"EM sampler call, {a.steps} steps ({a.precision}, {a.members} members, {a.size}x{a.size}): span {span / a.steps:.1f} us per step, "
          f"sum of kernel durations {busy / a.steps:.1f} us per step over {len(evs) / a.steps:.1f} launches per step")
    for name, (us, cnt) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:a.top]:
        print(f"{us / a.steps:9.1f} us {100 * us / busy:5.1f}%  x{cnt / a.steps:5.1f}  {name}")
    if a.list:
        keys = [k for k in a.list.split(",") if k]
        # the last step = everything after the second-to-last predictor kernel
        pred = [i for i, ev in enumerate(evs) if "predictor_kernel" in ev.name]
        last = evs[pred[-2] + 1:] if len(pred) >= 2 else evs
        for i, ev in enumerate(last):
            if any(k in ev.name for k in keys):
                print(f"  #{i:3d} {ev.time_range.end - ev.time_range.start:8.1f} us  {ev.name.split('(')[0][:70]}")


if __name__ == "__main__":
    main()
