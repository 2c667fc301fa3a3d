"""Per-kernel GPU time of one score-UNet evaluation at the benchmark shape (warm caches, eager launch sequence) from
torch.profiler's CUDA activity records: sum of kernel durations per kernel name and grid -- compare with the graph-replayed
ms/NFE of tools/ablate.py to see how much of an evaluation is launch gaps.

    python tools/profile_forward.py [--precision bf16x3] [--batch 64] [--size 128] [--top 50]"""
import argparse
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile


@torch.no_grad()
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="bf16x3")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--top", type=int, default=50)
    a = ap.parse_args()
    from sbgm_danra_b200.synth import config_for, synth_batch, synth_state_dict
    from sbgm_danra_b200._smoke import build_model
    cfg = config_for(n_lr=1)
    net = build_model(cfg, synth_state_dict(cfg), a.precision, "cuda:0")
    b = synth_batch(batch=a.batch, size=a.size, n_lr=1, shared_cond=True)
    x, t, c = b.x.cuda(), b.t.cuda(), b.cond_img.cuda()
    for _ in range(3):
        net(x, t, None, c)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(a.iters):
            net(x, t, None, c)
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0.0, 0])
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            name = ev.name.replace("void ", "").replace("sbgm::", "")
            name = name.split("(")[0][:70]
            agg[name][0] += ev.device_time_total
            agg[name][1] += 1
    total = sum(v[0] for v in agg.values())
    print(f"kernel time {total / a.iters / 1e3:.3f} ms per evaluation over {sum(v[1] for v in agg.values()) // a.iters} launches "
          f"({a.precision}, batch {a.batch}, {a.size}x{a.size})")
    for name, (us, cnt) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:a.top]:
        print(f"{us / a.iters:9.1f} us {100 * us / total:5.1f}%  x{cnt // a.iters:4d}  {name}")


if __name__ == "__main__":
    main()
