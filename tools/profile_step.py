"""One (or a few) UNet evaluations at the benchmark shape, for ncu launch lists / captures.

    python tools/profile_step.py [--precision bf16x3] [--batch 64] [--size 128] [--iters 2] [--events]
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv python tools/profile_step.py ...

With --events it prints a per-kernel-family time table measured with CUDA events around every C-ABI
call (serialised; shares only)."""
import argparse
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sbgm_danra_b200.synth import config_for, synth_batch, synth_state_dict
from sbgm_danra_b200 import _lib
from sbgm_danra_b200._smoke import build_model


@torch.no_grad()
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="bf16x3")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--iters", type=int, default=2)
    ap.add_argument("--events", action="store_true")
    ap.add_argument("--warmup", type=int, default=2)
    a = ap.parse_args()
    cfg = config_for(n_lr=1)
    net = build_model(cfg, synth_state_dict(cfg), a.precision, "cuda:0")
    b = synth_batch(batch=a.batch, size=a.size, n_lr=1, shared_cond=True)
    x, t, c = b.x.cuda(), b.t.cuda(), b.cond_img.cuda()
    for _ in range(a.warmup):
        net(x, t, None, c)
    torch.cuda.synchronize()
    if a.events:
        rec = []
        orig = _lib.call

        def timed_call(name, *args):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            orig(name, *args)
            e.record()
            rec.append((name, args, s, e))
        import sbgm_danra_b200.engine as E
        E.call = timed_call
        for _ in range(a.iters):
            net(x, t, None, c)
        torch.cuda.synchronize()
        E.call = orig
        agg = collections.OrderedDict()
        for name, args, s, e in rec:
            key = name
            if name == "sbgm_conv2d_tc":
                n, h, w, cin, cout, kh, kw, stride = args[12:20]
                key = f"conv_tc {cin}->{cout} k{kh} s{stride} @{h}x{w}"
            agg.setdefault(key, [0.0, 0])
            agg[key][0] += s.elapsed_time(e)
            agg[key][1] += 1
        total = sum(v[0] for v in agg.values())
        print(f"total {total / a.iters:.3f} ms per forward (serialised events)")
        for k, (ms, cnt) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            print(f"{ms / a.iters:8.3f} ms  {100 * ms / total:5.1f}%  x{cnt // a.iters:3d}  {k}")
    else:
        torch.cuda.profiler.start()       # `ncu --profile-from-start off`: the launch list holds these evaluations only
        for _ in range(a.iters):
            net(x, t, None, c)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    print("done")


if __name__ == "__main__":
    main()
