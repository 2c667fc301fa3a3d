"""Per-kernel GPU time of one DSM training step (C4 shape), warm caches, eager engine (SBGM_B200_TRAIN_GRAPHS=0), via
torch.profiler's CUDA activity records.  Sum of kernel durations, not wall time: compare with tools/bench_train.py.

    python tools/profile_train.py [--precision bf16] [--batch 64] [--top 45]"""
import argparse
import collections
import os
import sys

if "--graphs" not in sys.argv:
    os.environ["SBGM_B200_TRAIN_GRAPHS"] = "0"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--top", type=int, default=45)
    ap.add_argument("--graphs", action="store_true", help="keep the CUDA-graph replay on and report the GPU idle gaps of a step")
    ap.add_argument("--list", default="", help="comma-separated kernel-name substrings: also print every matching launch of the last step, in order")
    a = ap.parse_args()
    from sbgm_danra_b200.synth import config_for, synth_batch, synth_state_dict
    from sbgm_danra_b200 import score_sampling
    from sbgm_danra_b200._smoke import build_model
    from sbgm_danra_b200.score_unet import loss_fn, marginal_prob_std_fn
    dev = "cuda:0"
    cfg = config_for(n_lr=2, geo=True, seasons=True)
    net = build_model(cfg, synth_state_dict(cfg), a.precision, dev).train()
    b = synth_batch(batch=a.batch, size=a.size, n_lr=2, geo=True, seasons=True, seed=1234)
    c = lambda v: None if v is None else v.to(dev)
    x, y, cond, lsm, topo, sdf = c(b.x), c(b.y), c(b.cond_img), c(b.lsm_cond), c(b.topo_cond), c(b.sdf_cond)
    from sbgm_danra_b200 import optim as sbgm_optim
    opt = sbgm_optim.Adam(net.parameters(), lr=1e-4)
    score_sampling.manual_seed(5)

    def step():
        opt.zero_grad(set_to_none=True)
        loss = loss_fn(net, x, marginal_prob_std_fn, y=y, cond_img=cond, lsm_cond=lsm, topo_cond=topo, sdf_cond=sdf)
        loss.backward()
        opt.step()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(a.steps):
            step()
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0.0, 0])
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            name = ev.name.split("(")[0][:90]
            agg[name][0] += ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
            agg[name][1] += 1
    total = sum(v[0] for v in agg.values())
    print(f"kernel time {total / a.steps / 1e3:.3f} ms per step over {sum(v[1] for v in agg.values()) // a.steps} launches ({a.precision}, batch {a.batch})")
    for name, (us, cnt) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:a.top]:
        print(f"{us / a.steps:9.1f} us {100 * us / total:5.1f}%  x{cnt // a.steps:4d}  {name}")
    if a.graphs:
        evs = sorted((ev for ev in prof.events() if ev.device_type == torch.autograd.DeviceType.CUDA), key=lambda ev: ev.time_range.start)
        evs = evs[len(evs) - len(evs) // a.steps:]
        span = evs[-1].time_range.end - evs[0].time_range.start
        busy = sum(ev.time_range.end - ev.time_range.start for ev in evs)
        gaps = [(evs[i + 1].time_range.start - evs[i].time_range.end, evs[i].name.split("(")[0][:50], evs[i + 1].name.split("(")[0][:50])
                for i in range(len(evs) - 1)]
        print(f"last step under graph replay: span {span / 1e3:.3f} ms, kernels {busy / 1e3:.3f} ms, idle {sum(max(g[0], 0) for g in gaps) / 1e3:.3f} ms "
              f"over {len(gaps)} boundaries (median gap {sorted(g[0] for g in gaps)[len(gaps) // 2]:.2f} us)")
        for g in sorted(gaps, key=lambda g: -g[0])[:12]:
            print(f"   gap {g[0]:8.1f} us  after {g[1]}  before {g[2]}")
    if a.list:
        keys = [k for k in a.list.split(",") if k]
        evs = sorted((ev for ev in prof.events() if ev.device_type == torch.autograd.DeviceType.CUDA), key=lambda ev: ev.time_range.start)
        evs = evs[len(evs) - len(evs) // a.steps:]
        for i, ev in enumerate(evs):
            if any(k in ev.name for k in keys):
                dur = ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
                print(f"  #{i:4d} {dur:8.1f} us  {ev.name.split('(')[0][:70]}")


if __name__ == "__main__":
    main()
