#!/usr/bin/env python
"""DSM training-step benchmark (BASELINE.json configs[3], "C4"): 128x128, global batch 64, Cin = 7 + season label,
loss_fn -> backward -> Adam, data-parallel over N GPUs with the bucketed NCCL gradient all-reduce of
sbgm_danra_b200.parallel.  One process per GPU (torchrun for N > 1).

    python tools/bench_train.py [--steps K] [--warmup W] [--precision bf16|bf16x3] [--batch 64] [--size 128] [--strong]

Prints one JSON line: samples/s (whole job), ms/step (CUDA events, max over ranks), algorithmic TFLOP/s
(3 x F_fwd x batch) and the forward / backward / optimizer split of rank 0."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist

FWD_FLOP_128_CIN7 = 5.313e9


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--batch", type=int, default=64, help="global batch (strong scaling) or per-GPU batch with --weak")
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--weak", action="store_true")
    ap.add_argument("--timeline", action="store_true", help="rank 0: profile three more steps and print the NCCL kernels' durations, "
                    "the span of a step and the largest gaps between consecutive kernels")
    ap.add_argument("--optimizer", default="one-launch", choices=["one-launch", "torch"],
                    help="sbgm_danra_b200.optim.Adam (torch.optim.Adam with a single-kernel step) or torch's own foreach Adam")
    args = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    from sbgm_danra_b200.synth import config_for, synth_batch, synth_state_dict
    from sbgm_danra_b200 import parallel, score_sampling
    from sbgm_danra_b200._smoke import build_model
    from sbgm_danra_b200.score_unet import loss_fn, marginal_prob_std_fn
    cfg = config_for(n_lr=2, geo=True, seasons=True)
    net = build_model(cfg, synth_state_dict(cfg), args.precision, dev).train()
    b_local = args.batch if args.weak else args.batch // world
    b = synth_batch(batch=b_local, size=args.size, n_lr=2, geo=True, seasons=True, seed=1234 + rank)
    c = lambda v: None if v is None else v.to(dev)
    x, y, cond, lsm, topo, sdf = c(b.x), c(b.y), c(b.cond_img), c(b.lsm_cond), c(b.topo_cond), c(b.sdf_cond)
    from sbgm_danra_b200 import optim as sbgm_optim
    opt = (torch.optim if args.optimizer == "torch" else sbgm_optim).Adam(net.parameters(), lr=1e-4)
    sync = None
    if world > 1:
        parallel.broadcast_parameters(net)
        sync = parallel.attach(net)
        score_sampling.set_ensemble_shard(first_member=rank * b_local, members_total=b_local * world)
    score_sampling.manual_seed(5)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    split = [0.0, 0.0, 0.0]

    def step(timed: bool) -> float:
        e = [ev() for _ in range(4)]
        opt.zero_grad(set_to_none=True)
        e[0].record()
        loss = loss_fn(net, x, marginal_prob_std_fn, y=y, cond_img=cond, lsm_cond=lsm, topo_cond=topo, sdf_cond=sdf)
        e[1].record()
        loss.backward()
        e[2].record()
        opt.step()
        e[3].record()
        if timed:
            torch.cuda.synchronize()
            for i in range(3):
                split[i] += e[i].elapsed_time(e[i + 1])
        return loss

    for _ in range(args.warmup):
        loss = step(False)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0, t1 = ev(), ev()
    t0.record()
    for _ in range(args.steps):
        loss = step(True)
    t1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([t0.elapsed_time(t1) / args.steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        gb = b_local * world
        scale = (args.size / 128) ** 2
        out = {"metric": "DSM training samples/sec", "value": gb / (ms.item() * 1e-3), "unit": "samples/s", "n_gpus": world,
               "ms_per_step": ms.item(), "global_batch": gb, "per_gpu_batch": b_local, "img_size": args.size, "precision": args.precision, "optimizer": args.optimizer,
               "scaling": "weak" if args.weak else "strong",
               "algorithmic_tflops": 3 * FWD_FLOP_128_CIN7 * scale * gb / (ms.item() * 1e-3) / 1e12,
               "rank0_ms": {"loss_fn_forward": split[0] / args.steps, "backward": split[1] / args.steps, "adam": split[2] / args.steps},
               "loss": float(loss), "grad_buckets": None if sync is None else sync.stats}
        print(json.dumps(out))
    if args.timeline and rank == 0:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(3):
                step(False)
            torch.cuda.synchronize()
        evs = sorted((ev for ev in prof.events() if ev.device_type == torch.autograd.DeviceType.CUDA), key=lambda ev: ev.time_range.start)
        evs = evs[len(evs) - len(evs) // 3:]
        span = evs[-1].time_range.end - evs[0].time_range.start
        nccl = [ev for ev in evs if "nccl" in ev.name.lower()]
        rest = [ev for ev in evs if "nccl" not in ev.name.lower()]
        busy = sum(ev.time_range.end - ev.time_range.start for ev in rest)
        print(f"[timeline] last step: span {span / 1e3:.3f} ms, non-NCCL kernel time {busy / 1e3:.3f} ms over {len(rest)} launches", file=sys.stderr)
        t0 = evs[0].time_range.start
        for ev in nccl:
            print(f"[timeline]   NCCL {ev.name.split('(')[0][:60]}: start {(ev.time_range.start - t0) / 1e3:.3f} ms, {(ev.time_range.end - ev.time_range.start):.1f} us",
                  file=sys.stderr)
        gaps = sorted(((rest[i + 1].time_range.start - rest[i].time_range.end, (rest[i].time_range.end - t0) / 1e3, rest[i].name.split("(")[0][:40],
                        rest[i + 1].name.split("(")[0][:40]) for i in range(len(rest) - 1)), reverse=True)[:8]
        for g, at, a_, b_ in gaps:
            print(f"[timeline]   gap {g:7.1f} us at {at:.3f} ms after {a_} before {b_}", file=sys.stderr)
    elif args.timeline and world > 1:
        for _ in range(3):
            step(False)
        torch.cuda.synchronize()
    if world > 1:
        # the captured backward graph holds NCCL all-reduces: drop it before the communicator goes away and skip the
        # interpreter teardown (destroying a communicator that graphs still reference can block)
        net._train_runners.clear()
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
