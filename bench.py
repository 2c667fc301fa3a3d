#!/usr/bin/env python
"""bench.py -- headline benchmark: EM-sampled 128x128 fields/sec (BASELINE.json), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision bf16x3|bf16|fp32]

Workload (BASELINE.json configs[1], "C2"): 128x128 temperature downscaling, single LR condition
(Cin = 2), 500-step Euler-Maruyama, 64-member ensemble per GPU (weak scaling: every rank samples
its own 64 members of a 64*N-member ensemble, no data-path collective), synthetic ERA5/DANRA-shaped
inputs, random-init weights of the reference architecture (19.06 M parameters).

A "step" is one complete sampler call (500 network evaluations + 500 fused updates) on one batch.
  value : fields/s with the conditioning already resident in HBM (CUDA events, max over ranks)
  e2e   : the same call with HOST buffers -- pinned-host conditioning copied H2D and the sampled
          ensemble copied D2H inside the timed region
  roofline     : the dominant kernel (tcgen05 implicit-GEMM convolution) timed live on the largest
                 layer of the network, algorithmic FLOPs / CUDA-event time vs MEASURED_PEAKS.json
  cpu_baseline : the oracle port of the reference's CPU path (torch fp32, all host threads) on a
                 bounded sample of the same workload (fewer members and steps; per-step cost is
                 step-independent), rank 0, N = 1 only
`--impl reference` prints the reference-arm line: the CPU path alone, same metric/config/unit.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "EM-sampled 128x128 fields/sec"
UNIT = "fields/s"
SIZE, MEMBERS, EM_STEPS, N_LR = 128, 64, 500, 1
FWD_FLOP = 5.146e9            # per sample per forward at 128x128, Cin = 2
DTYPE_NAMES = {"bf16x3": "bf16x3 (split-bf16 operands, fp32 accumulate; fp32-class)",
               "fp16x2": "fp16x2 (float16 activations x float16 hi|lo weights, fp32 accumulate; fp32-class: score rel-L2 < 1e-3)",
               "bf16": "bf16", "fp32": "f32"}
TENSOR_PRODUCTS = {"bf16x3": 3.0, "fp16x2": 2.0, "bf16": 1.0, "fp32": 1.0}     # tensor-core products per algorithmic product


def _config(args, world):
    return {"workload": "C2: 128x128 ERA5->DANRA temperature, Cin=2, Euler-Maruyama 500 steps, 64 members per GPU",
            "img_size": SIZE, "members_per_gpu": MEMBERS, "sampler_steps": EM_STEPS, "precision": args.precision,
            "global_members": MEMBERS * world, "parallelism": f"ensemble-shard x{world} (no collective)",
            "l2": "256 MiB L2 flush between timed steps; per-network-evaluation activation footprint (~1.5 GB) exceeds the 126 MB L2"}


def load_traffic():
    """DRAM bytes (read + write) of ONE launch of the dominant kernel from the committed `ncu --set full` capture
    (profiles/r01_dominant_kernel_ncu.json, written by tools/ncu_traffic.py), or None."""
    p = os.path.join(ROOT, "profiles", "r01_dominant_kernel_ncu.json")
    if not os.path.exists(p):
        return None
    with open(p) as f:
        d = json.load(f)
    return d["dram_bytes_read"] + d["dram_bytes_write"]


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = sorted(sm)[len(sm) // 4:] if sm else []
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---- CPU arm: oracle port of the reference ----------------------------------------------------------
def cpu_em_fields_per_s(members: int, steps: int, threads: int, seed: int = 0):
    """Times `steps` Euler-Maruyama steps of the oracle (torch fp32 on the host) for `members` members and
    extrapolates linearly to the 500-step sampler.  Returns (fields/s, seconds measured)."""
    import torch
    from oracle import samplers_ref, score_ref
    from sbgm_danra_b200.synth import config_for, synth_batch, synth_state_dict
    torch.set_num_threads(threads)
    cfg = config_for(n_lr=N_LR)
    sd = synth_state_dict(cfg, seed)
    b = synth_batch(batch=members, size=SIZE, n_lr=N_LR, shared_cond=True)

    def score(x, t):
        return score_ref.score_forward(sd, cfg, x, t, None, b.cond_img)

    t0 = time.perf_counter()
    samplers_ref.euler_maruyama(score, score_ref.marginal_prob_std, score_ref.diffusion_coeff, members, steps, img_size=SIZE)
    dt = time.perf_counter() - t0
    return members / (dt / steps * EM_STEPS), dt


def run_reference(args):
    """Reference arm: the reference's CPU implementation (oracle port) on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    threads = os.cpu_count() or 1
    members = 8
    _, t2 = cpu_em_fields_per_s(members, 2, threads)          # calibrate: ~5 s of CPU work per timed step
    steps = max(2, min(60, int(5.0 / max(t2 / 2, 1e-3))))
    for _ in range(max(args.warmup - 1, 0)):
        cpu_em_fields_per_s(members, 2, threads)
    vals, secs = [], []
    for _ in range(args.steps):
        v, dt = cpu_em_fields_per_s(members, steps, threads)
        vals.append(v); secs.append(dt)
    value = statistics.mean(vals)
    sample = f"{members} members x {steps} EM steps per timed step, extrapolated x{EM_STEPS // steps} to 500 steps"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * statistics.mean(secs), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": _config(args, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "torch_threads": torch.get_num_threads()}))


# ---- our arm -----------------------------------------------------------------------------------
def time_dominant_kernel(net, precision: str, iters: int = 20):
    """Largest single convolution of the network (decoder.final_layer.conv_up, 64->64 3x3 at 128x128,
    1.208 GFLOP/sample): CUDA events on the launching stream, L2 flushed between launches."""
    import torch
    from sbgm_danra_b200 import engine as E
    eng = net.engine()
    k, cw = eng.dec.k, eng.dec.final_up
    x = E.Act(eng.fmt, MEMBERS, SIZE, SIZE, 64, eng.device)
    x.buf.normal_()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=eng.device)
    # exactly the launch the sampler step makes: the persistent 64->64 kernel with the projection epilogue (tensor-core formats)
    proj = eng.dec.final_w[0] if (precision != "fp32" and eng.dec.out_channels == 1) else None
    for _ in range(3):
        k.conv(x, cw, pad=1, proj=proj)
    times = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        k.conv(x, cw, pad=1, proj=proj)
        b.record()
        b.synchronize()
        times.append(a.elapsed_time(b))
    ms = statistics.median(times)
    flops = 2.0 * MEMBERS * SIZE * SIZE * 64 * 64 * 9
    return flops, ms


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from sbgm_danra_b200.synth import config_for, synth_batch, synth_state_dict
    from sbgm_danra_b200 import _lib, score_sampling as ss
    from sbgm_danra_b200._smoke import build_model
    from sbgm_danra_b200.score_unet import diffusion_coeff_fn, marginal_prob_std_fn

    cfg = config_for(n_lr=N_LR)
    net = build_model(cfg, synth_state_dict(cfg, 0), args.precision, dev)
    cond_host = synth_batch(batch=MEMBERS, size=SIZE, n_lr=N_LR, shared_cond=True, seed=1234 + rank).cond_img.pin_memory()
    cond_dev = cond_host.to(dev)
    out_host = torch.empty((MEMBERS, 1, SIZE, SIZE), dtype=torch.float32).pin_memory()
    ss.set_ensemble_shard(rank * MEMBERS, MEMBERS * world)
    ss.manual_seed(4242)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def sample(cond):
        return ss.Euler_Maruyama_sampler(net, marginal_prob_std_fn, diffusion_coeff_fn, batch_size=MEMBERS,
                                         num_steps=EM_STEPS, device=dev, img_size=SIZE, cond_img=cond)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        times = []
        for _ in range(n):
            flush.zero_()
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            barrier()
            times.append(a.elapsed_time(b))
        t = torch.tensor([sum(times)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / n      # ms per step, max over ranks

    def e2e_step():
        c = cond_host.to(dev, non_blocking=True)
        out_host.copy_(sample(c), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(args.warmup):
        sample(cond_dev)
    with ClockSampler(local) as clocks:
        before = _lib.stats.launches
        ms = timed(lambda: sample(cond_dev), args.steps)
        launches = _lib.stats.launches - before
        ms_e2e = timed(e2e_step, args.steps)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    fields = MEMBERS * world
    print(f"[bench] {ms:.1f} ms per {EM_STEPS}-step sampler call, e2e {ms_e2e:.1f} ms", file=sys.stderr)
    value = fields / (ms * 1e-3)
    flops, kms = time_dominant_kernel(net, args.precision)
    achieved = flops / (kms * 1e-3) / 1e12
    fwd_tflops = MEMBERS * FWD_FLOP * EM_STEPS / (ms * 1e-3) / 1e12
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": DTYPE_NAMES[args.precision],
        "data": "synthetic", "config": _config(args, world),
        "e2e": {"value": fields / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": cond_host.numel() * 4,
                "d2h_bytes_per_step": out_host.numel() * 4},
        "gpu_launches": launches,
        "clocks": clocks.summary(),
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                     "frac": achieved / peaks["bf16_tflops"], "traffic": load_traffic(),
                     "kernel": "conv3x3_c64_kernel (persistent tcgen05 implicit GEMM, projection epilogue) on decoder.final_layer.conv_up 64->64 3x3 @128x128 x64",
                     "kernel_ms": kms, "algorithmic_flops_per_launch": flops, "peak_source": peaks["source"] + ", burst bf16",
                     "tensor_pipe_frac": TENSOR_PRODUCTS[args.precision] * achieved / peaks["bf16_tflops"],
                     "note": "achieved/frac count ALGORITHMIC FLOPs; bf16x3 issues 3 bf16 tensor-core products per algorithmic "
                             "product, so the tensor pipe is busy tensor_pipe_frac of the measured bf16 peak"},
        "unet_fwd_tflops": fwd_tflops, "unet_fwd_frac_of_sustained_bf16": fwd_tflops / peaks["bf16_tflops_sustained"],
    }
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        _, t2 = cpu_em_fields_per_s(8, 2, threads)            # warm-up + calibration
        n_cpu = max(3, min(100, int(15.0 / max(t2 / 2, 1e-3))))  # ~15 s of CPU work
        v, dt = cpu_em_fields_per_s(8, n_cpu, threads)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"8 members x {n_cpu} EM steps ({dt:.1f} s), extrapolated linearly to 500 steps"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16x3", choices=["bf16x3", "fp16x2", "bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
