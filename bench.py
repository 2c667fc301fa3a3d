#!/usr/bin/env python
"""bench.py -- headline benchmark: EM-sampled 128x128 fields/sec (BASELINE.json), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision fp16x2|bf16x3|bf16|fp32]

Workload (BASELINE.json configs[1], "C2"): 128x128 temperature downscaling, single LR condition
(Cin = 2), 500-step Euler-Maruyama, 64-member ensemble per GPU (weak scaling: every rank samples
its own 64 members of a 64*N-member ensemble, no data-path collective), synthetic ERA5/DANRA-shaped
inputs, random-init weights of the reference architecture (19.06 M parameters).

A "step" is one complete sampler call (500 network evaluations + 500 fused updates) on one batch.
  value : fields/s with the conditioning already resident in HBM (CUDA events, max over ranks)
  e2e   : the same call with HOST buffers -- pinned-host conditioning copied H2D and the sampled
          ensemble copied D2H inside the timed region
  roofline     : the largest single kernel (persistent 64->64 tcgen05 convolution on the final layer) timed live,
                 algorithmic FLOPs / CUDA-event time vs MEASURED_PEAKS.json; next to it `family` = every launch of the
                 generic tcgen05 convolution kernel of one evaluation (the time-dominant kernel family) timed one by one,
                 and `path_frac` = whole-path UNet forward TFLOP/s over the sustained bf16 peak
  extra        : the two collective-bearing BASELINE workloads in the same run, at the same N:
                 pc_c3 (128x128, Cin = 7 + seasons, predictor-corrector, bf16; 64 members sharded over the ranks with the
                 per-step all-gather of the gradient norms in the captured graph, and 64 members per rank) and
                 train_c4 (DSM step 128x128, bf16, DDP with the bucketed NCCL all-reduce; global batch 64 and 64 per rank)
  gpu_eager_baseline : the reference's own modules (baseline/_ref, unmodified) as stock torch eager on the same GPU,
                 TF32 off (the reference's setting) and on -- BASELINE.md's "kernel to beat"; N = 1 only
  cpu_baseline : the reference's CPU path (baseline/_ref modules; the oracle port if they are absent) on a bounded sample
                 of the same workload, all host threads, rank 0, N = 1 only
`--impl reference` prints the reference-arm line: that CPU path alone, same metric/config/unit.
"""
from __future__ import annotations

import argparse
import importlib.util
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
FWD_FLOP_CIN7 = 5.313e9       # Cin = 7 + seasons (C3 / C4)
DEFAULT_PRECISION = "fp16x2"
DTYPE_NAMES = {"bf16x3": "bf16x3 (split-bf16 operands, fp32 accumulate; fp32-class)",
               "fp16x2": "fp16x2 (float16 activations x float16 hi|lo weights, fp32 accumulate; fp32-class: score rel-L2 < 1e-3)",
               "bf16": "bf16", "fp32": "f32"}
TENSOR_PRODUCTS = {"bf16x3": 3.0, "fp16x2": 2.0, "bf16": 1.0, "fp32": 1.0}     # tensor-core products per algorithmic product
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def _config(args, world):
    return {"workload": "C2: 128x128 ERA5->DANRA temperature, Cin=2, Euler-Maruyama 500 steps, 64 members per GPU",
            "img_size": SIZE, "members_per_gpu": MEMBERS, "sampler_steps": EM_STEPS, "precision": args.precision,
            "global_members": MEMBERS * world, "parallelism": f"ensemble-shard x{world} (no collective)",
            "l2": "256 MiB L2 flush between timed steps; per-network-evaluation activation footprint (~1 GB) exceeds the 126 MB L2"}


def load_traffic(precision: str):
    """DRAM bytes (read + write) of ONE launch of the dominant kernel from the committed `ncu --set full` capture of this
    precision (profiles/*dominant_kernel_ncu*.json, written by tools/ncu_traffic.py), or None."""
    names = [f"r02_dominant_kernel_ncu_{precision}.json"] + (["r01_dominant_kernel_ncu.json"] if precision == "bf16x3" else [])
    for name in names:
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            with open(p) as f:
                d = json.load(f)
            return d["dram_bytes_read"] + d["dram_bytes_write"]
    return None


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


# ---- the reference's own modules (baseline/_ref: unmodified copies made by __graft_entry__.build()) ------------------------
class _TorchShim:
    """Stands in for the name `torch` inside the reference's score_sampling module: everything is torch's, except that the
    hard-coded (B, 1, 32, 32) initial state of Euler_Maruyama_sampler / ode_sampler (sbgm/score_sampling.py:94, :274 --
    the reference crashes at any other img_size) is drawn at the requested size.  The file itself stays
    byte-identical to the reference."""

    def __init__(self, torch_mod, size: int):
        self._t, self._size = torch_mod, size

    def __getattr__(self, name):
        return getattr(self._t, name)

    def randn(self, *shape, **kw):
        if len(shape) == 4 and tuple(shape[2:]) == (32, 32):
            shape = (shape[0], shape[1], self._size, self._size)
        return self._t.randn(*shape, **kw)


def load_reference_modules():
    """(score_unet, score_sampling) of the REAL reference from baseline/_ref/sbgm/, or None if they did not travel."""
    paths = [os.path.join(REF_DIR, "sbgm", f) for f in ("score_unet.py", "score_sampling.py")]
    if not all(os.path.exists(p) for p in paths):
        return None
    mods = []
    for name, p in zip(("_ref_score_unet", "_ref_score_sampling"), paths):
        spec = importlib.util.spec_from_file_location(name, p)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mods.append(mod)
    return tuple(mods)


def build_reference_net(ref_unet, cfg, sd, device):
    import torch.nn as nn
    act = {"relu": nn.ReLU, "silu": nn.SiLU, "gelu": nn.GELU}[cfg.activation]
    enc = ref_unet.Encoder(cfg.in_channels, cfg.time_embedding, block_layers=list(cfg.block_layers), n_heads=cfg.n_heads,
                           num_classes=cfg.num_classes, device=device)
    dec = ref_unet.Decoder(cfg.last_fmap_channels, cfg.out_channels, cfg.time_embedding, n_heads=cfg.n_heads, device=device,
                           use_resize_conv=cfg.use_resize_conv, norm=cfg.norm, gn_groups=cfg.gn_groups, activation=act)
    net = ref_unet.ScoreNet(ref_unet.marginal_prob_std_fn, enc, dec, device=device, debug_pre_sigma_div=False)
    net.load_state_dict(sd, strict=True)
    return net.eval()


class ReferenceEM:
    """The reference's Euler-Maruyama sampler (real modules when baseline/_ref is present, else the oracle port) on `device`."""

    def __init__(self, device="cpu", seed: int = 0):
        import torch
        from sbgm_danra_b200.synth import config_for, synth_batch, synth_state_dict
        os.environ.setdefault("TQDM_DISABLE", "1")
        self.torch, self.device = torch, device
        self.cfg = config_for(n_lr=N_LR)
        sd = synth_state_dict(self.cfg, seed)
        self.cond = synth_batch(batch=MEMBERS, size=SIZE, n_lr=N_LR, shared_cond=True).cond_img.to(device)
        mods = load_reference_modules()
        if mods is not None:
            self.kind = "reference"
            self.unet, self.samp = mods
            self.samp.torch = _TorchShim(torch, SIZE)
            self.net = build_reference_net(self.unet, self.cfg, sd, device)
        else:
            self.kind = "port"
            self.sd = {k: v.to(device) for k, v in sd.items()}

    def forward(self, x, t):
        if self.kind == "reference":
            return self.net(x, t, None, self.cond[:x.shape[0]])
        from oracle import score_ref
        return score_ref.score_forward(self.sd, self.cfg, x, t, None, self.cond[:x.shape[0]])

    def sample(self, members: int, steps: int):
        """`steps` EM steps of a `members`-member ensemble (the per-step cost does not depend on the step count)."""
        if self.kind == "reference":
            return self.samp.Euler_Maruyama_sampler(self.net, self.unet.marginal_prob_std_fn, self.unet.diffusion_coeff_fn,
                                                    batch_size=members, num_steps=steps, device=self.device, img_size=SIZE,
                                                    cond_img=self.cond[:members])
        from oracle import samplers_ref, score_ref
        return samplers_ref.euler_maruyama(self.forward, score_ref.marginal_prob_std, score_ref.diffusion_coeff, members, steps,
                                           img_size=SIZE, device=self.device)


def cpu_em_fields_per_s(ref: ReferenceEM, members: int, steps: int):
    """Times `steps` EM steps on the host and extrapolates linearly to the 500-step sampler.  Returns (fields/s, seconds)."""
    t0 = time.perf_counter()
    ref.sample(members, steps)
    dt = time.perf_counter() - t0
    return members / (dt / steps * EM_STEPS), dt


def _cpu_sample_plan(ref: ReferenceEM, seconds: float):
    """Members and steps of a bounded CPU sample worth about `seconds` of work (the full 64-member batch when two steps fit)."""
    members = MEMBERS
    _, t2 = cpu_em_fields_per_s(ref, members, 2)              # warm-up + calibration
    if t2 > seconds:                                           # two 64-member steps alone exceed the time box
        members = 8
        _, t2 = cpu_em_fields_per_s(ref, members, 2)
    return members, max(2, min(100, int(seconds / max(t2 / 2, 1e-3))))


def run_reference(args):
    """Reference arm: the reference's CPU implementation on this box's host cores, all threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    ref = ReferenceEM("cpu")
    members, steps = _cpu_sample_plan(ref, 5.0)               # ~5 s of CPU work per timed step
    for _ in range(max(args.warmup - 1, 0)):
        cpu_em_fields_per_s(ref, members, 2)
    vals, secs = [], []
    for _ in range(args.steps):
        v, dt = cpu_em_fields_per_s(ref, members, steps)
        vals.append(v); secs.append(dt)
    value = statistics.mean(vals)
    sample = f"{members} members x {steps} EM steps per timed step, extrapolated linearly to 500 steps"
    cfg = _config(args, 1)
    cfg["precision"] = "fp32 (torch CPU eager)"
    what = ("unmodified sbgm/score_unet.py + score_sampling.py from baseline/_ref (EM initial shape shimmed to img_size)"
            if ref.kind == "reference" else "oracle port of the reference (baseline/_ref absent)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * statistics.mean(secs), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": ref.kind, "sample": sample, "what": what},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "torch_threads": torch.get_num_threads()}))


# ---- our arm ---------------------------------------------------------------------------------------------
def _event_ms(torch, fn, iters, flush=None, reps: int = 1):
    """Median CUDA-event time of `fn` (per call when reps > 1).  A launch that precedes the first event keeps the GPU busy
    while the host enqueues (the L2 flush when given, else a ~100 us spin kernel): without it the events would bracket
    the HOST's launch latency of a microsecond-scale kernel, not the kernel."""
    times = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        else:
            torch.cuda._sleep(200000)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        b.synchronize()
        times.append(a.elapsed_time(b) / reps)
    return statistics.median(times)


def time_dominant_kernel(net, precision: str, iters: int = 20):
    """Largest single convolution of the network (decoder.final_layer.conv_up, 64->64 3x3 at 128x128,
    1.208 GFLOP/sample): CUDA events on the launching stream, L2 flushed between launches."""
    import torch
    from sbgm_danra_b200 import engine as E
    eng = net.engine()
    k, cw = eng.dec.k, eng.dec.final_up
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=eng.device)
    # exactly the launch the sampler step makes: the persistent 64->64 kernel with the projection epilogue (tensor-core formats),
    # in the single-plane formats with the bilinear upsample of its input inside the operand stage (the kernel then reads the
    # 64 x 64 tensor; the FLOPs counted are the convolution's alone)
    proj = eng.dec.final_w[0] if (precision != "fp32" and eng.dec.out_channels == 1) else None
    x_lo = E.Act(eng.fmt, MEMBERS, SIZE // 2, SIZE // 2, 64, eng.device)
    x_lo.buf.normal_()
    if k.up_fused_ok(x_lo, cw):
        fn, name = (lambda: k.conv_up_fused(x_lo, cw, proj=proj)), "conv3x3_c64_kernel, UP mode (persistent tcgen05 implicit GEMM, bilinear upsample in the operand stage, projection epilogue)"
    else:
        x = E.Act(eng.fmt, MEMBERS, SIZE, SIZE, 64, eng.device)
        x.buf.normal_()
        fn, name = (lambda: k.conv(x, cw, pad=1, proj=proj)), "conv3x3_c64_kernel (persistent tcgen05 implicit GEMM, projection epilogue)"
    for _ in range(3):
        fn()
    ms = _event_ms(torch, fn, iters, flush)
    flops = 2.0 * MEMBERS * SIZE * SIZE * 64 * 64 * 9
    return flops, ms, name


def time_conv_family(net, peaks, precision: str):
    """The time-dominant kernel FAMILY: every launch of the generic tcgen05 convolution kernel (`conv_tc_kernel`: strided,
    1x1 / Linear and 3x3 layers other than 64->64) of one 64-member evaluation, traced from a real forward and then timed one
    by one with its real epilogue -- cold (L2 flushed before each launch) and warm (same launch repeated)."""
    import torch
    from sbgm_danra_b200 import engine as E
    from sbgm_danra_b200.synth import synth_batch
    eng = net.engine()
    b = synth_batch(batch=MEMBERS, size=SIZE, n_lr=N_LR, shared_cond=True)
    dev = eng.device
    E.CONV_TRACE = []
    try:
        with torch.no_grad():
            eng.forward(b.x.to(dev), b.t.to(dev), None, b.cond_img.to(dev), torch.ones(MEMBERS, device=dev))
        trace = E.CONV_TRACE
    finally:
        E.CONV_TRACE = None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    k = eng.dec.k
    cold = warm = flops = 0.0
    launches = 0
    rows = []
    for rec in trace:
        if rec["c64"] or rec["proj"]:
            continue
        cw = rec["cw"]
        x = E.Act(eng.fmt, rec["n"], rec["h"], rec["w"], cw.cin, dev)
        x.buf.normal_()
        ho = (rec["h"] + 2 * rec["pad"] - cw.kh) // rec["stride"] + 1
        wo = (rec["w"] + 2 * rec["pad"] - cw.kw) // rec["stride"] + 1
        res = None
        if rec["residual"]:
            res = E.Act(eng.fmt, rec["n"], ho, wo, cw.cout, dev)
            res.buf.normal_()
        tp = torch.randn(rec["n"], cw.cout, device=dev) if rec["tproj"] else None

        def fn(rec=rec, cw=cw, x=x, res=res, tp=tp):
            return k.conv(x, cw, stride=rec["stride"], pad=rec["pad"], act=rec["act"], residual=res, tproj=tp, gn_stats=rec["gn_stats"])

        for _ in range(2):
            fn()
        c_ms, w_ms = _event_ms(torch, fn, 5, flush), _event_ms(torch, fn, 5, reps=4)
        fl = 2.0 * rec["n"] * ho * wo * cw.cin * cw.cout * cw.kh * cw.kw
        rows.append(f"{rec['n']:3d} x {rec['h']:3d}x{rec['w']:<5d} {cw.cin:4d}->{cw.cout:<4d} k{cw.kh} s{rec['stride']}  {fl / 1e9:7.2f} GFLOP  "
                    f"warm {w_ms * 1e3:7.1f} us ({fl / (w_ms * 1e-3) / 1e12:6.1f} TFLOP/s)  cold {c_ms * 1e3:7.1f} us")
        cold += c_ms
        warm += w_ms
        flops += fl
        launches += 1
    ach = flops / (warm * 1e-3) / 1e12
    if os.environ.get("SBGM_BENCH_FAMILY_TABLE"):             # per-layer table for profiles/
        with open(os.environ["SBGM_BENCH_FAMILY_TABLE"], "w") as f:
            f.write(f"conv_tc_kernel launches of one evaluation ({precision}, 64 members, 128x128), one by one; tokens x 1 = Linear layers\n")
            f.write("\n".join(rows) + f"\ntotal: {flops / 1e9:.1f} GFLOP, warm {warm * 1e3:.1f} us, cold {cold * 1e3:.1f} us\n")
    return {"kernel": "conv_tc_kernel (generic tcgen05 implicit GEMM: every strided / 1x1 / Linear / 3x3 layer of one evaluation except the 64->64 ones)",
            "launches": launches, "algorithmic_flops": flops, "ms_warm": warm, "ms_cold": cold,
            "achieved": ach, "unit": "TFLOP/s", "frac": ach / peaks["bf16_tflops"],
            "tensor_pipe_frac": TENSOR_PRODUCTS[precision] * ach / peaks["bf16_tflops"],
            "note": "sum of per-launch CUDA-event medians; warm = the same launch four times back to back behind a spin kernel "
                    "(operands L2-resident where they fit, launch overlap as in the captured graph), cold = 256 MiB L2 flush before "
                    "every launch; achieved/frac from the warm sum"}


def gpu_eager_baseline(dev):
    """Stock torch eager on the same GPU -- the reference's own modules when baseline/_ref travelled (else the oracle port):
    the GPU "kernel to beat" (the reference ships no GPU code; on a GPU it is cuDNN / cuBLAS / ATen eager)."""
    import torch
    ref = ReferenceEM(dev)
    b_x = torch.randn(MEMBERS, 1, SIZE, SIZE, device=dev)
    b_t = torch.rand(MEMBERS, device=dev) * 0.9 + 0.05
    torch.backends.cudnn.benchmark = True                       # sbgm/training_main.py:108-110
    out = {"kind": ref.kind, "members": MEMBERS, "fwd_ms": {}, "unet_fwd_tflops": {}, "em_fields_per_s": {}}
    saved = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    try:
        for mode in ("fp32", "tf32"):
            torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = (mode == "tf32")
            with torch.no_grad():
                for _ in range(3):
                    ref.forward(b_x, b_t)
                ms = _event_ms(torch, lambda: ref.forward(b_x, b_t), 7, flush=None)
                steps = 10
                a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                ref.sample(MEMBERS, steps)
                e.record()
                e.synchronize()
                em_ms = a.elapsed_time(e)
            out["fwd_ms"][mode] = ms
            out["unet_fwd_tflops"][mode] = MEMBERS * FWD_FLOP / (ms * 1e-3) / 1e12
            out["em_fields_per_s"][mode] = MEMBERS / (em_ms * 1e-3 / steps * EM_STEPS)
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = saved
    out["sample"] = "forward: median of 7 after 3 warm-ups; EM: 10 eager steps of the reference sampler, extrapolated linearly to 500"
    out["note"] = ("TF32 moves only the contraction kernels (4.9 -> 2.4 ms of the 22 ms forward); 15 ms of it is ATen's NCHW fp32 "
                   "upsample_bilinear2d kernel, 5 launches (profiles/r02_torch_eager_kernel_breakdown.txt)")
    return out


def bench_pc_c3(dev, rank, world, steps, warmup):
    """BASELINE C3: 128x128, Cin = 7 (two LR fields + land-sea mask + topography) + season labels, predictor-corrector 500
    steps (2 evaluations per step), bf16.  strong: 64 members in total sharded over the ranks, exact mode = one all-gather of
    the per-member gradient norms per step inside the captured graph; weak: 64 members per rank, same exchange."""
    import torch
    import torch.distributed as dist
    from sbgm_danra_b200 import score_sampling as ss
    from sbgm_danra_b200._smoke import build_model
    from sbgm_danra_b200.score_unet import diffusion_coeff_fn, marginal_prob_std_fn
    from sbgm_danra_b200.synth import config_for, synth_batch, synth_state_dict
    ck = dict(n_lr=2, geo=True, seasons=True)
    cfg = config_for(**ck)
    net = build_model(cfg, synth_state_dict(cfg, 0), "bf16", dev)
    out = {"workload": "C3: 128x128, Cin=7 + seasons, predictor-corrector 500 steps (2 NFE/step), bf16", "unit": UNIT}
    for mode in ("strong", "weak"):
        local = MEMBERS // world if mode == "strong" else MEMBERS
        total = local * world
        if local < 1 or (mode == "weak" and world == 1):      # at one rank the two are the same run
            continue
        b = synth_batch(batch=local, size=SIZE, shared_cond=True, seed=1234, **ck)
        c = lambda v: v.to(dev)
        y, cond, lsm, topo = c(b.y), c(b.cond_img), c(b.lsm_cond), c(b.topo_cond)
        ss.set_ensemble_shard(rank * local, total, dist.group.WORLD if world > 1 else None)
        ss.manual_seed(99)
        call = lambda: ss.pc_sampler(net, marginal_prob_std_fn, diffusion_coeff_fn, batch_size=local, num_steps=EM_STEPS, snr=0.16,
                                     device=dev, img_size=SIZE, y=y, cond_img=cond, lsm_cond=lsm, topo_cond=topo)
        for _ in range(warmup):
            call()
        times = []
        for _ in range(steps):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            call()
            e.record()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            times.append(a.elapsed_time(e))
        t = torch.tensor([sum(times) / len(times)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        out[mode] = {"value": total / (ms * 1e-3), "ms_per_call": ms, "members_total": total, "members_per_gpu": local,
                     "collective": ("all_gather_into_tensor of per-member score norms, once per step, inside the CUDA graph"
                                    if world > 1 else "none (1 rank)"),
                     "unet_fwd_tflops": total * FWD_FLOP_CIN7 * 2 * EM_STEPS / (ms * 1e-3) / 1e12}
        ss.clear_sampler_cache()
    ss.set_ensemble_shard(0, None, None)
    return out


def bench_train_c4(dev, rank, world, steps, warmup):
    """BASELINE C4: DSM training step 128x128, Cin = 7 + seasons, bf16, loss_fn -> backward -> Adam, DDP through
    sbgm_danra_b200.parallel (flat gradient buffer, bucketed NCCL all-reduce overlapped with backward).  strong: global batch
    64; weak: 64 per rank."""
    import torch
    import torch.distributed as dist
    from sbgm_danra_b200 import optim as sbgm_optim, parallel, score_sampling
    from sbgm_danra_b200._smoke import build_model
    from sbgm_danra_b200.score_unet import loss_fn, marginal_prob_std_fn
    from sbgm_danra_b200.synth import config_for, synth_batch, synth_state_dict
    ck = dict(n_lr=2, geo=True, seasons=True)
    cfg = config_for(**ck)
    out = {"workload": "C4: DSM training step 128x128, Cin=7 + seasons, bf16, Adam (sbgm_danra_b200.optim: torch.optim.Adam with a one-launch step), DDP bucketed all-reduce", "unit": "samples/s"}
    for mode in ("strong", "weak"):
        local = MEMBERS // world if mode == "strong" else MEMBERS
        if local < 1 or (mode == "weak" and world == 1):
            continue
        net = build_model(cfg, synth_state_dict(cfg), "bf16", dev).train()
        b = synth_batch(batch=local, size=SIZE, seed=1234 + rank, **ck)
        c = lambda v: v.to(dev)
        x, y, cond, lsm, topo, sdf = c(b.x), c(b.y), c(b.cond_img), c(b.lsm_cond), c(b.topo_cond), c(b.sdf_cond)
        opt = sbgm_optim.Adam(net.parameters(), lr=1e-4)      # torch.optim.Adam with a one-launch step (csrc/optim.cu)
        sync = None
        if world > 1:
            parallel.broadcast_parameters(net)
            sync = parallel.attach(net)
        score_sampling.set_ensemble_shard(first_member=rank * local, members_total=local * world)
        score_sampling.manual_seed(5)

        def step():
            opt.zero_grad(set_to_none=True)
            loss = loss_fn(net, x, marginal_prob_std_fn, y=y, cond_img=cond, lsm_cond=lsm, topo_cond=topo, sdf_cond=sdf)
            loss.backward()
            opt.step()
            return loss

        for _ in range(max(warmup, 4)):                       # steps 1-2 run eagerly, step 3 captures the two graphs
            step()
        exchange = "none (1 rank)"
        if sync is not None:                                   # what capture recorded: the graph replays exactly these NCCL calls
            exchange = dict(sync.last_plan or {})
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            loss = step()
        e.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(e) / steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        gb = local * world
        out[mode] = {"value": gb / (ms * 1e-3), "ms_per_step": ms, "global_batch": gb, "per_gpu_batch": local,
                     "algorithmic_tflops": 3 * FWD_FLOP_CIN7 * gb / (ms * 1e-3) / 1e12, "loss": float(loss), "grad_exchange": exchange}
        if world > 1:
            parallel.detach(net)
        net.__dict__.get("_train_runners", {}).clear()
        del net, opt
        torch.cuda.synchronize()
        torch.cuda.empty_cache()
    score_sampling.set_ensemble_shard(0, None, None)
    return out


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
    ss.clear_sampler_cache()
    ss.set_ensemble_shard(0, None, None)

    extra = {}
    if not args.no_extras:
        torch.cuda.empty_cache()
        for name, fn in (("pc_c3", lambda: bench_pc_c3(dev, rank, world, steps=2, warmup=1)),
                         ("train_c4", lambda: bench_train_c4(dev, rank, world, steps=10, warmup=4))):
            try:
                extra[name] = fn()
            except Exception as exc:                          # an extra leg must never take the headline line down
                if world > 1:
                    raise                                      # ...but ranks must not diverge around collectives
                extra[name] = {"error": f"{type(exc).__name__}: {str(exc)[:300]}"}
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
    if rank != 0:
        sys.stdout.flush()
        os._exit(0)        # captured graphs may still reference the NCCL communicator: skip the interpreter teardown

    peaks = load_peaks()
    fields = MEMBERS * world
    print(f"[bench] {ms:.1f} ms per {EM_STEPS}-step sampler call, e2e {ms_e2e:.1f} ms", file=sys.stderr)
    value = fields / (ms * 1e-3)
    flops, kms, kname = time_dominant_kernel(net, args.precision)
    achieved = flops / (kms * 1e-3) / 1e12
    fwd_tflops = MEMBERS * FWD_FLOP * EM_STEPS / (ms * 1e-3) / 1e12
    nprod = TENSOR_PRODUCTS[args.precision]
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
                     "frac": achieved / peaks["bf16_tflops"], "traffic": load_traffic(args.precision),
                     "kernel": kname + " on decoder.final_layer.conv_up 64->64 3x3 @128x128 x64",
                     "kernel_ms": kms, "algorithmic_flops_per_launch": flops, "peak_source": peaks["source"] + ", burst bf16",
                     "tensor_pipe_frac": nprod * achieved / peaks["bf16_tflops"],
                     "path_frac": fwd_tflops / peaks["bf16_tflops_sustained"],
                     "note": "achieved/frac count ALGORITHMIC FLOPs of the largest single kernel; this precision issues "
                             f"{nprod:.0f} tensor-core product(s) per algorithmic product, so the tensor pipe is busy tensor_pipe_frac of the "
                             "measured 16-bit peak.  path_frac = whole-path UNet forward TFLOP/s (every launch of an evaluation incl. the "
                             "bandwidth kernels) over the SUSTAINED bf16 peak; `family` = the time-dominant kernel family"},
        "unet_fwd_tflops": fwd_tflops, "unet_fwd_frac_of_sustained_bf16": fwd_tflops / peaks["bf16_tflops_sustained"],
    }
    if args.precision != "fp32":
        try:
            line["roofline"]["family"] = time_conv_family(net, peaks, args.precision)
        except Exception as exc:
            line["roofline"]["family"] = {"error": f"{type(exc).__name__}: {str(exc)[:300]}"}
    if extra:
        line["extra"] = extra
    if world == 1 and not args.no_cpu_baseline:
        try:
            line["gpu_eager_baseline"] = gpu_eager_baseline(dev)
        except Exception as exc:                              # a baseline leg must never take the bench line down
            line["gpu_eager_baseline"] = {"error": f"{type(exc).__name__}: {str(exc)[:300]}"}
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        ref = ReferenceEM("cpu")
        members, n_cpu = _cpu_sample_plan(ref, 15.0)          # ~15 s of CPU work
        v, dt = cpu_em_fields_per_s(ref, members, n_cpu)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": ref.kind,
                                "sample": f"{members} members x {n_cpu} EM steps ({dt:.1f} s), extrapolated linearly to 500 steps"}
    print(json.dumps(line))
    sys.stdout.flush()
    if world > 1:
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=DEFAULT_PRECISION, choices=["fp16x2", "bf16x3", "bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the cpu_baseline and gpu_eager_baseline legs")
    ap.add_argument("--no-extras", action="store_true", help="skip extra.pc_c3 / extra.train_c4")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
