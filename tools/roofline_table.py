#!/usr/bin/env python
"""Speed-of-light budget of ONE score-UNet evaluation, layer by layer, from the network's shapes alone (no GPU needed).

For every operator of the forward (sbgm/score_unet.py Encoder :247-364, DecoderBlock :559-627) at a BASELINE config:
algorithmic FLOPs (2 x MAC), algorithmic HBM bytes in the engine's storage format (read the input once, write the output
once; weights once), and the time each bound allows at the MEASURED peaks of this pool's B200s (MEASURED_PEAKS.json:
sustained bf16 tensor throughput, HBM copy bandwidth).  The sum is the floor the measured evaluation time is compared with
in DESIGN.md: how far the whole path -- not just the dominant kernel -- is from the machine.

    python tools/roofline_table.py [--size 128] [--cin 2] [--batch 64] [--precision bf16x3|bf16|fp16w2] [--measured-ms 1.63] [--train]
"""
from __future__ import annotations

import argparse
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def layers(size: int, cin: int, bytes_per: int = 4):
    """(name, kind, flops per sample, elements read, elements written, weight elements) per operator.
    kind: "tc" tensor-core contraction, "bw" bandwidth operator (normalisation, upsample, LayerNorm, softmax core)."""
    out = []

    def conv(name, ci, co, k, h_out, stride=1, h_in=None):
        h_in = h_in if h_in is not None else h_out * stride
        out.append((name, "tc", 2 * ci * co * k * k * h_out * h_out, ci * h_in * h_in, co * h_out * h_out, ci * co * k * k))

    def bw(name, c, h_in, h_out, reads=1):
        out.append((name, "bw", 0, reads * c * h_in * h_in, c * h_out * h_out, 0))

    def attn(name, c, h):
        s = h * h
        bw(f"{name}.ln1", c, h, h)
        out.append((f"{name}.qkv", "tc", 2 * s * c * 3 * c, s * c, s * 3 * c, 3 * c * c))
        out.append((f"{name}.core", "bw", 4 * s * s * c, s * 3 * c, s * c, 0))
        out.append((f"{name}.out_proj", "tc", 2 * s * c * c, 2 * s * c, s * c, c * c))
        bw(f"{name}.ln2", c, h, h)
        out.append((f"{name}.ff0", "tc", 2 * s * c * c, s * c, s * c, c * c))
        out.append((f"{name}.ff2", "tc", 2 * s * c * c, 2 * s * c, s * c, c * c))

    h = size // 2
    conv("encoder.conv1 (8x8 s2)", cin, 64, 8, h, 2)
    h //= 2
    conv("encoder.conv2 (8x8 s2)", 64, 64, 8, h, 2)
    ci = 64
    for li, co in enumerate((64, 128, 256, 512), start=1):
        stride = 1 if li == 1 else 2
        hin, h = h, h // stride
        for b in range(2):
            conv(f"layer{li}.{b}.conv1", ci if b == 0 else co, co, 3, h, stride if b == 0 else 1, hin if b == 0 else h)
            conv(f"layer{li}.{b}.conv2", co, co, 3, h)
            if b == 0 and (stride != 1 or ci != co):
                conv(f"layer{li}.0.downsample", ci, co, 1, h, stride, hin)
        ci = co
        if li >= 3:
            attn(f"encoder.attn{li}", co, h)
    plan = ((512, 256, True), (256, 128, True), (128, 64, False), (64, 64, False))
    for i, (ci, co, at) in enumerate(plan):
        bw(f"dec{i}.upsample", ci, h, 2 * h)
        h *= 2
        conv(f"dec{i}.conv_up", ci, ci, 3, h)
        bw(f"dec{i}.norm1+act", ci, h, h)
        conv(f"dec{i}.conv", ci, co, 3, h)
        bw(f"dec{i}.norm2+skip+time+act", co, h, h, reads=2)
        if at:
            attn(f"dec{i}.attn", co, h)
    bw("final.upsample", 64, h, 2 * h)
    h *= 2
    # projection epilogue: conv_up stores 9 tap-wise dot products + padding = 12 fp32 (48 B) per pixel instead of 64 channels
    proj = 12 * h * h * 4 // bytes_per
    out.append(("final.conv_up (+ projected 64->1 conv)", "tc", 2 * 64 * 64 * 9 * h * h, 64 * h * h, proj, 64 * 64 * 9))
    out.append(("final.conv 64->1 (gather)", "bw", 2 * 64 * 9 * h * h, proj, h * h * 4 // bytes_per, 0))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--cin", type=int, default=2)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--precision", default="bf16x3", choices=["bf16x3", "bf16", "fp16w2"])
    ap.add_argument("--measured-ms", type=float, default=None)
    ap.add_argument("--train", action="store_true",
                    help="DSM training step: forward + data gradient + weight gradient = 3x the contraction FLOPs; every\n"
                         "operator's backward re-reads its input and the incoming gradient and writes one gradient: ~3x the bytes")
    args = ap.parse_args()
    peaks = {"bf16_tflops_sustained": 1399.7, "hbm_gbs": 6547.8}
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            peaks.update({k: v for k, v in json.load(f).items() if k in peaks})
    products, bytes_per = {"bf16x3": (3, 4), "bf16": (1, 2), "fp16w2": (2, 2)}[args.precision]   # fp16w2: DESIGN "Next" item 1
    tot = {"flops": 0.0, "t_tc": 0.0, "t_bw": 0.0, "floor": 0.0, "bytes": 0.0}
    print(f"{'operator':44s} {'GFLOP':>8s} {'MB':>8s} {'tensor us':>10s} {'HBM us':>8s}  bound")
    for name, kind, fl, rd, wr, wt in layers(args.size, args.cin, bytes_per):
        flops = fl * args.batch * (3 if args.train else 1)
        nbytes = ((rd + wr) * args.batch * bytes_per + wt * bytes_per) * (3 if args.train else 1)
        t_tc = (flops * products / (peaks["bf16_tflops_sustained"] * 1e12) * 1e6) if kind == "tc" else 0.0
        t_bw = nbytes / (peaks["hbm_gbs"] * 1e9) * 1e6
        bound = "tensor" if t_tc > t_bw else "hbm"
        tot["flops"] += flops; tot["t_tc"] += t_tc; tot["t_bw"] += t_bw; tot["floor"] += max(t_tc, t_bw); tot["bytes"] += nbytes
        print(f"{name:44s} {flops / 1e9:8.2f} {nbytes / 1e6:8.1f} {t_tc:10.1f} {t_bw:8.1f}  {bound}")
    print(f"{'TOTAL':44s} {tot['flops'] / 1e9:8.2f} {tot['bytes'] / 1e6:8.1f} {tot['t_tc']:10.1f} {tot['t_bw']:8.1f}")
    print(f"algorithmic FLOPs per sample: {tot['flops'] / args.batch / 1e9:.3f} G (SURVEY section 2.2: 5.146 G at 128x128, Cin 2)")
    print(f"floor = sum over operators of max(tensor, HBM) = {tot['floor'] / 1e3:.3f} ms per {'training step (fwd + bwd, optimizer excluded)' if args.train else 'evaluation'} "
          f"({args.precision}: {products} tensor product(s) per algorithmic product, {bytes_per} B per activation element; "
          f"peaks {peaks['bf16_tflops_sustained']:.0f} TFLOP/s sustained bf16, {peaks['hbm_gbs']:.0f} GB/s)")
    if args.measured_ms:
        print(f"measured {args.measured_ms:.3f} ms per {'step' if args.train else 'evaluation'} -> {tot['floor'] / 1e3 / args.measured_ms:.0%} of the floor's speed")


if __name__ == "__main__":
    main()
