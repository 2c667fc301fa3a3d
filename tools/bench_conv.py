"""Time individual convolution launches (CUDA events, L2 flushed) at the benchmark's shapes.

    python tools/bench_conv.py [--precision bf16x3] [--only NAME] [--iters 10]
"""
import argparse
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sbgm_danra_b200 import engine as E

SHAPES = {
    # name: (n, h, w, cin, cout, k, stride, pad, mode)
    "d4_conv_up_proj": (64, 128, 128, 64, 64, 3, 1, 1, "proj"),
    "d4_conv_up_store": (64, 128, 128, 64, 64, 3, 1, 1, "plain"),
    "d3_conv_gn": (64, 64, 64, 64, 64, 3, 1, 1, "gn"),
    "e3_block": (64, 32, 32, 64, 64, 3, 1, 1, "plain"),
    "conv2_8x8": (64, 64, 64, 64, 64, 8, 2, 3, "plain"),
    "d2_conv_up": (64, 32, 32, 128, 128, 3, 1, 1, "plain"),
    "d1_conv_up": (64, 16, 16, 256, 256, 3, 1, 1, "plain"),
    "d0_conv_up": (64, 8, 8, 512, 512, 3, 1, 1, "plain"),
    "e6_block": (64, 4, 4, 512, 512, 3, 1, 1, "plain"),
    "attn_ff_256": (1, 1, 4096, 256, 256, 1, 1, 0, "plain"),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="bf16x3")
    ap.add_argument("--only", default=None)
    ap.add_argument("--iters", type=int, default=10)
    a = ap.parse_args()
    fmt = E.PRECISIONS[a.precision]
    dev = torch.device("cuda")
    k = E.Kernels(fmt, dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for name, (n, h, w, cin, cout, ks, stride, pad, mode) in SHAPES.items():
        if a.only and a.only not in name:
            continue
        g = torch.Generator().manual_seed(0)
        wt = torch.randn(cout, cin, ks, ks, generator=g) * 0.05
        cw = E._Packer({"w": wt, "b": torch.zeros(cout)}, fmt, dev).conv("w", "b")
        x = E.Act(fmt, n, h, w, cin, dev)
        x.buf.normal_()
        pw = torch.randn(9, 64, device=dev) if mode == "proj" else None
        kw = dict(stride=stride, pad=pad)
        if mode == "proj":
            kw["proj"] = pw
        if mode == "gn":
            kw["gn_stats"] = True
        for _ in range(3):
            k.conv(x, cw, **kw)
        ts = []
        for _ in range(a.iters):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            k.conv(x, cw, **kw)
            e.record()
            e.synchronize()
            ts.append(s.elapsed_time(e))
        ms = statistics.median(ts)
        ho, wo = (h + 2 * pad - ks) // stride + 1, (w + 2 * pad - ks) // stride + 1
        fl = 2.0 * n * ho * wo * cout * cin * ks * ks
        print(f"{name:18s} {ms * 1e3:8.1f} us  {fl / ms / 1e9:7.1f} TFLOP/s (algorithmic)  min {min(ts) * 1e3:.1f} us")


if __name__ == "__main__":
    main()
