"""The 64 -> 64 convolution with the bilinear upsample in its operand stage (sbgm_conv3x3_c64_up) at the final layer's C2 shape,
next to the two launches it replaces: CUDA-event times of both (for ncu: `ncu -k regex:conv3x3_c64 ...` captures the kernels).

    python tools/run_up_kernel.py [--precision fp16x2] [--members 64] [--low 64] [--iters 20] [--proj 1]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="fp16x2")
    ap.add_argument("--members", type=int, default=64)
    ap.add_argument("--low", type=int, default=64)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--proj", type=int, default=1)
    a = ap.parse_args()
    from sbgm_danra_b200 import engine as E
    dev = torch.device("cuda:0")
    fmt = E.PRECISIONS[a.precision]
    k = E.Kernels(fmt, dev)
    g = torch.Generator().manual_seed(0)
    r = lambda *s, scale=1.0: (torch.randn(*s, generator=g) * scale)
    n, hl = a.members, a.low
    pk = E._Packer({"w1": r(64, 64, 3, 3, scale=1 / 24), "b1": r(64, scale=0.1)}, fmt, dev)
    cw1 = pk.conv("w1", "b1")
    y = E.Act.from_nchw(r(n, 64, hl, hl).to(dev), fmt)
    pw = r(9, 64, scale=0.1).to(dev) if a.proj else None

    def two():
        return k.conv(k.upsample2x(y), cw1, pad=1, proj=pw)

    def one():
        return k.conv_up_fused(y, cw1, proj=pw)

    for name, fn in (("two launches", two), ("fused", one)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        print(f"{name:15s} {1e3 * e0.elapsed_time(e1) / a.iters:8.1f} us per call ({a.precision}, {n} x {2 * hl} x {2 * hl})")


if __name__ == "__main__":
    main()
