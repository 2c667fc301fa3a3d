#!/usr/bin/env python
"""Where stock torch eager spends the reference forward on the B200 (C2 shape: 64 x 128x128, Cin = 2), TF32 off and on.

Answers VERDICT r1 weak-9(iv): why TF32 buys the eager baseline only ~12 %.  The timed code is the REAL reference module from
baseline/_ref (bench.ReferenceEM; the oracle port if it did not travel) -- a baseline measurement, none of this repo's kernels.

    python tools/profile_torch_eager.py > profiles/r02_torch_eager_kernel_breakdown.txt
"""
import os
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
from torch.profiler import ProfilerActivity, profile

import bench


def classify(name: str) -> str:
    n = name.lower()
    if any(k in n for k in ("cudnn", "conv", "implicit_gemm", "xmma", "cutlass", "sgemm", "gemm", "wgrad", "dgrad", "nchwtonhwc", "nhwctonchw")):
        return "contraction (cuDNN conv / cuBLAS gemm incl. layout transposes)"
    if any(k in n for k in ("batch_norm", "group_norm", "layer_norm", "rowwisemoments", "welford")):
        return "normalisation"
    if "upsample" in n:
        return "bilinear upsample"
    if any(k in n for k in ("softmax", "attention", "fmha", "flash")):
        return "attention core"
    if "memcpy" in n or "memset" in n:
        return "memcpy / memset"
    return "elementwise / reduce / other ATen"


def main():
    dev = torch.device("cuda", 0)
    ref = bench.ReferenceEM(dev)
    x = torch.randn(bench.MEMBERS, 1, bench.SIZE, bench.SIZE, device=dev)
    t = torch.rand(bench.MEMBERS, device=dev) * 0.9 + 0.05
    torch.backends.cudnn.benchmark = True
    print(f"torch eager forward of the reference ScoreNet ({ref.kind}), 64 x 128x128, Cin = 2 -- torch {torch.__version__}, cuDNN {torch.backends.cudnn.version()}")
    for mode in ("fp32", "tf32"):
        torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = (mode == "tf32")
        with torch.no_grad():
            for _ in range(3):
                ref.forward(x, t)
            torch.cuda.synchronize()
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                ref.forward(x, t)
            e.record()
            e.synchronize()
            wall = a.elapsed_time(e) / 5
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                for _ in range(3):
                    ref.forward(x, t)
                torch.cuda.synchronize()
        per, cnt, cls = defaultdict(float), defaultdict(int), defaultdict(float)
        for ev in prof.events():
            if ev.device_type == torch.autograd.DeviceType.CUDA:
                per[ev.name] += ev.device_time / 3.0
                cnt[ev.name] += 1
                cls[classify(ev.name)] += ev.device_time / 3.0
        total = sum(per.values())
        print(f"\n== {mode}: {wall:.2f} ms per forward by CUDA events; sum of kernel durations {total / 1e3:.2f} ms over {sum(cnt.values()) // 3} launches "
              f"(GPU idle between launches: {max(wall - total / 1e3, 0.0):.2f} ms)")
        for k, v in sorted(cls.items(), key=lambda kv: -kv[1]):
            print(f"   {v / 1e3:8.3f} ms  {100 * v / total:5.1f}%  {k}")
        print("   top kernels:")
        for k, v in sorted(per.items(), key=lambda kv: -kv[1])[:14]:
            print(f"   {v:9.1f} us  x{cnt[k] // 3:<4d} {k[:130]}")


if __name__ == "__main__":
    main()
