"""Probe: does a SWIZZLE_64B K-major A descriptor accept a start address that is an odd multiple of 64 B?"""
import os, sys, importlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.nn.functional as F
sys.path.insert(0, "tests")
from sbgm_danra_b200 import engine as E

def run(prec, probe):
    os.environ["SBGM_C64_PROBE"] = str(probe)
    fmt = {"bf16": 1, "bf16x3": 2}[prec]
    g = torch.Generator().manual_seed(1)
    n, h, w = 3, 32, 64
    x = torch.randn(n, 64, h, w, generator=g)
    w1 = torch.randn(64, 64, 3, 3, generator=g) * 0.05
    b1 = torch.randn(64, generator=g) * 0.1
    want = F.conv2d(x, w1, b1, padding=1)
    kern = E.Kernels(fmt, torch.device("cuda"))
    cw = E._Packer({"w": w1, "b": b1}, fmt, torch.device("cuda")).conv("w", "b")
    a = E.Act.from_nchw(x.cuda(), fmt)
    y = kern.conv(a, cw, pad=1).to_nchw().cpu()
    d = (y - want)
    cols = torch.arange(w)
    inner = (cols % 16) != 15
    e_in = float(d[..., inner].norm() / want[..., inner].norm())
    e_all = float(d.norm() / want.norm())
    print(f"{prec} probe={probe}: rel-L2 inner cols {e_in:.3e}, all cols {e_all:.3e}", flush=True)

for prec in ("bf16x3", "bf16"):
    for probe in (0, 1, 2):
        run(prec, probe)
