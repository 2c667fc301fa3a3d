"""Golden vectors of the reference's inverse transforms: imports /root/reference/sbgm/special_transforms.py (this
container only) and stores its float32 outputs for the cases of oracle/transforms_ref.py.

    python tests/golden/make_transforms_golden.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, "/root/reference")
from oracle import transforms_ref as tr          # noqa: E402
from sbgm import special_transforms as ref       # noqa: E402


def main():
    x = tr.case_input()
    out = {"x": x}
    for name, (kind, kw) in tr.CASES.items():
        t = torch.from_numpy(x)
        if kind == "zscore":
            y = ref.ZScoreBackTransform(kw["mean"], kw["std"])(t)
        elif kind == "scale":
            y = ref.ScaleBackTransform(kw["in_low"], kw["in_high"], kw["data_min"], kw["data_max"])(t)
        else:
            y = ref.PrcpLogBackTransform(**kw)(t)
        out[name] = y.numpy().astype(np.float32)
        print(name, float(y.mean()))
    np.savez_compressed(os.path.join(HERE, "transforms_golden.npz"), **out)


if __name__ == "__main__":
    main()
