"""Golden vectors of the reference's host->device boundary: imports /root/reference (this container only) and stores
(i) the float32 outputs of its forward transform classes (sbgm/special_transforms.py: Scale, ZScoreTransform,
PrcpLogTransform) for the cases of oracle/batch_ref.py and (ii) the 9-tuple of its `extract_samples`
(sbgm/utils.py:405-480, device='cpu') on oracle.batch_ref.sample_dict().

    python tests/golden/make_batch_golden.py"""
import os
import sys
from unittest import mock

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
for name in ("zarr", "netCDF4", "matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.gridspec", "matplotlib.patches",
             "matplotlib.dates", "matplotlib.ticker", "matplotlib.cm", "mpl_toolkits", "mpl_toolkits.axes_grid1", "omegaconf", "optuna",
             "cartopy", "seaborn", "cmocean"):
    m = mock.MagicMock(name=name)
    m.__path__, m.__spec__ = [], None
    sys.modules.setdefault(name, m)
sys.path.insert(0, "/root/reference")
from oracle import batch_ref as br               # noqa: E402
from sbgm import special_transforms as ref       # noqa: E402
from sbgm.utils import extract_samples           # noqa: E402

NAMES = ("hr", "classifier", "lr", "lsm_hr", "lsm", "sdf", "topo", "hr_point", "lr_point")


def main():
    out = {}
    for name, (kind, kw, inp) in br.FWD_CASES.items():
        x = br.fwd_case_input(inp)
        t = torch.from_numpy(x)
        if kind == "zscore":
            y = ref.ZScoreTransform(kw["mean"], kw["std"])(t)
        elif kind == "scale":
            y = ref.Scale(kw["in_low"], kw["in_high"], kw["data_min_in"], kw["data_max_in"])(t)
        else:
            y = ref.PrcpLogTransform(**kw)(t)
        out[f"fwd/{name}"] = y.numpy().astype(np.float32)
        print(name, float(y.mean()))
    for two in (True, False):
        res = extract_samples(br.sample_dict(two_lr=two), device=torch.device("cpu"))
        for nm, v in zip(NAMES, res):
            out[f"extract{int(two)}/{nm}"] = np.zeros(0, np.float32) if v is None else v.numpy()
    np.savez_compressed(os.path.join(HERE, "batch_golden.npz"), **out)


if __name__ == "__main__":
    main()
