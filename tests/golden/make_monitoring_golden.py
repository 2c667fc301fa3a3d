"""Pins oracle/monitoring_ref.py to the REAL reference: imports sbgm/utils.py from /root/reference (with inert stand-ins for
the plotting / file-format packages this image lacks), runs its `report_precip_extremes` on the seeded cases of
oracle.monitoring_ref.cases() and writes tests/golden/monitoring_golden.json.  Run here (needs /root/reference):

    python tests/golden/make_monitoring_golden.py
"""
import json
import os
import sys
from unittest import mock

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"
for name in ["zarr", "netCDF4", "matplotlib", "matplotlib.pyplot", "matplotlib.gridspec", "matplotlib.colors", "matplotlib.patches",
             "matplotlib.dates", "matplotlib.ticker", "matplotlib.cm", "mpl_toolkits", "mpl_toolkits.axes_grid1", "omegaconf",
             "cartopy", "seaborn", "cmocean"]:
    m = mock.MagicMock(name=name)
    m.__path__, m.__spec__ = [], None
    sys.modules[name] = m
sys.path.insert(0, REF)
import sbgm.utils as ref_utils  # noqa: E402

from oracle import monitoring_ref  # noqa: E402

out = {}
for name, x in monitoring_ref.cases().items():
    for cap in (500.0, 50.0):
        msgs = []
        out[f"{name}/cap{cap:g}"] = {"result": ref_utils.report_precip_extremes(x, name=name, cap_mm_day=cap, logger=msgs.append),
                                     "messages": msgs}
with open(os.path.join(ROOT, "tests", "golden", "monitoring_golden.json"), "w") as f:
    json.dump(out, f, indent=1)
print("wrote", len(out), "cases")
