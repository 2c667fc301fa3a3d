"""Per-launch table of the memory-bound kernels in an `ncu --set full` raw-CSV export: duration, DRAM bytes, achieved GB/s against
the measured HBM peak (MEASURED_PEAKS.json), L2 hit rate, achieved occupancy.

    python tools/ncu_bandwidth_table.py gpurun_out/raw.csv > profiles/<name>.txt"""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}


def main(path):
    peak = 6547.8
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        peak = float(d.get("hbm_gbps", d.get("hbm_gb_s", peak))) if isinstance(d, dict) else peak
    rows = list(csv.reader(open(path)))
    h, units = rows[0], rows[1]
    col = {k: i for i, k in enumerate(h)}

    def val(r, k, default=float("nan")):
        if k not in col or r[col[k]] in ("", "n/a"):
            return default
        return float(r[col[k]].replace(",", "")) * UNIT.get(units[col[k]], 1.0)

    print(f"{os.path.basename(path)}: ncu --set full --clock-control none (cold caches, serialised launches); HBM peak {peak:.0f} GB/s (measured)")
    print(f"{'#':>3} {'dur us':>8} {'rd MB':>8} {'wr MB':>8} {'GB/s':>7} {'of peak':>8} {'L2 hit%':>8} {'occ%':>6} {'regs':>5}  kernel")
    for i, r in enumerate(rows[2:]):
        name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void sbgm::", "").replace("sbgm::", "").replace("void ", "")
        d = val(r, "gpu__time_duration.sum")
        rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
        gbs = (rd + wr) / (d * 1e-6) / 1e9
        print(f"{i:3d} {d:8.1f} {rd / 1e6:8.2f} {wr / 1e6:8.2f} {gbs:7.0f} {gbs / peak:8.2f} {val(r, 'lts__t_sector_hit_rate.pct'):8.1f} "
              f"{val(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):6.1f} {val(r, 'launch__registers_per_thread'):5.0f}  {name[:70]}")


if __name__ == "__main__":
    main(sys.argv[1])
