"""Per-launch table of an `ncu --set full` report: duration, tensor-pipe activity, DRAM bytes, achieved occupancy.

    python tools/ncu_kernel_table.py gpurun_out/prof.ncu-rep > profiles/<name>.txt"""
import csv
import re
import subprocess
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}


def main(rep):
    if rep.endswith(".csv"):       # already exported on the GPU box (`ncu -i x.ncu-rep --page raw --csv`): full reports exceed the 64 MiB return limit
        out = open(rep).read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    col = {k: i for i, k in enumerate(h)}

    def val(r, k, default=float("nan")):
        if k not in col or r[col[k]] == "":
            return default
        return float(r[col[k]].replace(",", "")) * UNIT.get(units[col[k]], 1.0)

    print(f"{rep}: ncu --set full --clock-control none (cold caches, serialised launches: read shares and percentages, not absolutes)")
    print(f"{'#':>3} {'dur us':>8} {'tensor%':>8} {'memtens%':>8} {'dram rd MB':>10} {'dram wr MB':>10} {'dram %':>7} {'regs':>5} {'grid':>7}  kernel")
    tot = 0.0
    for i, r in enumerate(rows[2:]):
        name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("void sbgm::", "").replace("sbgm::", "")
        d = val(r, "gpu__time_duration.sum")
        tot += d
        grid = r[col["launch__grid_size"]] if "launch__grid_size" in col else "?"
        print(f"{i:3d} {d:8.1f} {val(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'):8.1f} "
              f"{val(r, 'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'):8.1f} "
              f"{val(r, 'dram__bytes_read.sum') / 1e6:10.2f} {val(r, 'dram__bytes_write.sum') / 1e6:10.2f} "
              f"{val(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):7.1f} {val(r, 'launch__registers_per_thread'):5.0f} {grid:>7}  {name[:90]}")
    print(f"total {tot:.1f} us over {len(rows) - 2} launches")


if __name__ == "__main__":
    main(sys.argv[1])
