"""Extract the dominant kernel's DRAM traffic from an `ncu --set full` report into profiles/r01_dominant_kernel_ncu.json.

    python tools/ncu_traffic.py gpurun_out/prof_dominant.ncu-rep [kernel-name-substring]"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    rep = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else "conv3x3_c64_kernel"
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    best = None
    for r in rows[2:]:
        d = dict(zip(h, r))
        if want not in d["Kernel Name"]:
            continue
        val = lambda k: float(d[k].replace(",", "")) * UNIT.get(units[h.index(k)], 1.0)
        rec = {"kernel": d["Kernel Name"], "duration_us": float(d["gpu__time_duration.sum"].replace(",", "")),
               "dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
               "sm__pipe_tensor_cycles_active_pct_of_peak_elapsed": float(d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "nan") or "nan"),
               "source": os.path.basename(rep) + " (ncu --set full --clock-control none)"}
        if best is None or rec["duration_us"] > best["duration_us"]:
            best = rec
    if best is None:
        sys.exit(f"no kernel matching {want!r} in {rep}")
    path = os.path.join(ROOT, "profiles", "r01_dominant_kernel_ncu.json")
    with open(path, "w") as f:
        json.dump(best, f, indent=1)
    print(json.dumps(best, indent=1))


if __name__ == "__main__":
    main()
