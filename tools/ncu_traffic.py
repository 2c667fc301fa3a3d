"""Extract the dominant kernel's DRAM traffic from an `ncu --set full` capture into profiles/<out>.json (read by bench.py's
`roofline.traffic`).  The capture is either the .ncu-rep itself or its `ncu -i ... --page raw --csv` export.

    python tools/ncu_traffic.py gpurun_out/prof.ncu-rep|raw.csv [kernel-name-substring] [out-name]
    e.g. python tools/ncu_traffic.py gpurun_out/r02_ncu_forward_fp16x2_raw.csv "conv3x3_c64_kernel<3, 6, 0, 1>" r02_dominant_kernel_ncu_fp16x2"""
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
    name = sys.argv[3] if len(sys.argv) > 3 else "r01_dominant_kernel_ncu"
    if rep.endswith(".csv"):
        out = open(rep).read()
    else:
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
    path = os.path.join(ROOT, "profiles", name + ".json")
    with open(path, "w") as f:
        json.dump(best, f, indent=1)
    print(json.dumps(best, indent=1))


if __name__ == "__main__":
    main()
