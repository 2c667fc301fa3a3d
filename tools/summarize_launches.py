"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (shares, not absolutes)."""
import collections
import csv
import re
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg, n = collections.OrderedDict(), 0
    for r in rows[hi + 1:]:
        if len(r) <= iv:
            continue
        name = re.sub(r"\(.*", "", r[ik]).replace("void sbgm::", "").replace("sbgm::", "")
        v = float(r[iv].replace(",", ""))
        v = v / 1000 if r[iu] == "ns" else (v * 1000 if r[iu] == "ms" else v)
        agg.setdefault(name, [0.0, 0])
        agg[name][0] += v
        agg[name][1] += 1
        n += 1
    tot = sum(v[0] for v in agg.values())
    print(f"launches {n}, total {tot:.1f} us (cold-cache, serialised under ncu: compare SHARES, not absolutes)")
    for k, (v, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"{v:9.1f} us {100 * v / tot:5.1f}%  x{c:3d}  {k[:110]}")


if __name__ == "__main__":
    main(sys.argv[1])
