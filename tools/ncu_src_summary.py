"""Per-instruction stall samples of an ncu report (--page source --csv), grouped so the warp roles can be
told apart: prints the top sampled instructions and totals per opcode."""
import collections
import csv
import subprocess
import sys


def main(rep, top=25):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = [i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r][0]
    hdr = rows[hi]
    i_src, i_s, i_ex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    data = []
    for r in rows[hi + 1:]:
        try:
            data.append((int(r[i_s]), int(r[i_ex]), r[i_src]))
        except (ValueError, IndexError):
            pass
    tot = sum(d[0] for d in data)
    print(rows[0][1] if len(rows[0]) > 1 else "")
    print(f"total samples {tot}, static instructions {len(data)}, executed warp-instructions {sum(d[1] for d in data)}")
    for s, e, src in sorted(data, key=lambda d: -d[0])[:top]:
        print(f"{s:7d} {100 * s / max(tot, 1):5.1f}%  exec={e:9d}  {src[:100]}")
    agg, ex = collections.Counter(), collections.Counter()
    for s, e, src in data:
        parts = src.split()
        op = (parts[1] if parts and parts[0].startswith("@") else parts[0] if parts else "?").split(".")[0]
        agg[op] += s
        ex[op] += e
    print("-- by opcode")
    for op, s in agg.most_common(14):
        print(f"{op:10s} {100 * s / max(tot, 1):5.1f}%  executed {ex[op]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
