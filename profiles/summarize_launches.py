"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total and mean
device time, share.  Usage: python profiles/summarize_launches.py gpurun_out/launches.csv [skip_first_n]"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rows = []
    with open(path, newline="") as fh:
        lines = [l for l in fh if l.startswith('"')]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        rows.append((r["Kernel Name"], float(r["Metric Value"]), r["Grid Size"], r["Block Size"]))
    rows = rows[skip:]
    agg = OrderedDict()
    for name, ns, grid, block in rows:
        short = re.sub(r"\(.*", "", name)
        short = re.sub(r"^void ", "", short)
        a = agg.setdefault(short, [0, 0.0, 1e30, 0.0])
        a[0] += 1; a[1] += ns; a[2] = min(a[2], ns); a[3] = max(a[3], ns)
    total = sum(a[1] for a in agg.values())
    print("launches %d, total device time %.3f ms (cold-cache, serialised by ncu: compare SHARES)" % (len(rows), total / 1e6))
    print("%-70s %7s %11s %9s %9s %9s %6s" % ("kernel", "count", "total_us", "mean_us", "min_us", "max_us", "share"))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-70s %7d %11.1f %9.2f %9.2f %9.2f %5.1f%%" % (k[:70], a[0], a[1] / 1e3, a[1] / a[0] / 1e3, a[2] / 1e3, a[3] / 1e3, 100 * a[1] / total))


if __name__ == "__main__":
    main()
