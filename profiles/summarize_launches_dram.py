"""Per-kernel device time AND DRAM traffic from an ncu launch list taken with
   --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv
Usage: python profiles/summarize_launches_dram.py gpurun_out/launches_r2_cfg3.csv"""
import collections
import csv
import re
import sys


def main():
    lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
    recs = collections.OrderedDict()
    for r in csv.DictReader(lines):
        d = recs.setdefault(r["ID"], {"name": re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("<unnamed>::", ""),
                                      "grid": r["Grid Size"]})
        v, u = float(r["Metric Value"].replace(",", "")), r["Metric Unit"]
        if r["Metric Name"].startswith("dram"):
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        else:
            v *= {"ns": 1, "us": 1e3, "ms": 1e6}.get(u, 1)
        d[r["Metric Name"]] = v
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for d in recs.values():
        a = agg[d["name"]]
        a[0] += 1; a[1] += d["gpu__time_duration.sum"]; a[2] += d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]
    tot = sum(a[1] for a in agg.values())
    print("launches %d, total device time %.3f ms (cold-cache, serialised by ncu: compare SHARES)" % (len(recs), tot / 1e6))
    print("%-36s %6s %11s %7s %10s %11s %10s" % ("kernel", "count", "total_us", "share", "mean_us", "dram_MB", "GB/s"))
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-36s %6d %11.1f %6.1f%% %10.1f %11.1f %10.1f" % (k[:36], a[0], a[1] / 1e3, 100 * a[1] / tot, a[1] / 1e3 / a[0], a[2] / 1e6, a[2] / a[1]))
    print("\nlargest single launches:")
    for d in sorted(recs.values(), key=lambda d: -d["gpu__time_duration.sum"])[:12]:
        b = d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]
        print("  %-28s %-16s %8.1f us  read %7.1f MB  write %7.1f MB  -> %6.0f GB/s" % (
            d["name"][:28], d["grid"], d["gpu__time_duration.sum"] / 1e3, d["dram__bytes_read.sum"] / 1e6, d["dram__bytes_write.sum"] / 1e6, b / d["gpu__time_duration.sum"]))


if __name__ == "__main__":
    main()
