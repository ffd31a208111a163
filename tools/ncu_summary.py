"""Summarise ncu output for profiles/:
   python tools/ncu_summary.py launches <launches.csv>
   python tools/ncu_summary.py raw <report.ncu-rep>
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
        "sm__cycles_active.avg"]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(row["Metric Unit"], 1e-6)
        name = row["Kernel Name"].split("(")[0]
        agg[name][0] += 1
        agg[name][1] += v * scale
    tot = sum(v[1] for v in agg.values())
    print("%-40s %6s %12s %10s %7s" % ("kernel", "n", "total_ms", "mean_ms", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-40s %6d %12.3f %10.4f %6.1f%%" % (k, v[0], v[1], v[1] / v[0], 100 * v[1] / tot))


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("== %s  (launch id %s)" % (r[hdr.index("Kernel Name")].split("(")[0], r[0]))
        for k in KEYS:
            if k in hdr:
                print("   %-62s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2])
