#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total and mean device time per kernel,
and every kernel's share of the profiled region.  usage: summarize_launches.py launches.csv [> summary.txt]"""
import collections
import csv
import sys


def main(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        a = agg.setdefault(r[ki].split("(")[0], [0, 0.0])
        a[0] += 1
        a[1] += v
    total = sum(t for _, t in agg.values())
    print("%-60s %8s %12s %12s %7s" % ("kernel", "launches", "total_ms", "mean_us", "share"))
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-60s %8d %12.3f %12.1f %6.1f%%" % (k[-60:], n, t / 1e6, t / n / 1e3, 100 * t / total))
    # the kernels of a step only (bench.py also launches its integer-peak micro-benchmarks and the L2 flush fill)
    step = {k: v for k, v in agg.items() if "me_" in k}
    st = sum(t for _, t in step.values())
    if st and st != total:
        print("\nshare of the step kernels only (me_*):")
        for k, (n, t) in sorted(step.items(), key=lambda kv: -kv[1][1]):
            print("%-60s %8d %12.3f %12.1f %6.1f%%" % (k[-60:], n, t / 1e6, t / n / 1e3, 100 * t / st))


if __name__ == "__main__":
    main(sys.argv[1])
