#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file x.csv
<command>`): launches, total and average duration, share of the GPU time.  This is how profiles/*_launches_summary.txt are
made from profiles/*_launches.csv.

usage: python tools/summarize_launches.py profiles/r2_launches.csv > profiles/r2_launches_summary.txt
"""
import collections
import csv
import sys


def main():
    rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
    rd = csv.DictReader(rows)
    tot, cnt, unit = collections.Counter(), collections.Counter(), "ns"
    for r in rd:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        unit = r["Metric Unit"]
        k = r["Kernel Name"]
        tot[k] += float(r["Metric Value"].replace(",", ""))
        cnt[k] += 1
    whole = sum(tot.values())
    print("unit: %s" % unit)
    for k, t in tot.most_common():
        print("%5d launches %13.1f total %10.1f avg %6.1f %%  %s" % (cnt[k], t, t / cnt[k], 100.0 * t / whole, k[:100]))


if __name__ == "__main__":
    main()
