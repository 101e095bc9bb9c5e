#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count,
total device time and share.  usage: summarize_launches.py launches.csv [skip_first_n]"""
import collections
import csv
import sys


def main(path, skip=0):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    n = 0
    for row in csv.DictReader(lines):
        n += 1
        if n <= skip:
            continue
        k = row["Kernel Name"].split("(")[0][:78]
        v = float(row["Metric Value"].replace(",", ""))
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(row["Metric Unit"], 1e-3)
        agg.setdefault(k, [0, 0.0])
        agg[k][0] += 1
        agg[k][1] += v * scale
    tot = sum(v[1] for v in agg.values())
    print("%5s %12s %7s  %s" % ("count", "total_us", "share", "kernel"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%5d %12.1f %6.1f%%  %s" % (v[0], v[1], 100 * v[1] / tot, k))
    print("total %.1f us over %d launches" % (tot, sum(v[0] for v in agg.values())))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
