#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (count, total, share, mean).

  python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/launches_rNN.txt
"""
import collections
import csv
import sys


def main(path):
    hdr, agg = None, collections.defaultdict(lambda: [0, 0.0])
    for r in csv.reader(open(path)):
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(d["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(d["Metric Unit"], 1.0)
        a = agg[d["Kernel Name"].split("(")[0]]
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot / 1e3:.2f} ms summed device time "
          "(ncu serialised, cold-cache: compare shares, not absolutes)")
    print(f"{'kernel':48s} {'launches':>8s} {'total_ms':>10s} {'share':>7s} {'mean_us':>10s}")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{k:48s} {v[0]:8d} {v[1] / 1e3:10.2f} {100 * v[1] / tot:6.1f}% {v[1] / v[0]:10.1f}")


if __name__ == "__main__":
    main(sys.argv[1])
