"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total and average duration and share
per kernel.  Usage: python tools/ncu_launch_summary.py launches.csv"""
import csv
import sys
from collections import defaultdict


def main():
    rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
    idx = {h: i for i, h in enumerate(rows[0])}
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[1:]:
        if r[idx["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = r[idx["Kernel Name"]].split("(")[0]
        v = float(r[idx["Metric Value"]].replace(",", ""))
        unit = r[idx["Metric Unit"]]
        us = v / 1000.0 if unit in ("ns", "nsecond") else (v * 1000.0 if unit in ("ms", "msecond") else v)
        tot[name] += us
        cnt[name] += 1
    total = sum(tot.values())
    print("%-62s %8s %12s %8s %8s" % ("kernel", "launches", "total_us", "avg_us", "share"))
    for name in sorted(tot, key=lambda k: -tot[k]):
        print("%-62s %8d %12.1f %8.1f %7.1f%%" % (name[:62], cnt[name], tot[name], tot[name] / cnt[name], 100 * tot[name] / total))


if __name__ == "__main__":
    main()
