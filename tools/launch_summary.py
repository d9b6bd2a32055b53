#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches and mean time per kernel."""
import collections
import csv
import sys


def main(path):
    hdr, agg = None, collections.OrderedDict()
    for r in csv.reader(open(path)):
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            try:
                v = float(d["Metric Value"].replace(",", ""))
            except ValueError:
                continue
            scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(d.get("Metric Unit", "ns"), 1e-3)
            agg.setdefault(d["Kernel Name"], []).append(v * scale)
    total = sum(sum(v) for v in agg.values())
    for k, v in agg.items():
        print("%-90s n=%4d mean=%9.1f us share=%5.1f%%" % (k[:90], len(v), sum(v) / len(v), 100 * sum(v) / total))


if __name__ == "__main__":
    main(sys.argv[1])
