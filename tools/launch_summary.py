"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count, mean, share."""
import collections
import csv
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        ns = v if unit.startswith("n") else v * 1e3 if unit.startswith("u") else v * 1e6
        agg.setdefault(row["Kernel Name"][:70], []).append(ns)
    tot = sum(sum(v) for v in agg.values())
    print(f"{'kernel':72s} {'n':>5s} {'mean us':>10s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k:72s} {len(v):5d} {sum(v) / len(v) / 1e3:10.2f} {sum(v) / tot * 100:6.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
