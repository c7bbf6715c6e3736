"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name (share of the step)."""
import csv
import re
import sys
from collections import defaultdict


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    tot = defaultdict(lambda: [0, 0.0])
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = re.sub(r"^void |pttspp::|\(anonymous namespace\)::", "", name)
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        tot[name][0] += 1
        tot[name][1] += us
    total = sum(v[1] for v in tot.values())
    print(f"# {path}: {sum(v[0] for v in tot.values())} launches, {total / 1e3:.2f} ms summed device time")
    print(f"{'kernel':70s} {'launches':>8s} {'ms':>10s} {'avg_us':>9s} {'share':>7s}")
    for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:70]:70s} {n:8d} {us / 1e3:10.3f} {us / n:9.2f} {us / total:7.1%}")


if __name__ == "__main__":
    main(sys.argv[1])
