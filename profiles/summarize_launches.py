"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launch count, total and share."""
import csv
import re
import sys
from collections import defaultdict


def main(path):
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.DictReader(lines)
    tot = defaultdict(lambda: [0, 0.0])
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        tot[name][0] += 1
        tot[name][1] += us
    total = sum(v[1] for v in tot.values())
    print(f"# {path}: {sum(v[0] for v in tot.values())} launches, {total / 1e3:.3f} ms total (cold-cache, serialised: compare shares)")
    print(f"{'kernel':60s} {'launches':>8s} {'total_us':>12s} {'share':>7s}")
    for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:60]:60s} {n:8d} {us:12.1f} {100 * us / total:6.2f}%")


if __name__ == "__main__":
    main(sys.argv[1])
