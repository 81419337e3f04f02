"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
tot = defaultdict(lambda: [0, 0.0])
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"].split("(")[0]
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    if unit in ("us", "usecond"):
        v *= 1e3
    elif unit in ("ms", "msecond"):
        v *= 1e6
    tot[name][0] += 1
    tot[name][1] += v
total = sum(v[1] for v in tot.values())
print(f"{'kernel':60s} {'launches':>8s} {'total_us':>10s} {'avg_us':>8s} {'share':>6s}")
for name, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{name[:60]:60s} {n:8d} {t/1e3:10.1f} {t/1e3/n:8.1f} {100*t/total:5.1f}%")
print(f"{'TOTAL':60s} {sum(v[0] for v in tot.values()):8d} {total/1e3:10.1f}")
