#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv, sys, re, collections
rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
agg = collections.OrderedDict()
tot = 0.0
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    if unit in ("us", "usecond"): v *= 1e3
    if unit in ("ms", "msecond"): v *= 1e6
    short = re.sub(r"\(.*", "", name)
    # keep template args compact
    short = short[:140]
    a = agg.setdefault(short, [0, 0.0])
    a[0] += 1; a[1] += v; tot += v
print("total %.3f ms over %d launches" % (tot / 1e6, sum(a[0] for a in agg.values())))
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%6.2f%%  %9.3f ms  n=%5d  avg %8.2f us  %s" % (100 * t / tot, t / 1e6, n, t / n / 1e3, k))
