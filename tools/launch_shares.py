#!/usr/bin/env python3
"""Per-kernel totals of an ncu launch list (gpu__time_duration.sum CSV).  usage: tools/launch_shares.py launches.csv"""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.OrderedDict()
for x in csv.DictReader(lines):
    name = x["Kernel Name"].split("(")[0].replace("void ", "").replace("nemb200::", "")
    agg.setdefault(name, []).append(float(x["Metric Value"].replace(",", "")) / 1000.0)
tot = sum(sum(v) for v in agg.values())
print(f"{'kernel':70s} {'n':>5s} {'mean_us':>9s} {'sum_us':>10s} share")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k[:70]:70s} {len(v):5d} {sum(v)/len(v):9.2f} {sum(v):10.1f} {sum(v)/tot:.3f}")
print("total us", tot)
