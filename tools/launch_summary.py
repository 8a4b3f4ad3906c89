#!/usr/bin/env python
"""Aggregates an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: python tools/launch_summary.py <launches.csv>"""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    a = agg.setdefault(r[4], [0, 0.0]); a[0] += 1; a[1] += float(r[-1]) / 1e6
tot = sum(t for _, t in agg.values())
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:70]:70s} n={n:3d} sum={t:9.3f} ms share={t / tot:6.3f} avg={t / n * 1e3:9.1f} us")
