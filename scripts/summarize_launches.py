#!/usr/bin/env python
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections
import csv
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    us = v / 1000.0 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1000.0)
    name = r[ki].split("(")[0]
    a = agg.setdefault(name, [0.0, 0])
    a[0] += us
    a[1] += 1
tot = sum(a[0] for a in agg.values())
print("# total %.1f us over %d launches" % (tot, sum(a[1] for a in agg.values())))
for k, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("%12.1f us %6.2f%%  n=%4d avg=%9.2f us  %s" % (t, 100 * t / tot, n, t / n, k))
