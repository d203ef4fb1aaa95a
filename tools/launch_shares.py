#!/usr/bin/env python3
"""Aggregate an ncu launch list (gpu__time_duration.sum per launch) by kernel: launches, total ms, share."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0].replace("void ", "").replace("rb::", "")
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += float(r[-1]) / 1e6
tot = sum(v[1] for v in agg.values())
print("%-34s %8s %10s %7s" % ("kernel", "launches", "total ms", "share"))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-34s %8d %10.3f %6.1f%%" % (k, v[0], v[1], 100 * v[1] / tot))
