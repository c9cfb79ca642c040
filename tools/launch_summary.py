#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (csv of `--metrics gpu__time_duration.sum`).  usage: launch_summary.py <csv> [skip_first_n_launches]"""
import collections
import csv
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
h = rows[0]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1 + skip:]:
    ms = float(r[vi].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}[r[ui]]
    name = r[ki].split("(")[0].replace("void ", "").replace("kasa::", "").replace("<unnamed>::", "")[:60]
    tot[name] += ms
    cnt[name] += 1
total = sum(tot.values())
print("%-62s %5s %12s %7s %10s" % ("kernel", "n", "total ms", "share", "avg ms"))
for k, v in tot.most_common():
    print("%-62s %5d %12.3f %6.1f%% %10.3f" % (k, cnt[k], v, 100 * v / total, v / cnt[k]))
print("%-62s %5d %12.3f" % ("all", sum(cnt.values()), total))
