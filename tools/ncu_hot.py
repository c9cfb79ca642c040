#!/usr/bin/env python
"""Hottest SASS instructions of a profiled kernel (stall samples + executed count). usage: ncu_hot.py rep [topN]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hi]
c = {n: h.index(n) for n in ("Address", "Source", "# Samples", "Instructions Executed", "Avg. Threads Executed", "stall_long_sb", "stall_short_sb", "stall_wait", "stall_mio", "stall_barrier", "stall_lg", "stall_branch_resolving")}
data = []
for r in rows[hi + 1:]:
    if len(r) <= c["# Samples"]: continue
    try: data.append((int(r[c["# Samples"]]), r))
    except ValueError: pass
tot = sum(d[0] for d in data)
print("total samples", tot, "instructions", len(data))
for i, (s, r) in enumerate(data):
    r.append(i)
for s, r in sorted(data, key=lambda x: -x[0])[:top]:
    print("%5.1f%% #%4d exec=%-10s thr=%-5s lsb=%-5s ssb=%-5s wait=%-5s mio=%-4s bar=%-4s lg=%-4s br=%-4s %s" % (
        100.0 * s / tot, r[-1], r[c["Instructions Executed"]], r[c["Avg. Threads Executed"]][:5], r[c["stall_long_sb"]], r[c["stall_short_sb"]],
        r[c["stall_wait"]], r[c["stall_mio"]], r[c["stall_barrier"]], r[c["stall_lg"]], r[c["stall_branch_resolving"]], r[c["Source"]][:70]))
