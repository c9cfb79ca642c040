#!/usr/bin/env python
"""One-line view of bench.py JSON lines: ms/step, stage times, roofline fraction, e2e.  usage: show.py <json> [...]"""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable:", e)
        continue
    st = {k: round(v, 1) for k, v in d.get("stage_ms_per_step", {}).items()}
    e2e = d.get("e2e") or {}
    print(f, "| %.1f ms/step | %.2f M reads/s | frac %.3f | e2e %.2f M |" % (d.get("ms_per_step", 0), d.get("value", 0) / 1e6,
          (d.get("roofline") or {}).get("frac", 0), (e2e.get("value") or 0) / 1e6), st)
