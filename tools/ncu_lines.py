#!/usr/bin/env python
"""Per-source-line cost of one profiled kernel: joins ncu's per-SASS-instruction metrics with nvdisasm line info.
usage: ncu_lines.py <rep> <object.o> <kernel substring> [topN]"""
import csv, os, re, subprocess, sys, tempfile, collections
rep, obj, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
lines, cur, inside = [], None, False
for l in dis.splitlines():
    if l.startswith("//---") and ".text." in l:
        inside = kname in l
        continue
    if not inside: continue
    m = re.search(r'//## File ".*?", line (\d+)', l)
    if m: cur = int(m.group(1)); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l): lines.append(cur)
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hi]; ci = h.index("Instructions Executed"); sa = h.index("# Samples"); ti = h.index("Avg. Threads Executed")
data = [r for r in rows[hi + 1:] if len(r) > ci and r[ci].isdigit()]
assert len(data) == len(lines), (len(data), len(lines))
agg = collections.defaultdict(lambda: [0, 0, 0.0])
for r, ln in zip(data, lines):
    e = int(r[ci]); agg[ln][0] += e; agg[ln][1] += int(r[sa]); agg[ln][2] += e * float(r[ti] or 0)
te = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
src = open(os.path.join(os.path.dirname(os.path.abspath(obj)), "..", os.path.basename(obj).replace(".o", ".cu"))).read().splitlines()
print("total warp instr %d, samples %d" % (te, ts))
for ln, v in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
    print("%5.1f%% samp %5.1f%% instr thr=%4.1f  L%-4s %s" % (100 * v[1] / ts, 100 * v[0] / te, v[2] / max(1, v[0]), ln, (src[ln - 1].strip()[:95] if ln and ln <= len(src) else "")))
