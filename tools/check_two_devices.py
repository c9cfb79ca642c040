#!/usr/bin/env python
"""Manual check on a box with >= 2 GPUs (not a pytest: the driver's test box has one): one process driving two devices —
the host CLI with --gpus 2, and a prefix-sharded LocalShardGroup over two devices (peer-access path of kasa_shard_connect)."""
import subprocess, sys, os, tempfile, pathlib
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np
import golden_util as G
from test_host_cli import _write_index, _g, CLI
tmp = pathlib.Path(tempfile.mkdtemp())
idx = _write_index("k12", tmp)
outs = []
for gpus in ("1", "2"):
    q = str(tmp / ("q" + gpus)); p = str(tmp / ("p" + gpus))
    r = subprocess.run([CLI, "identify", "-c", _g("k12", "content.txt"), "-d", idx, "-i", _g("k12", "reads.fq.gz"), "--six", "--jsonl", "-b", "100000", "-p", p, "-q", q, "--batch-reads", "40", "--gpus", gpus], capture_output=True)
    print("cli gpus", gpus, "rc", r.returncode, r.stderr[-200:])
    outs.append(open(q).read())
print("CLI: two real GPUs give the same per-read output:", outs[0] == outs[1], len(outs[0]))
# one process, two devices, prefix-sharded (peer access path of kasa_shard_connect)
from kasa_b200 import Context, Params, capi
from kasa_b200.dist import LocalShardGroup, shard_bounds
_, taxids, _ = G.content("k12")
index = G.index("k12")
reads = [s for _, s in G.fastq("k12")]
ref = Context(Params(frames=6)); ref.load_index(index, taxids)
nk, want, _ = ref.identify_batch(reads); ref.close()
eb, pb = capi.shard_plan(index, 7, 2)
ctxs = []
for r in range(2):
    c = Context(Params(device=r, frames=6)); c.load_index(index[int(eb[r]):int(eb[r + 1])], taxids); ctxs.append(c)
grp = LocalShardGroup(ctxs, pb)
lo0, hi0 = shard_bounds(len(reads), 2, 0)
out = grp.identify_batches([reads[:hi0], reads[hi0:]])
got = np.concatenate([o[1] for o in out])
print("one process / two devices, prefix-sharded == unsharded:", np.array_equal(got, want), sum(o[0] for o in out) == nk)
