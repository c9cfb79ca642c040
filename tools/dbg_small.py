import sys, numpy as np
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import test_gpu_parity as t
from oracle import oracle as O
from kasa_b200 import capi
for K, kh, kl, nt in [(12, 12, 7, 10), (12, 12, 7, 9)]:
    taxids, index, reads = t._shared_region_case(K, nt, 120, K * 1000 + nt)
    for mode in (capi.SCORE_NONE, capi.SCORE_ORDERED):
        ctx = t._ctx(highest_k=K, k_high=kh, k_low=kl, frames=6, max_batch_kmers=1 << 20, score_mode=mode)
        ctx.load_index(index, taxids)
        n_kmers, scores, _ = ctx.identify_batch(reads)
        ca, cu, _ = ctx.identify_finish()
        ctx.close()
        ora = O.identify_reads([[r] for r in reads], index, K, kh, kl, True, taxids)
        print("taxa", nt, "mode", mode, "unique equal", np.array_equal(cu, ora.count_unique), "ca close", np.allclose(ca, ora.count_all, rtol=1e-6))
        if not np.array_equal(cu, ora.count_unique):
            d = cu.astype(np.int64) - ora.count_unique.astype(np.int64)
            print(" diff unique (rows = k_high..k_low):\n", d)
            print(" diff all:\n", np.round(ca - ora.count_all, 3))
        if scores is not None:
            mine = O.identify_reads_ordered([[r] for r in reads], index, K, kh, kl, True, taxids)
            bad = np.argwhere(scores != mine.scores)
            print(" score cells differing from ordered oracle:", len(bad), bad[:5].tolist())
