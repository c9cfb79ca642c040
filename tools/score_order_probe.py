"""Which fp32 summation order of the per-read score reproduces the reference's running sum?  (design input, CPU only)

For a golden case, compute per instance the (taxon, levels, level sizes) the lookup produces, then accumulate
score[r][t] += float32(w(k) * (1.f / n)) sequentially in float32 under several candidate orders and compare with
the oracle's scores (which are bit-identical to the reference binary's, tests/test_oracle_vs_reference.py).
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from oracle import oracle as O
import golden_util as G


def instance_info(kmers, index_k, index_t, T, K, k_high, k_low):
    """per instance: hcnt[i, t] = number of levels (from k_low up) taxon t matches; nsz[i, j] = |taxa_j|"""
    n = len(kmers)
    num_k = k_high - k_low + 1
    hcnt = np.zeros((n, T), dtype=np.int32)
    nsz = np.zeros((n, num_k), dtype=np.int32)
    cum = np.zeros((T, len(index_k) + 1), dtype=np.int64)
    for t in range(T):
        cum[t, 1:] = np.cumsum(index_t == t)
    alive = np.ones(n, dtype=bool)
    for j, k in enumerate(range(k_low, k_high + 1)):
        sh = 5 * (K - k)
        qp = [int(x) >> sh for x in kmers]
        ip = [int(x) >> sh for x in index_k]
        caret = np.array([(p & 31) == 30 for p in qp])
        alive &= ~caret
        a = np.searchsorted(np.array(ip, dtype=object).astype(np.float64) if False else np.array(ip, dtype=np.uint64), np.array(qp, dtype=np.uint64), "left")
        b = np.searchsorted(np.array(ip, dtype=np.uint64), np.array(qp, dtype=np.uint64), "right")
        alive &= b > a
        for t in range(T):
            present = (cum[t, b] - cum[t, a] > 0) & alive
            hcnt[present, t] += 1
            nsz[present, j] += 1
    return hcnt, nsz


def accumulate(order, read_id, hcnt, nsz, n_reads, T, k_low):
    w = [np.float32(k * k) / np.float32(625.0) for k in range(k_low, k_low + nsz.shape[1])]
    sc = np.zeros((n_reads, T), dtype=np.float32)
    one = np.float32(1.0)
    for i in order:
        ts = np.nonzero(hcnt[i])[0]
        if len(ts) == 0:
            continue
        r = read_id[i]
        vals = [np.float32(w[j] * (one / np.float32(nsz[i, j]))) if nsz[i, j] else np.float32(0) for j in range(nsz.shape[1])]
        for t in ts:
            acc = sc[r, t]
            for j in range(hcnt[i, t]):
                acc = np.float32(acc + vals[j])
            sc[r, t] = acc
    return sc


def main():
    case, tag, K, k_high, k_low, six = "k12", "six_12_7", 12, 12, 7, True
    if len(sys.argv) > 1 and sys.argv[1] == "three":
        tag, k_high, k_low, six = "three_11_8", 11, 8, False
    _, taxids, _ = G.content(case)
    idx = G.index(case)
    reads = [[s] for _, s in G.fastq(case)]
    kmers, rid = O.read_kmers(reads, K, k_low, six)
    ora = O.identify_kmers(kmers, rid, len(reads), idx, K, k_high, k_low, taxids)
    T = len(taxids)
    dense = {int(t): i for i, t in enumerate(taxids)}
    index_t = np.array([dense[int(t)] for t in idx["tax"]])
    klo = kmers[:, 0].copy()
    hcnt, nsz = instance_info(klo, idx["lo"], index_t, T, K, k_high, k_low)
    n = len(klo)
    emission = np.arange(n)
    orders = {
        "kmer_full_then_emission": np.lexsort((emission, klo)),
        "kmer_top32_then_emission": np.lexsort((emission, klo >> np.uint64(28))),
        "kmer_top40_then_emission": np.lexsort((emission, klo >> np.uint64(20))),
        "emission": emission,
    }
    ref = ora.scores
    for name, order in orders.items():
        sc = accumulate(order, rid, hcnt, nsz, len(reads), T, k_low)
        nz = ref != 0
        assert np.array_equal(sc != 0, nz), name
        rel = np.abs(sc[nz].astype(np.float64) - ref[nz]) / ref[nz]
        print("%-28s cells %d  bit-exact %.2f%%  max rel %.3g  >1e-6: %d  >4e-7: %d" % (name, nz.sum(), 100.0 * np.mean(sc[nz] == ref[nz]), rel.max(), (rel > 1e-6).sum(), (rel > 4e-7).sum()))
    # exact (float64) sum
    sc64 = np.zeros_like(ref, dtype=np.float64)
    w = [np.float32(k * k) / np.float32(625.0) for k in range(k_low, k_high + 1)]
    for i in range(n):
        for t in np.nonzero(hcnt[i])[0]:
            for j in range(hcnt[i, t]):
                sc64[rid[i], t] += float(np.float32(w[j] * (np.float32(1.0) / np.float32(nsz[i, j]))))
    rel = np.abs(sc64[nz] - ref[nz]) / ref[nz]
    print("%-28s max rel %.3g  >1e-6: %d" % ("float64", rel.max(), (rel > 1e-6).sum()))


if __name__ == "__main__":
    main()
