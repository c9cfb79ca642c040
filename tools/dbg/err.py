import sys, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import golden_util as G
from kasa_b200 import Context, Params, capi
from oracle import oracle as O
_, taxids, _ = G.content('k12'); idx = G.index('k12')
reads = [[s] for _, s in G.fastq('k12')]
for six, kh, kl in ((True, 12, 7), (False, 11, 8)):
    ctx = Context(Params(k_high=kh, k_low=kl, frames=6 if six else 3))
    ctx.load_index(idx, taxids)
    n, sc, _ = ctx.identify_batch([r[0] for r in reads])
    ca, cu, _ = ctx.identify_finish()
    ora = O.identify_reads(reads, idx, 12, kh, kl, six, taxids)
    m = ora.scores > 0
    rel = np.abs(sc[m] - ora.scores[m]) / ora.scores[m]
    print('six', six, 'same support', np.array_equal(sc > 0, m), 'max rel', rel.max(), 'n>4e-6', (rel > 4e-6).sum(), 'of', m.sum(),
          'cu equal', np.array_equal(cu, ora.count_unique), 'ca maxrel', np.abs(ca - ora.count_all).max() / ora.count_all.max())
    w = np.argmax(np.where(m, np.abs(sc - ora.scores) / np.maximum(ora.scores, 1e-30), 0)); r, t = divmod(w, sc.shape[1])
    print('  worst', r, t, sc[r, t], ora.scores[r, t])
    ctx.close()
