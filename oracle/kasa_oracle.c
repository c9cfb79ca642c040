/*
 * kasa_oracle.c — CPU restatement of kASA's identify/build hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under kasa_b200/ (the product) may link,
 * import or execute this file; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it, as the checker.
 *
 * Parity status: PINNED against the reference itself.  tests/golden/ holds index,
 * trie, frequency, profile and per-read outputs produced by oracle/_ref/kASA_scalar
 * (the unmodified reference sources, scalar scoring path, -n 1) and
 * kASA_scalar_fix128 on seeded synthetic inputs (oracle/make_golden.py);
 * tests/test_oracle_vs_reference.py checks this restatement against them
 * (byte-exact files, exact unique counts, bit-exact fp32 k-mer scores).
 *
 * Each function cites the reference file:line it restates (paths relative to
 * /root/reference/source).  The lookup is restated from the *set semantics* of
 * Compare.hpp:679-1069 plus the flush order that fixes the
 * fp32 summation order; it is not a transcription of the state machine.
 *
 * All k-mers are handled as unsigned __int128 so one code path serves the
 * 64-bit (K=12) and 128-bit (K=25) encodings.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

typedef unsigned __int128 kmer_t;

/* ------------------------------------------------------------------------- */
/* Codon table: kASA.hpp:621-667 (_sAminoAcids_bs).  Index = c0<<6|c1<<3|c2 with
 * the 3-bit base codes of kASA.hpp:70-75 (A0 C1 T2 G3 X4 Z5).  Rebuilt here from
 * the standard genetic code: stop TAA/TAG -> '[', TGA -> ']', any codon holding
 * Z -> '_', otherwise any codon holding X -> '^'.  Unused slots are ' '.       */
void ko_default_codon_table(uint8_t table[512])
{
    /* standard code in TCAG order (NCBI table 1) */
    static const char *aa_tcag = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
    static const int tcag_to_code[4] = { 2, 1, 0, 3 };  /* T C A G -> kASA code */
    memset(table, ' ', 512);
    for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b)
            for (int c = 0; c < 4; ++c) {
                char aa = aa_tcag[a * 16 + b * 4 + c];
                int ia = tcag_to_code[a], ib = tcag_to_code[b], ic = tcag_to_code[c];
                if (aa == '*') aa = (a == 0 && b == 3 && c == 2) ? ']' : '[';   /* TGA -> ']' */
                table[ia << 6 | ib << 3 | ic] = (uint8_t)aa;
            }
    for (int a = 0; a < 6; ++a)
        for (int b = 0; b < 6; ++b)
            for (int c = 0; c < 6; ++c) {
                if (a < 4 && b < 4 && c < 4) continue;
                int z = (a == 5) || (b == 5) || (c == 5);
                table[a << 6 | b << 3 | c] = z ? '_' : '^';
            }
}

/* kASA.hpp:69-87 dnaToAminoacid: (c & 14) >> 1 per base, 9-bit table index. */
static inline uint32_t aa_at(const uint8_t *table, const char *s, uint64_t p)
{
    uint32_t idx = ((uint32_t)(s[p] & 14) << 5) | ((uint32_t)(s[p + 1] & 14) << 2) | ((uint32_t)(s[p + 2] & 14) >> 1);
    return table[idx] & 31u;    /* kASA.hpp:340 letter & 31 */
}

/* Read.hpp:36-57 calculatekMerCount (DNA, three/six frame branch). */
uint64_t ko_kmer_count(uint64_t len, int K)
{
    return (len > (uint64_t)(3 * K + 1)) ? len - 3 * K + 1 : 0;
}

/* kASA.hpp:214-221 reverseComplement with _aRevComp (kASA.hpp:54): index (c>>1)&7. */
void ko_revcomp(const char *s, uint64_t len, char *out)
{
    static const char rc[8] = { 'T', 'G', 'A', 'C', 'X', 'Z', '?', '?' };
    for (uint64_t j = 0; j < len; ++j) out[len - 1 - j] = rc[(s[j] >> 1) & 7];
}

/* Read.hpp:657-675 searchAndReplaceLettersOfRead (DNA): non-ACGTacgt -> 'Z'.
 * Returns -1 on space/tab (the reference throws).                            */
int ko_clean(char *s, uint64_t len)
{
    for (uint64_t i = 0; i < len; ++i) {
        char c = s[i];
        if (c == ' ' || c == '\t') return -1;
        if (c != 'A' && c != 'C' && c != 'G' && c != 'T' && c != 'a' && c != 'c' && c != 'g' && c != 't') s[i] = 'Z';
    }
    return 0;
}

static inline kmer_t kmer_mask(int K)
{
    return (K * 5 >= 128) ? ~(kmer_t)0 : (((kmer_t)1 << (5 * K)) - 1);
}

/* Read.hpp:84-220 convert_dnaTokMer (identify) — emits kMerCount k-mers in offset
 * order; kmer(p) = sum aa(p+3i) << 5(K-1-i)  (kASA.hpp:333-379).
 * out: 2 x u64 per k-mer {low, high} (uint128_t.hpp:65-75 layout).            */
uint64_t ko_extract(const uint8_t *table, const char *s, uint64_t len, int K, uint64_t *out)
{
    uint64_t n = ko_kmer_count(len, K);
    kmer_t mask = kmer_mask(K);
    kmer_t frame[3] = { 0, 0, 0 };
    for (uint64_t p = 0; p < n; ++p) {
        int f = (int)(p % 3);
        if (p < 3) {
            kmer_t v = 0;
            for (int i = 0; i < K; ++i) v = (v << 5) | aa_at(table, s, p + 3 * (uint64_t)i);
            frame[f] = v;
        } else {
            frame[f] = ((frame[f] << 5) & mask) | aa_at(table, s, p + 3 * (uint64_t)(K - 1));
        }
        out[2 * p] = (uint64_t)frame[f];
        out[2 * p + 1] = (uint64_t)(frame[f] >> 64);
    }
    return n;
}

/* Read.hpp:1991-2139 dnaTokMers (build): one k-mer per window of S; windows holding
 * letter 31 ('_') are dropped (:2016-2024, :2048-2059) and so are k-mers == 0
 * (:2097, :2122).  Build strings have iMaxRange = len - 3K + 1 >= 1 (no +1 rule).
 * Returns number of k-mers written ({low,high} pairs).                        */
uint64_t ko_extract_build_step(const uint8_t *table, const char *s, uint64_t len, int K, int step, uint64_t *out);
uint64_t ko_extract_build(const uint8_t *table, const char *s, uint64_t len, int K, uint64_t *out)
{
    return ko_extract_build_step(table, s, len, K, 1, out);
}

/* step = 3: `build --one` (Read.hpp:2143-2200 dnaTokMers_oneFrame): windows at the offsets 0, 3, 6, ... < len - 3K + 1 only. */
uint64_t ko_extract_build_step(const uint8_t *table, const char *s, uint64_t len, int K, int step, uint64_t *out)
{
    if (len < (uint64_t)(3 * K)) return 0;
    uint64_t n = len - 3 * K + 1, w = 0;
    for (uint64_t p = 0; p < n; p += (uint64_t)step) {
        kmer_t v = 0;
        int bad = 0;
        for (int i = 0; i < K; ++i) {
            uint32_t a = aa_at(table, s, p + 3 * (uint64_t)i);
            if (a == 31) bad = 1;
            v = (v << 5) | a;
        }
        if (bad || v == 0) continue;
        out[2 * w] = (uint64_t)v;
        out[2 * w + 1] = (uint64_t)(v >> 64);
        ++w;
    }
    return w;
}

/* ------------------------------------------------------------------------- */
/* Index entries: packedPairs.hpp:107-111 (12 B: u64 kmer, u32 taxid) and
 * :132-136 (20 B: u64 low, u64 high, u32 taxid).  Order: kmer then taxid
 * (packedPairs.hpp:123-125).                                                 */
typedef struct { kmer_t k; uint32_t tax; } entry_t;

static int cmp_entry(const void *a, const void *b)
{
    const entry_t *x = a, *y = b;
    if (x->k != y->k) return x->k < y->k ? -1 : 1;
    if (x->tax != y->tax) return x->tax < y->tax ? -1 : 1;
    return 0;
}

static inline size_t rec_size(int bits) { return bits == 128 ? 20 : 12; }

static void store_entry(uint8_t *dst, int bits, kmer_t k, uint32_t tax)
{
    uint64_t lo = (uint64_t)k, hi = (uint64_t)(k >> 64);
    memcpy(dst, &lo, 8);
    if (bits == 128) { memcpy(dst + 8, &hi, 8); memcpy(dst + 16, &tax, 4); }
    else memcpy(dst + 8, &tax, 4);
}

static entry_t load_entry(const uint8_t *src, int bits)
{
    entry_t e; uint64_t lo, hi = 0;
    memcpy(&lo, src, 8);
    if (bits == 128) { memcpy(&hi, src + 8, 8); memcpy(&e.tax, src + 16, 4); }
    else memcpy(&e.tax, src + 8, 4);
    e.k = ((kmer_t)hi << 64) | lo;
    return e;
}

/* Build.hpp:305-356 IntToExtPart (sort, unique) + :360-477 mergeTemporaries (merge, dedup):
 * the index is the sorted set of distinct (kmer, taxid).  kmers: {low,high} pairs.
 * Writes packed records to out (capacity n), returns the number kept.        */
uint64_t ko_sort_dedup(const uint64_t *kmers, const uint32_t *tax, uint64_t n, int bits, uint8_t *out)
{
    entry_t *v = malloc((n ? n : 1) * sizeof(entry_t));
    for (uint64_t i = 0; i < n; ++i) { v[i].k = ((kmer_t)kmers[2 * i + 1] << 64) | kmers[2 * i]; v[i].tax = tax[i]; }
    qsort(v, n, sizeof(entry_t), cmp_entry);
    uint64_t w = 0;
    for (uint64_t i = 0; i < n; ++i) {
        if (i && v[i].k == v[i - 1].k && v[i].tax == v[i - 1].tax) continue;
        store_entry(out + w * rec_size(bits), bits, v[i].k, v[i].tax);
        ++w;
    }
    free(v);
    return w;
}

/* Trie.hpp:366-394 SaveToStxxlVec: run-length encode the first 6 letters.
 * Records: packedBigPairTrie {u64 count, u32 prefix30} 12 B (packedPairs.hpp:157-166).
 * Returns number of records written (out capacity must be >= n).             */
uint64_t ko_trie(const uint8_t *index, uint64_t n, int bits, int K, uint8_t *out)
{
    if (!n) return 0;
    int sh = 5 * (K - 6);
    uint64_t w = 0, cnt = 0; uint32_t cur = 0;
    for (uint64_t i = 0; i < n; ++i) {
        uint32_t p = (uint32_t)(load_entry(index + i * rec_size(bits), bits).k >> sh);
        if (i && p != cur) { memcpy(out + 12 * w, &cnt, 8); memcpy(out + 12 * w + 8, &cur, 4); ++w; cnt = 0; }
        cur = p; ++cnt;
    }
    memcpy(out + 12 * w, &cnt, 8); memcpy(out + 12 * w + 8, &cur, 4); ++w;
    return w;
}

/* kASA.hpp:449-575 GetFrequencyK: freq[idx][j], j = 0..K-1 <-> letter position K-j
 * (shift 5*j), counts entries whose letter there != 30.  dense idx by taxid_of_idx. */
static int cmp_u32pair(const void *a, const void *b)
{
    const uint32_t *x = a, *y = b;
    return x[0] < y[0] ? -1 : (x[0] > y[0]);
}

typedef struct { uint32_t *map; uint32_t n; } taxmap_t;     /* sorted (taxid, idx) */

static taxmap_t taxmap_make(const uint32_t *taxid_of_idx, uint32_t T)
{
    taxmap_t m; m.n = T; m.map = malloc(8 * (size_t)(T ? T : 1));
    for (uint32_t i = 0; i < T; ++i) { m.map[2 * i] = taxid_of_idx[i]; m.map[2 * i + 1] = i; }
    /* idx 0 is the "non_unique" pseudo taxon (Compare.hpp:111-180); keep real ones searchable */
    qsort(m.map + 2, T ? T - 1 : 0, 8, cmp_u32pair);
    return m;
}

static int64_t taxmap_find(const taxmap_t *m, uint32_t taxid)
{
    int64_t lo = 1, hi = (int64_t)m->n - 1;
    while (lo <= hi) {
        int64_t mid = (lo + hi) / 2;
        uint32_t t = m->map[2 * mid];
        if (t == taxid) return m->map[2 * mid + 1];
        if (t < taxid) lo = mid + 1; else hi = mid - 1;
    }
    return -1;
}

int ko_frequency(const uint8_t *index, uint64_t n, int bits, int K,
                 const uint32_t *taxid_of_idx, uint32_t T, uint64_t *freq /* [T][K] */)
{
    taxmap_t m = taxmap_make(taxid_of_idx, T);
    memset(freq, 0, sizeof(uint64_t) * (size_t)T * K);
    int rc = 0;
    for (uint64_t i = 0; i < n; ++i) {
        entry_t e = load_entry(index + i * rec_size(bits), bits);
        int64_t idx = taxmap_find(&m, e.tax);
        if (idx < 0) { rc = -1; break; }
        for (int j = 0; j < K; ++j)
            if (((uint32_t)(e.k >> (5 * j)) & 31u) != 30u) ++freq[(size_t)idx * K + j];
    }
    free(m.map);
    return rc;
}

/* ------------------------------------------------------------------------- */
/* identify: Compare.hpp:1074-1260 (sort + range) and :679-1069 (lookup).       */

typedef struct { kmer_t k; uint32_t read; uint32_t seq; } query_t;

static int cmp_query(const void *a, const void *b)
{
    const query_t *x = a, *y = b;
    if (x->k != y->k) return x->k < y->k ? -1 : 1;
    /* the reference's sort is unstable on ties (Compare.hpp:1077); results do not
       depend on the tie order, we fix one to be deterministic */
    if (x->seq != y->seq) return x->seq < y->seq ? -1 : 1;
    return 0;
}

/* first index position whose pre_shift(kmer) >= v */
static uint64_t lower_pre(const entry_t *idx, uint64_t lo, uint64_t hi, kmer_t v, int sh)
{
    while (lo < hi) {
        uint64_t mid = lo + (hi - lo) / 2;
        if ((idx[mid].k >> sh) < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

typedef struct {
    kmer_t prefix; int active;
    uint64_t nhits, cap; uint32_t *hits;      /* read ids (Compare.hpp:846 addToMatchedReadID) */
    uint32_t ntaxa; uint32_t *taxa;           /* distinct dense idx in insertion order (BitArray.hpp:18-147) */
    uint8_t *mark;
} level_t;

typedef struct {
    int numK, T; const float *w; double *count_all; uint64_t *count_unique; float *scores; uint64_t *count_total;
} acc_t;

/* Compare.hpp:516-532 scoreMatchNonAVX — the scalar path is the spec. */
static void flush_level(level_t *L, int ik, acc_t *A)
{
    if (L->active && L->ntaxa) {
        uint32_t n = L->ntaxa;
        float score = A->w[ik] * (1.f / n);                 /* Compare.hpp:924 */
        double counts = (double)L->nhits / n;               /* Compare.hpp:925 */
        for (uint32_t j = 0; j < n; ++j) {
            uint32_t t = L->taxa[j];
            size_t ci = (size_t)ik * A->T + t;
            A->count_all[ci] += counts;
            if (A->count_total) A->count_total[ci] += 1;
            if (n == 1) A->count_unique[ci] += L->nhits;
            if (A->scores)
                for (uint64_t h = 0; h < L->nhits; ++h) A->scores[(size_t)L->hits[h] * A->T + t] += score;
        }
    }
    for (uint32_t j = 0; j < L->ntaxa; ++j) L->mark[L->taxa[j]] = 0;
    L->ntaxa = 0; L->nhits = 0; L->active = 0;
}

/*
 * kmers: {low,high} pairs per query, read_id per query.  index: packed records.
 * taxid_of_idx[T] (idx 0 = non_unique).  K = highestK (12|25).  Levels ik = 0..numK-1
 * <-> k = kHigh - ik (kASA.hpp:299-302).  Outputs (caller-zeroed, accumulated into):
 *   count_all [numK][T] double, count_unique [numK][T] u64, scores [nreads][T] float (may be NULL),
 *   count_total [numK][T] (may be NULL, --coverage).
 * Returns 0, or -1 if an index taxid is missing from taxid_of_idx.
 */
int ko_identify(const uint64_t *kmers, const uint32_t *read_id, uint64_t nq,
                const uint8_t *index, uint64_t nidx, int bits, int K, int kHigh, int kLow,
                const uint32_t *taxid_of_idx, uint32_t T,
                double *count_all, uint64_t *count_unique, float *scores, uint64_t *count_total)
{
    int numK = kHigh - kLow + 1;
    /* Compare.hpp:392 arrWeightingFactors[25-k] = k^2/625 as float literals "k*k.f / 625.f" */
    float w[32];
    for (int ik = 0; ik < numK; ++ik) { int k = kHigh - ik; w[ik] = (float)(k * k) / 625.f; }

    entry_t *idx = malloc((nidx ? nidx : 1) * sizeof(entry_t));
    uint32_t *dense = malloc((nidx ? nidx : 1) * sizeof(uint32_t));
    taxmap_t m = taxmap_make(taxid_of_idx, T);
    for (uint64_t i = 0; i < nidx; ++i) {
        idx[i] = load_entry(index + i * rec_size(bits), bits);
        int64_t d = taxmap_find(&m, idx[i].tax);
        if (d < 0) { free(idx); free(dense); free(m.map); return -1; }
        dense[i] = (uint32_t)d;
    }
    query_t *q = malloc((nq ? nq : 1) * sizeof(query_t));
    for (uint64_t i = 0; i < nq; ++i) { q[i].k = ((kmer_t)kmers[2 * i + 1] << 64) | kmers[2 * i]; q[i].read = read_id[i]; q[i].seq = (uint32_t)i; }
    qsort(q, nq, sizeof(query_t), cmp_query);       /* Compare.hpp:1077 / 1130 */

    level_t *L = calloc(numK, sizeof(level_t));
    for (int ik = 0; ik < numK; ++ik) { L[ik].taxa = malloc(4 * (size_t)T); L[ik].mark = calloc(T, 1); L[ik].cap = 64; L[ik].hits = malloc(4 * 64); }
    acc_t A = { numK, (int)T, w, count_all, count_unique, scores, count_total };

    /* Range = trie range of the first min(6,kLow) letters (Compare.hpp:1103-1109,
     * Trie.hpp:494-520); absent prefix <=> range == UINT64_MAX. */
    int Lr = kLow < 6 ? kLow : 6;
    int shR = 5 * (K - Lr);
    int have_range = 0; kmer_t cur_range = 0; uint64_t rlo = 0, rhi = 0;
    kmer_t prevq = 0; int have_prev = 0;     /* memo of the last range lookup */
    uint64_t prlo = 0, prhi = 0;

    for (uint64_t i = 0; i < nq; ++i) {
        kmer_t key = q[i].k;
        kmer_t pr = key >> shR;
        uint64_t lo, hi;
        if (have_prev && pr == prevq) { lo = prlo; hi = prhi; }
        else {
            lo = lower_pre(idx, 0, nidx, pr, shR);
            hi = lower_pre(idx, lo, nidx, pr + 1, shR);
            prevq = pr; have_prev = 1; prlo = lo; prhi = hi;
        }
        if (lo == hi) continue;                       /* range == MAX: skipped (Compare.hpp:753, 787) */
        if (!have_range || pr != cur_range) {
            /* new range: flush everything of the old one, kLow first (Compare.hpp:1032-1041) */
            if (have_range) for (int ik = numK - 1; ik >= 0; --ik) flush_level(&L[ik], ik, &A);
            have_range = 1; cur_range = pr; rlo = lo; rhi = hi;
        }
        /* '^' (30) at position kLow kills the k-mer (Compare.hpp:836) */
        if (((uint32_t)(key >> (5 * (K - kLow))) & 31u) == 30u) continue;

        for (int ik = numK - 1; ik >= 0; --ik) {      /* ascending k (Compare.hpp:865-995) */
            int k = kHigh - ik, sh = 5 * (K - k);
            kmer_t P = key >> sh;
            if (((uint32_t)P & 31u) == 30u) break;     /* '^' at position k (Compare.hpp:897) */
            if (L[ik].active && L[ik].prefix == P) {   /* same group: one more hit instance */
                if (L[ik].nhits == L[ik].cap) { L[ik].cap *= 2; L[ik].hits = realloc(L[ik].hits, 4 * L[ik].cap); }
                L[ik].hits[L[ik].nhits++] = q[i].read;
                continue;
            }
            uint64_t a = lower_pre(idx, rlo, rhi, P, sh);
            if (a == rhi || (idx[a].k >> sh) != P) break;          /* absent at level k => absent above */
            /* present: the previous group of this level is scored now (Compare.hpp:917-955) */
            flush_level(&L[ik], ik, &A);
            L[ik].active = 1; L[ik].prefix = P;
            L[ik].hits[0] = q[i].read; L[ik].nhits = 1;
            for (uint64_t e = a; e < rhi && (idx[e].k >> sh) == P; ++e) {   /* markTaxIDs (Compare.hpp:396-441) */
                uint32_t t = dense[e];
                if (!L[ik].mark[t]) { L[ik].mark[t] = 1; L[ik].taxa[L[ik].ntaxa++] = t; }
            }
        }
    }
    if (have_range) for (int ik = numK - 1; ik >= 0; --ik) flush_level(&L[ik], ik, &A);

    for (int ik = 0; ik < numK; ++ik) { free(L[ik].taxa); free(L[ik].mark); free(L[ik].hits); }
    free(L); free(q); free(idx); free(dense); free(m.map);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* KASA_SCORE_ORDERED of the CUDA library, restated on the CPU so that the GPU path can be checked bit for bit:
 * the same set semantics, but score[read][taxon] is summed in float32 in the order
 * (position of the instance in the stable sort of the queries on their leading `sort_bits` key bits, k ascending).
 * For instances that the reference flushes level by level in k-mer order this is the reference's own order
 * (Compare.hpp:917-955, 1032-1041); the value added is the reference's float(w(k) * (1.f / n)) (Compare.hpp:924).
 * count_all is accumulated in double per instance (the GPU uses 2^-32 fixed point; compare at 1e-6).            */
typedef struct { kmer_t top; uint32_t seq; } oq_t;
static int cmp_oq(const void *a, const void *b)
{
    const oq_t *x = a, *y = b;
    if (x->top != y->top) return x->top < y->top ? -1 : 1;
    return x->seq < y->seq ? -1 : (x->seq > y->seq);
}

int ko_identify_ordered(const uint64_t *kmers, const uint32_t *read_id, uint64_t nq,
                        const uint8_t *index, uint64_t nidx, int bits, int K, int kHigh, int kLow, int sort_bits,
                        const uint32_t *taxid_of_idx, uint32_t T,
                        double *count_all, uint64_t *count_unique, float *scores)
{
    int numK = kHigh - kLow + 1;
    float w[32];
    for (int j = 0; j < numK; ++j) { int k = kLow + j; w[j] = (float)(k * k) / 625.f; }      /* level j = 0 <-> kLow */
    entry_t *idx = malloc((nidx ? nidx : 1) * sizeof(entry_t));
    uint32_t *dense = malloc((nidx ? nidx : 1) * sizeof(uint32_t));
    taxmap_t m = taxmap_make(taxid_of_idx, T);
    for (uint64_t i = 0; i < nidx; ++i) {
        idx[i] = load_entry(index + i * rec_size(bits), bits);
        int64_t d = taxmap_find(&m, idx[i].tax);
        if (d < 0) { free(idx); free(dense); free(m.map); return -1; }
        dense[i] = (uint32_t)d;
    }
    int drop = 5 * K - sort_bits; if (drop < 0) drop = 0;
    oq_t *q = malloc((nq ? nq : 1) * sizeof(oq_t));
    for (uint64_t i = 0; i < nq; ++i) { kmer_t k = ((kmer_t)kmers[2 * i + 1] << 64) | kmers[2 * i]; q[i].top = k >> drop; q[i].seq = (uint32_t)i; }
    qsort(q, nq, sizeof(oq_t), cmp_oq);
    uint32_t *h = calloc(T ? T : 1, sizeof(uint32_t)), *touched = malloc(4 * (size_t)(T ? T : 1)), nj[32];
    int shLow = 5 * (K - kLow);
    for (uint64_t s = 0; s < nq; ++s) {
        uint64_t i = q[s].seq;
        kmer_t key = ((kmer_t)kmers[2 * i + 1] << 64) | kmers[2 * i];
        /* last level not killed by '^' (Compare.hpp:836, 897) */
        int kv = kLow - 1;
        for (int k = kLow; k <= kHigh; ++k) { if (((uint32_t)(key >> (5 * (K - k))) & 31u) == 30u) break; kv = k; }
        if (kv < kLow) continue;
        uint64_t a = lower_pre(idx, 0, nidx, key >> shLow, shLow);
        uint32_t nt = 0;
        for (uint64_t e = a; e < nidx && (idx[e].k >> shLow) == (key >> shLow); ++e) {
            kmer_t x = idx[e].k ^ key;
            int c = K;                                  /* common leading letters */
            for (int k = 1; k <= K; ++k) if ((x >> (5 * (K - k))) != 0) { c = k - 1; break; }
            if (c > kv) c = kv;
            uint32_t lev = (uint32_t)(c - kLow + 1), t = dense[e];
            if (!h[t]) touched[nt++] = t;
            if (lev > h[t]) h[t] = lev;
        }
        if (!nt) continue;
        for (int j = 0; j < numK; ++j) nj[j] = 0;
        for (uint32_t x = 0; x < nt; ++x) for (uint32_t j = 0; j < h[touched[x]]; ++j) ++nj[j];
        for (uint32_t x = 0; x < nt; ++x) {
            uint32_t t = touched[x];
            for (uint32_t j = 0; j < h[t]; ++j) {
                size_t ci = (size_t)(numK - 1 - j) * T + t;                  /* ik = 0 <-> kHigh */
                count_all[ci] += 1.0 / nj[j];
                if (nj[j] == 1) count_unique[ci] += 1;
                if (scores) scores[(size_t)read_id[i] * T + t] += w[j] * (1.f / nj[j]);
            }
            h[t] = 0;
        }
    }
    free(h); free(touched); free(q); free(idx); free(dense); free(m.map);
    return 0;
}

/* Compare.hpp:1452-1481 calculateBestScore (DNA): float accumulation, k ascending. */
float ko_best_score(uint32_t read_len, int kHigh, int kLow, int frames)
{
    float best = 0.f;
    for (int k = kLow; k <= kHigh; ++k) {
        float wk = (float)(k * k) / 625.f;
        if (frames == 1) best += (read_len / 3 - k + 1) * wk;
        else if (frames == 3) best += (read_len - k * 3 + 1) * wk;
        else best += 2 * (read_len - k * 3 + 1) * wk;
    }
    return best;
}

/* Compare.hpp:1510 relative score: kmerScore / (1 + log2(freq[t][0] * (Length - 3*highestK + 1))). */
double ko_relative_score(float kmer_score, uint64_t freq_highest, uint32_t read_len, int K)
{
    return kmer_score / (1.0 + log2((double)freq_highest * (double)(read_len - K * 3 + 1)));
}
