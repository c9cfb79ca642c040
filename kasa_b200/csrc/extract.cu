// extract.cu — see extract.cuh.
#include <algorithm>

#include "extract.cuh"

namespace kasa {

// 5-bit letter codes of the codon table (kASA.hpp:621-667: letter & 31), index c0<<6|c1<<3|c2.  Every lane looks up a
// different codon, which constant memory would serialise (it broadcasts one address per request): the table lives in
// global memory (one per context: `-a` tables differ between contexts) and each CTA stages its 512 bytes into shared
// memory with one 16-byte load per thread of its first warp.
void upload_codon_table(const uint8_t table512[512], uint8_t *d_table, cudaStream_t s)
{
    uint8_t t[512];
    for (int i = 0; i < 512; ++i) t[i] = table512[i] & 31u;
    KASA_CUDA_TRY(cudaMemcpyAsync(d_table, t, 512, cudaMemcpyHostToDevice, s));
    KASA_CUDA_TRY(cudaStreamSynchronize(s));
}

namespace {

constexpr int kExtractWarps = 4;      // shared memory per warp (presort buffers) sets the occupancy: small CTAs waste less of it (8 warps: 23.9 vs 21.2 ms)
constexpr int kSeg = 256;              // output positions per shared-memory segment
constexpr int kAaPad = 3 * 24 + 8;     // 3*(K-1) extra letters for K <= 25

__global__ void pack_sizes_kernel(const uint32_t *__restrict__ seq_len, uint32_t n, uint64_t *__restrict__ out)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = ((uint64_t)seq_len[i] + 31u) & ~31ull;
}

// one warp per sequence; lane l packs 32-base groups l, l+32, ...
__global__ void __launch_bounds__(256) pack_ascii_kernel(const char *__restrict__ ascii, const uint64_t *__restrict__ astart,
                                                         const uint32_t *__restrict__ seq_len, uint32_t n_seqs,
                                                         const uint64_t *__restrict__ pstart, uint32_t *__restrict__ bases2,
                                                         uint32_t *__restrict__ nmask)
{
    uint32_t s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (s >= n_seqs) return;
    const uint32_t L = seq_len[s];
    const char *src = ascii + astart[s];
    const uint64_t g0 = pstart[s] >> 5;           // first 32-base group
    const uint32_t groups = (L + 31) >> 5;
    for (uint32_t g = lane_id(); g < groups; g += 32) {
        uint32_t lo = 0, hi = 0, m = 0;
#pragma unroll 8
        for (int j = 0; j < 32; ++j) {
            uint32_t pos = g * 32 + j;
            uint32_t code = 0, bad = 0;
            if (pos < L) {
                char c = src[pos];
                // Read.hpp:657-675: only ACGTacgt survive, everything else becomes 'Z'
                bool ok = (c == 'A') | (c == 'C') | (c == 'G') | (c == 'T') | (c == 'a') | (c == 'c') | (c == 'g') | (c == 't');
                code = ok ? ((uint32_t)(c & 14) >> 1) : 0u;     // kASA.hpp:70-75: A0 C1 T2 G3
                bad = ok ? 0u : 1u;
            }
            if (j < 16) lo |= code << (2 * j); else hi |= code << (2 * (j - 16));
            m |= bad << j;
        }
        bases2[(g0 + g) * 2] = lo;
        bases2[(g0 + g) * 2 + 1] = hi;
        nmask[g0 + g] = m;
    }
}


__global__ void kmer_counts_kernel(const uint32_t *__restrict__ seq_len, uint32_t n, KmerParams p, uint64_t *__restrict__ out)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (uint64_t)p.strands * kmer_count((uint64_t)seq_len[i] + 3u * (p.K - p.k_low), p.K, p.one);
}

// 3-bit base code of the prepared string (forward: read + marker; reverse: revcomp(read) + marker)
__device__ __forceinline__ uint32_t code_at(const uint32_t *__restrict__ b2, const uint32_t *__restrict__ nm, uint64_t start,
                                            uint32_t L, int strand, uint32_t pos)
{
    if (pos >= L) return 4u;                                  // marker / padding 'X'
    uint64_t g = start + (strand ? (L - 1 - pos) : pos);
    uint32_t two = (__ldg(b2 + (g >> 4)) >> (2 * (uint32_t)(g & 15))) & 3u;
    uint32_t isn = (__ldg(nm + (g >> 5)) >> (uint32_t)(g & 31)) & 1u;
    return isn ? 5u : (strand ? (two ^ 2u) : two);            // complement A0<->T2, C1<->G3 (kASA.hpp:54)
}

// One warp per READ (all its sequences: both mates of a pair).  For each sequence, strand and segment of kSegR output positions:
//   1. base codes of the prepared string (read / reverse complement + 'X' marker, Read.hpp:612-695) into shared
//      memory — straight from the ASCII bytes (coalesced byte loads, cleaning of Read.hpp:657-675 on the fly) or
//      from the 2-bit planes;
//   2. letters: aa[j] = codon table[c_j, c_j+1, c_j+2];
//   3. rolling k-mers: lane l owns 5 positions of one reading frame (j0, j0+3, ...): the first k-mer is assembled
//      from K letters, each following one is (kmer << 5 | next letter) & mask — 3 instructions instead of 3 K;
//   4. the k-mers leave through shared memory as coalesced 8-byte stores; inst[i] = i.
// PRESORTED EMISSION.  The reference scores a read's k-mers in the order of its sorted k-mer vector (Compare.hpp:1077, 917-955),
// and a float32 running sum depends on that order (score.cuh).  A read of at most kPresortMax k-mers is therefore emitted with its
// k-mers ALREADY in ascending order: all of them are collected in the warp's shared memory and ordered there (512 buckets on the
// leading nine key bits, then ranks inside each bucket by counting — ~2 comparisons per k-mer for ordinary reads, quadratic only
// for a read that is one repeated letter) before they are stored.  The batch sort is stable, so the match records of such a read
// come back in slot order = k-mer order, and the scoring stage replays them without ordering anything.  Longer reads keep the
// reference's emission order (forward strand, then reverse strand, per sequence) and are ordered by rank later (order_ops_kernel).
constexpr int kRoll = 5;                         // same-frame positions per lane
constexpr int kSegR = 30 * kRoll;                // 150 output positions per segment: one segment per strand of a 150 bp read
#ifndef KASA_SORT_BUCKET_BITS
#define KASA_SORT_BUCKET_BITS 9                  // 9: 4 KB less shared memory per warp than 10 (= two letters), 3 instead of 2 CTAs per SM
#endif
constexpr int kSortBucketBits = KASA_SORT_BUCKET_BITS;
constexpr int kSortBuckets = 1 << kSortBucketBits;

// dynamic shared memory of one warp: 2 x [cap] low words, 2 x [cap] high words (WIDE), [kSortBuckets + 4] bucket counters, [cap] u32, codes, letters
__host__ __device__ inline size_t extract_warp_bytes(bool wide, uint32_t cap)
{
    size_t b = (size_t)cap * (wide ? 32 : 16) + (size_t)4 * (kSortBuckets + 4) + (size_t)4 * cap + 2 * (size_t)(kSegR + kAaPad + 8);
    return (b + 15) & ~(size_t)15;
}

template <bool WIDE, bool ASCII>
__global__ void __launch_bounds__(kExtractWarps * 32) extract_identify_kernel(PackedSeqs seqs, const char *__restrict__ ascii,
                                                                              const uint64_t *__restrict__ astart, KmerParams p,
                                                                              const uint64_t *__restrict__ koff,
                                                                              uint64_t *__restrict__ key_lo,
                                                                              uint64_t *__restrict__ key_hi,
                                                                              uint32_t *__restrict__ inst,
                                                                              const uint8_t *__restrict__ codon,
                                                                              const uint32_t *__restrict__ read_first, uint32_t n_reads,
                                                                              uint32_t cap, uint32_t wbytes)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(16) uint8_t s_tab[512];
    if (threadIdx.x < 32) reinterpret_cast<uint4 *>(s_tab)[threadIdx.x] = reinterpret_cast<const uint4 *>(codon)[threadIdx.x];
    __syncthreads();
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t r = blockIdx.x * (blockDim.x >> 5) + w;
    if (r >= n_reads) return;
    unsigned char *base = smem_raw + (size_t)w * wbytes;
    uint64_t *klo = reinterpret_cast<uint64_t *>(base);                        // (cap is a multiple of 8: every array is 16-byte aligned)
    uint64_t *khi = klo + (WIDE ? cap : 0);
    uint64_t *glo = khi + cap;
    uint64_t *ghi = glo + (WIDE ? cap : 0);
    uint32_t *bcnt = reinterpret_cast<uint32_t *>(ghi + cap);
    uint32_t *ba = bcnt + kSortBuckets + 4;
    uint8_t *code = reinterpret_cast<uint8_t *>(ba + cap);
    uint8_t *aa = code + (kSegR + kAaPad + 8);

    const uint32_t s_begin = read_first ? read_first[r] : r, s_end = read_first ? read_first[r + 1] : r + 1;
    const uint64_t out0 = koff[s_begin];
    const uint32_t n_read = (uint32_t)min((uint64_t)0xffffffffu, koff[s_end] - out0);
    if (n_read == 0) return;
    if (n_read > (uint32_t)kPresortMax && n_read <= (uint32_t)presort_limit(WIDE)) return;   // ordered by a whole CTA: extract_long_kernel
    const bool presort = n_read <= (uint32_t)kPresortMax && n_read <= cap;      // cap >= min(largest read, kPresortMax): see the launcher
    const int K = p.K;
    const uint64_t mask_lo = WIDE ? ~0ull : ((1ull << (5 * K)) - 1);
    const uint64_t mask_hi = WIDE ? ((1ull << (5 * K - 64)) - 1) : 0ull;
    const int j0 = (lane / 3) * (3 * kRoll) + (lane % 3);
    uint32_t roff = 0;                                                            // k-mers of the read emitted so far
    for (uint32_t s = s_begin; s < s_end; ++s) {
        const uint32_t L = seqs.seq_len[s];
        const uint64_t start = ASCII ? astart[s] : seqs.seq_start[s];
        const uint32_t n = (uint32_t)kmer_count((uint64_t)L + 3u * (K - p.k_low), K, p.one);
        if (n == 0) continue;
        const uint32_t npos = p.one ? 3u * n : n;                 // base offsets walked: --one keeps every third of them (frame 0)
        for (int strand = 0; strand < p.strands; ++strand) {
            for (uint32_t p0 = 0; p0 < npos; p0 += kSegR) {
                const uint32_t cnt = min((uint32_t)kSegR, npos - p0);
                const uint32_t ocnt = p.one ? cnt / 3u : cnt;      // k-mers of this segment (kSegR and npos are multiples of three)
                const uint32_t naa = cnt + 3 * (K - 1);
                for (uint32_t j = lane; j < naa + 2; j += 32) {
                    const uint32_t pos = p0 + j;
                    uint32_t c;
                    if (!ASCII) c = code_at(seqs.bases2, seqs.nmask, start, L, strand, pos);
                    else if (pos >= L) c = 4u;                                      // marker / padding 'X'
                    else {
                        const char ch = ascii[start + (strand ? (L - 1 - pos) : pos)];
                        // Read.hpp:657-675: only ACGTacgt survive, everything else becomes 'Z'
                        const bool ok = (ch == 'A') | (ch == 'C') | (ch == 'G') | (ch == 'T') | (ch == 'a') | (ch == 'c') | (ch == 'g') | (ch == 't');
                        const uint32_t two = (uint32_t)(ch & 14) >> 1;              // kASA.hpp:70-75: A0 C1 T2 G3
                        c = ok ? (strand ? (two ^ 2u) : two) : 5u;                  // complement A0<->T2, C1<->G3 (kASA.hpp:54)
                    }
                    code[j] = (uint8_t)c;
                }
                __syncwarp();
                for (uint32_t j = lane; j < naa; j += 32) aa[j] = s_tab[(uint32_t)code[j] << 6 | (uint32_t)code[j + 1] << 3 | code[j + 2]];
                __syncwarp();
                // presorted reads collect all their k-mers; the others pass segment by segment through the first kSegR slots
                uint64_t *dlo = presort ? klo + roff : klo, *dhi = presort ? khi + roff : khi;
                if (lane < 30 && (uint32_t)j0 < cnt && (!p.one || lane % 3 == 0)) {
                    uint64_t lo = 0, hi = 0;
                    for (int i = 0; i < K; ++i) {
                        if (WIDE) hi = (hi << 5) | (lo >> 59);
                        lo = (lo << 5) | aa[j0 + 3 * i];
                    }
                    const int d0 = p.one ? j0 / 3 : j0;
                    dlo[d0] = lo;
                    if (WIDE) dhi[d0] = hi;
#pragma unroll
                    for (int t = 1; t < kRoll; ++t) {
                        const int j = j0 + 3 * t;
                        if ((uint32_t)j >= cnt) break;
                        if (WIDE) hi = ((hi << 5) | (lo >> 59)) & mask_hi;
                        lo = ((lo << 5) | aa[j + 3 * (K - 1)]) & mask_lo;
                        const int d = p.one ? j / 3 : j;
                        dlo[d] = lo;
                        if (WIDE) dhi[d] = hi;
                    }
                }
                __syncwarp();
                if (!presort) {
                    const uint64_t obase = out0 + roff;
                    for (uint32_t j = lane; j < ocnt; j += 32) {
                        const uint64_t o = obase + j;
                        key_lo[o] = klo[j];
                        if (WIDE) key_hi[o] = khi[j];
                        inst[o] = (uint32_t)o;
                    }
                    __syncwarp();
                }
                roff += ocnt;
            }
        }
    }
    if (!presort) return;
    // ---- order the read's k-mers: bucket on the leading bits of the key, then rank inside the bucket
    const uint32_t m = roff;
    const int bsh = WIDE ? 5 * K - 64 - kSortBucketBits : 5 * K - kSortBucketBits;
    {
        uint4 *z = reinterpret_cast<uint4 *>(bcnt);
        for (uint32_t i = lane; i < kSortBuckets / 4 + 1; i += 32) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncwarp();
    for (uint32_t x = lane; x < m; x += 32) {
        const uint32_t bk = (uint32_t)((WIDE ? khi[x] : klo[x]) >> bsh) & (kSortBuckets - 1);
        ba[x] = bk | (atomicAdd(&bcnt[bk], 1u) << kSortBucketBits);                           // bucket | arrival index inside the bucket
    }
    __syncwarp();
    {   // exclusive scan of the bucket counts
        constexpr uint32_t per = kSortBuckets / 32;
        uint32_t c[per], sum = 0;
#pragma unroll
        for (uint32_t i = 0; i < per; i += 4) {
            const uint4 v = *reinterpret_cast<const uint4 *>(bcnt + lane * per + i);
            c[i] = v.x; c[i + 1] = v.y; c[i + 2] = v.z; c[i + 3] = v.w;
            sum += v.x + v.y + v.z + v.w;
        }
        uint32_t incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(kFull, incl, d);
            if (lane >= d) incl += y;
        }
        uint32_t run = incl - sum;
#pragma unroll
        for (uint32_t i = 0; i < per; i += 4) {
            uint4 v;
            v.x = run; run += c[i]; v.y = run; run += c[i + 1]; v.z = run; run += c[i + 2]; v.w = run; run += c[i + 3];
            *reinterpret_cast<uint4 *>(bcnt + lane * per + i) = v;
        }
        if (lane == 31) bcnt[kSortBuckets] = run;
    }
    __syncwarp();
    // the k-mers grouped by bucket, contiguous (second copy)
    for (uint32_t x = lane; x < m; x += 32) {
        const uint32_t e = ba[x], g = bcnt[e & (kSortBuckets - 1)] + (e >> kSortBucketBits);
        glo[g] = klo[x];
        if (WIDE) ghi[g] = khi[x];
    }
    __syncwarp();
    // rank inside the bucket by counting; the first copy receives the k-mers in their final order
    for (uint32_t g = lane; g < m; g += 32) {
        const uint64_t xl = glo[g], xh = WIDE ? ghi[g] : 0ull;
        const uint32_t bk = (uint32_t)((WIDE ? xh : xl) >> bsh) & (kSortBuckets - 1);
        const uint32_t b0 = bcnt[bk], b1 = bcnt[bk + 1];
        // the high 32-bit word of the key decides almost always; only equal high words need the whole k-mer
        const uint32_t *top = reinterpret_cast<const uint32_t *>(WIDE ? ghi : glo);
        const uint32_t xt = top[2 * g + 1];
        uint32_t less = 0;
        for (uint32_t y = b0; y < b1; ++y) {
            const uint32_t ot = top[2 * y + 1];
            less += ot < xt ? 1u : 0u;
            if (ot == xt && y != g) {
                const uint64_t ol = glo[y];
                bool lt;
                if (WIDE) { const uint64_t oh = ghi[y]; lt = oh != xh ? oh < xh : (ol != xl ? ol < xl : y < g); }
                else lt = ol != xl ? ol < xl : y < g;              // equal k-mers: any fixed order (their records are identical)
                less += lt ? 1u : 0u;
            }
        }
        klo[b0 + less] = xl;
        if (WIDE) khi[b0 + less] = xh;
    }
    __syncwarp();
    for (uint32_t j = lane; j < m; j += 32) {
        const uint64_t o = out0 + j;
        key_lo[o] = klo[j];
        if (WIDE) key_hi[o] = khi[j];
        inst[o] = (uint32_t)o;
    }
}

// ---- reads of more than kPresortMax and at most presort_limit() k-mers: ONE CTA PER READ.  The warps extract the read's segments
// side by side into one shared-memory array; the CTA then orders it like the warp kernel does, with 4096 buckets on the leading
// twelve key bits and the read's own output region in global memory (L2-resident: <= 192 KB) as the bucket-grouped second copy:
//   count -> scan -> group into global -> rank inside the bucket (comparing against the grouped copy) into shared memory -> store.
constexpr int kLongWarps = 16;                 // one CTA per SM (192 KB of keys): as many warps as the rest of shared memory allows
constexpr int kLongBucketBits = 12;
constexpr int kLongBuckets = 1 << kLongBucketBits;

template <bool WIDE, bool ASCII>
__global__ void __launch_bounds__(kLongWarps * 32) extract_long_kernel(PackedSeqs seqs, const char *__restrict__ ascii,
                                                                       const uint64_t *__restrict__ astart, KmerParams p,
                                                                       const uint64_t *__restrict__ koff, uint64_t *key_lo, uint64_t *key_hi,
                                                                       uint32_t *__restrict__ inst, const uint8_t *__restrict__ codon,
                                                                       const uint32_t *__restrict__ read_first, uint32_t n_reads)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(16) uint8_t s_tab[512];
    __shared__ uint32_t s_scan[kLongWarps];
    const uint32_t r = blockIdx.x;
    const uint32_t s_begin = read_first ? read_first[r] : r, s_end = read_first ? read_first[r + 1] : r + 1;
    const uint64_t out0 = koff[s_begin];
    const uint64_t n64 = koff[s_end] - out0;
    constexpr uint32_t cap = (uint32_t)presort_limit(WIDE);
    if (n64 <= (uint64_t)kPresortMax || n64 > (uint64_t)cap) return;          // the warp kernel's reads
    const uint32_t m = (uint32_t)n64;
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    uint64_t *klo = reinterpret_cast<uint64_t *>(smem_raw);
    uint64_t *khi = klo + (WIDE ? cap : 0);
    uint32_t *bcnt = reinterpret_cast<uint32_t *>(khi + cap);                  // [kLongBuckets + 4]
    uint8_t *code = reinterpret_cast<uint8_t *>(bcnt + kLongBuckets + 4) + (size_t)w * 2 * (kSegR + kAaPad + 8);
    uint8_t *aa = code + (kSegR + kAaPad + 8);
    if (tid < 32) reinterpret_cast<uint4 *>(s_tab)[tid] = reinterpret_cast<const uint4 *>(codon)[tid];
    for (int i = tid; i < kLongBuckets + 4; i += blockDim.x) bcnt[i] = 0;
    __syncthreads();
    // ---- the read's k-mers into shared memory, segment g by warp g % kLongWarps (same steps as extract_identify_kernel)
    const int K = p.K;
    const uint64_t mask_lo = WIDE ? ~0ull : ((1ull << (5 * K)) - 1);
    const uint64_t mask_hi = WIDE ? ((1ull << (5 * K - 64)) - 1) : 0ull;
    const int j0 = (lane / 3) * (3 * kRoll) + (lane % 3);
    uint32_t roff = 0, g = 0;
    for (uint32_t s = s_begin; s < s_end; ++s) {
        const uint32_t L = seqs.seq_len[s];
        const uint64_t start = ASCII ? astart[s] : seqs.seq_start[s];
        const uint32_t n = (uint32_t)kmer_count((uint64_t)L + 3u * (K - p.k_low), K, p.one);
        const uint32_t npos = p.one ? 3u * n : n;
        for (int strand = 0; strand < p.strands && n; ++strand) {
            for (uint32_t p0 = 0; p0 < npos; p0 += kSegR, ++g) {
                const uint32_t cnt = min((uint32_t)kSegR, npos - p0);
                const uint32_t ocnt = p.one ? cnt / 3u : cnt;
                if ((int)(g % kLongWarps) == w) {
                    const uint32_t naa = cnt + 3 * (K - 1);
                    for (uint32_t j = lane; j < naa + 2; j += 32) {
                        const uint32_t pos = p0 + j;
                        uint32_t c;
                        if (!ASCII) c = code_at(seqs.bases2, seqs.nmask, start, L, strand, pos);
                        else if (pos >= L) c = 4u;
                        else {
                            const char ch = ascii[start + (strand ? (L - 1 - pos) : pos)];
                            const bool ok = (ch == 'A') | (ch == 'C') | (ch == 'G') | (ch == 'T') | (ch == 'a') | (ch == 'c') | (ch == 'g') | (ch == 't');
                            const uint32_t two = (uint32_t)(ch & 14) >> 1;
                            c = ok ? (strand ? (two ^ 2u) : two) : 5u;
                        }
                        code[j] = (uint8_t)c;
                    }
                    __syncwarp();
                    for (uint32_t j = lane; j < naa; j += 32) aa[j] = s_tab[(uint32_t)code[j] << 6 | (uint32_t)code[j + 1] << 3 | code[j + 2]];
                    __syncwarp();
                    if (lane < 30 && (uint32_t)j0 < cnt && (!p.one || lane % 3 == 0)) {
                        uint64_t lo = 0, hi = 0;
                        for (int i = 0; i < K; ++i) {
                            if (WIDE) hi = (hi << 5) | (lo >> 59);
                            lo = (lo << 5) | aa[j0 + 3 * i];
                        }
                        const int d0 = p.one ? j0 / 3 : j0;
                        klo[roff + d0] = lo;
                        if (WIDE) khi[roff + d0] = hi;
#pragma unroll
                        for (int t = 1; t < kRoll; ++t) {
                            const int j = j0 + 3 * t;
                            if ((uint32_t)j >= cnt) break;
                            if (WIDE) hi = ((hi << 5) | (lo >> 59)) & mask_hi;
                            lo = ((lo << 5) | aa[j + 3 * (K - 1)]) & mask_lo;
                            const int d = p.one ? j / 3 : j;
                            klo[roff + d] = lo;
                            if (WIDE) khi[roff + d] = hi;
                        }
                    }
                    __syncwarp();
                }
                roff += ocnt;
            }
        }
    }
    __syncthreads();
    // ---- order: count, scan, group (into the read's output region), rank, store
    const int bsh = WIDE ? 5 * K - 64 - kLongBucketBits : 5 * K - kLongBucketBits;
    auto bucket_of = [&](uint64_t lo, uint64_t hi) -> uint32_t { return (uint32_t)((WIDE ? hi : lo) >> bsh) & (kLongBuckets - 1); };
    for (uint32_t x = tid; x < m; x += blockDim.x) atomicAdd(&bcnt[bucket_of(klo[x], WIDE ? khi[x] : 0ull)], 1u);
    __syncthreads();
    {   // exclusive scan of the bucket counts: every thread owns kLongBuckets / 256 consecutive counters
        constexpr uint32_t per = kLongBuckets / (kLongWarps * 32);
        uint32_t c[per], sum = 0;
#pragma unroll
        for (uint32_t i = 0; i < per; ++i) { c[i] = bcnt[tid * per + i]; sum += c[i]; }
        uint32_t incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(kFull, incl, d);
            if (lane >= d) incl += y;
        }
        if (lane == 31) s_scan[w] = incl;
        __syncthreads();
        uint32_t base = 0;
        for (int i = 0; i < w; ++i) base += s_scan[i];
        uint32_t run = base + incl - sum;
#pragma unroll
        for (uint32_t i = 0; i < per; ++i) { bcnt[tid * per + i] = run; run += c[i]; }
    }
    __syncthreads();
    // group: bcnt[b] walks from the start of bucket b to its end (= the start of bucket b + 1)
    for (uint32_t x = tid; x < m; x += blockDim.x) {
        const uint64_t lo = klo[x], hi = WIDE ? khi[x] : 0ull;
        const uint32_t pos = atomicAdd(&bcnt[bucket_of(lo, hi)], 1u);
        key_lo[out0 + pos] = lo;
        if (WIDE) key_hi[out0 + pos] = hi;
    }
    __syncthreads();                                                            // (also makes the CTA's global writes visible to the CTA)
    for (uint32_t gpos = tid; gpos < m; gpos += blockDim.x) {
        const uint64_t xl = __ldcg(key_lo + out0 + gpos), xh = WIDE ? __ldcg(key_hi + out0 + gpos) : 0ull;
        const uint32_t bk = bucket_of(xl, xh);
        const uint32_t b0 = bk ? bcnt[bk - 1] : 0u, b1 = bcnt[bk];
        uint32_t less = 0;
        for (uint32_t y = b0; y < b1; ++y) {
            const uint64_t ol = __ldcg(key_lo + out0 + y), oh = WIDE ? __ldcg(key_hi + out0 + y) : 0ull;
            const bool lt = WIDE && oh != xh ? oh < xh : (ol != xl ? ol < xl : y < gpos);       // equal k-mers: any fixed order
            less += lt ? 1u : 0u;
        }
        klo[b0 + less] = xl;
        if (WIDE) khi[b0 + less] = xh;
    }
    __syncthreads();
    for (uint32_t j = tid; j < m; j += blockDim.x) {
        const uint64_t o = out0 + j;
        key_lo[o] = klo[j];
        if (WIDE) key_hi[o] = khi[j];
        inst[o] = (uint32_t)o;
    }
}

__global__ void build_counts_kernel(const uint32_t *__restrict__ seq_len, uint32_t n, int K, int one, uint64_t *__restrict__ out)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t range = seq_len[i] >= (uint32_t)(3 * K) ? (uint64_t)seq_len[i] - 3 * K + 1 : 0;   // Read.hpp:1995
    out[i] = one ? (range + 2) / 3 : range;             // --one: the offsets 0, 3, 6, ... < range (Read.hpp:2143-2200)
}

template <bool WIDE>
__global__ void __launch_bounds__(256) extract_build_kernel(const char *__restrict__ ascii, const uint64_t *__restrict__ seq_start,
                                                            const uint32_t *__restrict__ taxid, uint32_t n_seqs, int K, int one,
                                                            const uint64_t *__restrict__ woff, uint64_t n_windows,
                                                            uint64_t *__restrict__ key_lo, uint64_t *__restrict__ key_hi,
                                                            uint32_t *__restrict__ val, const uint8_t *__restrict__ codon)
{
    __shared__ __align__(16) uint8_t s_tab[512];
    if (threadIdx.x < 32) reinterpret_cast<uint4 *>(s_tab)[threadIdx.x] = reinterpret_cast<const uint4 *>(codon)[threadIdx.x];
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t f = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; f < n_windows; f += stride) {
        // sequence of this window: last s with woff[s] <= f
        uint32_t lo_s = 0, hi_s = n_seqs;
        while (hi_s - lo_s > 1) {
            uint32_t mid = (lo_s + hi_s) >> 1;
            if (__ldg(woff + mid) <= f) lo_s = mid; else hi_s = mid;
        }
        const char *ptr = ascii + seq_start[lo_s] + (f - __ldg(woff + lo_s)) * (one ? 3u : 1u);
        uint64_t lo = 0, hi = 0;
        bool bad = false;
        for (int i = 0; i < K; ++i) {
            uint32_t c0 = (uint8_t)ptr[3 * i], c1 = (uint8_t)ptr[3 * i + 1], c2 = (uint8_t)ptr[3 * i + 2];
            uint32_t a = s_tab[((c0 & 14u) << 5) | ((c1 & 14u) << 2) | ((c2 & 14u) >> 1)];   // kASA.hpp:75
            bad |= (a == kLetterUnder);
            if (WIDE) hi = (hi << 5) | (lo >> 59);
            lo = (lo << 5) | a;
        }
        if (bad || (lo == 0 && hi == 0)) { lo = ~0ull; hi = ~0ull; }
        key_lo[f] = lo;
        if (WIDE) key_hi[f] = hi;
        val[f] = taxid[lo_s];
    }
}

}  // namespace

void launch_pack_sizes(cudaStream_t s, const uint32_t *seq_len, uint32_t n_seqs, uint64_t *pstart, uint64_t *launches)
{
    if (!n_seqs) return;
    pack_sizes_kernel<<<div_up(n_seqs, 256u), 256, 0, s>>>(seq_len, n_seqs, pstart);
    ++*launches;
}

void launch_pack_ascii(cudaStream_t s, const char *ascii, const uint64_t *ascii_start, const uint32_t *seq_len,
                       uint32_t n_seqs, const uint64_t *pstart, uint32_t *bases2, uint32_t *nmask, uint64_t *launches)
{
    if (!n_seqs) return;
    pack_ascii_kernel<<<div_up(n_seqs, 8u), 256, 0, s>>>(ascii, ascii_start, seq_len, n_seqs, pstart, bases2, nmask);
    ++*launches;
}

void launch_kmer_counts(cudaStream_t s, const uint32_t *seq_len, uint32_t n_seqs, KmerParams p, uint64_t *koff, uint64_t *launches)
{
    if (!n_seqs) return;
    kmer_counts_kernel<<<div_up(n_seqs, 256u), 256, 0, s>>>(seq_len, n_seqs, p, koff);
    ++*launches;
}

// shared memory per warp follows the largest read of the batch (a resource choice only: which reads are presorted depends on
// their own length, never on the batch)
template <class Kern>
static void launch_extract_kernel(Kern kern, cudaStream_t s, bool wide, PackedSeqs seqs, const char *ascii, const uint64_t *astart, KmerParams p,
                                  const uint8_t *codon, const uint64_t *koff, uint64_t *key_lo, uint64_t *key_hi, uint32_t *inst,
                                  const uint32_t *read_first, uint32_t n_reads, uint32_t max_inst)
{
    uint32_t cap = std::min<uint32_t>(std::max<uint32_t>(max_inst, (uint32_t)kSegR), (uint32_t)kPresortMax);
    cap = (cap + 7u) & ~7u;
    const size_t wbytes = extract_warp_bytes(wide, cap);
    uint32_t warps = kExtractWarps;                                     // fewer warps per CTA when the reads are long (shared memory)
    while (warps > 1 && wbytes * warps > (size_t)200 * 1024) warps >>= 1;
    const size_t sm = wbytes * warps;
    KASA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    kern<<<div_up(n_reads, warps), warps * 32, sm, s>>>(seqs, ascii, astart, p, koff, key_lo, key_hi, inst, codon, read_first, n_reads, cap, (uint32_t)wbytes);
}

// reads between kPresortMax and presort_limit() k-mers: one CTA each (launched only when the batch has such reads; CTAs of other reads leave at once)
template <class Kern>
static void launch_long_kernel(Kern kern, cudaStream_t s, bool wide, PackedSeqs seqs, const char *ascii, const uint64_t *astart, KmerParams p,
                               const uint8_t *codon, const uint64_t *koff, uint64_t *key_lo, uint64_t *key_hi, uint32_t *inst,
                               const uint32_t *read_first, uint32_t n_reads)
{
    const size_t sm = (size_t)presort_limit(wide) * (wide ? 16 : 8) + (size_t)4 * (kLongBuckets + 4) + (size_t)kLongWarps * 2 * (kSegR + kAaPad + 8);
    KASA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    kern<<<n_reads, kLongWarps * 32, sm, s>>>(seqs, ascii, astart, p, koff, key_lo, key_hi, inst, codon, read_first, n_reads);
}

void launch_extract_identify(cudaStream_t s, PackedSeqs seqs, KmerParams p, const uint8_t *codon, const uint64_t *koff, uint64_t *key_lo,
                             uint64_t *key_hi, uint32_t *inst, const uint32_t *read_first, uint32_t n_reads, uint32_t max_inst, uint64_t *launches)
{
    if (!seqs.n_seqs || !n_reads) return;
    if (p.K > 12) launch_extract_kernel(extract_identify_kernel<true, false>, s, true, seqs, nullptr, nullptr, p, codon, koff, key_lo, key_hi, inst, read_first, n_reads, max_inst);
    else launch_extract_kernel(extract_identify_kernel<false, false>, s, false, seqs, nullptr, nullptr, p, codon, koff, key_lo, key_hi, inst, read_first, n_reads, max_inst);
    ++*launches;
    if (max_inst > (uint32_t)kPresortMax) {
        if (p.K > 12) launch_long_kernel(extract_long_kernel<true, false>, s, true, seqs, nullptr, nullptr, p, codon, koff, key_lo, key_hi, inst, read_first, n_reads);
        else launch_long_kernel(extract_long_kernel<false, false>, s, false, seqs, nullptr, nullptr, p, codon, koff, key_lo, key_hi, inst, read_first, n_reads);
        ++*launches;
    }
}

void launch_extract_identify_ascii(cudaStream_t s, const char *ascii, const uint64_t *ascii_start, const uint32_t *seq_len,
                                   uint32_t n_seqs, KmerParams p, const uint8_t *codon, const uint64_t *koff, uint64_t *key_lo,
                                   uint64_t *key_hi, uint32_t *inst, const uint32_t *read_first, uint32_t n_reads, uint32_t max_inst,
                                   uint64_t *launches)
{
    if (!n_seqs || !n_reads) return;
    PackedSeqs seqs{nullptr, nullptr, nullptr, seq_len, n_seqs};
    if (p.K > 12) launch_extract_kernel(extract_identify_kernel<true, true>, s, true, seqs, ascii, ascii_start, p, codon, koff, key_lo, key_hi, inst, read_first, n_reads, max_inst);
    else launch_extract_kernel(extract_identify_kernel<false, true>, s, false, seqs, ascii, ascii_start, p, codon, koff, key_lo, key_hi, inst, read_first, n_reads, max_inst);
    ++*launches;
    if (max_inst > (uint32_t)kPresortMax) {
        if (p.K > 12) launch_long_kernel(extract_long_kernel<true, true>, s, true, seqs, ascii, ascii_start, p, codon, koff, key_lo, key_hi, inst, read_first, n_reads);
        else launch_long_kernel(extract_long_kernel<false, true>, s, false, seqs, ascii, ascii_start, p, codon, koff, key_lo, key_hi, inst, read_first, n_reads);
        ++*launches;
    }
}

void launch_build_counts(cudaStream_t s, const uint32_t *seq_len, uint32_t n_seqs, int K, int one, uint64_t *woff, uint64_t *launches)
{
    if (!n_seqs) return;
    build_counts_kernel<<<div_up(n_seqs, 256u), 256, 0, s>>>(seq_len, n_seqs, K, one, woff);
    ++*launches;
}

void launch_extract_build(cudaStream_t s, const char *ascii, const uint64_t *seq_start, const uint32_t *seq_len,
                          const uint32_t *taxid, uint32_t n_seqs, int K, int one, const uint8_t *codon, const uint64_t *woff, uint64_t n_windows,
                          uint64_t *key_lo, uint64_t *key_hi, uint32_t *val, uint64_t *launches)
{
    (void)seq_len;
    if (!n_windows) return;
    uint64_t want = div_up<uint64_t>(n_windows, 256);
    unsigned grid = (unsigned)(want < (uint64_t)kSMs * 32 ? want : (uint64_t)kSMs * 32);
    if (K > 12)
        extract_build_kernel<true><<<grid, 256, 0, s>>>(ascii, seq_start, taxid, n_seqs, K, one, woff, n_windows, key_lo, key_hi, val, codon);
    else
        extract_build_kernel<false><<<grid, 256, 0, s>>>(ascii, seq_start, taxid, n_seqs, K, one, woff, n_windows, key_lo, key_hi, val, codon);
    ++*launches;
}

}  // namespace kasa
