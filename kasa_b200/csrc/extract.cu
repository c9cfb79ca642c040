// extract.cu — see extract.cuh.
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

constexpr int kExtractWarps = 8;
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

__host__ __device__ __forceinline__ uint64_t kmer_count(uint64_t len, int K)
{
    return len > (uint64_t)(3 * K + 1) ? len - 3 * K + 1 : 0;   // Read.hpp:50-52
}

__global__ void kmer_counts_kernel(const uint32_t *__restrict__ seq_len, uint32_t n, KmerParams p, uint64_t *__restrict__ out)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (uint64_t)p.strands * kmer_count((uint64_t)seq_len[i] + 3u * (p.K - p.k_low), p.K);
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

// One warp per sequence.  For each strand and each segment of kSegR output positions:
//   1. base codes of the prepared string (read / reverse complement + 'X' marker, Read.hpp:612-695) into shared
//      memory — straight from the ASCII bytes (coalesced byte loads, cleaning of Read.hpp:657-675 on the fly) or
//      from the 2-bit planes;
//   2. letters: aa[j] = codon table[c_j, c_j+1, c_j+2];
//   3. rolling k-mers: lane l owns 8 positions of one reading frame (j0, j0+3, ...): the first k-mer is assembled
//      from K letters, each following one is (kmer << 5 | next letter) & mask — 3 instructions instead of 3 K;
//   4. the segment's k-mers leave through shared memory as coalesced 8-byte stores; inst[i] = i.
constexpr int kRoll = 5;                         // same-frame positions per lane
constexpr int kSegR = 30 * kRoll;                // 150 output positions per segment: one segment per strand of a 150 bp read

template <bool WIDE, bool ASCII>
__global__ void __launch_bounds__(kExtractWarps * 32) extract_identify_kernel(PackedSeqs seqs, const char *__restrict__ ascii,
                                                                              const uint64_t *__restrict__ astart, KmerParams p,
                                                                              const uint64_t *__restrict__ koff,
                                                                              uint64_t *__restrict__ key_lo,
                                                                              uint64_t *__restrict__ key_hi,
                                                                              uint32_t *__restrict__ inst,
                                                                              const uint8_t *__restrict__ codon)
{
    __shared__ uint8_t s_code[kExtractWarps][kSegR + kAaPad + 8];
    __shared__ uint8_t s_aa[kExtractWarps][kSegR + kAaPad];
    __shared__ uint64_t s_klo[kExtractWarps][kSegR];
    __shared__ uint64_t s_khi[WIDE ? kExtractWarps : 1][WIDE ? kSegR : 1];
    __shared__ __align__(16) uint8_t s_tab[512];
    if (threadIdx.x < 32) reinterpret_cast<uint4 *>(s_tab)[threadIdx.x] = reinterpret_cast<const uint4 *>(codon)[threadIdx.x];
    __syncthreads();
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t s = blockIdx.x * kExtractWarps + w;
    if (s >= seqs.n_seqs) return;
    const uint32_t L = seqs.seq_len[s];
    const uint64_t start = ASCII ? astart[s] : seqs.seq_start[s];
    const int K = p.K;
    const uint32_t n = (uint32_t)kmer_count((uint64_t)L + 3u * (K - p.k_low), K);
    if (n == 0) return;
    const uint64_t out0 = koff[s];
    uint8_t *code = s_code[w], *aa = s_aa[w];
    const uint64_t mask_lo = WIDE ? ~0ull : ((1ull << (5 * K)) - 1);
    const uint64_t mask_hi = WIDE ? ((1ull << (5 * K - 64)) - 1) : 0ull;
    const int j0 = (lane / 3) * (3 * kRoll) + (lane % 3);
    for (int strand = 0; strand < p.strands; ++strand) {
        const uint64_t obase = out0 + (uint64_t)strand * n;
        for (uint32_t p0 = 0; p0 < n; p0 += kSegR) {
            const uint32_t cnt = min((uint32_t)kSegR, n - p0);
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
            if (lane < 30 && (uint32_t)j0 < cnt) {
                uint64_t lo = 0, hi = 0;
                for (int i = 0; i < K; ++i) {
                    if (WIDE) hi = (hi << 5) | (lo >> 59);
                    lo = (lo << 5) | aa[j0 + 3 * i];
                }
                s_klo[w][j0] = lo;
                if (WIDE) s_khi[w][j0] = hi;
#pragma unroll
                for (int t = 1; t < kRoll; ++t) {
                    const int j = j0 + 3 * t;
                    if ((uint32_t)j >= cnt) break;
                    if (WIDE) hi = ((hi << 5) | (lo >> 59)) & mask_hi;
                    lo = ((lo << 5) | aa[j + 3 * (K - 1)]) & mask_lo;
                    s_klo[w][j] = lo;
                    if (WIDE) s_khi[w][j] = hi;
                }
            }
            __syncwarp();
            for (uint32_t j = lane; j < cnt; j += 32) {
                const uint64_t o = obase + p0 + j;
                key_lo[o] = s_klo[w][j];
                if (WIDE) key_hi[o] = s_khi[w][j];
                inst[o] = (uint32_t)o;
            }
            __syncwarp();
        }
    }
}

__global__ void build_counts_kernel(const uint32_t *__restrict__ seq_len, uint32_t n, int K, uint64_t *__restrict__ out)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = seq_len[i] >= (uint32_t)(3 * K) ? (uint64_t)seq_len[i] - 3 * K + 1 : 0;   // Read.hpp:1995
}

template <bool WIDE>
__global__ void __launch_bounds__(256) extract_build_kernel(const char *__restrict__ ascii, const uint64_t *__restrict__ seq_start,
                                                            const uint32_t *__restrict__ taxid, uint32_t n_seqs, int K,
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
        const char *ptr = ascii + seq_start[lo_s] + (f - __ldg(woff + lo_s));
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

void launch_extract_identify(cudaStream_t s, PackedSeqs seqs, KmerParams p, const uint8_t *codon, const uint64_t *koff, uint64_t *key_lo,
                             uint64_t *key_hi, uint32_t *inst, uint64_t *launches)
{
    if (!seqs.n_seqs) return;
    unsigned grid = div_up(seqs.n_seqs, (uint32_t)kExtractWarps);
    if (p.K > 12)
        extract_identify_kernel<true, false><<<grid, kExtractWarps * 32, 0, s>>>(seqs, nullptr, nullptr, p, koff, key_lo, key_hi, inst, codon);
    else
        extract_identify_kernel<false, false><<<grid, kExtractWarps * 32, 0, s>>>(seqs, nullptr, nullptr, p, koff, key_lo, key_hi, inst, codon);
    ++*launches;
}

void launch_extract_identify_ascii(cudaStream_t s, const char *ascii, const uint64_t *ascii_start, const uint32_t *seq_len,
                                   uint32_t n_seqs, KmerParams p, const uint8_t *codon, const uint64_t *koff, uint64_t *key_lo,
                                   uint64_t *key_hi, uint32_t *inst, uint64_t *launches)
{
    if (!n_seqs) return;
    PackedSeqs seqs{nullptr, nullptr, nullptr, seq_len, n_seqs};
    unsigned grid = div_up(n_seqs, (uint32_t)kExtractWarps);
    if (p.K > 12)
        extract_identify_kernel<true, true><<<grid, kExtractWarps * 32, 0, s>>>(seqs, ascii, ascii_start, p, koff, key_lo, key_hi, inst, codon);
    else
        extract_identify_kernel<false, true><<<grid, kExtractWarps * 32, 0, s>>>(seqs, ascii, ascii_start, p, koff, key_lo, key_hi, inst, codon);
    ++*launches;
}

void launch_build_counts(cudaStream_t s, const uint32_t *seq_len, uint32_t n_seqs, int K, uint64_t *woff, uint64_t *launches)
{
    if (!n_seqs) return;
    build_counts_kernel<<<div_up(n_seqs, 256u), 256, 0, s>>>(seq_len, n_seqs, K, woff);
    ++*launches;
}

void launch_extract_build(cudaStream_t s, const char *ascii, const uint64_t *seq_start, const uint32_t *seq_len,
                          const uint32_t *taxid, uint32_t n_seqs, int K, const uint8_t *codon, const uint64_t *woff, uint64_t n_windows,
                          uint64_t *key_lo, uint64_t *key_hi, uint32_t *val, uint64_t *launches)
{
    (void)seq_len;
    if (!n_windows) return;
    uint64_t want = div_up<uint64_t>(n_windows, 256);
    unsigned grid = (unsigned)(want < (uint64_t)kSMs * 32 ? want : (uint64_t)kSMs * 32);
    if (K > 12)
        extract_build_kernel<true><<<grid, 256, 0, s>>>(ascii, seq_start, taxid, n_seqs, K, woff, n_windows, key_lo, key_hi, val, codon);
    else
        extract_build_kernel<false><<<grid, 256, 0, s>>>(ascii, seq_start, taxid, n_seqs, K, woff, n_windows, key_lo, key_hi, val, codon);
    ++*launches;
}

}  // namespace kasa
