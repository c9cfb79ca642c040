// score.cu — see score.cuh.
#include <algorithm>

#include "score.cuh"

namespace kasa {
namespace {

constexpr uint64_t kNib1 = 0x1111111111111111ull;

// nibble j (j < h) = 1: adding it per (taxon, levels = h) pair counts the level sizes |taxa_j| (<= 9 per nibble for inline records)
__device__ __forceinline__ void nib_below(uint32_t h, uint64_t &lo, uint64_t &hi)
{
    lo = h >= 16 ? kNib1 : (((1ull << (4 * h)) - 1) & kNib1);
    hi = h > 16 ? (((1ull << (4 * (h - 16))) - 1) & kNib1) : 0ull;
}
__device__ __forceinline__ uint32_t nib_get(uint64_t lo, uint64_t hi, uint32_t j)
{
    return (uint32_t)((j < 16 ? lo >> (4 * j) : hi >> (4 * (j - 16))) & 15u);
}

__device__ __forceinline__ uint32_t inv_fx(uint32_t n) { return (uint32_t)((4294967296ull + (n >> 1)) / n); }   // 2^32 / n, n >= 2

// pair k (taxon | levels << 16) of an inline list record / of the up-pairs of a mask record
__device__ __forceinline__ uint32_t rec_pair(const uint64_t (&pw)[3], int k) { return (uint32_t)(pw[k / 3] >> (kPairBits * (k % 3))) & kPairMask; }
__device__ __forceinline__ uint32_t up_pair(const uint64_t (&pw)[3], int k) { return (uint32_t)(pw[1 + k / 3] >> (kPairBits * (k % 3))) & kPairMask; }

// levels of taxon t in a mask record (1 unless one of the up-pairs names it)
__device__ __forceinline__ uint32_t mask_levels(const uint64_t (&pw)[3], uint32_t t)
{
    uint32_t h = 1;
#pragma unroll
    for (int k = 0; k < kRecUpPairs; ++k) { const uint32_t p = up_pair(pw, k); if ((p & 0xffffu) == t) h = p >> 16; }
    return h;
}

// header words of an ordered-stream group from nibble-packed level sizes (levels >= 1 in lo / hi; score.cuh)
__device__ __forceinline__ uint32_t header_word(uint32_t n0, uint64_t lo) { return kOpHeader | ((n0 < kOpNExt ? n0 : kOpNExt) << 21) | (((uint32_t)(lo >> 4) & 0xFFFFFu) << 1); }
__device__ __forceinline__ uint32_t header_extra(int k, uint64_t lo, uint64_t hi)
{
    return k == 0 ? (uint32_t)(lo >> 24) : (k == 1 ? (uint32_t)((lo >> 56) | (hi << 8)) : (uint32_t)(hi >> 24));
}

// ---------------------------------------------------------------------------------------------------------------
// A: order the match records of a read and write the op stream
constexpr int kAMaxWarps = 4;

struct WarpMem {
    ulonglong2 *rlo, *rhi;   // [cap] record words (w0, w1) / (w2, w3) by slot
    uint64_t *key;           // [icap]  sort key << 32 | level << 16 | slot
    uint32_t *bcnt;          // [nb + 1]
    uint32_t *un;            // [num_k][64] unique hits of this warp's reads (COUNTS only)
    uint16_t *tmp, *ord, *arr;   // [icap]
};

__host__ __device__ inline size_t warp_mem_bytes(uint32_t cap, uint32_t icap, uint32_t nb, int count_levels)
{
    size_t b = (size_t)32 * cap + (size_t)8 * icap + (size_t)4 * (nb + 1) + (count_levels ? (size_t)256 * count_levels + 512 : 0) + (size_t)4 * icap;
    return (b + 15) & ~(size_t)15;
}

// nibble-packed sizes of the levels >= 1 of an inline / mask record
__device__ __forceinline__ void level_sizes(const uint64_t (&pw)[3], uint32_t count, bool mask, uint64_t &lo, uint64_t &hi)
{
    lo = hi = 0;
    if (mask) {
#pragma unroll
        for (int k = 0; k < kRecUpPairs; ++k) {
            const uint32_t p = up_pair(pw, k);
            if (p) { uint64_t a, b; nib_below(p >> 16, a, b); lo += a & ~15ull; hi += b; }
        }
    } else {
#pragma unroll
        for (int k = 0; k < kRecInline; ++k) {
            if ((uint32_t)k < count) { uint64_t a, b; nib_below(rec_pair(pw, k) >> 16, a, b); lo += a & ~15ull; hi += b; }
        }
    }
}

// sum of the level counts of a record's pairs (= number of (taxon, level) contributions)
__device__ __forceinline__ uint32_t record_sumh(uint64_t w0, const uint64_t (&pw)[3], const OverflowPeers &peers)
{
    const uint32_t c = rec_count(w0);
    uint32_t sumh = 0;
    if (rec_is_mask(w0)) {
        sumh = c;
#pragma unroll
        for (int k = 0; k < kRecUpPairs; ++k) { const uint32_t p = up_pair(pw, k); if (p) sumh += (p >> 16) - 1; }
    } else if (!rec_overflow(w0)) {
#pragma unroll
        for (int k = 0; k < kRecInline; ++k) if ((uint32_t)k < c) sumh += rec_pair(pw, k) >> 16;
    } else {
        const uint32_t *list = peers.base[rec_owner(w0)] + pw[0];
        for (uint32_t k = 0; k < c; ++k) sumh += __ldg(list + k) >> 16;
    }
    return sumh;
}

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}

// STREAM: 0 = pair stream in rank order, 1 = pair stream in the reference's flush order (EXACT), 2 = mask stream in rank order
// (<= 64 taxa and <= 6 levels: groups of 16 bytes, see accumulate_mask_kernel).
// COUNTS (content files of <= 64 taxa): the per-taxon counts are accumulated here, order-free, lane = taxon (lane l owns the taxa
// l+1 and l+33) over the records the warp holds in shared memory: level k_low hits as 2^-32 fixed point / integers in registers,
// the levels above through packed per-read counters; kernel B then only replays the float32 scores.
template <int STREAM, bool COUNTS>
__global__ void __launch_bounds__(kAMaxWarps * 32) order_ops_kernel(const uint64_t *__restrict__ rec, const uint64_t *__restrict__ read_koff, uint32_t r0,
                                                                    uint32_t r1, uint64_t n_kmers, LevelParams lp, OverflowPeers peers, OpsWorkspace ws,
                                                                    const uint32_t *__restrict__ trig, uint32_t *__restrict__ err_flag, uint32_t cap,
                                                                    uint32_t icap, uint32_t nb, uint32_t wbytes, unsigned long long *__restrict__ cnt_fx,
                                                                    unsigned long long *__restrict__ cnt_unique, bool emit)
{
    constexpr bool EXACT = STREAM == 1, MASKS = STREAM == 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int w = threadIdx.x >> 5;
    const uint32_t lane = threadIdx.x & 31;
    const int T = lp.n_taxa, NK = lp.num_k;
    WarpMem M;
    {
        unsigned char *base = smem_raw + (size_t)w * wbytes;
        M.rlo = reinterpret_cast<ulonglong2 *>(base);
        M.rhi = M.rlo + cap;
        M.key = reinterpret_cast<uint64_t *>(M.rhi + cap);
        M.bcnt = reinterpret_cast<uint32_t *>(M.key + icap);
        M.un = M.bcnt + nb + 1;
        M.tmp = reinterpret_cast<uint16_t *>(M.un + (COUNTS ? 64 * NK + 128 : 0));
        M.ord = M.tmp + icap;
        M.arr = M.ord;          // the arrival index of an item is consumed before its final position is written
    }
    const uint32_t n_warps = blockDim.x >> 5;
    const uint32_t warps_total = gridDim.x * n_warps;
    const uint32_t NX = STREAM == 0 ? (uint32_t)ops_extra_words(NK) : 0u;
    // level k_low: 2^-32 sums and unique hits of this lane's two taxa over all the warp's reads
    unsigned long long fx0a = 0, fx0b = 0, un0a = 0, un0b = 0;
    uint32_t *s_invfx_w = M.un + 64 * NK;                              // [128] 2^32 / n (COUNTS only)
    if (COUNTS) {
        for (int i = lane; i < NK * 64; i += 32) M.un[i] = 0;
        for (int i = lane; i < 128; i += 32) s_invfx_w[i] = i >= 2 ? inv_fx((uint32_t)i) : 0u;
    }
    __syncwarp();

    // words the group of a record takes in the stream (ordered streams)
    auto group_words = [&](uint64_t w0, uint64_t w2, uint64_t w3) -> uint32_t {
        const uint32_t c = rec_count(w0);
        if (!MASKS) return 1 + NX + c;
        uint32_t nup;
        if (rec_is_mask(w0)) {
            nup = 0;
#pragma unroll
            for (int k = 0; k < kRecUpPairs; ++k) nup += ((uint32_t)((k < 3 ? w2 : w3) >> (kPairBits * (k % 3))) & kPairMask) ? 1u : 0u;
        } else {
            nup = c;                                           // list records: every pair is written as an up pair (levels >= 1 too)
        }
        return 4u + (nup > 1u ? 4u * ((nup - 1u + 3u) / 4u) : 0u);
    };

    auto count_slots = [&](uint32_t n_slots) {
        // lane = taxon (lane l owns the taxa l+1 and l+33); warp-uniform walk over the records in shared memory
        unsigned long long upa = 0, upb = 0;                        // unique hits of the levels 1..8, 8-bit fields (flushed every 128 records)
        auto flush_up = [&]() {
            for (int j = 1; j < NK && j <= 8; ++j) {
                const uint32_t ua = (uint32_t)(upa >> (8 * (j - 1))) & 255u, ub = (uint32_t)(upb >> (8 * (j - 1))) & 255u;
                if (ua) M.un[j * 64 + lane] += ua;
                if (ub) M.un[j * 64 + 32 + lane] += ub;
            }
            upa = upb = 0;
        };
        auto add_upper_unique = [&](uint32_t t, uint32_t h) {       // taxon t (0-based) is alone at each of the levels 1 .. h-1
            const uint32_t top = h - 1 < 8u ? h - 1 : 8u;
            const unsigned long long f = (top >= 8u ? ~0ull : ((1ull << (8 * top)) - 1)) & 0x0101010101010101ull;
            if (t >> 5) upb += f; else upa += f;
            for (uint32_t j = 9; j < h; ++j) M.un[j * 64 + t] += 1u;
        };
#pragma unroll 2
        for (uint32_t i = 0; i < n_slots; ++i) {
            const ulonglong2 a = M.rlo[i];
            const uint64_t w0 = a.x;
            const uint32_t c = rec_count(w0), hmax = rec_hmax(w0);
            if (!c) continue;
            if (rec_is_mask(w0)) {
                const uint64_t mk = a.y;
                const uint32_t in0 = (uint32_t)(mk >> lane) & 1u, in1 = (uint32_t)(mk >> (lane + 32)) & 1u;
                const uint32_t one = c == 1u ? 1u : 0u, f = s_invfx_w[c & 127u];
                un0a += in0 & one; un0b += in1 & one;
                fx0a += (unsigned long long)(in0 * f); fx0b += (unsigned long long)(in1 * f);
                if (hmax > 1u) {
                    const ulonglong2 b = M.rhi[i];
                    if ((b.x >> kPairBits) == 0 && b.y == 0) {
                        // the usual case: one taxon above k_low, alone at each of its upper levels
                        const uint32_t p = (uint32_t)b.x & kPairMask, t = (p & 0xffffu) - 1u;
                        if ((t & 31u) == lane) add_upper_unique(t, p >> 16);
                    } else {
                        const uint64_t pw[3] = {mk, b.x, b.y};
                        uint64_t lo, hi;
                        level_sizes(pw, c, true, lo, hi);
#pragma unroll
                        for (int k = 0; k < kRecUpPairs; ++k) {
                            const uint32_t p = up_pair(pw, k);
                            if (!p) break;
                            const uint32_t t = (p & 0xffffu) - 1u, h = p >> 16;
                            if ((t & 31u) == lane) {
                                for (uint32_t j = 1; j < h; ++j) {
                                    const uint32_t nn = nib_get(lo, hi, j);
                                    if (nn == 1u) M.un[j * 64 + t] += 1u;
                                    else atomicAdd(cnt_fx + (size_t)j * T + t + 1, (unsigned long long)s_invfx_w[nn & 127u]);
                                }
                            }
                        }
                    }
                }
            } else {
                // list records (more than six taxa above k_low, or a span read from HBM): pair by pair
                const bool ovf = rec_overflow(w0);
                const ulonglong2 b = M.rhi[i];
                const uint64_t pw[3] = {a.y, b.x, b.y};
                const uint32_t *list = ovf ? peers.base[rec_owner(w0)] + pw[0] : nullptr;
                uint32_t nj[32];
                if (hmax > 1u) {
                    for (uint32_t j = 0; j < hmax; ++j) nj[j] = 0;
                    for (uint32_t k = 0; k < c; ++k) {
                        const uint32_t h = (ovf ? __ldg(list + k) : (uint32_t)(pw[k / 3] >> (kPairBits * (k % 3))) & kPairMask) >> 16;
                        for (uint32_t j = 0; j < h; ++j) ++nj[j];
                    }
                }
                const uint32_t f0 = c >= 2u ? (c < 128u ? s_invfx_w[c] : inv_fx(c)) : 0u;
                for (uint32_t k = 0; k < c; ++k) {
                    const uint32_t e = ovf ? __ldg(list + k) : (uint32_t)(pw[k / 3] >> (kPairBits * (k % 3))) & kPairMask;
                    const uint32_t t = (e & 0xffffu) - 1u, h = e >> 16;
                    if ((t & 31u) != lane) continue;
                    if (c == 1u) { if (t >> 5) un0b += 1; else un0a += 1; }
                    else { if (t >> 5) fx0b += f0; else fx0a += f0; }
                    for (uint32_t j = 1; j < h; ++j) {
                        if (nj[j] == 1u) M.un[j * 64 + t] += 1u;
                        else atomicAdd(cnt_fx + (size_t)j * T + t + 1, (unsigned long long)(nj[j] < 128u ? s_invfx_w[nj[j]] : inv_fx(nj[j])));
                    }
                }
            }
            if ((i & 127u) == 127u) flush_up();
        }
        flush_up();
    };

    // sorts the `m` items in M.key and writes their groups at (o_cur, e_cur), which advance; keys in [lo, lo + range)
    auto sort_emit = [&](uint32_t m, uint64_t lo, uint64_t range, uint64_t &o_cur, uint64_t &e_cur) {
        if (!m) return;
        for (uint32_t i = lane; i <= nb; i += 32) M.bcnt[i] = 0;
        __syncwarp();
        const float fscale = (float)nb / (float)(range ? range : 1) * 0.999999f;
        auto bucket = [&](uint64_t key) {
            const uint64_t pk = key >> 32;
            const float f = (float)(pk > lo ? pk - lo : 0ull) * fscale;
            const uint32_t b = f >= (float)nb ? nb - 1 : (uint32_t)f;
            return b < nb ? b : nb - 1;
        };
        for (uint32_t x = lane; x < m; x += 32) M.arr[x] = (uint16_t)atomicAdd(&M.bcnt[bucket(M.key[x])], 1u);
        __syncwarp();
        {   // exclusive scan of the bucket counts (nb is a multiple of 32)
            const uint32_t per = nb / 32;
            uint32_t sum = 0;
            for (uint32_t i = 0; i < per; ++i) sum += M.bcnt[lane * per + i];
            uint32_t incl = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t y = __shfl_up_sync(kFull, incl, d);
                if ((int)lane >= d) incl += y;
            }
            uint32_t run = incl - sum;
            for (uint32_t i = 0; i < per; ++i) { const uint32_t c = M.bcnt[lane * per + i]; M.bcnt[lane * per + i] = run; run += c; }
            if (lane == 31) M.bcnt[nb] = run;
        }
        __syncwarp();
        for (uint32_t x = lane; x < m; x += 32) M.tmp[M.bcnt[bucket(M.key[x])] + M.arr[x]] = (uint16_t)x;
        __syncwarp();
        for (uint32_t x = lane; x < m; x += 32) {
            const uint64_t kx = M.key[x];
            const uint32_t b = bucket(kx), b0 = M.bcnt[b], b1 = M.bcnt[b + 1];
            uint32_t less = 0;
            for (uint32_t y = b0; y < b1; ++y) less += M.key[M.tmp[y]] < kx ? 1u : 0u;
            M.ord[b0 + less] = (uint16_t)x;
        }
        __syncwarp();
        // ---- emission in sorted order: lane = item
        for (uint32_t p0 = 0; p0 < m; p0 += 32) {
            const uint32_t p = p0 + lane;
            const bool valid = p < m;
            const uint64_t key = valid ? M.key[M.ord[p]] : 0ull;
            const uint32_t slot = (uint32_t)key & 0xffffu, lev = (uint32_t)(key >> 16) & 0xffffu;
            uint32_t c = 0, hmax = 0;
            bool ovf = false, msk = false;
            uint64_t w0 = 0, pw[3] = {0, 0, 0}, nlo = 0, nhi = 0;
            uint32_t opc = 0, extc = 0, nlev = 0;                  // nlev: |taxa_lev| (EXACT)
            const uint32_t *list = nullptr;
            if (valid) {
                const ulonglong2 a = M.rlo[slot], b = M.rhi[slot];
                w0 = a.x; pw[0] = a.y; pw[1] = b.x; pw[2] = b.y;
                c = rec_count(w0); hmax = rec_hmax(w0); ovf = rec_overflow(w0); msk = rec_is_mask(w0);
                if (ovf) list = peers.base[rec_owner(w0)] + pw[0];
                else if (!MASKS) level_sizes(pw, c, msk, nlo, nhi);
                if (!EXACT) {
                    opc = group_words(w0, pw[1], pw[2]);
                    extc = (!MASKS && ovf) ? hmax : 0u;             // overflow records escape every level size (deterministic demand)
                } else {
                    if (!ovf) nlev = lev ? nib_get(nlo, nhi, lev) : c;
                    else for (uint32_t k = 0; k < c; ++k) nlev += (__ldg(list + k) >> 16) > lev ? 1u : 0u;
                    opc = 1 + nlev;
                    extc = ovf ? 1u : 0u;
                }
            }
            uint32_t io = opc, ie = extc;                  // inclusive scans
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t yo = __shfl_up_sync(kFull, io, d), ye = __shfl_up_sync(kFull, ie, d);
                if ((int)lane >= d) { io += yo; ie += ye; }
            }
            const uint32_t tot_o = __shfl_sync(kFull, io, 31), tot_e = __shfl_sync(kFull, ie, 31);
            uint32_t *op = ws.ops + o_cur + (io - opc);
            uint16_t *ex = ws.ext + e_cur + (ie - extc);
            if (valid && opc) {
                if (MASKS) {
                    // 16-byte groups: [header | n0 << 24 | nup << 17][mask lo][mask hi][up pair 0], then the other up pairs four per block
                    uint4 g;
                    if (msk) {
                        uint32_t nup = 0;
#pragma unroll
                        for (int k = 0; k < kRecUpPairs; ++k) nup += up_pair(pw, k) ? 1u : 0u;
                        g = make_uint4(kOpHeader | (c << 24) | (nup << 17), (uint32_t)pw[0], (uint32_t)(pw[0] >> 32), up_pair(pw, 0));
                        *reinterpret_cast<uint4 *>(op) = g;
                        if (nup > 1u) {
                            uint32_t e4[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
                            for (int k = 1; k < kRecUpPairs; ++k) e4[k - 1] = up_pair(pw, k);
                            *reinterpret_cast<uint4 *>(op + 4) = make_uint4(e4[0], e4[1], e4[2], e4[3]);
                            if (nup > 5u) *reinterpret_cast<uint4 *>(op + 8) = make_uint4(e4[4], 0, 0, 0);
                        }
                    } else {
                        // list record of a small taxonomy: the mask from the pairs, every pair as an up pair
                        uint64_t mk = 0;
                        for (uint32_t k = 0; k < c; ++k) {
                            const uint32_t e = ovf ? __ldg(list + k) : (uint32_t)(pw[k / 3] >> (kPairBits * (k % 3))) & kPairMask;
                            mk |= 1ull << ((e & 0xffffu) - 1u);
                        }
                        const uint32_t first = ovf ? __ldg(list) : (uint32_t)pw[0] & kPairMask;
                        *reinterpret_cast<uint4 *>(op) = make_uint4(kOpHeader | (c << 24) | (c << 17), (uint32_t)mk, (uint32_t)(mk >> 32), first);
                        for (uint32_t k = 1; k < c; k += 4) {
                            uint32_t e4[4] = {0, 0, 0, 0};
                            for (uint32_t i = 0; i < 4 && k + i < c; ++i)
                                e4[i] = ovf ? __ldg(list + k + i) : (uint32_t)(pw[(k + i) / 3] >> (kPairBits * ((k + i) % 3))) & kPairMask;
                            *reinterpret_cast<uint4 *>(op + 4 + (k - 1)) = make_uint4(e4[0], e4[1], e4[2], e4[3]);
                        }
                    }
                } else if (!EXACT) {
                    if (!ovf) {
                        *op++ = header_word(c, nlo);
                        for (uint32_t k = 0; k < NX; ++k) *op++ = header_extra((int)k, nlo, nhi);
                        if (msk) {
                            for (uint64_t left = pw[0]; left; left &= left - 1) {
                                const uint32_t t = (uint32_t)__ffsll((long long)left);
                                *op++ = t | (mask_levels(pw, t) << 16);
                            }
                        } else {
#pragma unroll
                            for (int k = 0; k < kRecInline; ++k) {
                                if ((uint32_t)k >= c) break;
                                *op++ = rec_pair(pw, k);                  // already taxon | levels << 16
                            }
                        }
                    } else {
                        // level sizes by counting; all of them go through the ext stream: |taxa_klow|, then the levels 1 .. hmax-1
                        uint16_t nj[32];
                        for (uint32_t j = 0; j < 32; ++j) nj[j] = 0;
                        for (uint32_t k = 0; k < c; ++k) { const uint32_t h = __ldg(list + k) >> 16; for (uint32_t j = 0; j < h; ++j) ++nj[j]; }
                        uint64_t lo2 = 0, hi2 = 0;
                        *ex++ = (uint16_t)c;
                        for (uint32_t j = 1; j < hmax; ++j) {
                            if (j < 16) lo2 |= (uint64_t)kOpNibExt << (4 * j); else hi2 |= (uint64_t)kOpNibExt << (4 * (j - 16));
                            *ex++ = nj[j];
                        }
                        *op++ = header_word(kOpNExt, lo2);
                        for (uint32_t k = 0; k < NX; ++k) *op++ = header_extra((int)k, lo2, hi2);
                        for (uint32_t k = 0; k < c; ++k) *op++ = __ldg(list + k);
                    }
                } else {
                    *op++ = kOpHeader | ((ovf || nlev >= kOpNExt ? kOpNExt : nlev) << 21) | lev;
                    if (ovf) *ex++ = (uint16_t)nlev;
                    if (msk) {
                        for (uint64_t left = pw[0]; left; left &= left - 1) {
                            const uint32_t t = (uint32_t)__ffsll((long long)left);
                            if (mask_levels(pw, t) > lev) *op++ = t;
                        }
                    } else {
                        for (uint32_t k = 0; k < c; ++k) {
                            uint32_t e;
                            if (!ovf) e = (uint32_t)(pw[k / 3] >> (kPairBits * (k % 3))) & kPairMask;
                            else e = __ldg(list + k);
                            if ((e >> 16) > lev) *op++ = e & 0xffffu;
                        }
                    }
                }
            }
            o_cur += tot_o; e_cur += tot_e;
        }
        __syncwarp();
    };

#pragma unroll 1
    for (uint32_t r = r0 + blockIdx.x * n_warps + w; r < r1; r += warps_total) {
        const uint64_t s0 = read_koff[r];
        const uint32_t ni = (uint32_t)(read_koff[r + 1] - s0);
        const bool single = ni <= cap;
        uint32_t m = 0, my_ops = 0, my_ext = 0;
        // demand of one record on the op / ext streams
        auto demand = [&](uint64_t w0, uint64_t w1, uint64_t w2, uint64_t w3) {
            if (!EXACT) { my_ops += group_words(w0, w2, w3); if (!MASKS) my_ext += rec_overflow(w0) ? rec_hmax(w0) : 0u; }
            else { const uint64_t pw[3] = {w1, w2, w3}; my_ops += rec_hmax(w0) + record_sumh(w0, pw, peers); my_ext += rec_overflow(w0) ? rec_hmax(w0) : 0u; }
        };
        if (single) {
            // ---- the whole read into shared memory with asynchronous copies (all its sectors in flight at once), slot = instance
            for (uint32_t x = lane; x < ni; x += 32) {
                const uint64_t *g = rec + (s0 + x) * kRecWords;
                cp_async16(&M.rlo[x], g);
                cp_async16(&M.rhi[x], g + 2);
            }
            asm volatile("cp.async.wait_all;" ::: "memory");
            __syncwarp();
            for (uint32_t x0 = 0; x0 < ni; x0 += 32) {
                const uint32_t x = x0 + lane;
                const bool sel = x < ni && rec_count(M.rlo[x].x) != 0;
                const uint32_t bal = __ballot_sync(kFull, sel);
                if (sel) {
                    const ulonglong2 a = M.rlo[x], b = M.rhi[x];
                    demand(a.x, a.y, b.x, b.y);
                    if (!EXACT) M.key[m + __popc(bal & lanemask_lt())] = ((uint64_t)rec_rank(a.x) << 32) | x;
                }
                m += __popc(bal);
            }
            __syncwarp();
        } else {
            // ---- long read: only the totals now; the records come in rank slices below
            for (uint32_t x0 = 0; x0 < ni; x0 += 32) {
                const uint32_t x = x0 + lane;
                uint64_t w0 = 0, w1 = 0, w2 = 0, w3 = 0;
                if (x < ni) asm volatile("ld.global.nc.L1::no_allocate.v4.b64 {%0, %1, %2, %3}, [%4];" : "=l"(w0), "=l"(w1), "=l"(w2), "=l"(w3) : "l"(rec + (s0 + x) * kRecWords));
                if (rec_count(w0)) demand(w0, w1, w2, w3);
            }
        }
        // records of the rank slice [lo, hi) of a long read into shared memory, compacted; m = their number (may exceed cap: retry)
        auto load_slice = [&](uint64_t lo, uint64_t hi) {
            m = 0;
            for (uint32_t x0 = 0; x0 < ni; x0 += 32) {
                const uint32_t x = x0 + lane;
                uint64_t w0 = 0, w1 = 0, w2 = 0, w3 = 0;
                if (x < ni) asm volatile("ld.global.nc.L1::no_allocate.v4.b64 {%0, %1, %2, %3}, [%4];" : "=l"(w0), "=l"(w1), "=l"(w2), "=l"(w3) : "l"(rec + (s0 + x) * kRecWords));
                const uint32_t rank = rec_rank(w0);
                const bool sel = rec_count(w0) != 0 && (uint64_t)rank >= lo && (uint64_t)rank < hi;
                const uint32_t bal = __ballot_sync(kFull, sel);
                const uint32_t slot = m + __popc(bal & lanemask_lt());
                if (sel && slot < cap) {
                    M.rlo[slot] = make_ulonglong2(w0, w1); M.rhi[slot] = make_ulonglong2(w2, w3);
                    M.key[slot] = ((uint64_t)rank << 32) | slot;
                }
                m += __popc(bal);
            }
            __syncwarp();
        };
        const uint64_t rmax = n_kmers ? n_kmers : 1;
        if (COUNTS && !emit) {
            // profile only (no per-read scores wanted): the counts are all there is to do
            if (single) count_slots(ni);
            else {
                uint64_t lo = 0, width = rmax / (uint64_t)(2ull * ni / cap + 1) + 1;
                while (lo < rmax) {
                    const uint64_t hi = lo + width < rmax ? lo + width : rmax;
                    load_slice(lo, hi);
                    if (m > cap) { width = (hi - lo) >> 1; if (!width) width = 1; continue; }
                    count_slots(m);
                    lo = hi;
                }
            }
            continue;
        }
        uint32_t tot_ops = my_ops, tot_ext = my_ext;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) { tot_ops += __shfl_xor_sync(kFull, tot_ops, d); tot_ext += __shfl_xor_sync(kFull, tot_ext, d); }
        // space in the streams: one atomic per read, asked for before the counts so that its latency is hidden behind them
        uint64_t o_base = 0, e_base = 0;
        if (lane == 0) {
            o_base = atomicAdd(ws.cursors + 0, (unsigned long long)((tot_ops + 3u) & ~3u));
            e_base = tot_ext ? atomicAdd(ws.cursors + 1, (unsigned long long)tot_ext) : 0ull;
        }
        if (COUNTS && single) count_slots(ni);
        o_base = __shfl_sync(kFull, o_base, 0);
        e_base = __shfl_sync(kFull, e_base, 0);
        bool ok = o_base + ((tot_ops + 3u) & ~3u) <= ws.ops_cap && e_base + tot_ext <= ws.ext_cap;
        if (!ok && lane == 0) atomicOr(err_flag, 32u);
        uint32_t n_items = m;
        if (ok && EXACT) {
            if (!single) { ok = false; if (lane == 0) atomicOr(err_flag, 64u); }
            else {
                // items = (instance, level) pairs with at least one taxon; key = flush position
                uint32_t base_i = 0;
                for (uint32_t x0 = 0; x0 < ni && ok; x0 += 32) {
                    const uint32_t x = x0 + lane;
                    const uint32_t hm = x < ni && rec_count(M.rlo[x].x) ? rec_hmax(M.rlo[x].x) : 0u;
                    uint32_t incl = hm;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(kFull, incl, d); if ((int)lane >= d) incl += y; }
                    const uint32_t tot = __shfl_sync(kFull, incl, 31);
                    if (base_i + tot > icap) { ok = false; if (lane == 0) atomicOr(err_flag, 64u); break; }
                    const uint32_t at = base_i + incl - hm;
                    for (uint32_t j = 0; j < hm; ++j)
                        M.key[at + j] = ((uint64_t)trig[(s0 + x) * (uint64_t)NK + j] << 32) | ((uint64_t)j << 16) | x;
                    base_i += tot;
                }
                n_items = base_i;
                __syncwarp();
            }
        }
        if (lane == 0) { ws.ops_start[r] = o_base; ws.ops_count[r] = ok ? tot_ops : 0u; ws.ext_start[r] = e_base; }
        if (!ok) continue;
        uint64_t o_cur = o_base, e_cur = e_base;
        if (single) {
            sort_emit(n_items, 0, EXACT ? 0x100000000ull : rmax, o_cur, e_cur);
        } else {
            // long read: rank slices small enough for shared memory, in ascending order (its records are re-read from L2)
            const uint64_t w0r = rmax / (uint64_t)(2ull * ni / cap + 1);
            const uint64_t width0 = w0r ? w0r : 1;
            uint64_t lo = 0, width = width0;
            while (lo < rmax) {
                const uint64_t hi = lo + width < rmax ? lo + width : rmax;
                load_slice(lo, hi);
                if (m > cap) { width = (hi - lo) >> 1; if (!width) width = 1; continue; }
                if (COUNTS) count_slots(m);
                sort_emit(m, lo, hi - lo, o_cur, e_cur);
                lo = hi;
                if (m < cap / 4 && width < rmax) width *= 2;
            }
        }
        // pad the stream to a multiple of four words (B loads 16 bytes at a time)
        if (lane < ((4u - (tot_ops & 3u)) & 3u)) ws.ops[o_base + tot_ops + lane] = 0;
    }
    if (COUNTS) {
        __syncwarp();
        const int ta = (int)lane + 1, tb = (int)lane + 33;
        if (ta < T) {
            if (fx0a) atomicAdd(cnt_fx + ta, fx0a);
            if (un0a) atomicAdd(cnt_unique + ta, un0a);
            for (int j = 1; j < NK; ++j) { const uint32_t u = M.un[j * 64 + ta - 1]; if (u) atomicAdd(cnt_unique + (size_t)j * T + ta, (unsigned long long)u); }
        }
        if (tb < T) {
            if (fx0b) atomicAdd(cnt_fx + tb, fx0b);
            if (un0b) atomicAdd(cnt_unique + tb, un0b);
            for (int j = 1; j < NK; ++j) { const uint32_t u = M.un[j * 64 + tb - 1]; if (u) atomicAdd(cnt_unique + (size_t)j * T + tb, (unsigned long long)u); }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// B for the mask stream (<= 64 taxa, <= 6 levels, rank order): thread per read, a 16-byte group per k-mer instance.
//   [header | n0 << 24 | nup << 17][mask lo][mask hi][up pair 0]   (+ blocks of four further up pairs when nup > 1)
// Every taxon of the mask gets float(w_0 * (1.f / n0)); a taxon named by an up pair (taxon | h << 16) then gets the levels
// 1 .. h-1, each float(w_j * (1.f / n_j)) with n_j = number of up pairs reaching level j.  Scores and the k_low share of the
// counts are two columns of shared-memory cells per thread (nothing is atomic, nothing crosses lanes); the rarer events —
// a k-mer with a single taxon, taxa above k_low — go to small per-CTA tables with shared-memory atomics.
constexpr int kMaskBlock = 128;
constexpr int kFxShift = 8;           // per-thread k_low sums in 2^-24 units (a read's sum fits 32 bits), the totals in 2^-32

template <int SLOTS>
__global__ void __launch_bounds__(kMaskBlock) accumulate_mask_kernel(OpsWorkspace ws, uint32_t r0, uint32_t r1, LevelParams lp, float *__restrict__ scores,
                                                                     unsigned long long *__restrict__ cnt_fx, unsigned long long *__restrict__ cnt_unique)
{
    constexpr int NC = 32 * SLOTS, BT = kMaskBlock;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, T = lp.n_taxa, NK = lp.num_k;
    // [NC][BT] per thread and taxon: .x = score of the thread's read (float bits), .y = sum of 1 / n0 over the read's k-mers with
    // n0 >= 2 in 2^-24 units — one 8-byte shared-memory access reads or writes both
    uint2 *cell = reinterpret_cast<uint2 *>(smem_raw);
    __shared__ float s_inv[72];
    __shared__ uint32_t s_q[72];                                        // 2^24 / n rounded to nearest (0 for n < 2)
    __shared__ uint32_t s_q32[72];                                      // 2^32 / n
    __shared__ unsigned long long s_fx[6 * 64];                         // CTA totals [level][taxon - 1], 2^-32 units
    __shared__ uint32_t s_un[6 * 64];
    __shared__ uint32_t s_hist[64 * 8];                                 // [taxon - 1][h]: k-mers where the taxon is alone at its levels 1 .. h-1
    for (int i = tid; i < NC * BT; i += BT) cell[i] = make_uint2(0u, 0u);
    for (int i = tid; i < 6 * 64; i += BT) { s_fx[i] = 0; s_un[i] = 0; }
    for (int i = tid; i < 64 * 8; i += BT) s_hist[i] = 0;
    if (tid < 72) {
        s_inv[tid] = tid ? __fdiv_rn(1.f, (float)tid) : 0.f;             // Compare.hpp:924: 1.f / numOfEntries, up to 64 taxa
        s_q[tid] = tid >= 2 ? (uint32_t)(((1u << 24) + (tid >> 1)) / tid) : 0u;
        s_q32[tid] = tid >= 2 ? inv_fx((uint32_t)tid) : 0u;
    }
    __syncthreads();
    float wj[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) { const int k = lp.k_low + j; wj[j] = (float)(k * k) / 625.f; }      // Compare.hpp:392

    const uint32_t stride = gridDim.x * BT;
#pragma unroll 1
    for (uint32_t base = r0 + blockIdx.x * BT; base < r1; base += stride) {
        const uint32_t r = base + tid;
        const bool have = r < r1;
        const uint4 *cur = reinterpret_cast<const uint4 *>(ws.ops + (have ? ws.ops_start[r] : 0));
        uint32_t left = have ? ws.ops_count[r] / 4u : 0u;               // 16-byte blocks
        uint4 nxt = make_uint4(0, 0, 0, 0);
        if (left) nxt = __ldg(cur);
        uint32_t since = 0;                                              // groups since the k_low sums were last handed on
        while (__any_sync(kFull, left > 0)) {
            if (left) {
                if (++since == 256u) {
                    // a k-mer adds at most 2^23 units to a k_low sum: long reads hand their sums on before 32 bits could overflow
                    since = 0;
                    for (int c = 0; c < NC && c < T - 1; ++c) {
                        uint2 x = cell[c * BT + tid];
                        if (x.y) { atomicAdd(&s_fx[c], (unsigned long long)x.y << kFxShift); x.y = 0u; cell[c * BT + tid] = x; }
                    }
                }
                const uint4 g = nxt;
                const uint32_t n0 = (g.x >> 24) & 127u, nup = (g.x >> 17) & 127u;
                const uint32_t blocks = 1u + (nup > 1u ? (nup - 1u + 3u) / 4u : 0u);
                if (left > blocks) nxt = __ldg(cur + blocks);            // next group: on its way while this one is applied
                const float v0 = __fmul_rn(wj[0], s_inv[n0]);
                const uint32_t q0 = s_q[n0];
                uint64_t mk = ((uint64_t)g.z << 32) | g.y;
                if (n0 == 1u) {
                    const int c = __ffsll((long long)mk) - 1;
                    uint2 x = cell[c * BT + tid];
                    x.x = __float_as_uint(__fadd_rn(__uint_as_float(x.x), v0));
                    cell[c * BT + tid] = x;
                    atomicAdd(&s_un[c], 1u);
                } else {
                    // four taxa per iteration: distinct cells, independent read-modify-write chains (Compare.hpp:530, level k_low)
                    while (mk) {
                        int c[4];
                        bool on[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) { on[u] = mk != 0; c[u] = on[u] ? __ffsll((long long)mk) - 1 : 0; mk &= mk - 1; }
                        uint2 x[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) x[u] = cell[c[u] * BT + tid];
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (on[u]) cell[c[u] * BT + tid] = make_uint2(__float_as_uint(__fadd_rn(__uint_as_float(x[u].x), v0)), x[u].y + q0);
                    }
                }
                if (nup == 1u) {
                    const uint32_t c = (g.w & 0xffffu) - 1u, h = g.w >> 16;
                    uint2 x = cell[c * BT + tid];
                    float acc = __uint_as_float(x.x);
#pragma unroll
                    for (int j = 1; j < 6; ++j) if ((uint32_t)j < h) acc = __fadd_rn(acc, wj[j]);       // w * (1.f / 1) == w
                    x.x = __float_as_uint(acc);
                    cell[c * BT + tid] = x;
                    atomicAdd(&s_hist[c * 8 + (h & 7u)], 1u);
                } else if (nup > 1u) {
                    // several taxa above k_low: level sizes by counting the up pairs, then pair by pair
                    const uint32_t *ups = reinterpret_cast<const uint32_t *>(cur + 1);
                    uint64_t nlev = 0;                                   // 8 bits per level j = 1..5: number of up pairs with h > j
                    for (uint32_t k = 0; k < nup; ++k) {
                        const uint32_t h = (k ? __ldg(ups + k - 1) : g.w) >> 16;
                        nlev += (h >= 6u ? 0x0101010101ull : ((1ull << (8 * (h - 1))) - 1) & 0x0101010101ull);
                    }
                    for (uint32_t k = 0; k < nup; ++k) {
                        const uint32_t e = k ? __ldg(ups + k - 1) : g.w;
                        const uint32_t c = (e & 0xffffu) - 1u, h = e >> 16;
                        uint2 x = cell[c * BT + tid];
                        float acc = __uint_as_float(x.x);
                        for (uint32_t j = 1; j < h && j < 6; ++j) {
                            const uint32_t nn = (uint32_t)(nlev >> (8 * (j - 1))) & 127u;
                            acc = __fadd_rn(acc, __fmul_rn(wj[j], s_inv[nn]));
                            if (nn == 1u) atomicAdd(&s_un[j * 64 + c], 1u);
                            else atomicAdd(&s_fx[j * 64 + c], (unsigned long long)s_q32[nn]);
                        }
                        x.x = __float_as_uint(acc);
                        cell[c * BT + tid] = x;
                    }
                }
                cur += blocks; left -= blocks;
            }
        }
        // ---- end of the read: scores out, its k_low sums into the CTA's totals
        if (have) {
            for (int c = 0; c < NC && c < T - 1; ++c) {
                const uint2 x = cell[c * BT + tid];
                if (scores) scores[(uint64_t)r * T + c + 1] = __uint_as_float(x.x);
                if (x.x | x.y) cell[c * BT + tid] = make_uint2(0u, 0u);
                if (x.y) atomicAdd(&s_fx[c], (unsigned long long)x.y << kFxShift);
            }
            if (scores) scores[(uint64_t)r * T] = 0.f;
        }
    }
    __syncthreads();
    for (int i = tid; i < NK * 64; i += BT) {
        const int j = i / 64, c = i % 64;
        if (c >= T - 1) continue;
        unsigned long long u = s_un[i];
        if (j >= 1) for (int h = j + 1; h < 8; ++h) u += s_hist[c * 8 + h];     // alone at the levels 1 .. h-1
        if (u) atomicAdd(cnt_unique + (size_t)j * T + c + 1, u);
        if (s_fx[i]) atomicAdd(cnt_fx + (size_t)j * T + c + 1, s_fx[i]);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// B: replay the op streams.  Thread per read; every thread owns a column of float32 cells (one per taxon) in shared memory,
// so additions need neither atomics nor any cross-lane step, and a warp streams 32 reads' words with 16-byte loads.
// (The counts of these small taxonomies are taken by kernel A.)
constexpr int kInvTab = 128;
constexpr int kDenseBlock = 128;

template <int SLOTS>
__global__ void __launch_bounds__(kDenseBlock) accumulate_dense_kernel(OpsWorkspace ws, uint32_t r0, uint32_t r1, LevelParams lp, bool exact_stream,
                                                                       float *__restrict__ scores)
{
    constexpr int NC = 32 * SLOTS, BT = kDenseBlock;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, T = lp.n_taxa, NK = lp.num_k;
    float *cellS = reinterpret_cast<float *>(smem_raw);                 // [NC][BT]
    float *vbuf = cellS + (size_t)NC * BT;                              // [NK][BT] value of each level of the current group when an upper level is shared
    float *s_w = vbuf + (size_t)NK * BT;                                // [32]
    float *s_inv = s_w + 32;                                            // [kInvTab]
    for (int i = tid; i < NC * BT; i += BT) cellS[i] = 0.f;
    if (tid < 32) { const int k = lp.k_low + tid; s_w[tid] = (float)(k * k) / 625.f; }      // Compare.hpp:392
    for (int i = tid; i < kInvTab; i += BT) s_inv[i] = i ? __fdiv_rn(1.f, (float)i) : 0.f;   // Compare.hpp:924: 1.f / numOfEntries
    __syncthreads();
    const uint32_t NX = exact_stream ? 0u : (uint32_t)ops_extra_words(NK);
    auto inv_of = [&](uint32_t n) { return n < (uint32_t)kInvTab ? s_inv[n] : __fdiv_rn(1.f, (float)n); };

    const uint32_t stride = gridDim.x * BT;
#pragma unroll 1
    for (uint32_t base = r0 + blockIdx.x * BT; base < r1; base += stride) {
        const uint32_t r = base + tid;
        const bool have = r < r1;
        uint64_t cur = have ? ws.ops_start[r] : 0;
        uint32_t left = have ? ws.ops_count[r] : 0;
        uint64_t ecur = have ? ws.ext_start[r] : 0;
        uint32_t it = 0;
        uint4 buf = make_uint4(0, 0, 0, 0), nxt = make_uint4(0, 0, 0, 0);
        if (left > 0) nxt = __ldg(reinterpret_cast<const uint4 *>(ws.ops + cur));
        uint32_t pairs_left = 0, hdr_left = 0;                          // state of the current group
        float v0 = 0.f;
        bool uniq = true;
        while (__any_sync(kFull, left > 0)) {
            if ((it & 3u) == 0u) {
                buf = nxt;                                               // software pipeline: the next 16 bytes are already on their way
                if (left > 4) nxt = __ldg(reinterpret_cast<const uint4 *>(ws.ops + cur + 4));
            }
            const uint32_t sel = it & 3u;
            const uint32_t word = sel == 0 ? buf.x : (sel == 1 ? buf.y : (sel == 2 ? buf.z : buf.w));
            if (left > 0) {
                if (pairs_left == 0 && hdr_left == 0) {
                    // ---- header: sizes of the group's levels
                    uint32_t n0 = (word >> 21) & 1023u;
                    if (n0 == kOpNExt) n0 = ws.ext[ecur++];
                    pairs_left = n0;
                    if (exact_stream) {
                        v0 = __fmul_rn(s_w[word & 31u], inv_of(n0));                          // Compare.hpp:924
                    } else {
                        hdr_left = NX;
                        v0 = __fmul_rn(s_w[0], inv_of(n0));
                        uniq = NX == 0 && ((word >> 1) & 0xEEEEEu) == 0u;
                        if (!uniq) {
                            for (int j = 1; j <= 5 && j < NK; ++j) {
                                uint32_t nn = (word >> (1 + 4 * (j - 1))) & 15u;
                                if (nn == kOpNibExt) nn = ws.ext[ecur++];
                                vbuf[j * BT + tid] = nn ? __fmul_rn(s_w[j], inv_of(nn)) : 0.f;
                            }
                        }
                    }
                } else if (hdr_left > 0) {
                    // ---- extra header word: levels 6 + 8 k .. (escapes of the levels are in level order, so they can be resolved here)
                    const uint32_t k = NX - hdr_left;
                    for (int i = 0; i < 8; ++i) {
                        const int j = 6 + 8 * (int)k + i;
                        if (j < NK) {
                            uint32_t nn = (word >> (4 * i)) & 15u;
                            if (nn == kOpNibExt) nn = ws.ext[ecur++];
                            vbuf[j * BT + tid] = nn ? __fmul_rn(s_w[j], inv_of(nn)) : 0.f;
                        }
                    }
                    --hdr_left;
                } else {
                    // ---- pair: taxon (| levels)
                    const uint32_t cell = (word & 0xffffu) - 1u;
                    float acc = __fadd_rn(cellS[cell * BT + tid], v0);                             // Compare.hpp:530
                    if (!exact_stream) {
                        const uint32_t h = (word >> 16) & 31u;
                        if (uniq) { for (uint32_t j = 1; j < h; ++j) acc = __fadd_rn(acc, s_w[j]); }        // w * (1.f / 1) == w
                        else { for (uint32_t j = 1; j < h; ++j) acc = __fadd_rn(acc, vbuf[j * BT + tid]); }
                    }
                    cellS[cell * BT + tid] = acc;
                    --pairs_left;
                }
                ++cur; --left;
            }
            ++it;
        }
        // ---- end of the read: scores out
        if (have) {
            for (int c = 0; c < NC && c < T - 1; ++c) {
                if (scores) scores[(uint64_t)r * T + c + 1] = cellS[c * BT + tid];
                cellS[c * BT + tid] = 0.f;
            }
            if (scores) scores[(uint64_t)r * T] = 0.f;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// B for any number of taxa: warp per read, group after group; the taxa of a group are distinct, so its pairs are
// applied in parallel (a lane per pair) to the cells of the read's taxa, which live in a shared-memory hash table.
constexpr int kHBlock = 128;
constexpr int kHWarps = kHBlock / 32;
constexpr uint32_t kHashSlots = 1024;
constexpr uint32_t kHashEmpty = 0xffffffffu;

__global__ void __launch_bounds__(kHBlock) accumulate_hash_kernel(OpsWorkspace ws, uint32_t r0, uint32_t r1, LevelParams lp, bool exact_stream,
                                                                  float *__restrict__ scores, HitsOut hits, unsigned long long *__restrict__ cnt_fx,
                                                                  unsigned long long *__restrict__ cnt_unique, uint32_t *__restrict__ err_flag)
{
    __shared__ uint32_t s_key[kHWarps][kHashSlots];
    __shared__ float s_acc[kHWarps][kHashSlots];
    __shared__ uint16_t s_used[kHWarps][kHashSlots];
    __shared__ uint32_t s_nused[kHWarps];
    __shared__ float s_w[32];
    __shared__ float s_v[kHWarps][32];          // value of each level of the current group
    __shared__ uint32_t s_n[kHWarps][32];       // ... and its size
    const int w = threadIdx.x >> 5;
    const uint32_t lane = threadIdx.x & 31;
    const int T = lp.n_taxa, NK = lp.num_k;
    for (uint32_t i = threadIdx.x; i < kHWarps * kHashSlots; i += kHBlock) { (&s_key[0][0])[i] = kHashEmpty; (&s_acc[0][0])[i] = 0.f; }
    if (threadIdx.x < 32) { const int k = lp.k_low + threadIdx.x; s_w[threadIdx.x] = (float)(k * k) / 625.f; }
    if (lane == 0) s_nused[w] = 0;
    __syncthreads();
    uint32_t *key = s_key[w];
    float *accs = s_acc[w];
    const uint32_t NX = exact_stream ? 0u : (uint32_t)ops_extra_words(NK);
    const uint32_t warps_total = gridDim.x * kHWarps;
#pragma unroll 1
    for (uint32_t r = r0 + blockIdx.x * kHWarps + w; r < r1; r += warps_total) {
        const uint64_t o0 = ws.ops_start[r];
        const uint32_t n_ops = ws.ops_count[r];
        uint64_t ecur = ws.ext_start[r];
        if (scores) for (int t = lane; t < T; t += 32) scores[(uint64_t)r * T + t] = 0.f;
        bool full = false;
        // a read with more ops than the table safely holds taxa is replayed once per taxon range of <= kHashSlots / 2 taxa
        const uint32_t width = (n_ops <= kHashSlots * 5 / 8 || (uint32_t)T <= kHashSlots * 5 / 8) ? (uint32_t)T : kHashSlots / 2;
        const uint32_t n_pass = ((uint32_t)T + width - 1) / width;
        uint64_t hb = 0;
        uint32_t n_hits_read = 0;
        // header of the group at `at`: group size, escapes consumed from the ext stream; fills the warp's level tables when `fill`
        auto read_header = [&](uint32_t at, uint64_t &ec, uint32_t &lev, bool fill) {
            const uint32_t hw = __ldg(ws.ops + o0 + at);
            uint32_t n0 = (hw >> 21) & 1023u;
            if (n0 == kOpNExt) n0 = ws.ext[ec++];
            if (exact_stream) {
                lev = hw & 31u;
                if (fill && lane == 0) { s_n[w][lev] = n0; s_v[w][lev] = __fmul_rn(s_w[lev], __fdiv_rn(1.f, (float)n0)); }
            } else {
                // lane j decodes level j; escaped sizes come from the ext stream in level order
                uint32_t nn = 0;
                const int j = (int)lane;
                if (j == 0) nn = n0;
                else if (j < NK) {
                    if (j <= 5) nn = (hw >> (1 + 4 * (j - 1))) & 15u;
                    else nn = (__ldg(ws.ops + o0 + at + 1 + (j - 6) / 8) >> (4 * ((j - 6) % 8))) & 15u;
                }
                const bool escaped = j > 0 && j < NK && nn == kOpNibExt;
                const uint32_t esc = __ballot_sync(kFull, escaped);
                if (escaped) nn = ws.ext[ec + __popc(esc & lanemask_lt())];
                ec += __popc(esc);
                if (fill && j < NK) { s_n[w][j] = nn; s_v[w][j] = nn ? __fmul_rn(s_w[j], __fdiv_rn(1.f, (float)nn)) : 0.f; }
            }
            __syncwarp();
            return n0;
        };
        if (hits.taxon && n_pass > 1) {
            // the hits of a read are contiguous in the output: count its distinct taxa first, range by range, with the table alone
            for (uint32_t pass = 0; pass < n_pass; ++pass) {
                const uint32_t t_lo = pass * width, t_hi = t_lo + width;
                uint64_t ec = ws.ext_start[r];
                uint32_t at = 0, lev = 0;
                while (at < n_ops) {
                    const uint32_t n0 = read_header(at, ec, lev, false);
                    at += 1 + NX;
                    for (uint32_t p0 = 0; p0 < n0; p0 += 32) {
                        const bool active = p0 + lane < n0;
                        const uint32_t t = active ? __ldg(ws.ops + o0 + at + p0 + lane) & 0xffffu : 0u;
                        if (active && t >= t_lo && t < t_hi) {
                            uint32_t slot = (t * 2654435761u) >> 22;
                            for (;;) {
                                const uint32_t k = key[slot];
                                if (k == t) break;
                                if (k == kHashEmpty) {
                                    const uint32_t old = atomicCAS(&key[slot], kHashEmpty, t);
                                    if (old == kHashEmpty) { s_used[w][atomicAdd(&s_nused[w], 1u)] = (uint16_t)slot; break; }
                                    if (old == t) break;
                                }
                                slot = (slot + 1) & (kHashSlots - 1);
                            }
                        }
                        __syncwarp();
                    }
                    at += n0;
                }
                __syncwarp();
                const uint32_t nu = s_nused[w];
                n_hits_read += nu;
                for (uint32_t i = lane; i < nu; i += 32) key[s_used[w][i]] = kHashEmpty;
                __syncwarp();
                if (lane == 0) s_nused[w] = 0;
                __syncwarp();
            }
        }
        uint32_t hits_done = 0;
        bool hits_ok = false;
        for (uint32_t pass = 0; pass < n_pass; ++pass) {
        const uint32_t t_lo = pass * width, t_hi = t_lo + width;
        ecur = ws.ext_start[r];
        uint32_t at = 0;
        while (at < n_ops) {
            // ---- header (the same words for every lane)
            uint32_t lev = 0;
            const uint32_t n0 = read_header(at, ecur, lev, true);
            at += 1 + NX;
            // ---- pairs: a lane each
            for (uint32_t p0 = 0; p0 < n0; p0 += 32) {
                const bool active = p0 + lane < n0;
                const uint32_t word = active ? __ldg(ws.ops + o0 + at + p0 + lane) : 0u;
                const uint32_t t = word & 0xffffu, h = (word >> 16) & 31u;
                if (active && t >= t_lo && t < t_hi) {
                    uint32_t slot = (t * 2654435761u) >> 22;
                    for (uint32_t probes = 0;; ++probes) {
                        const uint32_t k = key[slot];
                        if (k == t) break;
                        if (k == kHashEmpty) {
                            const uint32_t old = atomicCAS(&key[slot], kHashEmpty, t);
                            if (old == kHashEmpty) { s_used[w][atomicAdd(&s_nused[w], 1u)] = (uint16_t)slot; break; }
                            if (old == t) break;
                        }
                        slot = (slot + 1) & (kHashSlots - 1);
                        if (probes > kHashSlots) { full = true; break; }
                    }
                    float acc = accs[slot];
                    const uint32_t j0 = exact_stream ? lev : 0u, j1 = exact_stream ? lev + 1u : h;
                    for (uint32_t j = j0; j < j1; ++j) {
                        acc = __fadd_rn(acc, s_v[w][j]);                                   // Compare.hpp:924, 530
                        const uint32_t nn = s_n[w][j];
                        if (nn == 1u) atomicAdd(cnt_unique + (size_t)j * T + t, 1ull);
                        else atomicAdd(cnt_fx + (size_t)j * T + t, (unsigned long long)inv_fx(nn));
                    }
                    accs[slot] = acc;
                }
                __syncwarp();
                if (s_nused[w] > kHashSlots * 3 / 4) full = true;
            }
            at += n0;
            __syncwarp();
        }
        // ---- results of this taxon range, then clear its cells
        const uint32_t nu = s_nused[w];
        if (hits.taxon && pass == 0) {
            const uint32_t total = n_pass > 1 ? n_hits_read : nu;
            if (lane == 0) {
                hb = atomicAdd(hits.cursor, (unsigned long long)total);
                if (hb + total > hits.cap) { atomicOr(err_flag, 128u); hits.begin[r] = 0; hits.count[r] = 0; }
                else { hits.begin[r] = hb; hits.count[r] = total; }
            }
            hb = __shfl_sync(kFull, hb, 0);
            hits_ok = hb + total <= hits.cap;
        }
        for (uint32_t i = lane; i < nu; i += 32) {
            const uint32_t slot = s_used[w][i], t = key[slot];
            const float v = accs[slot];
            if (scores) scores[(uint64_t)r * T + t] = v;
            if (hits.taxon && hits_ok) { hits.taxon[hb + hits_done + i] = t; hits.score[hb + hits_done + i] = v; }
            key[slot] = kHashEmpty;
            accs[slot] = 0.f;
        }
        hits_done += nu;
        __syncwarp();
        if (lane == 0) s_nused[w] = 0;
        __syncwarp();
        }   // taxon ranges
        if (__any_sync(kFull, full) && lane == 0) atomicOr(err_flag, 256u);
    }
}

__global__ void max_instances_kernel(const uint64_t *__restrict__ read_koff, uint32_t n_reads, uint32_t *out)
{
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n_reads) atomicMax(out, (uint32_t)(read_koff[r + 1] - read_koff[r]));
}

template <int SLOTS>
void launch_dense(cudaStream_t s, OpsWorkspace ws, uint32_t r0, uint32_t r1, LevelParams lp, bool exact_stream, float *scores)
{
    constexpr int NC = 32 * SLOTS;
    const size_t sm = (size_t)NC * kDenseBlock * 4 + (size_t)lp.num_k * kDenseBlock * 4 + 32 * 4 + kInvTab * 4;
    auto kern = accumulate_dense_kernel<SLOTS>;
    KASA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    int per_sm = 1;
    KASA_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kDenseBlock, sm));
    unsigned grid = (unsigned)div_up<uint64_t>(r1 - r0, kDenseBlock);
    const unsigned resident = (unsigned)(kSMs * (per_sm > 0 ? per_sm : 1));
    if (grid > resident) grid = resident;
    kern<<<grid, kDenseBlock, sm, s>>>(ws, r0, r1, lp, exact_stream, scores);
}

// ---------------------------------------------------------------------------------------------------------------
// PRESORTED READS (extract.cu emits a read of <= kPresortMax k-mers in ascending k-mer order, the batch sort is stable): the
// match records of a read already lie in HBM in the order in which the reference adds them up, so for content files of <= 64 taxa
// and <= 6 levels scores and counts come straight from the records — no ordering pass, no op stream.
//   LANES threads per read, each owning NT taxa (taxon - 1 in [lane * NT, lane * NT + NT)); every thread reads the read's records
//   (one 32-byte sector per k-mer, the LANES loads of a read coalesce into one request, four records in flight per thread) and keeps
//   the float32 score and the k_low share of the count of its taxa IN REGISTERS:
//     level k_low   a predicated scan over the thread's slice of the record's taxa mask: NT independent add chains, no divergence;
//     levels above  the (few) up pairs of the record: the owner selects the taxon's register through a select tree, adds the
//                   levels in order and writes it back through predicated moves — registers cannot be indexed, and keeping such
//                   taxa in shared-memory cells instead costs more: against a saturated index a read meets ~10 of 50 taxa above
//                   k_low somewhere, so most k_low adds would take the slow path.
//   Rarer events (a k-mer with a single taxon, counts above k_low) go to small per-CTA tables with shared-memory atomics.
constexpr int kSortedBlock = 128;
constexpr int kSortedDepth = 4;         // records in flight per thread

// sc[s] for a run-time s < NT: select tree on the bits of s (NT - 1 selects)
template <int LO, int N>
__device__ __forceinline__ float tree_select(const float *sc, uint32_t s)
{
    if constexpr (N == 1) return sc[LO];
    else {
        constexpr int H = 1 << (31 - __builtin_clz((unsigned)(N - 1)));     // largest power of two < N
        const float lo = tree_select<LO, H>(sc, s);
        const float hi = tree_select<LO + H, N - H>(sc, s);
        return (s & (uint32_t)H) ? hi : lo;
    }
}

template <int LANES, int NT>
__global__ void __launch_bounds__(kSortedBlock) score_sorted_kernel(const uint64_t *__restrict__ rec, const uint64_t *__restrict__ read_koff, uint32_t n_reads,
                                                                     LevelParams lp, OverflowPeers peers, float *__restrict__ scores,
                                                                     unsigned long long *__restrict__ cnt_fx, unsigned long long *__restrict__ cnt_unique)
{
    static_assert(NT <= 32 && LANES * NT <= 64 + NT, "slice of the taxa mask must fit 32 bits");
    constexpr int BT = kSortedBlock, D = kSortedDepth;
    constexpr uint32_t kSlice = NT >= 32 ? 0xffffffffu : ((1u << NT) - 1u);
    __shared__ float2 s_vq[72];                                         // [n]: .x = float(w_0 * (1.f / n)), .y = bits of 2^24 / n (0 for n < 2)
    __shared__ float s_inv[72];
    __shared__ uint32_t s_q32[72];                                      // 2^32 / n
    __shared__ unsigned long long s_fx[6 * 64];                         // CTA totals [level][taxon - 1], 2^-32 units
    __shared__ uint32_t s_un[6 * 64];
    __shared__ uint32_t s_hist[64 * 8];                                 // [taxon - 1][h]: k-mers where the taxon is alone at its levels 1 .. h-1
    const int tid = threadIdx.x, T = lp.n_taxa, NK = lp.num_k;
    const int q = tid / LANES, l = tid % LANES, tbase = l * NT;
    float wj[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) { const int k = lp.k_low + j; wj[j] = (float)(k * k) / 625.f; }      // Compare.hpp:392
    for (int i = tid; i < 6 * 64; i += BT) { s_fx[i] = 0; s_un[i] = 0; }
    for (int i = tid; i < 64 * 8; i += BT) s_hist[i] = 0;
    if (tid < 72) {
        const float inv = tid ? __fdiv_rn(1.f, (float)tid) : 0.f;       // Compare.hpp:924: 1.f / numOfEntries
        s_inv[tid] = inv;
        s_vq[tid] = make_float2(__fmul_rn(wj[0], inv), __uint_as_float(tid >= 2 ? (uint32_t)(((1u << 24) + (tid >> 1)) / tid) : 0u));
        s_q32[tid] = tid >= 2 ? inv_fx((uint32_t)tid) : 0u;
    }
    __syncthreads();

    float sc[NT];
    uint32_t fq[NT];
    // adds the levels 1 .. h-1 of taxon c (0-based) to its score; nj = level sizes (nullptr: the taxon is alone at each of them)
    auto add_upper = [&](uint32_t c, uint32_t h, const uint32_t *nj) {
        const uint32_t s = c - (uint32_t)tbase;
        if (s >= (uint32_t)NT || h < 2u) return;
        float v = tree_select<0, NT>(sc, s);
#pragma unroll
        for (int j = 1; j < 6; ++j) {
            if ((uint32_t)j >= h) break;
            if (!nj) v = __fadd_rn(v, wj[j]);                            // w * (1.f / 1) == w
            else {
                v = __fadd_rn(v, __fmul_rn(wj[j], s_inv[nj[j]]));
                if (nj[j] == 1u) atomicAdd(&s_un[j * 64 + c], 1u);
                else atomicAdd(&s_fx[j * 64 + c], (unsigned long long)s_q32[nj[j]]);
            }
        }
        if (!nj) atomicAdd(&s_hist[c * 8 + (h & 7u)], 1u);
#pragma unroll
        for (int t = 0; t < NT; ++t) sc[t] = s == (uint32_t)t ? v : sc[t];
    };
    auto flush_fq = [&]() {
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            if (fq[t]) atomicAdd(&s_fx[tbase + t], (unsigned long long)fq[t] << kFxShift);
            fq[t] = 0u;
        }
    };
    // one match record
    auto apply = [&](const uint64_t (&w)[4]) {
        const uint32_t n0 = rec_count(w[0]);                             // 0: no hit (mask and pairs are zero too)
        const uint32_t hmax = rec_hmax(w[0]);
        if (n0 && !rec_is_mask(w[0])) {
            // ---- list record inside a small taxonomy (more than six taxa above k_low, or a span read from HBM): pair by pair
            const bool ovf = rec_overflow(w[0]);
            const uint32_t *list = ovf ? peers.base[rec_owner(w[0])] + w[1] : nullptr;
            auto pair_at = [&](uint32_t k) -> uint32_t { return ovf ? __ldg(list + k) : (uint32_t)(w[1 + k / 3] >> (kPairBits * (k % 3))) & kPairMask; };
            uint32_t nj[6] = {0, 0, 0, 0, 0, 0};
            for (uint32_t k = 0; k < n0; ++k) {
                const uint32_t h = pair_at(k) >> 16;
#pragma unroll
                for (int j = 0; j < 6; ++j) nj[j] += (uint32_t)j < h ? 1u : 0u;
            }
            const float2 vq = s_vq[n0 < 72u ? n0 : 0u];
            for (uint32_t k = 0; k < n0; ++k) {
                const uint32_t e = pair_at(k), c = (e & 0xffffu) - 1u, h = e >> 16, s = c - (uint32_t)tbase;
                if (s >= (uint32_t)NT) continue;
#pragma unroll
                for (int t = 0; t < NT; ++t)
                    if (s == (uint32_t)t) { sc[t] = __fadd_rn(sc[t], vq.x); fq[t] += __float_as_uint(vq.y); }
                if (n0 == 1u) atomicAdd(&s_un[c], 1u);
                add_upper(c, h, nj);
            }
            return;
        }
        // ---- mask record, level k_low (Compare.hpp:530): every taxon of the mask gets float(w_0 * (1.f / n0))
        const float2 vq = s_vq[n0];
        const float v0 = vq.x;
        const uint32_t q0 = __float_as_uint(vq.y);
        const uint32_t bits = (uint32_t)(w[1] >> tbase) & kSlice;
#pragma unroll
        for (int t = 0; t < NT; ++t)
            if ((bits >> t) & 1u) { sc[t] = __fadd_rn(sc[t], v0); fq[t] += q0; }
        if (n0 == 1u && bits) atomicAdd(&s_un[tbase + __ffs((int)bits) - 1], 1u);           // alone at k_low: a unique hit
        // ---- levels above k_low: up pairs (taxon | levels << 16), 21 bits each, three per word
        if (hmax > 1u) {
            const uint32_t p0 = (uint32_t)w[2] & kPairMask;
            if ((w[2] >> kPairBits) == 0 && w[3] == 0) add_upper((p0 & 0xffffu) - 1u, p0 >> 16, nullptr);
            else {
                const uint64_t pw[3] = {w[1], w[2], w[3]};
                uint32_t nj[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
                for (int k = 0; k < kRecUpPairs; ++k) {
                    const uint32_t h = up_pair(pw, k) >> 16;                                    // 0 for an empty pair
#pragma unroll
                    for (int j = 1; j < 6; ++j) nj[j] += (uint32_t)j < h ? 1u : 0u;
                }
#pragma unroll 1
                for (int k = 0; k < kRecUpPairs; ++k) {
                    const uint32_t e = (uint32_t)(pw[1 + k / 3] >> (kPairBits * (k % 3))) & kPairMask;
                    if (!e) break;
                    add_upper((e & 0xffffu) - 1u, e >> 16, nj);
                }
            }
        }
    };
    auto load_rec = [&](const uint64_t *p, uint64_t (&w)[4]) {
        asm volatile("ld.global.nc.L1::no_allocate.v4.b64 {%0, %1, %2, %3}, [%4];" : "=l"(w[0]), "=l"(w[1]), "=l"(w[2]), "=l"(w[3]) : "l"(p));
    };

#pragma unroll 1
    for (uint32_t r = blockIdx.x * (BT / LANES) + q; r < n_reads; r += gridDim.x * (BT / LANES)) {
        const uint64_t s0 = read_koff[r];
        const uint32_t ni = (uint32_t)(read_koff[r + 1] - s0);
        const uint64_t *p = rec + s0 * kRecWords;
#pragma unroll
        for (int t = 0; t < NT; ++t) { sc[t] = 0.f; fq[t] = 0u; }
        uint64_t buf[D][4];
#pragma unroll
        for (int d = 0; d < D; ++d) {
            buf[d][0] = buf[d][1] = buf[d][2] = buf[d][3] = 0;
            if ((uint32_t)d < ni) load_rec(p + (uint64_t)d * kRecWords, buf[d]);
        }
#pragma unroll 1
        for (uint32_t i = 0; i < ni; i += D) {
#pragma unroll
            for (int d = 0; d < D; ++d) {
                if (i + d < ni) {
                    uint64_t w[4] = {buf[d][0], buf[d][1], buf[d][2], buf[d][3]};
                    if (i + d + D < ni) load_rec(p + (uint64_t)(i + d + D) * kRecWords, buf[d]);
                    apply(w);
                }
            }
            // a k-mer adds at most 2^23 units to a k_low sum: hand the sums to the CTA's totals before 32 bits could overflow
            if ((i & 255u) == 256u - D) flush_fq();
        }
        // ---- end of the read: scores out, its k_low sums into the CTA's totals
#pragma unroll
        for (int t = 0; t < NT; ++t) {
            const int c = tbase + t;
            if (c < T - 1 && scores) scores[(uint64_t)r * T + c + 1] = sc[t];
        }
        flush_fq();
        if (scores && l == 0) scores[(uint64_t)r * T] = 0.f;
    }
    __syncthreads();
    for (int i = tid; i < NK * 64; i += BT) {
        const int j = i / 64, c = i % 64;
        if (c >= T - 1) continue;
        unsigned long long u = s_un[i];
        if (j >= 1) for (int h = j + 1; h < 8; ++h) u += s_hist[c * 8 + h];     // alone at the levels 1 .. h-1
        if (u) atomicAdd(cnt_unique + (size_t)j * T + c + 1, u);
        if (s_fx[i]) atomicAdd(cnt_fx + (size_t)j * T + c + 1, s_fx[i]);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// PRESORTED READS, second organisation: WARP PER READ, 32 k-mers at a time.
//   setup   lane = k-mer: the 32 records of the block arrive as one coalesced 1 KB load; every lane publishes its k-mer's value
//           float(w_0 * (1.f / n0)) and its up pairs in shared memory; the 32 taxa masks are TRANSPOSED with warp shuffles (five
//           butterfly stages per 32 x 32 bit block), after which
//   main    lane = taxon holds, per taxon word, the bit vector of the block's k-mers that name ITS taxon — and walks only those
//           bits, in k-mer order, adding into its own float32 register.  No lane touches a k-mer that does not concern it: a
//           k-mer names ~4.5 of 50 taxa against a saturated index, so the dense organisations above (one predicated add per
//           taxon and k-mer) spend > 90 % of their issue slots on nothing.  The walk is as long as the busiest lane's (the read's
//           own taxon: ~15 of 32 k-mers, most of them with upper levels); branches on record shape are rare and cheap.
// Counts: the k_low shares ride along in the walk; unique k_low hits are a popcount against the ballot of "n0 == 1".
constexpr int kBlkWarps = 8;

__device__ __forceinline__ uint32_t transpose32(uint32_t x, uint32_t lane)
{
    // row `lane` in, column `lane` out: bit i of the result = bit `lane` of lane i's input
#pragma unroll
    for (int j = 16; j >= 1; j >>= 1) {
        const uint32_t m = j == 16 ? 0x0000ffffu : j == 8 ? 0x00ff00ffu : j == 4 ? 0x0f0f0f0fu : j == 2 ? 0x33333333u : 0x55555555u;   // columns with (col & j) == 0
        const uint32_t y = __shfl_xor_sync(kFull, x, j);
        x = (lane & (uint32_t)j) ? ((x & ~m) | ((y & ~m) >> j)) : ((x & m) | ((y & m) << j));
    }
    return x;
}

struct BlkWarpMem {
    float2 vq[32];            // per k-mer of the block: .x = float(w_0 * (1.f / n0)), .y = bits of 2^24 / n0 (0 for n0 < 2)
    float val[32][8];         // k-mers with taxa above k_low: [j - 1] = float(w_j * (1.f / n_j)) for the levels j = 1 .. 5
    uint8_t h[32][64];        // ... and [taxon - 1] = number of levels that taxon matches (only the entries of up pairs are written)
};

template <bool TWO>
__global__ void __launch_bounds__(kBlkWarps * 32) score_blocks_kernel(const uint64_t *__restrict__ rec, const uint64_t *__restrict__ read_koff, uint32_t n_reads,
                                                                       LevelParams lp, OverflowPeers peers, float *__restrict__ scores,
                                                                       unsigned long long *__restrict__ cnt_fx, unsigned long long *__restrict__ cnt_unique)
{
    __shared__ __align__(16) BlkWarpMem s_w[kBlkWarps];
    __shared__ float2 s_vqtab[72];
    __shared__ float s_inv[72];
    __shared__ uint32_t s_q32[72];                                      // 2^32 / n
    __shared__ unsigned long long s_fx[6 * 64];                         // CTA totals [level][taxon - 1], 2^-32 units (levels above k_low)
    __shared__ uint32_t s_un[6 * 64];
    __shared__ uint32_t s_hist[64 * 8];                                 // [taxon - 1][h]: k-mers where the taxon is alone at its levels 1 .. h-1
    const int tid = threadIdx.x, T = lp.n_taxa, NK = lp.num_k;
    const uint32_t lane = tid & 31u;
    const int w = tid >> 5;
    BlkWarpMem &M = s_w[w];
    float wj[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) { const int k = lp.k_low + j; wj[j] = (float)(k * k) / 625.f; }      // Compare.hpp:392
    for (int i = tid; i < 6 * 64; i += blockDim.x) { s_fx[i] = 0; s_un[i] = 0; }
    for (int i = tid; i < 64 * 8; i += blockDim.x) s_hist[i] = 0;
    if (tid < 72) {
        const float inv = tid ? __fdiv_rn(1.f, (float)tid) : 0.f;       // Compare.hpp:924: 1.f / numOfEntries
        s_inv[tid] = inv;
        s_vqtab[tid] = make_float2(__fmul_rn(wj[0], inv), __uint_as_float(tid >= 2 ? (uint32_t)(((1u << 24) + (tid >> 1)) / tid) : 0u));
        s_q32[tid] = tid >= 2 ? inv_fx((uint32_t)tid) : 0u;
    }
    __syncthreads();
    unsigned long long fa64 = 0, fb64 = 0;                              // k_low sums of this lane's taxa (lane + 1, lane + 33), 2^-24 units
    uint32_t ua = 0, ub = 0;                                            // their unique k_low hits

    // the levels 1 .. h-1 of taxon c (0-based) at k-mer i of the block: values and level count were laid out by the k-mer's lane
    auto upper = [&](uint32_t i, uint32_t c, float acc) -> float {
        const uint32_t h = M.h[i][c];
        const float4 v = *reinterpret_cast<const float4 *>(&M.val[i][0]);
        const float v5 = M.val[i][4];
        if (h > 1u) acc = __fadd_rn(acc, v.x);
        if (h > 2u) acc = __fadd_rn(acc, v.y);
        if (h > 3u) acc = __fadd_rn(acc, v.z);
        if (h > 4u) acc = __fadd_rn(acc, v.w);
        if (h > 5u) acc = __fadd_rn(acc, v5);
        return acc;
    };

    const uint32_t warps_total = gridDim.x * kBlkWarps;
#pragma unroll 1
    for (uint32_t r = blockIdx.x * kBlkWarps + w; r < n_reads; r += warps_total) {
        const uint64_t s0 = read_koff[r];
        const uint32_t ni = (uint32_t)(read_koff[r + 1] - s0);
        const uint64_t *p = rec + s0 * kRecWords;
        float sa = 0.f, sb = 0.f;
        uint32_t fa = 0, fb = 0;
        uint64_t nx[4] = {0, 0, 0, 0};
        if (lane < ni) asm volatile("ld.global.nc.L1::no_allocate.v4.b64 {%0, %1, %2, %3}, [%4];" : "=l"(nx[0]), "=l"(nx[1]), "=l"(nx[2]), "=l"(nx[3]) : "l"(p + (uint64_t)lane * kRecWords));
#pragma unroll 1
        for (uint32_t b0 = 0; b0 < ni; b0 += 32) {
            // ---- setup, lane = k-mer b0 + lane
            const uint64_t w0 = nx[0], w1 = nx[1], w2 = nx[2], w3 = nx[3];
            nx[0] = nx[1] = nx[2] = nx[3] = 0;
            if (b0 + 32 + lane < ni)
                asm volatile("ld.global.nc.L1::no_allocate.v4.b64 {%0, %1, %2, %3}, [%4];" : "=l"(nx[0]), "=l"(nx[1]), "=l"(nx[2]), "=l"(nx[3]) : "l"(p + (uint64_t)(b0 + 32 + lane) * kRecWords));
            const uint32_t n0 = rec_count(w0), hmax = rec_hmax(w0);
            uint64_t mk = w1, um = 0;
            if (n0 && hmax > 1u) {
                // taxa above k_low: every up pair (taxon | levels << 16) leaves its level count, the k-mer its level values, and the
                // counts of the levels above k_low are taken here (all lanes at once) rather than in the walk (one lane at a time)
                const bool is_mask = rec_is_mask(w0);
                const uint32_t p0 = (uint32_t)w2 & kPairMask;
                if (is_mask && (w2 >> kPairBits) == 0 && w3 == 0) {
                    const uint32_t c = (p0 & 0xffffu) - 1u, h = p0 >> 16;                     // the usual case: one taxon, alone at its levels
                    um = 1ull << c;
                    M.h[lane][c] = (uint8_t)h;
                    *reinterpret_cast<float4 *>(&M.val[lane][0]) = make_float4(wj[1], wj[2], wj[3], wj[4]);      // w * (1.f / 1) == w
                    M.val[lane][4] = wj[5];
                    atomicAdd(&s_hist[c * 8 + (h & 7u)], 1u);
                } else {
                    const uint32_t *list = !is_mask && rec_overflow(w0) ? peers.base[rec_owner(w0)] + w1 : nullptr;
                    const uint32_t np = is_mask ? (uint32_t)kRecUpPairs : n0;
                    auto pair_at = [&](uint32_t k) -> uint32_t {
                        if (list) return __ldg(list + k);
                        const uint32_t word = is_mask ? 3u + k : k;                            // mask records keep their pairs in w2, w3
                        return (uint32_t)((word < 3u ? w1 : word < 6u ? w2 : w3) >> (kPairBits * (word % 3u))) & kPairMask;
                    };
                    uint32_t nj[6] = {0, 0, 0, 0, 0, 0};
                    for (uint32_t k = 0; k < np; ++k) {
                        const uint32_t h = pair_at(k) >> 16;                                   // 0 for an empty pair
#pragma unroll
                        for (int j = 1; j < 6; ++j) nj[j] += (uint32_t)j < h ? 1u : 0u;
                    }
                    float vj[6];
#pragma unroll
                    for (int j = 1; j < 6; ++j) vj[j] = __fmul_rn(wj[j], s_inv[nj[j] < 72u ? nj[j] : 0u]);
                    *reinterpret_cast<float4 *>(&M.val[lane][0]) = make_float4(vj[1], vj[2], vj[3], vj[4]);
                    M.val[lane][4] = vj[5];
                    for (uint32_t k = 0; k < np; ++k) {
                        const uint32_t e = pair_at(k), c = (e & 0xffffu) - 1u, h = e >> 16;
                        if (h < 2u) continue;
                        um |= 1ull << c;
                        M.h[lane][c] = (uint8_t)h;
#pragma unroll
                        for (int j = 1; j < 6; ++j) {
                            if ((uint32_t)j >= h) break;
                            if (nj[j] == 1u) atomicAdd(&s_un[j * 64 + c], 1u);
                            else atomicAdd(&s_fx[j * 64 + c], (unsigned long long)s_q32[nj[j] < 72u ? nj[j] : 0u]);
                        }
                    }
                }
            }
            if (n0 && !rec_is_mask(w0)) {
                // list record inside a small taxonomy (more than six taxa above k_low, or a span read from HBM): mask from its pairs
                const uint32_t *list = rec_overflow(w0) ? peers.base[rec_owner(w0)] + w1 : nullptr;
                mk = 0;
                for (uint32_t k = 0; k < n0; ++k) {
                    const uint32_t e = list ? __ldg(list + k) : (uint32_t)((k < 3 ? w1 : k < 6 ? w2 : w3) >> (kPairBits * (k % 3))) & kPairMask;
                    mk |= 1ull << ((e & 0xffffu) - 1u);
                }
            }
            M.vq[lane] = s_vqtab[n0 < 72u ? n0 : 0u];
            const uint32_t uni = __ballot_sync(kFull, n0 == 1u);
            const bool any_up = __any_sync(kFull, um != 0);
            // ---- transposes: lane = taxon from here on
            const uint32_t tlo = transpose32((uint32_t)mk, lane);
            const uint32_t thi = TWO ? transpose32((uint32_t)(mk >> 32), lane) : 0u;
            uint32_t xlo = 0, xhi = 0;
            if (any_up) { xlo = transpose32((uint32_t)um, lane); if (TWO) xhi = transpose32((uint32_t)(um >> 32), lane); }
            __syncwarp();
            ua += __popc(tlo & uni);
            if (TWO) ub += __popc(thi & uni);
            // ---- main: this lane's k-mers only, in order (Compare.hpp:530 for level k_low, then the levels above)
            for (uint32_t bits = tlo; bits; bits &= bits - 1u) {
                const uint32_t i = (uint32_t)__ffs((int)bits) - 1u;
                const float2 vq = M.vq[i];
                sa = __fadd_rn(sa, vq.x);
                fa += __float_as_uint(vq.y);
                if ((xlo >> i) & 1u) sa = upper(i, lane, sa);
            }
            if (TWO) {
                for (uint32_t bits = thi; bits; bits &= bits - 1u) {
                    const uint32_t i = (uint32_t)__ffs((int)bits) - 1u;
                    const float2 vq = M.vq[i];
                    sb = __fadd_rn(sb, vq.x);
                    fb += __float_as_uint(vq.y);
                    if ((xhi >> i) & 1u) sb = upper(i, lane + 32u, sb);
                }
            }
            // a k-mer adds at most 2^23 units to a k_low sum: widen before 32 bits could overflow
            if ((b0 & 255u) == 224u) { fa64 += fa; fb64 += fb; fa = fb = 0; }
            __syncwarp();
        }
        fa64 += fa; fb64 += fb;
        if (scores) {
            float *row = scores + (uint64_t)r * T;
            if (lane == 0) row[0] = 0.f;
            if ((int)lane + 1 < T) row[lane + 1] = sa;
            if (TWO && (int)lane + 33 < T) row[lane + 33] = sb;
        }
    }
    if ((int)lane + 1 < T) {
        if (fa64) atomicAdd(&s_fx[lane], fa64 << kFxShift);
        if (ua) atomicAdd(&s_un[lane], ua);
    }
    if (TWO && (int)lane + 33 < T) {
        if (fb64) atomicAdd(&s_fx[lane + 32], fb64 << kFxShift);
        if (ub) atomicAdd(&s_un[lane + 32], ub);
    }
    __syncthreads();
    for (int i = tid; i < NK * 64; i += blockDim.x) {
        const int j = i / 64, c = i % 64;
        if (c >= T - 1) continue;
        unsigned long long u = s_un[i];
        if (j >= 1) for (int h = j + 1; h < 8; ++h) u += s_hist[c * 8 + h];     // alone at the levels 1 .. h-1
        if (u) atomicAdd(cnt_unique + (size_t)j * T + c + 1, u);
        if (s_fx[i]) atomicAdd(cnt_fx + (size_t)j * T + c + 1, s_fx[i]);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// PRESORTED READS, third organisation: THREAD PER READ, ONE (k-mer, taxon) ITEM PER ITERATION.
// The two organisations above waste issue slots in opposite ways: the dense scan executes one predicated add per taxon and k-mer
// (a k-mer names ~4.5 of 50 taxa), the warp-per-read walk runs at the pace of the read's own taxon while 31 lanes idle.  Here
// every lane works through ITS OWN read as a stream of items — pop the next taxon of the current record's mask, add into that
// taxon's cell, fetch the next record when the mask is empty — so every lane does useful work in every iteration, whatever the
// taxa sets look like.  The per-taxon cells (float32 score + k_low share of the count) are a column of shared memory per thread:
// dynamic indexing without bank conflicts; records are 32-byte sectors read one ahead.  Records with several taxa above k_low
// lay out their level values and a bit-plane map taxon -> levels when they are fetched.
constexpr int kItemBlock = 128;

template <bool TWO>
__global__ void __launch_bounds__(kItemBlock) score_items_kernel(const uint64_t *__restrict__ rec, const uint64_t *__restrict__ read_koff, uint32_t n_reads,
                                                                  LevelParams lp, OverflowPeers peers, float *__restrict__ scores,
                                                                  unsigned long long *__restrict__ cnt_fx, unsigned long long *__restrict__ cnt_unique)
{
    constexpr int BT = kItemBlock;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, T = lp.n_taxa, NK = lp.num_k;
    uint2 *cell = reinterpret_cast<uint2 *>(smem_raw);                  // [T - 1][BT]: .x = score bits, .y = k_low sum (2^-24 units)
    float *s_val = reinterpret_cast<float *>(cell + (size_t)(T - 1) * BT);   // [5][BT]: level values of the thread's current record (several up pairs)
    __shared__ float2 s_vq[72];                                         // [n]: .x = float(w_0 * (1.f / n)), .y = bits of 2^24 / n (0 for n < 2)
    __shared__ float s_inv[72];
    __shared__ uint32_t s_q32[72];                                      // 2^32 / n
    __shared__ unsigned long long s_fx[6 * 64];                         // CTA totals [level][taxon - 1], 2^-32 units
    __shared__ uint32_t s_un[6 * 64];
    __shared__ uint32_t s_hist[64 * 8];                                 // [taxon - 1][h]: k-mers where the taxon is alone at its levels 1 .. h-1
    float wj[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) { const int k = lp.k_low + j; wj[j] = (float)(k * k) / 625.f; }      // Compare.hpp:392
    for (int i = tid; i < (T - 1) * BT; i += BT) cell[i] = make_uint2(0u, 0u);
    for (int i = tid; i < 6 * 64; i += BT) { s_fx[i] = 0; s_un[i] = 0; }
    for (int i = tid; i < 64 * 8; i += BT) s_hist[i] = 0;
    if (tid < 72) {
        const float inv = tid ? __fdiv_rn(1.f, (float)tid) : 0.f;       // Compare.hpp:924: 1.f / numOfEntries
        s_inv[tid] = inv;
        s_vq[tid] = make_float2(__fmul_rn(wj[0], inv), __uint_as_float(tid >= 2 ? (uint32_t)(((1u << 24) + (tid >> 1)) / tid) : 0u));
        s_q32[tid] = tid >= 2 ? inv_fx((uint32_t)tid) : 0u;
    }
    __syncthreads();
    constexpr uint32_t kNone = 0xffffffffu;

#pragma unroll 1
    for (uint32_t base = blockIdx.x * BT; base < n_reads; base += gridDim.x * BT) {
        const uint32_t r = base + tid;
        const bool have = r < n_reads;
        const uint64_t s0 = have ? read_koff[r] : 0;
        const uint32_t ni = have ? (uint32_t)(read_koff[r + 1] - s0) : 0u;
        const uint64_t *p = rec + s0 * kRecWords;
        uint32_t i = 0;                                                  // records fetched so far
        uint32_t mlo = 0, mhi = 0;                                       // taxa of the current record still to be added
        float v0 = 0.f;
        uint32_t q0 = 0;
        uint32_t upc = kNone, uph = 0;                                   // the current record's only up pair (taxon 0-based, levels)
        bool cplx = false;                                               // ... or several: levels of taxon c = bit c of the three planes
        uint64_t pl0 = 0, pl1 = 0, pl2 = 0;
        uint64_t nx[4] = {0, 0, 0, 0};
        if (ni) asm volatile("ld.global.nc.L1::no_allocate.v4.b64 {%0, %1, %2, %3}, [%4];" : "=l"(nx[0]), "=l"(nx[1]), "=l"(nx[2]), "=l"(nx[3]) : "l"(p));
        bool active = ni > 0;
        while (__any_sync(kFull, active)) {
            if (active && (mlo | mhi) == 0u) {
                if (i == ni) active = false;
                else {
                    // ---- fetch record i (the next one is requested right away)
                    const uint64_t w0 = nx[0], w1 = nx[1], w2 = nx[2], w3 = nx[3];
                    ++i;
                    if (i < ni) asm volatile("ld.global.nc.L1::no_allocate.v4.b64 {%0, %1, %2, %3}, [%4];" : "=l"(nx[0]), "=l"(nx[1]), "=l"(nx[2]), "=l"(nx[3]) : "l"(p + (uint64_t)i * kRecWords));
                    if ((i & 255u) == 0u) {
                        // a k-mer adds at most 2^23 units to a k_low sum: hand the sums on before 32 bits could overflow
                        for (int t = 0; t < T - 1; ++t) {
                            uint2 x = cell[t * BT + tid];
                            if (x.y) { atomicAdd(&s_fx[t], (unsigned long long)x.y << kFxShift); x.y = 0u; cell[t * BT + tid] = x; }
                        }
                    }
                    const uint32_t n0 = rec_count(w0), hmax = rec_hmax(w0);
                    upc = kNone; cplx = false;
                    if (n0) {
                        const float2 vq = s_vq[n0 < 72u ? n0 : 0u];
                        v0 = vq.x; q0 = __float_as_uint(vq.y);
                        if (rec_is_mask(w0)) {
                            mlo = (uint32_t)w1; mhi = (uint32_t)(w1 >> 32);
                            if (n0 == 1u) atomicAdd(&s_un[w1 ? __ffsll((long long)w1) - 1 : 0], 1u);       // alone at k_low: a unique hit
                            if (hmax > 1u) {
                                const uint32_t e0 = (uint32_t)w2 & kPairMask;
                                if ((w2 >> kPairBits) == 0 && w3 == 0) {
                                    upc = (e0 & 0xffffu) - 1u; uph = e0 >> 16;                 // the usual case: one taxon, alone at its levels
                                    atomicAdd(&s_hist[upc * 8 + (uph & 7u)], 1u);
                                } else {
                                    // several taxa above k_low: level sizes (nibbles), level values, planes, counts
                                    const uint32_t e[kRecUpPairs] = {e0, (uint32_t)(w2 >> kPairBits) & kPairMask, (uint32_t)(w2 >> (2 * kPairBits)) & kPairMask,
                                                                     (uint32_t)w3 & kPairMask, (uint32_t)(w3 >> kPairBits) & kPairMask, (uint32_t)(w3 >> (2 * kPairBits)) & kPairMask};
                                    uint32_t njp = 0;                                          // nibble j - 1 = number of pairs reaching level j (<= 6)
#pragma unroll
                                    for (int k = 0; k < kRecUpPairs; ++k) { const uint32_t h = e[k] >> 16; njp += h ? ((1u << (4 * (h - 1u))) - 1u) & 0x11111u : 0u; }
                                    pl0 = pl1 = pl2 = 0;
#pragma unroll
                                    for (int j = 1; j < 6; ++j) s_val[(j - 1) * BT + tid] = __fmul_rn(wj[j], s_inv[(njp >> (4 * (j - 1))) & 15u]);
#pragma unroll
                                    for (int k = 0; k < kRecUpPairs; ++k) {
                                        const uint32_t h = e[k] >> 16, c = (e[k] & 0xffffu) - 1u;
                                        if (h < 2u) continue;
                                        pl0 |= (uint64_t)(h & 1u) << c; pl1 |= (uint64_t)((h >> 1) & 1u) << c; pl2 |= (uint64_t)(h >> 2) << c;
#pragma unroll
                                        for (int j = 1; j < 6; ++j) {
                                            if ((uint32_t)j >= h) break;
                                            const uint32_t nn = (njp >> (4 * (j - 1))) & 15u;
                                            if (nn == 1u) atomicAdd(&s_un[j * 64 + c], 1u);
                                            else atomicAdd(&s_fx[j * 64 + c], (unsigned long long)s_q32[nn]);
                                        }
                                    }
                                    cplx = true;
                                }
                            }
                        } else {
                            // list record inside a small taxonomy (more than six taxa above k_low, or a span read from HBM): the same from its pairs
                            const uint32_t *list = rec_overflow(w0) ? peers.base[rec_owner(w0)] + w1 : nullptr;
                            auto pair_at = [&](uint32_t k) -> uint32_t {
                                return list ? __ldg(list + k) : (uint32_t)((k < 3u ? w1 : k < 6u ? w2 : w3) >> (kPairBits * (k % 3u))) & kPairMask;
                            };
                            uint32_t nj[6] = {0, 0, 0, 0, 0, 0};
                            uint64_t mk = 0;
                            pl0 = pl1 = pl2 = 0;
                            for (uint32_t k = 0; k < n0; ++k) {
                                const uint32_t e = pair_at(k), h = e >> 16, c = (e & 0xffffu) - 1u;
                                mk |= 1ull << c;
#pragma unroll
                                for (int j = 1; j < 6; ++j) nj[j] += (uint32_t)j < h ? 1u : 0u;
                                if (h >= 2u) { pl0 |= (uint64_t)(h & 1u) << c; pl1 |= (uint64_t)((h >> 1) & 1u) << c; pl2 |= (uint64_t)((h >> 2) & 1u) << c; }
                            }
                            mlo = (uint32_t)mk; mhi = (uint32_t)(mk >> 32);
                            if (n0 == 1u) atomicAdd(&s_un[__ffsll((long long)mk) - 1], 1u);
#pragma unroll
                            for (int j = 1; j < 6; ++j) s_val[(j - 1) * BT + tid] = __fmul_rn(wj[j], s_inv[nj[j] < 72u ? nj[j] : 0u]);
                            for (uint32_t k = 0; k < n0 && hmax > 1u; ++k) {
                                const uint32_t e = pair_at(k), h = e >> 16, c = (e & 0xffffu) - 1u;
                                for (uint32_t j = 1; j < h && j < 6u; ++j) {
                                    if (nj[j] == 1u) atomicAdd(&s_un[j * 64 + c], 1u);
                                    else atomicAdd(&s_fx[j * 64 + c], (unsigned long long)s_q32[nj[j] < 72u ? nj[j] : 0u]);
                                }
                            }
                            cplx = (pl0 | pl1 | pl2) != 0;
                        }
                    }
                }
            }
            if (active && (mlo | mhi) != 0u) {
                // ---- one item: the next taxon of the record (Compare.hpp:530 for level k_low, then its levels above, in order)
                uint32_t c;
                if (!TWO || mlo) { c = (uint32_t)__ffs((int)mlo) - 1u; mlo &= mlo - 1u; }
                else { c = 32u + (uint32_t)__ffs((int)mhi) - 1u; mhi &= mhi - 1u; }
                const uint2 x = cell[c * BT + tid];
                float acc = __fadd_rn(__uint_as_float(x.x), v0);
                if (c == upc) {
#pragma unroll
                    for (int j = 1; j < 6; ++j) if ((uint32_t)j < uph) acc = __fadd_rn(acc, wj[j]);          // alone there: w * (1.f / 1) == w
                } else if (cplx) {
                    const uint32_t h = ((uint32_t)(pl0 >> c) & 1u) | (((uint32_t)(pl1 >> c) & 1u) << 1) | (((uint32_t)(pl2 >> c) & 1u) << 2);
#pragma unroll
                    for (int j = 1; j < 6; ++j) if ((uint32_t)j < h) acc = __fadd_rn(acc, s_val[(j - 1) * BT + tid]);
                }
                cell[c * BT + tid] = make_uint2(__float_as_uint(acc), x.y + q0);
            }
        }
        // ---- end of the reads of this pass: scores out, k_low sums into the CTA's totals
        if (have) {
            for (int t = 0; t < T - 1; ++t) {
                const uint2 x = cell[t * BT + tid];
                if (scores) scores[(uint64_t)r * T + t + 1] = __uint_as_float(x.x);
                if (x.x | x.y) cell[t * BT + tid] = make_uint2(0u, 0u);
                if (x.y) atomicAdd(&s_fx[t], (unsigned long long)x.y << kFxShift);
            }
            if (scores) scores[(uint64_t)r * T] = 0.f;
        }
    }
    __syncthreads();
    for (int i = tid; i < NK * 64; i += BT) {
        const int j = i / 64, c = i % 64;
        if (c >= T - 1) continue;
        unsigned long long u = s_un[i];
        if (j >= 1) for (int h = j + 1; h < 8; ++h) u += s_hist[c * 8 + h];     // alone at the levels 1 .. h-1
        if (u) atomicAdd(cnt_unique + (size_t)j * T + c + 1, u);
        if (s_fx[i]) atomicAdd(cnt_fx + (size_t)j * T + c + 1, s_fx[i]);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// PRESORTED READS, fourth organisation: WARP PER READ, LANE PER TAXON, UNIFORM CONTROL FLOW.
//   setup  lane = k-mer: the 32 records of a block arrive as one coalesced 1 KB load; each lane publishes, in shared memory, its
//          k-mer's taxa mask, its k_low value float(w_0 * (1.f / n0)) and what its up pairs need (levels, level values);
//   main   lane = taxon (lane l owns the taxa l + 1 and l + 33): the 32 k-mers are replayed in order with BROADCAST loads; a lane adds
//          when its bit is set (two predicated adds per k-mer), the lane named by an up pair adds the levels above — every branch is
//          warp-uniform, nothing is indexed dynamically, nothing depends on how many taxa a k-mer names.
// The sparse organisations above execute fewer useful lanes' worth of work but pay full price for every rare path a single lane takes;
// against a saturated index (C2: 0.6 up pairs per k-mer, 10 % of the records with several) this one issues the fewest instructions.
// one lane's share of a k-mer at one level: if its taxon's bit is set, the value joins the score and 1 / n the count — as two
// predicated instructions (the compiler's own choice was a select and an unconditional add for the count)
__device__ __forceinline__ void dense_add(float &score, uint32_t &share, float v, uint32_t q, uint32_t mask, uint32_t bit)
{
    asm("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\tand.b32 t, %4, %5;\n\tsetp.ne.u32 p, t, 0;\n\t@p add.rn.f32 %0, %0, %2;\n\t@p add.u32 %1, %1, %3;\n\t}"
        : "+f"(score), "+r"(share) : "f"(v), "r"(q), "r"(mask), "r"(bit));
}
// a taxon alone at the levels 2 .. 5 of the six-level instance (uc & 31 = the lane that holds its sum; the same uc in every lane): the
// level weights join that sum in order
__device__ __forceinline__ void dense_all(float &s, uint32_t uc, uint32_t lane, float w2, float w3, float w4, float w5)
{
    asm("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\tand.b32 t, %5, 31;\n\tsetp.eq.u32 p, t, %6;\n\t@p add.rn.f32 %0, %0, %1;\n\t@p add.rn.f32 %0, %0, %2;\n\t"
        "@p add.rn.f32 %0, %0, %3;\n\t@p add.rn.f32 %0, %0, %4;\n\t}"
        : "+f"(s) : "f"(w2), "f"(w3), "f"(w4), "f"(w5), "r"(uc), "r"(lane));
}

// the same for a taxon alone at the levels 2 .. h - 1, 3 <= h <= 6 (uc bits 8..12 = h)
__device__ __forceinline__ void dense_some(float &s, uint32_t uc, uint32_t lane, float w2, float w3, float w4, float w5)
{
    const uint32_t h = uc & 0x1f00u;
    if ((uc & 31u) == lane) {
        s = __fadd_rn(s, w2);
        if (h > (3u << 8)) s = __fadd_rn(s, w3);
        if (h > (4u << 8)) s = __fadd_rn(s, w4);
        if (h > (5u << 8)) s = __fadd_rn(s, w5);
    }
}

// a taxon among several at three or more levels (six-level instance): the lane that holds its sum adds the level values in order
__device__ __forceinline__ void dense_vals(float &s, uint32_t c, uint32_t h, uint32_t lane, float4 v)
{
    if ((c & 31u) == lane) {
        s = __fadd_rn(s, v.x);                                                                 // h >= 3
        if (h > 3u) s = __fadd_rn(s, v.y);
        if (h > 4u) s = __fadd_rn(s, v.z);
        if (h > 5u) s = __fadd_rn(s, v.w);
    }
}

// shared-memory loads that stay where they are written (the compiler would sink them to their first use behind a branch)
__device__ __forceinline__ uint4 lds_v4(const uint4 *p)
{
    uint4 r;
    asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
    return r;
}
__device__ __forceinline__ uint32_t lds_u32(const uint32_t *p)
{
    uint32_t r;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(r) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
    return r;
}

struct DenseKmer {            // 48 bytes: one address register reaches a, b and up of the k-mers i, i + 1, i + 2
    uint4 a;                  // taxa mask of level k_low (lo, hi), bits of float(w_0 * (1.f / n0)), 2^24 / n0 (0 for n0 < 2)
    uint4 b;                  // the same for level k_low + 1 (two-mask records, lookup.cuh)
    uint32_t up;              // what the k-mer adds above level k_low (kUp*): 0 = nothing; kUpMask = b only; with b: kUpAllA | lane / kUpAllB | lane = one taxon alone
                              // at the levels 2 .. 5 of the six-level instance, held by this lane in sa / sb (a read's own taxon); kUpOne |
                              // levels << 8 | taxon - 1 = one taxon alone at its levels 2 .. h - 1; kUpSeveral (pairs[], val[]); kUpList
                              // (list record: head[], pairs[], every level above k_low)
    uint32_t pad[3];
};
template <int ML>
struct DenseWarpMem {
    DenseKmer k[33];          // (k[32]: the loop loads one k-mer ahead)
    uint4 pairs[32];          // .xy = the record's pair word w3 when it holds several pairs; list records: the words w2, w3
    uint4 head[32];           // list records: the words w0, w1
    float val[32][ML <= 6 ? 4 : 24];   // float(w_j * (1.f / n_j)) for the levels j = 2 .. ML - 1 of a record with several pairs
};

// ML = 6: the default k = 12..7 (six levels, chains unrolled); ML = 26: up to 26 levels (k = 25..7 and every other range kASA allows)
// (measured: 64 registers / 4 CTAs per SM 93 ms, 85 registers / 2 CTAs 93 ms, the compiler's own 74 registers / 3 CTAs 77 ms;
//  4 instead of 8 warps per CTA 97 ms)
template <bool TWO, int ML>
__global__ void __launch_bounds__(kBlkWarps * 32) score_dense_kernel(const uint64_t *__restrict__ rec, const uint64_t *__restrict__ read_koff, uint32_t n_reads,
                                                                      LevelParams lp, OverflowPeers peers, float *__restrict__ scores,
                                                                      unsigned long long *__restrict__ cnt_fx, unsigned long long *__restrict__ cnt_unique,
                                                                      uint32_t *__restrict__ err_flag)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DenseWarpMem<ML> *s_w = reinterpret_cast<DenseWarpMem<ML> *>(smem_raw);        // [kBlkWarps]
    __shared__ float2 s_vq0[72], s_vq1[72];                             // [n]: .x = float(w_j * (1.f / n)) for j = 0, 1; .y = bits of 2^24 / n resp. 2^32 / n (0 for n < 2)
    __shared__ float s_inv[72];
    __shared__ uint32_t s_q32[72];                                      // 2^32 / n
    __shared__ unsigned long long s_fx[ML * 64];                        // CTA totals [level][taxon - 1], 2^-32 units
    __shared__ uint32_t s_un[ML * 64];
    __shared__ uint32_t s_hist[64 * 32];                                // [taxon - 1][h]: k-mers where the taxon is alone at its levels 2 .. h-1
    __shared__ float s_wl[32];                                          // level weights (Compare.hpp:392)
    const int tid = threadIdx.x, T = lp.n_taxa, NK = lp.num_k;
    const uint32_t lane = tid & 31u;
    const int w = tid >> 5;
    DenseWarpMem<ML> &M = s_w[w];
    float wj[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) { const int k = lp.k_low + j; wj[j] = (float)(k * k) / 625.f; }      // Compare.hpp:392
    for (int i = tid; i < ML * 64; i += blockDim.x) { s_fx[i] = 0; s_un[i] = 0; }
    for (int i = tid; i < 64 * 32; i += blockDim.x) s_hist[i] = 0;
    if (tid < 32) { const int kk = lp.k_low + tid; s_wl[tid] = (float)(kk * kk) / 625.f; }
    if (tid < 72) {
        const float inv = tid ? __fdiv_rn(1.f, (float)tid) : 0.f;       // Compare.hpp:924: 1.f / numOfEntries
        const float q = __uint_as_float(tid >= 2 ? (uint32_t)(((1u << 24) + (tid >> 1)) / tid) : 0u);
        s_inv[tid] = inv;
        s_vq0[tid] = make_float2(__fmul_rn(wj[0], inv), q);
        s_vq1[tid] = make_float2(__fmul_rn(wj[1], inv), q);
        s_q32[tid] = tid >= 2 ? inv_fx((uint32_t)tid) : 0u;
    }
    __syncthreads();
    // sums of 1 / n of this lane's taxa (lane + 1, lane + 33) at k_low and k_low + 1, 2^-24 units (shared-memory atomics per k-mer
    // for the second level were measured: + 8 ms per step)
    // (64-bit, widened from the per-read 32-bit registers; in shared memory, one private word per warp, sum and lane: registers are
    // what keeps three CTAs on an SM)
    __shared__ unsigned long long s_acc[kBlkWarps][4][32];
    for (int j = 0; j < 4; ++j) s_acc[w][j][lane] = 0;              // (every thread its own words: no barrier needed)
    const uint32_t lbit = 1u << lane;
    constexpr uint32_t kUpAllA = 0xE0000000u, kUpAllB = 0xC0000000u, kUpOne = 0x40000000u, kUpSeveral = 0x20000000u, kUpList = 0x10000000u, kUpMask = 0x08000000u;

    const uint32_t warps_total = gridDim.x * kBlkWarps;
    // The records of a read arrive one block of 32 ahead of their use, and the first block of the warp's NEXT read is fetched during
    // the last block of the current one (its offsets at the start of the current one): no load latency lies bare between reads.
    uint32_t r = blockIdx.x * kBlkWarps + w;
    uint64_t s0 = 0;
    uint32_t ni = 0;
    if (r < n_reads) { s0 = read_koff[r]; ni = (uint32_t)(read_koff[r + 1] - s0); }
    uint64_t nx[4] = {0, 0, 0, 0};
    if (lane < ni) asm volatile("ld.global.nc.L1::no_allocate.v4.b64 {%0, %1, %2, %3}, [%4];" : "=l"(nx[0]), "=l"(nx[1]), "=l"(nx[2]), "=l"(nx[3]) : "l"(rec + (s0 + lane) * kRecWords));
#pragma unroll 1
    for (; r < n_reads; r += warps_total) {
        const uint32_t rn = r + warps_total;
        uint64_t s0n = 0;
        uint32_t nin = 0;
        if (rn < n_reads) { s0n = __ldg(read_koff + rn); nin = (uint32_t)(__ldg(read_koff + rn + 1) - s0n); }
        const uint64_t *p = rec + s0 * kRecWords;
        float sa = 0.f, sb = 0.f;
        uint32_t fa = 0, fb = 0, ga = 0, gb = 0;
        if (ni == 0 && lane < nin) asm volatile("ld.global.nc.L1::no_allocate.v4.b64 {%0, %1, %2, %3}, [%4];" : "=l"(nx[0]), "=l"(nx[1]), "=l"(nx[2]), "=l"(nx[3]) : "l"(rec + (s0n + lane) * kRecWords));
#pragma unroll 1
        for (uint32_t b0 = 0; b0 < ni; b0 += 32) {
            // ---- setup, lane = k-mer b0 + lane
            const uint64_t w0 = nx[0], w1 = nx[1], w2 = nx[2], w3 = nx[3];
            nx[0] = nx[1] = nx[2] = nx[3] = 0;
            {
                const bool last = b0 + 32 >= ni;
                const uint64_t *q = last ? rec + (s0n + lane) * kRecWords : p + (uint64_t)(b0 + 32 + lane) * kRecWords;
                if (last ? lane < nin : b0 + 32 + lane < ni)
                    asm volatile("ld.global.nc.L1::no_allocate.v4.b64 {%0, %1, %2, %3}, [%4];" : "=l"(nx[0]), "=l"(nx[1]), "=l"(nx[2]), "=l"(nx[3]) : "l"(q));
            }
            const uint32_t n0 = rec_count(w0), hmax = rec_hmax(w0);
            uint64_t mk = n0 ? w1 : 0ull, m1 = 0;
            uint32_t upd = 0;
            if (n0 && !rec_is_mask(w0)) {
                // list record inside a small taxonomy (more than three taxa at three or more levels, or a span read from HBM; rare): mask
                // from its pairs here, its levels above k_low in the main pass (every lane looks its own taxa up in the pairs)
                const uint32_t *list = rec_overflow(w0) ? peers.base[rec_owner(w0)] + w1 : nullptr;
                mk = 0;
                for (uint32_t k = 0; k < n0; ++k) {
                    const uint32_t e = list ? __ldg(list + k) : (uint32_t)((k < 3u ? w1 : k < 6u ? w2 : w3) >> (kPairBits * (k % 3u))) & kPairMask;
                    mk |= 1ull << ((e & 0xffffu) - 1u);
                }
                if (hmax > 1u) {
                    upd = kUpList;
                    M.head[lane] = make_uint4((uint32_t)w0, (uint32_t)(w0 >> 32), (uint32_t)w1, (uint32_t)(w1 >> 32));
                    M.pairs[lane] = make_uint4((uint32_t)w2, (uint32_t)(w2 >> 32), (uint32_t)w3, (uint32_t)(w3 >> 32));
                }
            } else if (n0 && hmax > 1u) {
                if (!rec_is_mask2(w0)) atomicOr(err_flag, 512u);                              // this kernel reads two-mask records only
                m1 = w2;
                const uint32_t n1 = (uint32_t)__popcll(m1);
                if (n1 == 1u) atomicAdd(&s_un[64 + __ffsll((long long)m1) - 1], 1u);           // alone at k_low + 1
                upd = m1 ? kUpMask : 0u;
                const uint32_t e0 = m1 ? (uint32_t)w3 & kPairMask : 0u;
                if (e0 && (w3 >> kPairBits) == 0) {
                    // the usual case: one taxon, alone at its levels 2 ..
                    const uint32_t c0 = (e0 & 0xffffu) - 1u, h0 = e0 >> 16;
                    upd = (ML <= 6 && h0 == 6u) ? ((c0 & 32u) ? kUpAllB : kUpAllA) | (c0 & 31u) : kUpOne | (h0 << 8) | c0;
                    atomicAdd(&s_hist[((e0 & 0xffffu) - 1u) * 32 + ((e0 >> 16) & 31u)], 1u);
                } else if (e0) {
                    // several taxa at three or more levels: level sizes (nibbles), level values, counts
                    const uint32_t e[3] = {e0, (uint32_t)(w3 >> kPairBits) & kPairMask, (uint32_t)(w3 >> (2 * kPairBits)) & kPairMask};
                    uint64_t nlo = 0, nhi = 0;                                                // nibble j = number of pairs reaching level j (<= 3)
#pragma unroll
                    for (int k = 0; k < 3; ++k) { uint64_t x, y; nib_below(e[k] >> 16, x, y); nlo += x; nhi += y; }
                    const uint32_t hm = max(e[0] >> 16, max(e[1] >> 16, e[2] >> 16));
                    for (uint32_t j = 2; j < hm && j < (uint32_t)ML; ++j) M.val[lane][j - 2] = __fmul_rn(s_wl[j], s_inv[nib_get(nlo, nhi, j)]);
                    M.pairs[lane] = make_uint4((uint32_t)w3, (uint32_t)(w3 >> 32), 0u, 0u);
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const uint32_t h = e[k] >> 16, c = (e[k] & 0xffffu) - 1u;
                        for (uint32_t j = 2; j < h && j < (uint32_t)ML; ++j) {
                            const uint32_t nn = nib_get(nlo, nhi, j);
                            if (nn == 1u) atomicAdd(&s_un[j * 64 + c], 1u);
                            else atomicAdd(&s_fx[j * 64 + c], (unsigned long long)s_q32[nn]);
                        }
                    }
                    upd = kUpSeveral;
                }
            }
            if (n0 == 1u && mk) atomicAdd(&s_un[__ffsll((long long)mk) - 1], 1u);             // alone at k_low: a unique hit
            {
                const float2 vq = s_vq0[n0 < 72u ? n0 : 0u];
                M.k[lane].a = make_uint4((uint32_t)mk, (uint32_t)(mk >> 32), __float_as_uint(vq.x), __float_as_uint(vq.y));
                {
                    const float2 v1 = s_vq1[__popcll(m1)];                                    // (an empty mask adds nothing)
                    M.k[lane].b = make_uint4((uint32_t)m1, (uint32_t)(m1 >> 32), __float_as_uint(v1.x), __float_as_uint(v1.y));
                }
                M.k[lane].up = upd;
            }
            __syncwarp();
            // ---- main, lane = taxon: the block's k-mers in order (Compare.hpp:530 for level k_low, then the levels above)
            const uint32_t nblk = min(32u, ni - b0);
            // One k-mer (lane = taxon).  The loop is a chain of dependent instructions per warp — what it waits for is latency, not issue
            // slots — so the k-mer's second level and the NEXT k-mer's words are loaded before this k-mer's adds (two register sets,
            // alternating: no moves).
            auto kmer = [&](const uint32_t i, const uint4 a, const uint32_t uc, const uint4 b) {
                const float v0 = __uint_as_float(a.z);
                dense_add(sa, fa, v0, a.w, a.x, lbit);
                if (TWO) dense_add(sb, fb, v0, a.w, a.y, lbit);
                if (uc == 0u) return;                                   // more than half of C2's k-mers: no taxon above k_low
                {   // level k_low + 1: a second mask (empty in a list record)
                    const float v1 = __uint_as_float(b.z);
                    dense_add(sa, ga, v1, b.w, b.x, lbit);
                    if (TWO) dense_add(sb, gb, v1, b.w, b.y, lbit);
                }
                if (ML <= 6 && uc >= kUpAllB) {
                    // a read's own taxon, alone at the levels 2 .. 5: w * (1.f / 1) == w joins the sum of the lane (uc & 31), in sa or sb
                    if (!TWO || uc >= kUpAllA) dense_all(sa, uc, lane, wj[2], wj[3], wj[4], wj[5]);
                    else dense_all(sb, uc, lane, wj[2], wj[3], wj[4], wj[5]);
                    return;
                }
                if (uc == kUpMask) return;
                if (uc & kUpOne) {
                    // one taxon alone at its levels 2 .. h - 1 (h >= 3, the same in every lane; uc = kUpOne | h << 8 | taxon - 1)
                    if (ML <= 6) {
                        if (TWO && (uc & 32u)) dense_some(sb, uc, lane, wj[2], wj[3], wj[4], wj[5]);
                        else dense_some(sa, uc, lane, wj[2], wj[3], wj[4], wj[5]);
                    } else {
                        const bool mine = (uc & 31u) == lane;
                        const uint32_t h = (uc >> 8) & 31u;
                        float x = (TWO && (uc & 32u)) ? sb : sa;
                        for (uint32_t j = 2; j < h; ++j) x = __fadd_rn(x, s_wl[j]);
                        if (mine) { if (TWO && (uc & 32u)) sb = x; else sa = x; }
                    }
                    return;
                }
                if (uc == kUpList) {
                    // list record: level sizes by counting its pairs (every lane, the same values), this lane's taxa looked up on the way
                    const uint4 hd = M.head[i], pw = M.pairs[i];
                    const uint64_t l0 = ((uint64_t)hd.y << 32) | hd.x, l1 = ((uint64_t)hd.w << 32) | hd.z;
                    const uint64_t l2 = ((uint64_t)pw.y << 32) | pw.x, l3 = ((uint64_t)pw.w << 32) | pw.z;
                    const uint32_t np = rec_count(l0);
                    const uint32_t *list = rec_overflow(l0) ? peers.base[rec_owner(l0)] + l1 : nullptr;
                    uint32_t nj[ML], ha = 0, hb = 0;
#pragma unroll
                    for (int j = 0; j < ML; ++j) nj[j] = 0;
                    for (uint32_t k = 0; k < np; ++k) {
                        const uint32_t e = list ? __ldg(list + k) : (uint32_t)((k < 3u ? l1 : k < 6u ? l2 : l3) >> (kPairBits * (k % 3u))) & kPairMask;
                        const uint32_t c = (e & 0xffffu) - 1u, h = e >> 16;
#pragma unroll
                        for (int j = 1; j < ML; ++j) nj[j] += (uint32_t)j < h ? 1u : 0u;
                        if (c == lane) ha = h;
                        if (c == lane + 32u) hb = h;
                    }
#pragma unroll
                    for (int j = 1; j < ML; ++j) {
                        const uint32_t nn = nj[j] < 72u ? nj[j] : 0u;
                        const float vj = __fmul_rn(s_wl[j], s_inv[nn]);
                        if ((uint32_t)j < ha) {
                            sa = __fadd_rn(sa, vj);
                            if (nn == 1u) atomicAdd(&s_un[j * 64 + lane], 1u); else atomicAdd(&s_fx[j * 64 + lane], (unsigned long long)s_q32[nn]);
                        }
                        if ((uint32_t)j < hb) {
                            sb = __fadd_rn(sb, vj);
                            if (nn == 1u) atomicAdd(&s_un[j * 64 + lane + 32], 1u); else atomicAdd(&s_fx[j * 64 + lane + 32], (unsigned long long)s_q32[nn]);
                        }
                    }
                    return;
                }
                {   // several taxa at three or more levels
                    const uint4 pw = M.pairs[i];
                    const uint64_t p3 = ((uint64_t)pw.y << 32) | pw.x;
                    const float4 vv = *reinterpret_cast<const float4 *>(&M.val[i][0]);          // (six-level instance: the values of the levels 2 .. 5)
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const uint32_t e = (uint32_t)(p3 >> (kPairBits * k)) & kPairMask;
                        if (!e) break;                                                         // uniform: pairs fill from slot 0
                        const uint32_t c = (e & 0xffffu) - 1u, h = e >> 16;
                        if (ML <= 6) {
                            if (TWO && c >= 32u) dense_vals(sb, c, h, lane, vv); else dense_vals(sa, c, h, lane, vv);
                        } else {
                            const bool mine = (c & 31u) == lane;
                            float x = (TWO && c >= 32u) ? sb : sa;
                            for (uint32_t j = 2; j < h && j < (uint32_t)ML; ++j) x = __fadd_rn(x, M.val[i][j - 2]);
                            if (mine) { if (TWO && c >= 32u) sb = x; else sa = x; }
                        }
                    }
                }
            };
            uint4 a0 = lds_v4(&M.k[0].a), a1;
            uint32_t u0 = lds_u32(&M.k[0].up), u1;
#pragma unroll 1
            for (uint32_t i = 0; i < nblk; i += 2) {
                uint4 b = lds_v4(&M.k[i].b);
                a1 = lds_v4(&M.k[i + 1].a);
                u1 = lds_u32(&M.k[i + 1].up);
                kmer(i, a0, u0, b);
                if (i + 1 >= nblk) break;
                b = lds_v4(&M.k[i + 1].b);
                a0 = lds_v4(&M.k[i + 2].a);
                u0 = lds_u32(&M.k[i + 2].up);
                kmer(i + 1, a1, u1, b);
            }
            // a k-mer adds at most 2^23 units to a sum: widen before 32 bits could overflow
            if ((b0 & 255u) == 224u) { s_acc[w][0][lane] += fa; s_acc[w][1][lane] += fb; s_acc[w][2][lane] += ga; s_acc[w][3][lane] += gb; fa = fb = ga = gb = 0; }
            __syncwarp();
        }
        s_acc[w][0][lane] += fa; s_acc[w][1][lane] += fb; s_acc[w][2][lane] += ga; s_acc[w][3][lane] += gb;
        if (scores) {
            float *row = scores + (uint64_t)r * T;
            if (lane == 0) row[0] = 0.f;
            if ((int)lane + 1 < T) row[lane + 1] = sa;
            if (TWO && (int)lane + 33 < T) row[lane + 33] = sb;
        }
        s0 = s0n; ni = nin;
    }
    {
        const unsigned long long fa64 = s_acc[w][0][lane], fb64 = s_acc[w][1][lane], ga64 = s_acc[w][2][lane], gb64 = s_acc[w][3][lane];
        if ((int)lane + 1 < T) {
            if (fa64) atomicAdd(&s_fx[lane], fa64 << kFxShift);
            if (ga64) atomicAdd(&s_fx[64 + lane], ga64 << kFxShift);
        }
        if (TWO && (int)lane + 33 < T) {
            if (fb64) atomicAdd(&s_fx[lane + 32], fb64 << kFxShift);
            if (gb64) atomicAdd(&s_fx[64 + lane + 32], gb64 << kFxShift);
        }
    }
    __syncthreads();
    for (int i = tid; i < NK * 64; i += blockDim.x) {
        const int j = i / 64, c = i % 64;
        if (c >= T - 1) continue;
        unsigned long long u = s_un[i];
        if (j >= 2) for (int h = j + 1; h < 32; ++h) u += s_hist[c * 32 + h];   // alone at the levels 2 .. h-1
        if (u) atomicAdd(cnt_unique + (size_t)j * T + c + 1, u);
        if (s_fx[i]) atomicAdd(cnt_fx + (size_t)j * T + c + 1, s_fx[i]);
    }
}

}  // namespace

void launch_order_ops(cudaStream_t s, const uint64_t *rec, const uint64_t *read_koff, uint32_t r0, uint32_t r1, uint64_t n_kmers, uint32_t max_inst,
                      LevelParams lp, OverflowPeers peers, OpsWorkspace ws, const uint32_t *trig, bool counts, bool emit, unsigned long long *cnt_fx,
                      unsigned long long *cnt_unique, uint32_t *err_flag, uint64_t *launches)
{
    if (r1 <= r0) return;
    const bool exact = trig != nullptr;
    KASA_REQUIRE(!counts || lp.n_taxa <= kMaskTaxa, -6, "counts in the ordering kernel need <= 64 taxa");
    // instances per shared-memory segment: the whole read when it is short, else a slice of the rank range (ORDERED only);
    // four warps (reads) per CTA, fewer when a read needs more shared memory
    uint32_t cap = max_inst < 32 ? 32 : (max_inst + 31) & ~31u;
    uint32_t icap = 0, nb = 0, warps = kAMaxWarps;
    size_t wbytes = 0, sm = 0;
    const size_t budget = 224 * 1024;
    const int count_levels = counts && !(!exact && ops_mask_stream(lp) && emit) ? lp.num_k : 0;
    KASA_REQUIRE(!exact || cap <= 4096, -1, "KASA_SCORE_EXACT: a read has too many k-mers for the shared-memory replay; use KASA_SCORE_ORDERED for long reads");
    // EXACT sorts (k-mer, level) items: at most cap * num_k of them, in practice far fewer (only levels with a hit); the kernel reports
    // a read that does not fit (error 64)
    uint32_t items = exact ? std::min<uint32_t>(cap * (uint32_t)lp.num_k, 65535u) : 0u;
    for (;;) {
        icap = exact ? items : cap;
        nb = 64;
        while (nb < icap && nb < 4096) nb <<= 1;
        if (nb > 256 && icap <= 512) nb = 256;                  // short reads: 256 buckets (about one item each) keep four CTAs on an SM
        wbytes = warp_mem_bytes(cap, icap, nb, count_levels);
        sm = wbytes * warps;
        if (sm <= budget) break;
        if (!exact && cap > 512) { cap = (cap / 2 + 31) & ~31u; continue; }
        if (warps > 1) { warps >>= 1; continue; }
        if (exact && items > 2 * cap) { items = items * 3 / 4; continue; }
        if (!exact && cap > 64) { cap = (cap / 2 + 31) & ~31u; continue; }
        break;
    }
    if (sm > budget)
        throw Error{-1, "KASA_SCORE_EXACT: a read has too many (k-mer, level) pairs for the shared-memory replay; use KASA_SCORE_ORDERED for long reads"};
    auto go = [&](auto kern) {
        KASA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        int per_sm = 1;
        KASA_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, (int)warps * 32, sm));
        unsigned grid = (unsigned)div_up<uint64_t>(r1 - r0, warps);
        const unsigned resident = (unsigned)(kSMs * (per_sm > 0 ? per_sm : 1));
        if (grid > resident) grid = resident;
        kern<<<grid, warps * 32, sm, s>>>(rec, read_koff, r0, r1, n_kmers, lp, peers, ws, trig, err_flag, cap, icap, nb, (uint32_t)wbytes, cnt_fx, cnt_unique, emit || !counts);
    };
    const bool masks = !exact && ops_mask_stream(lp) && counts && emit;       // mask stream: small taxonomy scored densely; counts in kernel B
    if (exact) { if (counts) go(order_ops_kernel<1, true>); else go(order_ops_kernel<1, false>); }
    else if (masks) go(order_ops_kernel<2, false>);
    else { if (counts) go(order_ops_kernel<0, true>); else go(order_ops_kernel<0, false>); }
    ++*launches;
    KASA_CUDA_TRY(cudaGetLastError());
}

void launch_accumulate(cudaStream_t s, OpsWorkspace ws, uint32_t r0, uint32_t r1, LevelParams lp, bool exact_stream, float *scores, HitsOut hits,
                       unsigned long long *cnt_fx, unsigned long long *cnt_unique, uint32_t *err_flag, uint64_t *launches)
{
    if (r1 <= r0) return;
    if (lp.n_taxa <= kMaskTaxa && !hits.taxon) {
        // the counts of these reads were taken by launch_order_ops(counts = true)
        if (!exact_stream && ops_mask_stream(lp)) {
            auto go = [&](auto kern, int nc) {
                const size_t sm = (size_t)nc * kMaskBlock * 8;
                KASA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
                int per_sm = 1;
                KASA_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kMaskBlock, sm));
                unsigned grid = (unsigned)div_up<uint64_t>(r1 - r0, kMaskBlock);
                const unsigned resident = (unsigned)(kSMs * (per_sm > 0 ? per_sm : 1));
                if (grid > resident) grid = resident;
                kern<<<grid, kMaskBlock, sm, s>>>(ws, r0, r1, lp, scores, cnt_fx, cnt_unique);
            };
            if (lp.n_taxa <= 33) go(accumulate_mask_kernel<1>, 32); else go(accumulate_mask_kernel<2>, 64);
        } else if (lp.n_taxa <= 33) launch_dense<1>(s, ws, r0, r1, lp, exact_stream, scores);
        else launch_dense<2>(s, ws, r0, r1, lp, exact_stream, scores);
    } else {
        unsigned grid = (unsigned)div_up<uint64_t>(r1 - r0, kHWarps);
        if (grid > (unsigned)kSMs * 8) grid = kSMs * 8;
        accumulate_hash_kernel<<<grid, kHBlock, 0, s>>>(ws, r0, r1, lp, exact_stream, scores, hits, cnt_fx, cnt_unique, err_flag);
    }
    ++*launches;
    KASA_CUDA_TRY(cudaGetLastError());
}

void launch_max_instances(cudaStream_t s, const uint64_t *read_koff, uint32_t n_reads, uint32_t *out, uint64_t *launches)
{
    KASA_CUDA_TRY(cudaMemsetAsync(out, 0, 4, s));
    if (!n_reads) return;
    max_instances_kernel<<<div_up(n_reads, 256u), 256, 0, s>>>(read_koff, n_reads, out);
    ++*launches;
}

void launch_score_sorted(cudaStream_t s, const uint64_t *rec, const uint64_t *read_koff, uint32_t n_reads, LevelParams lp, OverflowPeers peers,
                         float *scores, unsigned long long *cnt_fx, unsigned long long *cnt_unique, int lanes_hint, uint32_t *err_flag, uint64_t *launches)
{
    if (!n_reads) return;
    KASA_REQUIRE(lp.n_taxa <= kMaskTaxa, -6, "the presorted scoring kernels need <= 64 taxa");
    if (lp.num_k > 6) lanes_hint = 0;                  // only the dense organisation takes more than six levels
    auto go = [&](auto kern, int lanes) {
        int per_sm = 1;
        KASA_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kSortedBlock, 0));
        unsigned grid = (unsigned)div_up<uint64_t>(n_reads, kSortedBlock / lanes);
        const unsigned resident = (unsigned)(kSMs * (per_sm > 0 ? per_sm : 1));
        if (grid > resident) grid = resident;
        kern<<<grid, kSortedBlock, 0, s>>>(rec, read_koff, n_reads, lp, peers, scores, cnt_fx, cnt_unique);
    };
    const int taxa = lp.n_taxa - 1;
    if (lanes_hint == 100) {
        // thread per read, one (k-mer, taxon) item per iteration (score_items_kernel)
        auto goi = [&](auto kern) {
            const size_t sm = (size_t)std::max(taxa, 1) * kItemBlock * 8 + (size_t)5 * kItemBlock * 4;
            KASA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
            int per_sm = 1;
            KASA_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kItemBlock, sm));
            unsigned grid = (unsigned)div_up<uint64_t>(n_reads, kItemBlock);
            const unsigned resident = (unsigned)(kSMs * (per_sm > 0 ? per_sm : 1));
            if (grid > resident) grid = resident;
            kern<<<grid, kItemBlock, sm, s>>>(rec, read_koff, n_reads, lp, peers, scores, cnt_fx, cnt_unique);
        };
        if (taxa > 32) goi(score_items_kernel<true>); else goi(score_items_kernel<false>);
        ++*launches;
        KASA_CUDA_TRY(cudaGetLastError());
        return;
    }
    if (lanes_hint == 0 || lanes_hint == 64) {
        // warp per read, lane per taxon, uniform control flow over two-mask records (score_dense_kernel): the default
        auto god = [&](auto kern, size_t sm) {
            KASA_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
            int per_sm = 1;
            KASA_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kBlkWarps * 32, sm));
            unsigned grid = (unsigned)div_up<uint64_t>(n_reads, kBlkWarps);
            const unsigned resident = (unsigned)(kSMs * (per_sm > 0 ? per_sm : 1));
            if (grid > resident) grid = resident;
            kern<<<grid, kBlkWarps * 32, sm, s>>>(rec, read_koff, n_reads, lp, peers, scores, cnt_fx, cnt_unique, err_flag);
        };
        constexpr size_t sm6 = sizeof(DenseWarpMem<6>) * kBlkWarps, sm26 = sizeof(DenseWarpMem<26>) * kBlkWarps;
        if (lp.num_k <= 6) { if (taxa > 32) god(score_dense_kernel<true, 6>, sm6); else god(score_dense_kernel<false, 6>, sm6); }
        else { if (taxa > 32) god(score_dense_kernel<true, 26>, sm26); else god(score_dense_kernel<false, 26>, sm26); }
        ++*launches;
        KASA_CUDA_TRY(cudaGetLastError());
        return;
    }
    if (lanes_hint == 32) {
        // warp per read, 32 k-mers at a time (score_blocks_kernel)
        auto gow = [&](auto kern) {
            int per_sm = 1;
            KASA_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kBlkWarps * 32, 0));
            unsigned grid = (unsigned)div_up<uint64_t>(n_reads, kBlkWarps);
            const unsigned resident = (unsigned)(kSMs * (per_sm > 0 ? per_sm : 1));
            if (grid > resident) grid = resident;
            kern<<<grid, kBlkWarps * 32, 0, s>>>(rec, read_koff, n_reads, lp, peers, scores, cnt_fx, cnt_unique);
        };
        if (taxa > 32) gow(score_blocks_kernel<true>); else gow(score_blocks_kernel<false>);
        ++*launches;
        KASA_CUDA_TRY(cudaGetLastError());
        return;
    }
    // threads per read x taxa per thread (lanes_hint = 1, 2 or 4: tuning knob KASA_TUNE_SCORE_LANES)
    if (taxa <= 8) go(score_sorted_kernel<1, 8>, 1);
    else if (taxa <= 16) { if (lanes_hint == 1) go(score_sorted_kernel<1, 16>, 1); else go(score_sorted_kernel<2, 8>, 2); }
    else if (taxa <= 32) { if (lanes_hint == 1) go(score_sorted_kernel<1, 32>, 1); else if (lanes_hint == 2) go(score_sorted_kernel<2, 16>, 2); else go(score_sorted_kernel<4, 8>, 4); }
    else if (taxa <= 52) { if (lanes_hint == 2) go(score_sorted_kernel<2, 26>, 2); else go(score_sorted_kernel<4, 13>, 4); }
    else { if (lanes_hint == 2) go(score_sorted_kernel<2, 32>, 2); else go(score_sorted_kernel<4, 16>, 4); }
    ++*launches;
    KASA_CUDA_TRY(cudaGetLastError());
}

}  // namespace kasa
