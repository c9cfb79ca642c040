// radix_sort.cu — see radix_sort.cuh.
#include <cstdlib>

#include "radix_sort.cuh"

namespace kasa {

namespace {

constexpr uint64_t kFlagAgg = 1ull << 62;    // tile aggregate published
constexpr uint64_t kFlagIncl = 2ull << 62;   // inclusive prefix published
constexpr uint64_t kValMask = (1ull << 62) - 1;
constexpr int kLookBatch = 2;                // predecessor status words requested per look-back round trip
constexpr uint32_t kSpinLimit = 1u << 22;    // watchdog: ~seconds, never reached when healthy

struct PlanDev {
    int n_passes;
    int src[kMaxPasses];
    int shift[kMaxPasses];
    uint32_t mask[kMaxPasses];
};

// ---------------------------------------------------------------------------------------------
// Up-front histograms of every pass digit: one read of the record arrays that feed a digit.
// grid = multiple of 148 SMs; shared-memory counters, one global atomic per (block, pass, bin).
template <bool HAS_W1>
__global__ void __launch_bounds__(512) hist_kernel(const uint64_t *__restrict__ w0, const uint64_t *__restrict__ w1,
                                                   const uint32_t *__restrict__ v, uint64_t n, PlanDev plan,
                                                   bool need_w0, bool need_w1, bool need_v,
                                                   unsigned long long *__restrict__ ghist)
{
    extern __shared__ uint32_t sh[];   // [n_passes][256]
    const int P = plan.n_passes;
    for (int i = threadIdx.x; i < P * kRadix; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    // four independent loads per thread and iteration: the kernel is bound by bytes in flight, not by the atomics
    constexpr int U = 4;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += U * stride) {
        uint64_t a[U], b[U], c[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint64_t i = i0 + u * stride;
            const bool ok = i < n;
            a[u] = (need_w0 && ok) ? ldg_stream(w0 + i) : 0;
            b[u] = (HAS_W1 && need_w1 && ok) ? ldg_stream(w1 + i) : 0;
            c[u] = (need_v && ok) ? (uint64_t)ldg_stream(v + i) : 0;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (i0 + u * stride >= n) break;
#pragma unroll 4
            for (int p = 0; p < P; ++p) {
                uint64_t word = plan.src[p] == SRC_W0 ? a[u] : (plan.src[p] == SRC_W1 ? b[u] : c[u]);
                uint32_t d = (uint32_t)(word >> plan.shift[p]) & plan.mask[p];
                atomicAdd(&sh[p * kRadix + d], 1u);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < P * kRadix; i += blockDim.x)
        if (sh[i]) atomicAdd(ghist + i, (unsigned long long)sh[i]);
}

// The identify sort takes every digit from the low k-mer word (<= 8 passes): the same histograms with the pass list
// unrolled at compile time and the shifts / masks in registers — the general kernel above spends two thirds of its
// instructions decoding the plan per (record, pass).
template <int NP>
__global__ void __launch_bounds__(512) hist_w0_kernel(const uint64_t *__restrict__ w0, uint64_t n, PlanDev plan,
                                                      unsigned long long *__restrict__ ghist)
{
    __shared__ uint32_t sh[NP * kRadix];
    for (int i = threadIdx.x; i < NP * kRadix; i += blockDim.x) sh[i] = 0;
    int shift[NP];
    uint32_t mask[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) { shift[p] = plan.shift[p]; mask[p] = plan.mask[p]; }
    __syncthreads();
    constexpr int U = 4;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += U * stride) {
        uint64_t a[U];
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint64_t i = i0 + u * stride;
            ok[u] = i < n;
            a[u] = ok[u] ? ldg_stream(w0 + i) : 0;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!ok[u]) continue;
#pragma unroll
            for (int p = 0; p < NP; ++p) atomicAdd(&sh[p * kRadix + ((uint32_t)(a[u] >> shift[p]) & mask[p])], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NP * kRadix; i += blockDim.x)
        if (sh[i]) atomicAdd(ghist + i, (unsigned long long)sh[i]);
}

// exclusive scan of each pass histogram (one block of 256 threads per pass), in place
__global__ void __launch_bounds__(kRadix) hist_scan_kernel(unsigned long long *ghist)
{
    __shared__ unsigned long long s[kRadix];
    unsigned long long *h = ghist + (size_t)blockIdx.x * kRadix;
    s[threadIdx.x] = h[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long run = 0;
        for (int i = 0; i < kRadix; ++i) { unsigned long long c = s[i]; s[i] = run; run += c; }
    }
    __syncthreads();
    h[threadIdx.x] = s[threadIdx.x];
}

// ---------------------------------------------------------------------------------------------
// One radix pass.  Tile = BLOCK x ITEMS records, warp-striped so that loads coalesce and the
// (warp, item, lane) order equals the input order (needed for LSD stability).
//
// Resource plan (B200: 64 K registers, 227 KB shared memory per SM): 256 threads x 18 items at 80 registers
// -> 3 CTAs = 24 warps per SM; 68 KB dynamic shared memory per CTA.  The kernel is latency bound per tile
// (rank chain + look-back): many items per thread give independent work between the barriers, three
// resident CTAs overlap the phases of different tiles (shapes measured: radix_sort.cuh).
constexpr uint32_t kPrefetchAhead = kSortCtas * kSMs;   // CTAs resident at once

template <int BLOCK, int ITEMS, bool HAS_W1>
__global__ void __launch_bounds__(BLOCK, kSortCtas) onesweep_kernel(
    const uint64_t *__restrict__ in_w0, const uint64_t *__restrict__ in_w1, const uint32_t *__restrict__ in_v,
    uint64_t *__restrict__ out_w0, uint64_t *__restrict__ out_w1, uint32_t *__restrict__ out_v, uint64_t n, int src,
    int shift, uint32_t mask, const unsigned long long *__restrict__ digit_start, uint64_t *status,
    uint32_t *ticket, uint32_t *err_flag, uint32_t prefetch_ahead)
{
    constexpr int TILE = BLOCK * ITEMS;
    constexpr int NW = BLOCK / 32;
    static_assert(BLOCK >= kRadix, "one look-back thread per digit");
    static_assert(TILE <= 65536, "tile positions are kept in 16 bits");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *s_buf = reinterpret_cast<uint64_t *>(smem_raw);                       // TILE x 8 B exchange buffer
    uint32_t *s_hist = reinterpret_cast<uint32_t *>(smem_raw + (size_t)TILE * 8);   // [NW][256]
    uint8_t *s_dig = reinterpret_cast<uint8_t *>(s_hist + NW * kRadix);             // [TILE]
    __shared__ uint32_t s_dbase[kRadix];
    __shared__ uint64_t s_gbase[kRadix];
    __shared__ uint32_t s_wsum[kRadix / 32];
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_cnt[kRadix];

    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    if (tid < kRadix) s_cnt[tid] = 0;   // ticket order = start order: predecessors are resident or done
    {
        uint4 *h4 = reinterpret_cast<uint4 *>(s_hist + w * kRadix);          // 256 counters = 64 x 16 bytes per warp
        h4[lane] = make_uint4(0, 0, 0, 0);
        h4[lane + 32] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint64_t base = (uint64_t)tile * TILE;
    const uint32_t n_tile = (uint32_t)min((uint64_t)TILE, n - base);
    const uint64_t wbase = base + (uint64_t)w * (32 * ITEMS) + lane;
    const bool v_is_src = (src == SRC_V);

    // ---- pull the tile that will start when the resident ones retire into L2 (tickets are handed out in order, so
    // that is tile + resident CTAs): its loads then see L2 latency instead of HBM latency at the head of its critical path
    {
        const uint64_t pbase = ((uint64_t)tile + prefetch_ahead) * TILE;
        if (prefetch_ahead && pbase < n) {
            // 128-byte lines: TILE/16 of the 64-bit words, then TILE/32 of the payload
            const uint32_t m = (uint32_t)min((uint64_t)TILE, n - pbase);
            for (uint32_t l = tid; l < TILE / 16 + TILE / 32; l += BLOCK) {
                if (l < TILE / 16) {
                    if (l * 16 < m) {
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(in_w0 + pbase + l * 16));
                        if (HAS_W1) asm volatile("prefetch.global.L2 [%0];" ::"l"(in_w1 + pbase + l * 16));
                    }
                } else {
                    const uint32_t lv = l - TILE / 16;
                    if (lv * 32 < m) asm volatile("prefetch.global.L2 [%0];" ::"l"(in_v + pbase + lv * 32));
                }
            }
        }
    }

    // ---- load the digit source
    uint64_t key[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        uint64_t idx = wbase + (uint64_t)i * 32;
        bool ok = idx < n;
        if (src == SRC_W0) key[i] = ok ? ldg_stream(in_w0 + idx) : ~0ull;
        else if (HAS_W1 && src == SRC_W1) key[i] = ok ? ldg_stream(in_w1 + idx) : ~0ull;
        else key[i] = ok ? (uint64_t)ldg_stream(in_v + idx) : ~0ull;
    }

    // ---- the tile's digit counts first (shared-memory atomics) and out to the status word: successors doing their
    // look-back need nothing else from this tile, so its ranking below is off their critical path
    {
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) atomicAdd(&s_cnt[(uint32_t)(key[i] >> shift) & mask], 1u);
        __syncthreads();
        if (tid < kRadix) {
            uint32_t c = s_cnt[tid];
            if ((uint32_t)tid == mask) c -= (uint32_t)(TILE - n_tile);               // padding sorts last inside digit `mask`
            st_relaxed_gpu(status + (uint64_t)tile * kRadix + tid, (tile == 0 ? kFlagIncl : kFlagAgg) | (uint64_t)c);
        }
    }

    // ---- rank inside the warp, digit by digit (match-any multisplit); padding keys carry digit `mask`
    uint16_t rk[ITEMS];
    uint32_t *my_hist = s_hist + w * kRadix;
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        uint32_t d = (uint32_t)(key[i] >> shift) & mask;
        // lanes holding the same digit: eight ballots (VOTE is far cheaper than MATCH.ANY on the MIO pipe)
        uint32_t peers = kFull;
#pragma unroll
        for (int bit = 0; bit < kRadixBits; ++bit) {
            const bool one = (d >> bit) & 1u;
            const uint32_t vote = __ballot_sync(kFull, one);
            peers &= one ? vote : ~vote;
        }
        int leader = __ffs(peers) - 1;
        uint32_t before = __popc(peers & lanemask_lt());
        uint32_t prev = 0;
        if (lane == leader) {
            prev = my_hist[d];
            my_hist[d] = prev + __popc(peers);
        }
        prev = __shfl_sync(kFull, prev, leader);
        rk[i] = (uint16_t)(prev + before);
        __syncwarp();
    }
    __syncthreads();

    // ---- per digit: exclusive offsets over warps, tile count, tile-local digit base
    uint32_t count = 0;
    if (tid < kRadix) {
        uint32_t run = 0;
#pragma unroll
        for (int ww = 0; ww < NW; ++ww) {
            uint32_t c = s_hist[ww * kRadix + tid];
            s_hist[ww * kRadix + tid] = run;
            run += c;
        }
        count = run;
        if ((uint32_t)tid == mask) count -= (uint32_t)(TILE - n_tile);   // padding sorts last inside digit `mask`
    }
    uint32_t incl = count;
    if (tid < kRadix) {
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(kFull, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_wsum[w] = incl;
    }
    __syncthreads();
    if (tid < 32) {
        uint32_t x = (tid < kRadix / 32) ? s_wsum[tid] : 0;
        uint32_t xi = x;
#pragma unroll
        for (int o = 1; o < kRadix / 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(kFull, xi, o);
            if (lane >= o) xi += t;
        }
        if (tid < kRadix / 32) s_wsum[tid] = xi - x;
    }
    __syncthreads();

    // ---- the 32-bit payload is requested now: its latency hides behind the look-back
    uint32_t val[ITEMS];
    if (!v_is_src) {
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            uint64_t idx = wbase + (uint64_t)i * 32;
            val[i] = idx < n ? ldg_stream(in_v + idx) : 0u;
        }
    } else {
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) val[i] = (uint32_t)key[i];
    }

    // ---- decoupled look-back, one thread per digit
    if (tid < kRadix) {
        const uint32_t dbase = incl - count + s_wsum[w];
        uint64_t *my = status + (uint64_t)tile * kRadix + tid;
        uint64_t excl = 0;
        if (tile != 0) {
            // predecessors are read kLookBatch at a time (independent loads in flight); tiles before the first count
            // as "inclusive, 0"; an unpublished word restarts the batch at that tile
            int64_t t = (int64_t)tile - 1;
            uint32_t spins = 0;
            bool done = false;
            while (!done) {
                uint64_t sv[kLookBatch];
#pragma unroll
                for (int k = 0; k < kLookBatch; ++k)
                    sv[k] = (t - k >= 0) ? ld_relaxed_gpu(status + (uint64_t)(t - k) * kRadix + tid) : kFlagIncl;
                int used = 0;
#pragma unroll
                for (int k = 0; k < kLookBatch; ++k) {
                    const uint64_t f = sv[k] >> 62;
                    if (!done && used == k && f != 0) {
                        excl += sv[k] & kValMask;
                        ++used;
                        if (f == 2) done = true;
                    }
                }
                t -= used;
                if (!done && used < kLookBatch) {               // ran into a tile that has not published yet
                    if (++spins > kSpinLimit || ((spins & 1023u) == 0 && *(volatile uint32_t *)err_flag)) {
                        atomicExch(err_flag, 1u);
                        break;
                    }
                    __nanosleep(20);
                }
            }
            st_relaxed_gpu(my, kFlagIncl | (excl + count));
        }
        s_dbase[tid] = dbase;
        s_gbase[tid] = (uint64_t)digit_start[tid] + excl - dbase;
    }
    __syncthreads();

    // ---- exchange through shared memory: digit-sorted tile, then coalesced runs to HBM
    uint32_t *s_val = reinterpret_cast<uint32_t *>(s_dig + TILE);                    // TILE x 4 B payload buffer
    uint16_t pos[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        uint32_t d = (uint32_t)(key[i] >> shift) & mask;
        uint32_t p = s_dbase[d] + my_hist[d] + rk[i];
        pos[i] = (uint16_t)p;
        s_dig[p] = (uint8_t)d;
        s_val[p] = val[i];
        if (!v_is_src) s_buf[p] = key[i];
    }
    __syncthreads();
    {
        uint64_t *out_src = (HAS_W1 && src == SRC_W1) ? out_w1 : out_w0;
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            uint32_t p = tid + j * BLOCK;
            if (p < n_tile) {
                const uint64_t g = s_gbase[s_dig[p]] + p;
                if (!v_is_src) out_src[g] = s_buf[p];
                out_v[g] = s_val[p];
            }
        }
    }
    // the other 64-bit word(s)
    const int n_other = v_is_src ? (HAS_W1 ? 2 : 1) : (HAS_W1 ? 1 : 0);
    for (int o = 0; o < n_other; ++o) {
        const bool take_w1 = HAS_W1 && (v_is_src ? (o == 1) : (src == SRC_W0));
        const uint64_t *inp = take_w1 ? in_w1 : in_w0;
        uint64_t *outp = take_w1 ? out_w1 : out_w0;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            uint64_t idx = wbase + (uint64_t)i * 32;
            if (idx < n) s_buf[pos[i]] = ldg_stream(inp + idx);
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            uint32_t p = tid + j * BLOCK;
            if (p < n_tile) outp[s_gbase[s_dig[p]] + p] = s_buf[p];
        }
    }
}

template <int BLOCK, int ITEMS>
constexpr size_t onesweep_smem()
{
    return (size_t)BLOCK * ITEMS * 8 + (size_t)(BLOCK / 32) * kRadix * 4 + (size_t)BLOCK * ITEMS + (size_t)BLOCK * ITEMS * 4;
}

}  // namespace

SortPlan plan_for_bit_range(int bit_lo, int bit_hi)
{
    SortPlan p{};
    int done = bit_lo;
    while (done < bit_hi) {
        int src = done < 64 ? SRC_W0 : SRC_W1;
        int in_word = done < 64 ? done : done - 64;
        int bits = kRadixBits;
        if (in_word + bits > 64) bits = 64 - in_word;
        if (done + bits > bit_hi) bits = bit_hi - done;
        p.pass[p.n_passes++] = SortPass{src, in_word, (1u << bits) - 1u};
        done += bits;
    }
    return p;
}

SortPlan plan_for_kmer_bits(int key_bits) { return plan_for_bit_range(0, key_bits); }

int radix_sort(cudaStream_t stream, const SortBuffers &buf, int start_buf, uint64_t n, const SortPlan &plan,
               void *workspace, uint32_t *err_flag, uint64_t *launches, cudaEvent_t after_hist)
{
    if (n == 0 || plan.n_passes == 0) {
        if (after_hist) KASA_CUDA_TRY(cudaEventRecord(after_hist, stream));
        return start_buf;
    }
    const bool has_w1 = buf.w1[0] != nullptr;
    auto *ghist = reinterpret_cast<unsigned long long *>(workspace);
    auto *ticket = reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(workspace) + (size_t)kMaxPasses * kRadix * 8);
    auto *status = reinterpret_cast<uint64_t *>(reinterpret_cast<char *>(workspace) + (size_t)kMaxPasses * kRadix * 8 + 256);
    const uint64_t tiles = div_up<uint64_t>(n, kSortTile);
    KASA_REQUIRE(tiles < (1ull << 32), -1, "radix_sort: too many records for one call");

    PlanDev pd{};
    pd.n_passes = plan.n_passes;
    bool need[3] = {false, false, false};
    for (int p = 0; p < plan.n_passes; ++p) {
        pd.src[p] = plan.pass[p].src;
        pd.shift[p] = plan.pass[p].shift;
        pd.mask[p] = plan.pass[p].mask;
        need[plan.pass[p].src] = true;
    }
    KASA_CUDA_TRY(cudaMemsetAsync(ghist, 0, (size_t)kMaxPasses * kRadix * 8 + 256, stream));
    {
        int grid = kSMs * 4;
        uint64_t want = div_up<uint64_t>(n, 512);
        if (want < (uint64_t)grid) grid = (int)want;
        size_t sm = (size_t)plan.n_passes * kRadix * 4;
        const bool only_w0 = need[0] && !need[1] && !need[2];
        if (only_w0 && plan.n_passes == 4) hist_w0_kernel<4><<<grid, 512, 0, stream>>>(buf.w0[start_buf], n, pd, ghist);
        else if (only_w0 && plan.n_passes == 8) hist_w0_kernel<8><<<grid, 512, 0, stream>>>(buf.w0[start_buf], n, pd, ghist);
        else if (has_w1)
            hist_kernel<true><<<grid, 512, sm, stream>>>(buf.w0[start_buf], buf.w1[start_buf], buf.v[start_buf], n, pd,
                                                         need[0], need[1], need[2], ghist);
        else
            hist_kernel<false><<<grid, 512, sm, stream>>>(buf.w0[start_buf], nullptr, buf.v[start_buf], n, pd, need[0],
                                                          need[1], need[2], ghist);
        hist_scan_kernel<<<plan.n_passes, kRadix, 0, stream>>>(ghist);
        *launches += 2;
    }
    if (after_hist) KASA_CUDA_TRY(cudaEventRecord(after_hist, stream));
    constexpr size_t sm_pass = onesweep_smem<kSortBlock, kSortItems>();
    // function attributes are per device: one process may drive several GPUs (host CLI --gpus, LocalShardGroup)
    static bool attr_done[64] = {};
    int dev = 0;
    KASA_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_done[dev]) {
        KASA_CUDA_TRY(cudaFuncSetAttribute(onesweep_kernel<kSortBlock, kSortItems, false>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_pass));
        KASA_CUDA_TRY(cudaFuncSetAttribute(onesweep_kernel<kSortBlock, kSortItems, true>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_pass));
        KASA_CUDA_TRY(cudaFuncSetAttribute(onesweep_kernel<kSortBlock, kSortItems, false>,
                                           cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        KASA_CUDA_TRY(cudaFuncSetAttribute(onesweep_kernel<kSortBlock, kSortItems, true>,
                                           cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        if (dev >= 0 && dev < 64) attr_done[dev] = true;
    }
    static const uint32_t ahead = [] { const char *e = getenv("KASA_TUNE_SORT_PREFETCH"); return e ? (uint32_t)atoi(e) : kPrefetchAhead; }();
    int cur = start_buf;
    for (int p = 0; p < plan.n_passes; ++p) {
        KASA_CUDA_TRY(cudaMemsetAsync(ticket, 0, 4, stream));
        KASA_CUDA_TRY(cudaMemsetAsync(status, 0, tiles * kRadix * 8, stream));
        const int nxt = cur ^ 1;
        if (has_w1)
            onesweep_kernel<kSortBlock, kSortItems, true><<<(unsigned)tiles, kSortBlock, sm_pass, stream>>>(
                buf.w0[cur], buf.w1[cur], buf.v[cur], buf.w0[nxt], buf.w1[nxt], buf.v[nxt], n, plan.pass[p].src,
                plan.pass[p].shift, plan.pass[p].mask, ghist + (size_t)p * kRadix, status, ticket, err_flag, ahead);
        else
            onesweep_kernel<kSortBlock, kSortItems, false><<<(unsigned)tiles, kSortBlock, sm_pass, stream>>>(
                buf.w0[cur], nullptr, buf.v[cur], buf.w0[nxt], nullptr, buf.v[nxt], n, plan.pass[p].src,
                plan.pass[p].shift, plan.pass[p].mask, ghist + (size_t)p * kRadix, status, ticket, err_flag, ahead);
        *launches += 1;
        cur = nxt;
    }
    KASA_CUDA_TRY(cudaGetLastError());
    return cur;
}

}  // namespace kasa
